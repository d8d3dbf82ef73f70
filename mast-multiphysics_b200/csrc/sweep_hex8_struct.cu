// sweep_hex8_struct.cu -- HEX8 linear-elastic residual + Jacobian sweep for ijk-numbered (build_cube) meshes with
// implicit connectivity (sm_100a, fp64).
//
// Same device counterpart as sweep_hex8.cu -- the element loop of MAST::NonlinearImplicitAssembly::
// residual_and_jacobian (/root/reference/src/base/nonlinear_implicit_assembly.cpp:111-186) with
// StructuralElement3D::internal_residual (src/elasticity/solid_element_3d.cpp:135-423, linear strain) and the
// isotropic D of src/property_cards/isotropic_material_property_card.cpp:512-528 -- for the meshes libMesh's
// MeshTools::Generation::build_cube produces (examples/structural/base/bracket_3d_model.h:172-177): node id
// i + j(nx+1) + k(nx+1)(ny+1), element id i + j nx + k nx ny.  The host recognises that numbering (context.cu) and
// uses this kernel for the R/J sweep (nv = 3 or MAST's six-variable system, constant or SIMP modulus); any other mesh
// and the thermoelastic load take the plan-driven kernel of sweep_hex8.cu.
// Geometry is general (every element's Jacobian is evaluated at its 8 Gauss points); only the integer bookkeeping
// is implicit, so no plan is streamed from HBM.
//
// Row-owner streaming gather.  A work unit is a column of <= 7x7 owned nodes over a z-range; the CTA (16 warps) sweeps
// it plane by plane, visiting the <= 8x8 elements of each element plane that touch its nodes (elements on the column
// boundary are visited by the neighbouring columns too).  Per element plane:
//   compute : warp w owns a 2x2 block of elements.  Geometry: lane = (element, qp) -> J^-1 and w = JxW E staged in
//             shared memory.  Matrix: all 32 lanes on one element, S_ij[a][b] = sum_q w g_a[i] g_b[j] for the six
//             component pairs i<=j with 12 mma.sync.m8n8k4.f64 (rows/columns = the element's nodes in ijk order).
//             Accumulate: the 3x3 block S_AB of every node pair goes to structure-of-arrays accumulators in shared
//             memory, arena[(direction A->B, component)][node column of A], for the lexicographically first node A of
//             the pair -- addresses are lane constants plus the element's column index.  The four elements sharing a
//             node column accumulate in colour order (x parity + 2 y parity), enforced by one counter per column
//             (acquire/release): fixed summation order, no atomics, no CTA barrier inside the plane.
//   flush   : node plane s is complete once element plane s is done.  Eight lanes per owned node, four passes over its
//             <= 27 neighbour directions (per-pass constants and the node table live in shared memory, nothing is held
//             in registers across the element phase): the pair block of each neighbour (forward pairs stored at the
//             node, backward pairs at the neighbour, transposed) becomes K_AB = lam S + mu S^T + mu tr(S) I with libMesh's
//             per-element Dirichlet semantics and is stored -- every CSR value exactly once, consecutive lanes writing
//             consecutive column blocks of a row; f_A = sum_B K_AB u_B by a fixed-order butterfly over the eight lanes.
//             K_e is never formed.
//   zero    : the accumulators of the flushed plane are cleared for reuse two planes later.
// Nodal coordinates / displacements of plane s+2 are fetched with cp.async while plane s is processed.
#include <cuda_runtime.h>

#include "kernels.h"

namespace mastb200 {
namespace {

constexpr int SW = kStructWarps;       // warps per CTA: 4 x (SW / 4) blocks of 2x2 elements
constexpr int ST = SW * 32;
constexpr int TCX = 9;                 // node columns per tile row (8 elements + 1)
constexpr int TCY = kStructTileY + 1;  // node columns per tile column
constexpr int NS = TCX * TCY;          // node-column slots
constexpr int NSB = (NS + 3) & ~3;     // padded byte-array row
static_assert(NS == kStructSlots, "layout header out of date");
constexpr int QS = 10;                 // doubles staged per (element, qp): inv[3][3], w
constexpr int GEO_ROUNDS = kStructGeoRounds;  // geometry of a warp's 2x2 block in one round of four elements or two rounds of two
constexpr int STAGE_EL = 4 / GEO_ROUNDS;        // elements staged per warp at a time
constexpr double GP = 0.57735026918962576451;

// shared-memory carve-up (bytes)
constexpr int O_ARENA = 0;
constexpr int ARENA_BYTES = kStructArenaDoubles * 8;
constexpr int O_XYZ = O_ARENA + ARENA_BYTES;            // [3 planes][3][NS]
constexpr int O_U = O_XYZ + 3 * 3 * NS * 8;             // [4 planes][3][NS]
constexpr int O_STAGE = (O_U + 4 * 3 * NS * 8 + 15) & ~15;  // [SW][STAGE_EL el][8 qp][QS]
constexpr int O_CNT = O_STAGE + SW * STAGE_EL * 8 * QS * 8;    // [NS] int32 (+pad)
constexpr int O_CM = O_CNT + NSB * 4;                   // [4 planes][NSB] uint8
constexpr int O_TAB = O_CM + 4 * NSB;                   // [124] int32: array offsets within an IN / CR region
constexpr int O_RHO = O_TAB + 124 * 4;                  // [3 planes][NS] nodal density (SIMP variant)
constexpr int O_NODE = O_RHO + 3 * NS * 8;              // [64] int2: owned nodes of the unit (in-plane id, slot | boundary flags)
constexpr int O_FPK = O_NODE + 64 * 8;                  // [4 passes][8 lanes] packed flush constants
constexpr int O_ROT = O_FPK + 32 * 4;                   // [2 planes][NS] sum of 1e-20 * max(diag K_e) per node (nv = 6)
constexpr int O_MISC = O_ROT + 2 * NS * 8;
constexpr int SMEM_BYTES = O_MISC + 16;                 // variants without sensitivities end here
constexpr int O_ZERO = (SMEM_BYTES + 15) & ~15;         // nv = 6: zeros, the source of the bulk stores of the rotational rows
constexpr int ZERO_BYTES = 18 * 27 * 8;                 //   (three rows of 6 x 27 values: the longest run of one node)
constexpr int SMEM_BYTES_NV6 = O_ZERO + ZERO_BYTES;     // (the same bytes hold the adjoint planes of the sensitivity variant)
constexpr int O_ADJ = SMEM_BYTES;                       // [3 planes][3][NS] adjoint (fused density sensitivities)
constexpr int O_GACC = O_ADJ + 3 * 3 * NS * 8;          // [2 planes][NS] G_n = adj . dR/d(rho_n) accumulators
constexpr int O_GQ = O_GACC + 2 * NS * 8;               // [SW][4 el][8 nodes] nodal sums of the point values s_q
constexpr int SMEM_BYTES_SENS = O_GQ + SW * 4 * 8 * 8;
static_assert(O_STAGE % 16 == 0 && O_XYZ % 16 == 0 && O_RHO % 8 == 0 && O_NODE % 8 == 0 && O_ROT % 8 == 0, "alignment");
static_assert(SMEM_BYTES <= (kStructCtasPerSm == 1 ? 232448 : 115712), "shared-memory layout exceeds the per-CTA budget");
static_assert(kStructCtasPerSm > 1 || SMEM_BYTES_SENS <= 232448, "shared-memory layout of the sensitivity variant exceeds 227 KB");
static_assert(SMEM_BYTES_NV6 <= (kStructCtasPerSm == 1 ? 232448 : 115712), "shared-memory layout of the nv = 6 variant exceeds the per-CTA budget");

// Accumulator regions as 16-byte aligned ranges [lo, hi) of the arena (doubles): the bulk-copy clear may touch the padding
// between regions but nothing that is live.  kTmaClear is false for a layout without that padding.
constexpr int even_dn(int x) { return x & ~1; }
constexpr int even_up(int x) { return (x + 1) & ~1; }
constexpr int IN0_LO = 0, IN0_HI = even_up(42 * NS);
constexpr int IN1_LO = even_dn(kStructIn1Off), IN1_HI = even_up(kStructIn1Off + 42 * NS);
constexpr int CR0_LO = even_dn(kStructCr0Off), CR0_HI = even_up(kStructCr0Off + 81 * NS);
constexpr int CR1_LO = even_dn(kStructCr1Off), CR1_HI = even_up(kStructCr1Off + 81 * NS);
constexpr bool kTmaClear = IN0_HI <= kStructIn1Off && IN1_LO >= 42 * NS && IN1_HI <= kStructCr0Off &&
                           CR0_LO >= kStructIn1Off + 42 * NS && CR0_HI <= kStructCr1Off && CR1_LO >= kStructCr0Off + 81 * NS &&
                           CR1_HI <= kStructArenaDoubles && CR0_HI - CR0_LO <= kStructZeroDoubles && O_ARENA % 16 == 0;

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ int ld_acquire_s32(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_s32(int32_t* p, int v) {
  asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
// 32-bit shared-memory addressing for the accumulators: one integer add per access instead of 64-bit pointer arithmetic
__device__ __forceinline__ double lds64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void stg256(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void stg128(double* p, double a, double b) {
  asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// TMA bulk store shared -> global (async proxy; both addresses and the size multiples of 16 bytes), one bulk group per call
__device__ __forceinline__ void bulk_store(void* gmem_dst, unsigned smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\tcp.async.bulk.commit_group;"
               ::"l"(gmem_dst), "r"(smem_src), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier; mbarrier helpers (one arriving thread)
__device__ __forceinline__ void bulk_load(unsigned smem_dst, const void* gmem_src, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gmem_src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
  unsigned ok = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Sm {
  double* arena;
  double* xyz;
  double* u;
  double* stage;
  int32_t* cnt;
  uint8_t* cm;
  int32_t* tab;
  double* rho;
  int2* node;
  int32_t* fpk;
  double* rot;
  double* adj;
  double* gacc;
  double* gq;
  int* misc;
  unsigned zero32;  // nv = 6: shared-space address of the zero block
};

// in-plane pair direction (A -> B, B after A in ijk order): 0 self, 1 (+1,0), 2 (-1,+1), 3 (0,+1), 4 (+1,+1)
__device__ __forceinline__ int inplane_dir(int di, int dj) { return dj == 0 ? (di == 0 ? 0 : 1) : 3 + di; }
// index of the array (direction, component) within an IN region (42 arrays) or CR region (81 arrays, tab offset 42)
__device__ __forceinline__ int in_array(int dir, int comp) { return dir == 0 ? comp : 6 + (dir - 1) * 9 + comp; }
__device__ __forceinline__ int cr_array(int di, int dj, int comp) { return 42 + ((dj + 1) * 3 + (di + 1)) * 9 + comp; }

struct LaneK {
  double d0[2], d1[2], d2;   // reference gradients of N_m at qps q4 (h = 0) and q4 + 4 (h = 1)
  int off[2][12];            // accumulate BYTE offsets per plane parity: region + array + column offset of the pair's first node
  unsigned pred;             // bit bb*6+k: this lane adds C[k][bb]
};

// Row / column index m of the DMMA tile is the element node (dx = m&1, dy = (m>>1)&1, dz = m>>2).
__device__ __forceinline__ void make_lane(const StructParams& p, const int32_t* tab, int lane, LaneK& lk) {
  const int m = lane >> 2, q4 = lane & 3;
  const double sx = (m & 1) ? 1. : -1., sy = (m & 2) ? 1. : -1., sz = (m & 4) ? 1. : -1.;
  const double fx = 1. + sx * ((q4 & 1) ? GP : -GP), fy = 1. + sy * ((q4 & 2) ? GP : -GP);
  const double fz0 = 1. - sz * GP, fz1 = 1. + sz * GP;
  lk.d0[0] = 0.125 * sx * fy * fz0; lk.d0[1] = 0.125 * sx * fy * fz1;
  lk.d1[0] = 0.125 * sy * fx * fz0; lk.d1[1] = 0.125 * sy * fx * fz1;
  lk.d2 = 0.125 * sz * fx * fy;
  lk.pred = 0u;
  const int KI[6] = {0, 1, 2, 0, 0, 1}, KJ[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
  for (int bb = 0; bb < 2; ++bb) {
    const int b = 2 * q4 + bb;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      int lo = m, hi = b, ci = KI[k], cj = KJ[k];
      bool on = true;
      int comp;
      if (m == b) {
        comp = k;  // self block: the six components i <= j
      } else {
        if (m > b) {  // C_ij[m][b] = S_(b,m)[j][i]
          lo = b; hi = m; ci = KJ[k]; cj = KI[k];
          if (KI[k] == KJ[k]) on = false;  // the lane (b, m) holds the same value
        }
        comp = ci * 3 + cj;
      }
      const int lx = lo & 1, ly = (lo >> 1) & 1, lz = lo >> 2;
      const int ox = (hi & 1) - lx, oy = ((hi >> 1) & 1) - ly, oz = (hi >> 2) - lz;
      const int so = ly * TCX + lx;
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        int o;
        if (oz == 0) {
          const int plane_par = lz ? (par ^ 1) : par;  // node plane s (bottom) lives in IN[s&1], plane s+1 in the other
          o = (plane_par ? p.in1_off : 0) + tab[in_array(inplane_dir(ox, oy), comp)] + so;
        } else {
          o = (par ? p.cr1_off : p.cr0_off) + tab[cr_array(ox, oy, comp)] + so;
        }
        lk.off[par][bb * 6 + k] = o * 8;
      }
      if (on) lk.pred |= 1u << (bb * 6 + k);
    }
  }
}

// ---- geometry: lane = (element el of the warp's 2x2 block, quadrature point q) ------------------------------------
// SENS: the lane also evaluates s_q = JxW E'(rho_q) eps(adj)^T D_1 eps(u) (adj zeroed on constrained dofs, the reference's
// constrain_element_vector) and the eight lanes of an element turn the point values into the element's nodal sums
// g_a = sum_q N_a(q) s_q (q ascending), left in gq[a] for the ordered accumulation of the element phase: the density
// sensitivities G_n = adj . dR/d(rho_n) of all design variables come out of the same sweep as R and J
// (assembly_base.cpp:594-639, solid_element_3d.cpp:566-788 with dD/dp of the SIMP law).
template <bool SIMP, bool SENS>
__device__ __forceinline__ void geometry_phase(const StructParams& p, const Sm& sm, int s, int base, int q, double* out,
                                               double* gq, bool live) {
  const double xi = (q & 1) ? GP : -GP, eta = (q & 2) ? GP : -GP, zeta = (q & 4) ? GP : -GP;
  const double xm = 1. - xi, xp = 1. + xi, ym = 1. - eta, yp = 1. + eta, zm = 1. - zeta, zp = 1. + zeta;
  double Jm[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};  // Jm[r][c] = d x_c / d xi_r
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const double fx = (m & 1) ? xp : xm, fy = (m & 2) ? yp : ym, fz = (m & 4) ? zp : zm;
    const double d0 = ((m & 1) ? 0.125 : -0.125) * fy * fz, d1 = ((m & 2) ? 0.125 : -0.125) * fx * fz,
                 d2 = ((m & 4) ? 0.125 : -0.125) * fx * fy;
    const double* X = sm.xyz + (((m & 4) ? s + 1 : s) % 3) * 3 * NS + base + ((m >> 1) & 1) * TCX + (m & 1);
    const double x0 = X[0], x1 = X[NS], x2 = X[2 * NS];
    Jm[0][0] = fma(d0, x0, Jm[0][0]); Jm[0][1] = fma(d0, x1, Jm[0][1]); Jm[0][2] = fma(d0, x2, Jm[0][2]);
    Jm[1][0] = fma(d1, x0, Jm[1][0]); Jm[1][1] = fma(d1, x1, Jm[1][1]); Jm[1][2] = fma(d1, x2, Jm[1][2]);
    Jm[2][0] = fma(d2, x0, Jm[2][0]); Jm[2][1] = fma(d2, x1, Jm[2][1]); Jm[2][2] = fma(d2, x2, Jm[2][2]);
  }
  const double c00 = Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1];
  const double c01 = Jm[1][2] * Jm[2][0] - Jm[1][0] * Jm[2][2];
  const double c02 = Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0];
  const double det = Jm[0][0] * c00 + Jm[0][1] * c01 + Jm[0][2] * c02;
  const double idet = __drcp_rn(det);
  double inv[3][3];  // inv[c][r] = d xi_r / d x_c
  inv[0][0] = c00 * idet; inv[1][0] = c01 * idet; inv[2][0] = c02 * idet;
  inv[0][1] = (Jm[0][2] * Jm[2][1] - Jm[0][1] * Jm[2][2]) * idet;
  inv[1][1] = (Jm[0][0] * Jm[2][2] - Jm[0][2] * Jm[2][0]) * idet;
  inv[2][1] = (Jm[0][1] * Jm[2][0] - Jm[0][0] * Jm[2][1]) * idet;
  inv[0][2] = (Jm[0][1] * Jm[1][2] - Jm[0][2] * Jm[1][1]) * idet;
  inv[1][2] = (Jm[0][2] * Jm[1][0] - Jm[0][0] * Jm[1][2]) * idet;
  inv[2][2] = (Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0]) * idet;
  double2* o = reinterpret_cast<double2*>(out);
  if (live) {
    o[0] = make_double2(inv[0][0], inv[0][1]);
    o[1] = make_double2(inv[0][2], inv[1][0]);
    o[2] = make_double2(inv[1][1], inv[1][2]);
    o[3] = make_double2(inv[2][0], inv[2][1]);
  }
  double E = p.E0, rho_q = 0.;
  if (SIMP) {  // modulus of the interpolated nodal density (examples/structural/example_6/example_6.cpp:100-137)
    double r = 0.;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const double fx = (m & 1) ? xp : xm, fy = (m & 2) ? yp : ym, fz = (m & 4) ? zp : zm;
      r = fma(0.125 * fx * fy * fz, sm.rho[(((m & 4) ? s + 1 : s) % 3) * NS + base + ((m >> 1) & 1) * TCX + (m & 1)], r);
    }
    E = p.E0 * (p.rho_min + (1. - p.rho_min) * simp_pow(r, p.penalty));
    rho_q = r;
  }
  if (live) o[4] = make_double2(inv[2][2], det * E);  // Gauss weights are 1 for the 2x2x2 rule
  if (SENS) {
    double sq = 0.;
    if (live) {
      // displacement gradients through the reference gradients: Href[i][r] = sum_m u_m,i dN_m/dxi_r (72 fma), then
      // H[i][j] = sum_r Href[i][r] dxi_r/dx_j (27 fma) instead of physical gradients of all eight nodes; the adjoint's
      // gradient equals the state's when the adjoint IS the state (compliance) and no dof of the element is constrained
      const bool same = p.adj == p.X;
      const double* adjp = same ? sm.u : sm.adj;
      double Hr[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}}, Ar[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};
      unsigned cmany = 0u;
      if (p.cmask) {
#pragma unroll
        for (int m = 0; m < 8; ++m)
          cmany |= sm.cm[((((m & 4) ? s + 1 : s)) & 3) * NSB + base + ((m >> 1) & 1) * TCX + (m & 1)];
      }
      const bool need_adj = !same || cmany != 0u;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const double fx = (m & 1) ? xp : xm, fy = (m & 2) ? yp : ym, fz = (m & 4) ? zp : zm;
        const double d0 = ((m & 1) ? 0.125 : -0.125) * fy * fz, d1 = ((m & 2) ? 0.125 : -0.125) * fx * fz,
                     d2 = ((m & 4) ? 0.125 : -0.125) * fx * fy;
        const int z = (m & 4) ? s + 1 : s, col = base + ((m >> 1) & 1) * TCX + (m & 1);
        const double* U = sm.u + (z & 3) * 3 * NS + col;
        const double u0 = U[0], u1 = U[NS], u2 = U[2 * NS];
        Hr[0][0] = fma(u0, d0, Hr[0][0]); Hr[0][1] = fma(u0, d1, Hr[0][1]); Hr[0][2] = fma(u0, d2, Hr[0][2]);
        Hr[1][0] = fma(u1, d0, Hr[1][0]); Hr[1][1] = fma(u1, d1, Hr[1][1]); Hr[1][2] = fma(u1, d2, Hr[1][2]);
        Hr[2][0] = fma(u2, d0, Hr[2][0]); Hr[2][1] = fma(u2, d1, Hr[2][1]); Hr[2][2] = fma(u2, d2, Hr[2][2]);
        if (need_adj) {
          const double* L = adjp + (same ? (z & 3) : (z % 3)) * 3 * NS + col;
          const unsigned cm = p.cmask ? sm.cm[(z & 3) * NSB + col] : 0u;
          const double l0 = (cm & 1u) ? 0. : L[0], l1 = (cm & 2u) ? 0. : L[NS], l2 = (cm & 4u) ? 0. : L[2 * NS];
          Ar[0][0] = fma(l0, d0, Ar[0][0]); Ar[0][1] = fma(l0, d1, Ar[0][1]); Ar[0][2] = fma(l0, d2, Ar[0][2]);
          Ar[1][0] = fma(l1, d0, Ar[1][0]); Ar[1][1] = fma(l1, d1, Ar[1][1]); Ar[1][2] = fma(l1, d2, Ar[1][2]);
          Ar[2][0] = fma(l2, d0, Ar[2][0]); Ar[2][1] = fma(l2, d1, Ar[2][1]); Ar[2][2] = fma(l2, d2, Ar[2][2]);
        }
      }
      double H[3][3], A[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          H[i][j] = Hr[i][0] * inv[j][0] + Hr[i][1] * inv[j][1] + Hr[i][2] * inv[j][2];
          A[i][j] = need_adj ? Ar[i][0] * inv[j][0] + Ar[i][1] * inv[j][1] + Ar[i][2] * inv[j][2] : H[i][j];
        }
      const double e0 = H[0][0], e1 = H[1][1], e2 = H[2][2];
      const double e3 = H[0][1] + H[1][0], e4 = H[1][2] + H[2][1], e5 = H[0][2] + H[2][0];
      const double a0 = A[0][0], a1 = A[1][1], a2 = A[2][2];
      const double a3 = A[0][1] + A[1][0], a4 = A[1][2] + A[2][1], a5 = A[0][2] + A[2][0];
      const double dens = p.lam1 * (a0 + a1 + a2) * (e0 + e1 + e2) +
                          p.mu1 * (2. * (a0 * e0 + a1 * e1 + a2 * e2) + a3 * e3 + a4 * e4 + a5 * e5);
      const double pm1 = p.penalty - 1.;
      const double rpow = pm1 == 3. ? rho_q * rho_q * rho_q : (pm1 == 2. ? rho_q * rho_q : (pm1 == 1. ? rho_q : (pm1 == 0. ? 1. : pow(rho_q, pm1))));
      sq = det * p.E0 * (1. - p.rho_min) * p.penalty * rpow * dens;
    }
    // lane a of the element's eight lanes ends with sum_q N_a(q) s_q; a in ijk order (dx = a & 1, dy = (a >> 1) & 1, dz = a >> 2):
    // the shape functions are a tensor product, so the sum is three butterfly stages (x, y, z) over the eight lanes
    // (sweep_hex8_sens_struct.cu does the same); fixed order, bitwise reproducible
    const int a = q;
    double acc = sq;
    {
      const double ws = 0.5 * (1. + GP), wo = 0.5 * (1. - GP);
#pragma unroll
      for (int bit = 1; bit < 8; bit <<= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, acc, bit);
        acc = ws * acc + wo * other;
      }
    }
    if (live) gq[a] = acc;
  }
}

// ---- matrix + ordered accumulate: all 32 lanes on one element ------------------------------------------------------
template <int PAR, bool NV6, bool SENS>
__device__ __forceinline__ void element_phase(const StructParams& p, const Sm& sm, const LaneK& lk, unsigned arena32, int rankpack,
                                              const double* st, int lex, int ley, int color, int lane, const double* gq) {
  const int m = lane >> 2, q4 = lane & 3;
  double A[3][2], w[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const double2* s2 = reinterpret_cast<const double2*>(st + (q4 + 4 * h) * QS);
    const double2 v0 = s2[0], v1 = s2[1], v2 = s2[2], v3 = s2[3], v4 = s2[4];
    A[0][h] = v0.x * lk.d0[h] + v0.y * lk.d1[h] + v1.x * lk.d2;
    A[1][h] = v1.y * lk.d0[h] + v2.x * lk.d1[h] + v2.y * lk.d2;
    A[2][h] = v3.x * lk.d0[h] + v3.y * lk.d1[h] + v4.x * lk.d2;
    w[h] = v4.y;
  }
  double B[3][2], C[6][2];
#pragma unroll
  for (int j = 0; j < 3; ++j) { B[j][0] = A[j][0] * w[0]; B[j][1] = A[j][1] * w[1]; }
#pragma unroll
  for (int k = 0; k < 6; ++k) { C[k][0] = 0.; C[k][1] = 0.; }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    dmma8x8x4(C[0][0], C[0][1], A[0][h], B[0][h]);
    dmma8x8x4(C[1][0], C[1][1], A[1][h], B[1][h]);
    dmma8x8x4(C[2][0], C[2][1], A[2][h], B[2][h]);
    dmma8x8x4(C[3][0], C[3][1], A[0][h], B[1][h]);
    dmma8x8x4(C[4][0], C[4][1], A[0][h], B[2][h]);
    dmma8x8x4(C[5][0], C[5][1], A[1][h], B[2][h]);
  }
  double mx = 0.;
  if (NV6) {
    // 1e-20*max(diag K_e) goes on the rotational diagonal of the 6-variable system (solid_element_3d.cpp:413-417);
    // diag of K_e from S_aa: the lane with b == m holds it (the rotational diagonal itself is zero, so the max is >= 0)
    const int bself = m - 2 * q4;
    if (bself == 0 || bself == 1) {
      const double s0 = C[0][bself & 1], s1 = C[1][bself & 1], s2 = C[2][bself & 1];
      const double trm = p.mu1 * (s0 + s1 + s2);
      mx = fmax(mx, fmax((p.lam1 + p.mu1) * s0 + trm, fmax((p.lam1 + p.mu1) * s1 + trm, (p.lam1 + p.mu1) * s2 + trm)));
    }
    // warp maximum of non-negative doubles: their bit patterns order like unsigned integers, so two redux.sync (the high
    // words, then the low words of the lanes that hold the maximal high word) replace a five-stage shuffle butterfly
    {
      const unsigned hi = (unsigned)__double2hiint(mx), lo = (unsigned)__double2loint(mx);
      const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
      const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
      mx = __hiloint2double((int)mhi, (int)mlo);
    }
  }
  // ordered accumulation: the (up to) four elements around a node column add in colour order
  const int base = ley * TCX + lex;
  const bool checker = q4 == 0 && m < 4;
  const int rank = (rankpack >> (2 * color)) & 3;
  int32_t* cnt = sm.cnt + base + (m >> 1) * TCX + (m & 1);
  if (checker)
    while (ld_acquire_s32(cnt) != rank) __nanosleep(20);
  __syncwarp();
  // the twelve addresses of a lane are distinct: all loads are issued before the first store so that their latencies
  // overlap (a load-add-store chain per value would serialise twelve shared-memory round trips)
  const unsigned ar = arena32 + base * 8;
  double old[12];
#pragma unroll
  for (int t = 0; t < 12; ++t) {
    old[t] = 0.;
    if ((lk.pred >> t) & 1u) old[t] = lds64(ar + lk.off[PAR][t]);
  }
#pragma unroll
  for (int t = 0; t < 12; ++t)
    if ((lk.pred >> t) & 1u) sts64(ar + lk.off[PAR][t], old[t] + C[t % 6][t / 6]);
  if (SENS && q4 == 0) {  // node m of the element: its nodal sensitivity sum joins the node's accumulator (plane s or s+1)
    double* ga = sm.gacc + (PAR ^ (m >> 2)) * NS + base + ((m >> 1) & 1) * TCX + (m & 1);
    *ga += gq[m];
  }
  if (NV6 && checker) {  // both nodes of the column (node planes s and s+1)
    const int col = base + (m >> 1) * TCX + (m & 1);
    sm.rot[PAR * NS + col] += 1.0e-20 * mx;
    sm.rot[(PAR ^ 1) * NS + col] += 1.0e-20 * mx;
  }
  __syncwarp();
  if (checker) st_release_s32(cnt, rank + 1);
}

// ---- flush of node plane s: eight lanes per owned node, four passes over its <= 27 neighbour directions ---------------
// Consecutive lanes of a node write consecutive column blocks of its CSR rows (8 x 24 bytes per store instruction).
// Everything that depends on (lane, pass) only is decoded once per kernel into one packed word per pass:
//   bits 0-4 neighbour column shift + 10, 5-6 dk + 1, 7-12 boundary flags that remove the neighbour, 13 self,
//   14 transposed (backward pair, stored at the neighbour), 15 direction exists (d < 27), 16-23 first entry of the
//   direction's nine array offsets in the table, 24-28 d.
__device__ __forceinline__ int flush_pack(int t, int l8) {
  {
    const int d0 = 8 * t + l8;
    const int d = d0 < 27 ? d0 : 13;
    const int dk = d / 9 - 1, dj = (d / 3) % 3 - 1, di = d % 3 - 1;
    const bool self = d == 13, fwd = d > 13, tr = !(fwd || self);
    const int sdi = tr ? -di : di, sdj = tr ? -dj : dj;
    const int a0 = dk == 0 ? in_array(inplane_dir(sdi, sdj), 0) : cr_array(sdi, sdj, 0);
    const int kill = (di < 0 ? 1 : 0) | (di > 0 ? 2 : 0) | (dj < 0 ? 4 : 0) | (dj > 0 ? 8 : 0) | (dk < 0 ? 16 : 0) | (dk > 0 ? 32 : 0);
    return (dj * TCX + di + 10) | ((dk + 1) << 5) | (kill << 7) | (self ? 1 << 13 : 0) | (tr ? 1 << 14 : 0) |
           (d0 < 27 ? 1 << 15 : 0) | (a0 << 16) | (d << 24);
  }
}

template <bool NV6>
__device__ __forceinline__ void flush_plane(const StructParams& p, const Sm& sm, unsigned arena32, int s,
                                            int n_own, int64_t plane_off, int warp, int lane) {
  const int g = warp * 4 + (lane >> 3), l8 = lane & 7;
  const bool valid = g < n_own;
  const int2 ni = sm.node[valid ? g : 0];  // in-plane node id, slot | xy boundary flags << 8
  const int64_t nb0 = valid ? p.nbr_ptr[plane_off + ni.x - (int64_t)p.oz0 * (p.nx + 1) * (p.ny + 1)] : 0;
  const int slotA = ni.y & 0xff;
  const int flags = (ni.y >> 8) | (s == 0 ? 16 : 0) | (s == p.nz ? 32 : 0);  // x0, xN, y0, yN, z0, zN
  const int par = s & 1;
  const unsigned cmn = p.cmask ? sm.cm[(s & 3) * NSB + slotA] : 0u;
  // CSR numbering of the node's neighbours: interior nodes have all 27 (column block = direction)
  int x0 = 0, y0 = 0, z0 = 0, xcnt = 3, ycnt = 3, rs = 81;
  double nadj = 8.;
  if (flags) {
    x0 = flags & 1; y0 = (flags >> 2) & 1; z0 = (flags >> 4) & 1;
    xcnt = 3 - x0 - ((flags >> 1) & 1); ycnt = 3 - y0 - ((flags >> 3) & 1);
    const int zcnt = 3 - z0 - ((flags >> 5) & 1);
    rs = 3 * xcnt * ycnt * zcnt;
    nadj = double((xcnt - 1) * (ycnt - 1) * (zcnt - 1));
  }
  const unsigned regIn = arena32 + ((par ? p.in1_off : 0) + slotA) * 8;
  const unsigned regUp = arena32 + ((par ? p.cr1_off : p.cr0_off) + slotA) * 8;
  const unsigned regDn = arena32 + ((par ? p.cr0_off : p.cr1_off) + slotA) * 8;
  double* Jrow = p.J + (NV6 ? 36 : 9) * nb0;
  const double rotA = NV6 ? sm.rot[par * NS + slotA] : 0.;
  double r0 = 0., r1 = 0., r2 = 0.;
#pragma unroll 1
  for (int t = 0; t < 4; ++t) {
    const int pk = sm.fpk[t * 8 + l8];
    const bool on = valid && (pk & (1 << 15)) && !(flags & (pk >> 7) & 63);
    if (on) {
      const int nsh = (pk & 31) - 10, dk = ((pk >> 5) & 3) - 1;
      const bool self = (pk >> 13) & 1, tr = (pk >> 14) & 1;
      const unsigned blk = (dk == 0 ? regIn : (dk > 0 ? regUp : regDn)) + (tr ? nsh * 8 : 0);
      const int32_t* tb = sm.tab + ((pk >> 16) & 255);
      double S[3][3];
      if (self) {
        S[0][0] = lds64(blk + tb[0] * 8); S[1][1] = lds64(blk + tb[1] * 8); S[2][2] = lds64(blk + tb[2] * 8);
        S[0][1] = S[1][0] = lds64(blk + tb[3] * 8); S[0][2] = S[2][0] = lds64(blk + tb[4] * 8);
        S[1][2] = S[2][1] = lds64(blk + tb[5] * 8);
      } else {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) S[i][j] = lds64(blk + tb[i * 3 + j] * 8);
      }
      // the block holds S_(row,nbr) or its transpose: K = lam*S + mu*S^T + mu*tr(S)*I either way, with lam/mu swapped
      const double l1 = tr ? p.mu1 : p.lam1, l2 = tr ? p.lam1 : p.mu1;
      const double trm = p.mu1 * (S[0][0] + S[1][1] + S[2][2]);
      double K[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) K[i][j] = l1 * S[i][j] + l2 * S[j][i] + (i == j ? trm : 0.);
      const int slotB = slotA + nsh;
      if (p.R) {  // internal force of linear elasticity: f_a += K_ab u_b (unconstrained element matrices)
        const double* U = sm.u + ((s + dk) & 3) * 3 * NS + slotB;
        const double u0 = U[0], u1 = U[NS], u2 = U[2 * NS];
        r0 += K[0][0] * u0 + K[0][1] * u1 + K[0][2] * u2;
        r1 += K[1][0] * u0 + K[1][1] * u1 + K[1][2] * u2;
        r2 += K[2][0] * u0 + K[2][1] * u1 + K[2][2] * u2;
      }
      if (p.J) {
        const unsigned cmm = p.cmask ? sm.cm[((s + dk) & 3) * NSB + slotB] : 0u;
        if (cmn | cmm) {  // libMesh constrain_element_matrix_and_vector: row/column zeroed, diagonal 1 per adjacent element
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              double k = K[i][j];
              if (((cmn >> i) | (cmm >> j)) & 1u) k = 0.;
              if (self && i == j && ((cmn >> i) & 1u)) k = nadj;
              K[i][j] = k;
            }
        }
        int c = (pk >> 24) & 31;
        if (flags) {
          const int d = c, ddk = d / 9, ddj = (d / 3) % 3, ddi = d % 3;
          c = ((ddk - z0) * ycnt + (ddj - y0)) * xcnt + (ddi - x0);
        }
        if (!NV6) {
          double* dst = Jrow + 3 * c;
          if (p.j_aligned32) {
            // a row piece of three values starts at an even or an odd double (9 nb0 + i rs + 3 c): one 16-byte store of the
            // aligned pair and one 8-byte store of the third value, chosen without a branch -- two store instructions
            // per row piece instead of three, and every 32-byte sector is touched by two instructions instead of three
            const int par0 = ((int)nb0 + c) & 1;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              double* d = dst + i * rs;
              const bool odd = ((par0 + i * rs) & 1) != 0;
              stg128(odd ? d + 1 : d, odd ? K[i][1] : K[i][0], odd ? K[i][2] : K[i][1]);
              *(odd ? d : d + 2) = odd ? K[i][0] : K[i][2];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j) dst[i * rs + j] = K[i][j];
          }
        } else {
          // 6-variable system, translational rows of this column block: [K_i0 K_i1 K_i2 0 0 0].  A row piece starts at a
          // multiple of 48 bytes (36 nb0 + 6 c + i * 6 nnb doubles), i.e. on a 32-byte boundary for every other piece:
          // one 32-byte and one 16-byte store instead of six 8-byte ones.  The rotational rows follow after the passes.
          double* dst = Jrow + 6 * c;
          const int rl = 2 * rs;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            double* d = dst + i * rl;
            if (p.j_aligned32) {  // chosen without a branch: neighbouring lanes alternate between the two cases
              const bool odd = ((c + i * (rs / 3)) & 1) != 0;
              stg128(odd ? d : d + 4, odd ? K[i][0] : 0., odd ? K[i][1] : 0.);
              stg256(odd ? d + 2 : d, odd ? K[i][2] : K[i][0], odd ? 0. : K[i][1], odd ? 0. : K[i][2], 0.);
            } else {  // a caller's device pointer that is only 8-byte aligned
              d[0] = K[i][0]; d[1] = K[i][1]; d[2] = K[i][2]; d[3] = 0.; d[4] = 0.; d[5] = 0.;
            }
          }
        }
      }
    }
  }
  if (NV6 && valid && p.J) {
    // rotational rows 3-5 of the node: one contiguous run of 18 nnb doubles, zero but for the three diagonal entries
    // (1e-20 sum of max diag K_e, or the adjacent-element count on a constrained dof): the eight lanes write it in
    // 32-byte chunks
    const int nnb = rs / 3, L = 18 * nnb;
    const int cself = ((1 - z0) * ycnt + (1 - y0)) * xcnt + (1 - x0);
    const int dpos0 = 6 * cself + 3, dpos1 = 6 * nnb + 6 * cself + 4, dpos2 = 12 * nnb + 6 * cself + 5;
    const double dv0 = (cmn & 8u) ? nadj : rotA, dv1 = (cmn & 16u) ? nadj : rotA, dv2 = (cmn & 32u) ? nadj : rotA;
    auto at = [&](int x) { return x == dpos0 ? dv0 : (x == dpos1 ? dv1 : (x == dpos2 ? dv2 : 0.)); };
    double* base = Jrow + 18 * nnb;
    if (p.j_aligned32) {
      // The run starts on a 16-byte boundary (36 nb0 + 18 nnb doubles) and is zero but for three values: the three aligned
      // pairs that hold a diagonal entry go out as 16-byte stores, the four zero stretches between them as TMA bulk stores
      // from the zero block (cp.async.bulk shared -> global: half of the matrix bytes of the six-variable layout never
      // pass the LSU).  The bulk stores and the pair stores touch disjoint bytes: no ordering between the proxies needed.
      const int P0 = dpos0 - 1, P1 = dpos1, P2 = dpos2 - 1;  // even positions of the three pairs
      if (l8 < 4) {
        const int lo = l8 == 0 ? 0 : (l8 == 1 ? P0 + 2 : (l8 == 2 ? P1 + 2 : P2 + 2));
        const int hi = l8 == 0 ? P0 : (l8 == 1 ? P1 : (l8 == 2 ? P2 : L));
        if (hi > lo) bulk_store(base + lo, sm.zero32, (unsigned)(hi - lo) * 8u);
      } else if (l8 == 4) stg128(base + P0, 0., dv0);
      else if (l8 == 5) stg128(base + P1, dv1, 0.);
      else if (l8 == 6) stg128(base + P2, 0., dv2);
    } else {
      for (int x = l8; x < L; x += 8) base[x] = at(x);
    }
  }
  if (p.R) {
    // deterministic butterfly over the 8 lanes of a node for the three sums at once: lanes 0, 2 and 4 end with r0, r1, r2
    const bool hi4 = (lane & 4) != 0, hi2 = (lane & 2) != 0;
    const double g0 = __shfl_xor_sync(0xffffffffu, hi4 ? r0 : r2, 4);
    const double g1 = __shfl_xor_sync(0xffffffffu, hi4 ? r1 : 0., 4);
    const double x0s = (hi4 ? r2 : r0) + g0, x1s = (hi4 ? 0. : r1) + g1;
    double y = (hi2 ? x1s : x0s) + __shfl_xor_sync(0xffffffffu, hi2 ? x0s : x1s, 2);
    y += __shfl_xor_sync(0xffffffffu, y, 1);
    if (valid && (l8 & 1) == 0 && l8 < 6) {
      const int i = l8 >> 1;
      p.R[(plane_off + ni.x) * (NV6 ? 6 : 3) + i] = ((cmn >> i) & 1u) ? 0. : y;
      if (NV6) p.R[(plane_off + ni.x) * 6 + 3 + i] = 0.;  // rotational dofs: no residual
    }
  }
  if (!NV6 && p.G && valid && l8 == 7) p.G[plane_off + ni.x] = sm.gacc[par * NS + slotA];  // fused density sensitivities
}

template <bool SIMP, bool NV6, bool SENS>
__global__ void __launch_bounds__(ST, kStructCtasPerSm) hex8_struct_kernel(const StructParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Sm sm;
  sm.arena = reinterpret_cast<double*>(smem_raw + O_ARENA);
  sm.xyz = reinterpret_cast<double*>(smem_raw + O_XYZ);
  sm.u = reinterpret_cast<double*>(smem_raw + O_U);
  sm.stage = reinterpret_cast<double*>(smem_raw + O_STAGE);
  sm.cnt = reinterpret_cast<int32_t*>(smem_raw + O_CNT);
  sm.cm = smem_raw + O_CM;
  sm.tab = reinterpret_cast<int32_t*>(smem_raw + O_TAB);
  sm.rho = reinterpret_cast<double*>(smem_raw + O_RHO);
  sm.node = reinterpret_cast<int2*>(smem_raw + O_NODE);
  sm.fpk = reinterpret_cast<int32_t*>(smem_raw + O_FPK);
  sm.rot = reinterpret_cast<double*>(smem_raw + O_ROT);
  sm.misc = reinterpret_cast<int*>(smem_raw + O_MISC);
  sm.adj = reinterpret_cast<double*>(smem_raw + O_ADJ);
  sm.gacc = reinterpret_cast<double*>(smem_raw + O_GACC);
  sm.gq = reinterpret_cast<double*>(smem_raw + O_GQ);
  sm.zero32 = (unsigned)__cvta_generic_to_shared(smem_raw + O_ZERO);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (NV6) {  // written once (generic proxy), read by the bulk stores (async proxy) from the first flush on
    for (int i = tid; i < ZERO_BYTES / 8; i += ST) reinterpret_cast<double*>(smem_raw + O_ZERO)[i] = 0.;
    fence_proxy_async_smem();
  }
  // bulk-copy clear of the accumulators: the mbarrier lives in the second half of the misc words
  const unsigned clear_bar = (unsigned)__cvta_generic_to_shared(smem_raw + O_MISC + 8);
  const bool tma_clear = kTmaClear && p.zeros != nullptr;
  unsigned clear_phase = 0u;
  bool clear_pending = false;
  if (tma_clear && tid == 0) mbar_init(clear_bar, 1u);
  if (tid < 123) sm.tab[tid] = p.arr_off[tid];
  if (tid >= 128 && tid < 160) sm.fpk[tid - 128] = flush_pack((tid - 128) >> 3, (tid - 128) & 7);
  __syncthreads();
  LaneK lk;
  make_lane(p, sm.tab, lane, lk);
  const unsigned arena32 = (unsigned)__cvta_generic_to_shared(sm.arena);
  double* my_stage = sm.stage + warp * (STAGE_EL * 8 * QS);
  const int bx2 = (warp & 3) * 2, by2 = (warp >> 2) * 2;  // origin of the warp's 2x2 element block

  for (;;) {
    if (tid == 0) sm.misc[0] = atomicAdd(p.counter, 1);
    __syncthreads();
    const int unit = sm.misc[0];
    if (unit >= p.n_units) break;
    const StructUnit u = p.units[unit];
    const int s_begin = u.z0 > 0 ? u.z0 - 1 : 0;
    const int64_t npx = p.nx + 1, plane_nodes = npx * (p.ny + 1);
    // per-thread constants of the unit: the state item this thread stages every plane, its flush node
    const int ncx = u.nex + 1, ncols = ncx * (u.ney + 1);
    const bool ld_on = tid < ncols * 6;
    int ld_dst = 0;
    int64_t ld_src = 0;
    bool ld_u = false;
    {
      const int col = tid / 6, c = tid - col * 6;
      const int ly = col / ncx, lx = col - ly * ncx;
      const int64_t gid0 = (u.e0x + lx) + (u.e0y + ly) * npx;
      ld_u = c >= 3;
      ld_dst = (ld_u ? c - 3 : c) * NS + ly * TCX + lx;
      ld_src = gid0 * (ld_u && NV6 ? 6 : 3) + (ld_u ? c - 3 : c);
    }
    const bool cm_on = p.cmask != nullptr && tid < ncols;
    int cm_dst = 0;
    int64_t cm_src = 0;
    {
      const int ly = tid / ncx, lx = tid - ly * ncx;
      cm_dst = ly * TCX + lx;
      cm_src = (u.e0x + lx) + (u.e0y + ly) * npx;
    }
    auto load_plane = [&](int z) {  // coordinates, displacements (and density) of node plane z, asynchronously
      if (SIMP && tid < ncols) {
        const int ly = tid / ncx, lx = tid - ly * ncx;
        cp_async8(sm.rho + (z % 3) * NS + ly * TCX + lx, p.rho + (int64_t)z * plane_nodes + (u.e0x + lx) + (u.e0y + ly) * npx);
      }
      if (ld_on) {
        if (ld_u) cp_async8(sm.u + (z & 3) * 3 * NS + ld_dst, p.X + (int64_t)z * plane_nodes * (NV6 ? 6 : 3) + ld_src);
        else cp_async8(sm.xyz + (z % 3) * 3 * NS + ld_dst, p.xyz + (int64_t)z * plane_nodes * 3 + ld_src);
      }
      if (SENS && p.adj != p.X)  // adjoint planes of the fused sensitivities (three buffers, like the coordinates)
        for (int t = tid; t < ncols * 3; t += ST) {
          const int col = t / 3, c = t - col * 3;
          const int ly = col / ncx, lx = col - ly * ncx;
          cp_async8(sm.adj + (z % 3) * 3 * NS + c * NS + ly * TCX + lx,
                    p.adj + ((int64_t)z * plane_nodes + (u.e0x + lx) + (u.e0y + ly) * npx) * 3 + c);
        }
      if (NS * 6 > ST && tid + ST < ncols * 6) {  // small CTAs: the few items beyond one per thread
        const int t = tid + ST, col = t / 6, c = t - col * 6;
        const int ly = col / ncx, lx = col - ly * ncx;
        const int64_t gid0 = (u.e0x + lx) + (u.e0y + ly) * npx;
        const bool isu = c >= 3;
        const int dst = (isu ? c - 3 : c) * NS + ly * TCX + lx;
        const int64_t src = gid0 * (isu && NV6 ? 6 : 3) + (isu ? c - 3 : c);
        if (isu) cp_async8(sm.u + (z & 3) * 3 * NS + dst, p.X + (int64_t)z * plane_nodes * (NV6 ? 6 : 3) + src);
        else cp_async8(sm.xyz + (z % 3) * 3 * NS + dst, p.xyz + (int64_t)z * plane_nodes * 3 + src);
      }
    };
    // accumulation ranks of this lane's node column for the four elements of the warp's block (2 bits per colour)
    int rankpack = 0;
    {
      const int m = lane >> 2, dx = m & 1, dy = (m >> 1) & 1;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int lex = bx2 + (c & 1), ley = by2 + (c >> 1);
        const int ox = lex - 1 + 2 * dx, oy = ley - 1 + 2 * dy;  // the other elements around the column
        const bool ex = ox >= 0 && ox < u.nex, ey = oy >= 0 && oy < u.ney;
        const int rank = ((ex && (c & 1)) ? 1 : 0) + ((ey && (c & 2)) ? 1 : 0) + ((ex && ey && c >= 2) ? 1 : 0);
        rankpack |= rank << (2 * c);
      }
    }
    // flush: warp w writes the owned nodes 4 w .. 4 w + 3, eight lanes each (flush_plane)
    const int n_own = u.nox * u.noy;
    if (tid < n_own) {
      const int oyi = tid / u.nox, oxi = tid - oyi * u.nox;
      const int ax = u.o0x + oxi, ay = u.o0y + oyi;
      const int flags = (ax == 0 ? 1 : 0) | (ax == p.nx ? 2 : 0) | (ay == 0 ? 4 : 0) | (ay == p.ny ? 8 : 0);
      sm.node[tid] = make_int2((int)(ax + ay * npx), ((ay - u.e0y) * TCX + (ax - u.e0x)) | (flags << 8));
    }

    // prologue: clear the accumulators, stage the first two node planes
    for (int i = tid; i < kStructArenaDoubles; i += ST) sm.arena[i] = 0.;
    if (tid < NS) sm.cnt[tid] = 0;
    if (NV6 && tid < 2 * NS) sm.rot[tid] = 0.;
    if (SENS && tid < 2 * NS) sm.gacc[tid] = 0.;
    load_plane(s_begin);
    if (s_begin + 1 <= p.nz) load_plane(s_begin + 1);
    if (cm_on) {
      sm.cm[(s_begin & 3) * NSB + cm_dst] = p.cmask[(int64_t)s_begin * plane_nodes + cm_src];
      if (s_begin + 1 <= p.nz) sm.cm[((s_begin + 1) & 3) * NSB + cm_dst] = p.cmask[(int64_t)(s_begin + 1) * plane_nodes + cm_src];
    }
    cp_async_wait_all();
    __syncthreads();

    for (int s = s_begin; s < u.z1; ++s) {
      // fetched while this plane is processed: state of plane s+2
      unsigned cmv = 0u;
      if (s + 2 <= p.nz) {
        load_plane(s + 2);
        if (cm_on) cmv = p.cmask[(int64_t)(s + 2) * plane_nodes + cm_src];
      }
      if (s < p.nz) {
        // ---- compute: the warp's 2x2 block of elements of plane s ----
#pragma unroll 1
        for (int round = 0; round < GEO_ROUNDS; ++round) {
          const int c_lo = round * STAGE_EL;
          if (GEO_ROUNDS > 1) __syncwarp();  // the staged geometry of the previous round has been consumed
          {
            const int el = c_lo + (lane >> 3), q = lane & 7;
            const int lex = bx2 + (el & 1), ley = by2 + (el >> 1);
            const bool live = (lane >> 3) < STAGE_EL && lex < u.nex && ley < u.ney;
            if (SENS || live)  // with sensitivities all lanes go in: the point values are joined by shuffles
              geometry_phase<SIMP, SENS>(p, sm, s, ley * TCX + lex, q, my_stage + lane * QS, sm.gq + warp * 32 + (el & 3) * 8, live);
          }
          __syncwarp();
          if (clear_pending) {  // the zeros of the recycled accumulators have landed (issued after the previous flush)
            mbar_wait(clear_bar, clear_phase);
            clear_phase ^= 1u;
            clear_pending = false;
          }
          if (s & 1) {
#pragma unroll 1
            for (int c = c_lo; c < c_lo + STAGE_EL; ++c) {
              const int lex = bx2 + (c & 1), ley = by2 + (c >> 1);
              if (lex < u.nex && ley < u.ney) element_phase<1, NV6, SENS>(p, sm, lk, arena32, rankpack, my_stage + (c - c_lo) * 8 * QS, lex, ley, c, lane, sm.gq + warp * 32 + c * 8);
            }
          } else {
#pragma unroll 1
            for (int c = c_lo; c < c_lo + STAGE_EL; ++c) {
              const int lex = bx2 + (c & 1), ley = by2 + (c >> 1);
              if (lex < u.nex && ley < u.ney) element_phase<0, NV6, SENS>(p, sm, lk, arena32, rankpack, my_stage + (c - c_lo) * 8 * QS, lex, ley, c, lane, sm.gq + warp * 32 + c * 8);
            }
          }
        }
      }
      __syncthreads();
      const bool do_flush = s >= u.z0;
      if (clear_pending) {  // top node plane: no element phase has waited for the clear
        mbar_wait(clear_bar, clear_phase);
        clear_phase ^= 1u;
        clear_pending = false;
      }
      if (do_flush) flush_plane<NV6>(p, sm, arena32, s, u.nox * u.noy, (int64_t)s * plane_nodes, warp, lane);
      if (cm_on && s + 2 <= p.nz) sm.cm[((s + 2) & 3) * NSB + cm_dst] = (uint8_t)cmv;
      cp_async_wait_all();
      __syncthreads();
      // the accumulators of node plane s (-> plane s+2) and of the cross pairs s-1 -> s (-> s+1 -> s+2) start from zero
      {
        const int par = s & 1;
        if (tma_clear) {
          // two TMA bulk copies from a block of zeros in global memory (L2-resident) instead of 20 shared-memory stores
          // per thread: issued by one thread, they land while the CTA runs the geometry of the next element plane; every
          // thread waits on the mbarrier before its first accumulate.  The last plane of a unit needs none (the next
          // unit's prologue clears the whole arena).
          if (s + 1 < u.z1) {
            if (tid == 0) {
              fence_proxy_async_smem();  // the flush's reads (generic proxy, ordered by the barrier above) before the async writes
              const int ilo = par ? IN1_LO : IN0_LO, ihi = par ? IN1_HI : IN0_HI;
              const int clo = par ? CR0_LO : CR1_LO, chi = par ? CR0_HI : CR1_HI;
              mbar_arrive_expect_tx(clear_bar, (unsigned)(ihi - ilo + chi - clo) * 8u);
              bulk_load(arena32 + ilo * 8, p.zeros, (unsigned)(ihi - ilo) * 8u, clear_bar);
              bulk_load(arena32 + clo * 8, p.zeros, (unsigned)(chi - clo) * 8u, clear_bar);
            }
            clear_pending = true;
          }
        } else {
          double* zi = sm.arena + (par ? p.in1_off : 0);
          for (int i = tid; i < 42 * NS; i += ST) zi[i] = 0.;
          double* zc = sm.arena + (par ? p.cr0_off : p.cr1_off);
          for (int i = tid; i < 81 * NS; i += ST) zc[i] = 0.;
        }
        if (tid < NS) sm.cnt[tid] = 0;
        if (NV6 && tid < NS) sm.rot[par * NS + tid] = 0.;
        if (SENS && tid < NS) sm.gacc[par * NS + tid] = 0.;
      }
      __syncthreads();
    }
  }
  if (NV6) bulk_store_wait_all();  // the zero block must outlive the reads of the last bulk stores
}

}  // namespace

size_t hex8_struct_smem_bytes() { return SMEM_BYTES_NV6 > SMEM_BYTES ? SMEM_BYTES_NV6 : SMEM_BYTES; }
size_t hex8_struct_sens_smem_bytes() { return kStructCtasPerSm == 1 ? SMEM_BYTES_SENS : 0; }  // 0: variant not built
int hex8_struct_threads() { return ST; }

cudaError_t launch_hex8_struct(const StructParams& p, int n_ctas, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(p.counter, 0, sizeof(int32_t), stream);
  if (e != cudaSuccess) return e;
  static bool attr_set[5][64] = {};  // per kernel variant and device: the attribute is set once, not per launch
  int dev = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  auto go = [&](auto kern, int variant, int smem) -> cudaError_t {
    if (dev >= 64 || !attr_set[variant][dev]) {
      cudaError_t er = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (er != cudaSuccess) return er;
      if ((er = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return er;
      if (dev < 64) attr_set[variant][dev] = true;
    }
    kern<<<n_ctas, ST, smem, stream>>>(p);
    return cudaSuccess;
  };
  if (p.G) {  // R / J with the density sensitivities of the same state in one launch (3-dof system, SIMP modulus)
    if (p.nv != 3 || !p.rho || !p.adj || kStructCtasPerSm != 1) return cudaErrorInvalidValue;
    e = go(hex8_struct_kernel<true, false, true>, 4, SMEM_BYTES_SENS);
  } else if (p.nv == 6) {
    e = p.rho ? go(hex8_struct_kernel<true, true, false>, 3, SMEM_BYTES_NV6) : go(hex8_struct_kernel<false, true, false>, 2, SMEM_BYTES_NV6);
  } else {
    e = p.rho ? go(hex8_struct_kernel<true, false, false>, 1, SMEM_BYTES) : go(hex8_struct_kernel<false, false, false>, 0, SMEM_BYTES);
  }
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

}  // namespace mastb200
