// sweep_hex8_slots.cu -- HEX8 linear-elastic residual + Jacobian sweep for ijk-numbered (build_cube) meshes, 3-dof
// system: element results in shared-memory SLOTS, row-owner gather straight to the CSR (sm_100a, fp64).
//
// Same device counterpart as sweep_hex8_struct.cu -- the element loop of MAST::NonlinearImplicitAssembly::
// residual_and_jacobian (/root/reference/src/base/nonlinear_implicit_assembly.cpp:111-186) with
// StructuralElement3D::internal_residual (src/elasticity/solid_element_3d.cpp:135-423, linear strain) and the
// isotropic D of src/property_cards/isotropic_material_property_card.cpp:512-528 -- with a different division of
// labour between the two phases of an element plane.  The lock-step kernel of sweep_hex8_struct.cu scatter-ADDS every
// element's pair blocks into per-node-pair accumulators (a read-modify-write per value, four elements per node column
// ordered by counters, accumulators cleared after every plane) and its flush decodes (node, direction) per lane.  Here
//   element phase : warp w owns a 2x2 block of elements (same geometry lanes and the same 12 mma.sync.m8n8k4.f64 as
//                   the lock-step kernel: S_ij[a][b] = sum_q w g_a[i] g_b[j]); the pair blocks of an element are STORED
//                   once into the element's own slot, arrays [value id][element column] -- cross pairs (bottom node,
//                   top node; 144 values) plainly, in-plane pairs (78 unique values) added onto what the element below
//                   stored for the same node plane (the only read-modify-write; same warp, so no ordering protocol).
//                   No counters, no spinning, no clearing pass.
//   gather phase  : item = (owned node, neighbour direction); three consecutive lanes take the x-neighbours -1, 0, +1 of one
//                   (node, dy, dz), ten such triples per warp.  An item sums the <= 4 element columns that hold its pair
//                   (fixed order: y then x ascending), K_AB = lam S + mu S^T + mu tr(S) I with libMesh's per-element
//                   Dirichlet semantics goes straight to the CSR (three lanes = 72 contiguous bytes per row), and
//                   f_A += K_AB u_B is joined over the triple by shuffles and over the nine (dy, dz) by a fixed-order sum.
//                   At time s the items are: plane s with dz = 0, +1 and plane s+1 with dz = -1 (all of them read the
//                   element plane s that has just been computed), so a slot lives for one plane only.
// Two CTA barriers per plane.  Summation order of a CSR value: quadrature points inside the DMMA, then the element below
// before the element above (in-plane pairs), then element columns y-major -- fixed, so runs are bitwise reproducible.
// Serves nv = 3 without fused sensitivities; the six-variable layout and the fused density sensitivities stay with
// sweep_hex8_struct.cu.
#include <cuda_runtime.h>

#include "kernels.h"

namespace mastb200 {
namespace {

constexpr int SW = 16;                 // warps per CTA: 4 x 4 blocks of 2x2 elements
constexpr int ST = SW * 32;
constexpr int TCX = 9, TCY = 9;        // node columns of a tile (8 x 8 elements)
constexpr int NS = TCX * TCY;
constexpr int NSB = (NS + 3) & ~3;
constexpr int NEP = 69;                 // doubles per slot array: 64 element columns + 5: bank = 5 x value id + column (tools/design/slots_bank_model.py:
                                       // the gather costs 47 wavefronts per node at this stride, 68 at 65, 44 without conflicts)
constexpr int N_CR = 144, N_IN = 78;   // value arrays: cross pairs 9 x 16; in-plane 3 x 16 (i < j) + 3 x 10 (i == j, unordered)
constexpr int QS = 10;                 // doubles staged per (element, qp): inv[3][3], w
constexpr int MAXOWN = 49;
constexpr double GP = 0.57735026918962576451;

// shared-memory carve-up (bytes).  The in-plane regions start 8 doubles (mod 16) after the cross region so that the two
// kinds of lanes of one store instruction fall into different banks.
constexpr int O_CR = 0;
constexpr int CR_BYTES = N_CR * NEP * 8;
constexpr int IN_BYTES = N_IN * NEP * 8;
constexpr int O_IN0 = ((O_CR + CR_BYTES + 127) & ~127) + 64;
constexpr int O_IN1 = ((O_IN0 + IN_BYTES + 127) & ~127) + 64;
constexpr int O_XYZ = O_IN1 + IN_BYTES;                  // [3 planes][3][NS]
constexpr int O_U = O_XYZ + 3 * 3 * NS * 8;              // [4 planes][3][NS]
constexpr int O_STAGE = (O_U + 4 * 3 * NS * 8 + 15) & ~15;   // [SW][4 el][8 qp][QS]
constexpr int O_CM = O_STAGE + SW * 4 * 8 * QS * 8;      // [4 planes][NSB] uint8
constexpr int O_RHO = (O_CM + 4 * NSB + 7) & ~7;         // [3 planes][NS]
constexpr int O_NODE = O_RHO + 3 * NS * 8;               // [64] int2: in-plane node id, lx | ly << 4 | xy boundary flags << 8
constexpr int O_FP = O_NODE + 64 * 8;                    // [3 dz][MAXOWN][3] force partials of a plane's items
constexpr int O_CARRY = O_FP + 3 * MAXOWN * 3 * 8;       // [MAXOWN][3] dz = -1 partials of the plane above, from the step before
constexpr int O_MISC = O_CARRY + MAXOWN * 3 * 8;
constexpr int SMEM_BYTES = O_MISC + 16;
static_assert((O_CR / 8) % 16 == 0 && (O_IN0 / 8) % 16 == 8 && (O_IN1 / 8) % 16 == 8, "bank phases of the slot regions");
static_assert(O_XYZ % 8 == 0 && O_STAGE % 16 == 0 && O_RHO % 8 == 0 && O_NODE % 8 == 0 && O_FP % 8 == 0, "alignment");
static_assert(SMEM_BYTES <= 232448, "shared-memory layout exceeds 227 KB");
static_assert(2 * 8 * 20 <= 4 * 8 * QS, "the Green-Lagrange variant stages two elements in the same per-warp buffer");

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ double lds64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct Sm {
  double* xyz;
  double* u;
  double* stage;
  uint8_t* cm;
  double* rho;
  int2* node;
  double* fp;
  double* carry;
  int* misc;
};

// position of the pair (first node a, second node b; xy-local indices 0..3) within a 16-array group: the x bit of b moves
// by 8, so that the three lanes of a gather triple and the eight lanes of a store quadrant hit distinct banks
__device__ __forceinline__ int pm(int a, int b) { return (b & 1) * 8 + a * 2 + (b >> 1); }
// unordered pair a <= b of 0..3 -> 0..9, by diagonals: linear in a for a fixed difference (the gather's corners)
__device__ __forceinline__ int updiag(int d) { return d == 0 ? 0 : (d == 1 ? 4 : (d == 2 ? 7 : 9)); }
__device__ __forceinline__ int up(int a, int b) { return a + updiag(b - a); }

struct LaneK {
  double d0[2], d1[2], d2;   // reference gradients of N_m at qps q4 (h = 0) and q4 + 4 (h = 1)
  int off[12];               // BYTE offset of value t = bb * 6 + k within its region (array start; the column is added)
  unsigned pred;             // bit t: the lane stores value t (values it holds in duplicate with another lane are skipped)
  int kind;                  // 0 cross pairs, 1 in-plane pairs of the bottom nodes (added), 2 of the top nodes (stored)
};

// Row / column index m of the DMMA tile is the element node (dx = m&1, dy = (m>>1)&1, dz = m>>2); lane = (m, q4) holds
// C[k][bb] = S_(m,b)[KI[k]][KJ[k]] for b = 2 q4 + bb.
__device__ __forceinline__ void make_lane(int lane, LaneK& lk) {
  const int m = lane >> 2, q4 = lane & 3;
  const double sx = (m & 1) ? 1. : -1., sy = (m & 2) ? 1. : -1., sz = (m & 4) ? 1. : -1.;
  const double fx = 1. + sx * ((q4 & 1) ? GP : -GP), fy = 1. + sy * ((q4 & 2) ? GP : -GP);
  const double fz0 = 1. - sz * GP, fz1 = 1. + sz * GP;
  lk.d0[0] = 0.125 * sx * fy * fz0; lk.d0[1] = 0.125 * sx * fy * fz1;
  lk.d1[0] = 0.125 * sy * fx * fz0; lk.d1[1] = 0.125 * sy * fx * fz1;
  lk.d2 = 0.125 * sz * fx * fy;
  const int KI[6] = {0, 1, 2, 0, 0, 1}, KJ[6] = {0, 1, 2, 1, 2, 2};
  const int mz = m >> 2, bz = q4 >> 1, m4 = m & 3;
  lk.kind = mz == bz ? (mz == 0 ? 1 : 2) : 0;
  lk.pred = 0u;
#pragma unroll
  for (int bb = 0; bb < 2; ++bb) {
    const int b4 = (2 * q4 + bb) & 3;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      int idx = 0;
      bool on = true;
      if (mz == bz) {
        if (k >= 3) idx = (k - 3) * 16 + pm(m4, b4);
        else if (m4 <= b4) idx = 48 + k * 10 + up(m4, b4);
        else on = false;  // S_ii of an unordered pair: the lane (b, m) holds it too
      } else if (mz == 0) {  // bottom m, top b: S_(m,b)[KI][KJ]
        idx = (KI[k] * 3 + KJ[k]) * 16 + pm(m4, b4);
      } else {               // top m, bottom b: C = S_(m,b)[KI][KJ] = S_(b,m)[KJ][KI]; the diagonal components are the other lane's
        if (k >= 3) idx = (KJ[k] * 3 + KI[k]) * 16 + pm(b4, m4);
        else on = false;
      }
      lk.off[bb * 6 + k] = idx * NEP * 8;
      if (on) lk.pred |= 1u << (bb * 6 + k);
    }
  }
}

// ---- geometry: lane = (element el of the warp's 2x2 block, quadrature point q): J^-1 and w = JxW E to the stage -----
template <bool SIMP>
__device__ __forceinline__ void geometry_phase(const StructParams& p, const Sm& sm, int s, int base, int q, double* out) {
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
  double2* o = reinterpret_cast<double2*>(out);
  o[0] = make_double2(c00 * idet, (Jm[0][2] * Jm[2][1] - Jm[0][1] * Jm[2][2]) * idet);
  o[1] = make_double2((Jm[0][1] * Jm[1][2] - Jm[0][2] * Jm[1][1]) * idet, c01 * idet);
  o[2] = make_double2((Jm[0][0] * Jm[2][2] - Jm[0][2] * Jm[2][0]) * idet, (Jm[0][2] * Jm[1][0] - Jm[0][0] * Jm[1][2]) * idet);
  o[3] = make_double2(c02 * idet, (Jm[0][1] * Jm[2][0] - Jm[0][0] * Jm[2][1]) * idet);
  double E = p.E0;
  if (SIMP) {  // modulus of the interpolated nodal density (examples/structural/example_6/example_6.cpp:100-137)
    double r = 0.;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const double fx = (m & 1) ? xp : xm, fy = (m & 2) ? yp : ym, fz = (m & 4) ? zp : zm;
      r = fma(0.125 * fx * fy * fz, sm.rho[(((m & 4) ? s + 1 : s) % 3) * NS + base + ((m >> 1) & 1) * TCX + (m & 1)], r);
    }
    E = p.E0 * (p.rho_min + (1. - p.rho_min) * simp_pow(r, p.penalty));
  }
  o[4] = make_double2((Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0]) * idet, det * E);  // Gauss weights are 1 for the 2x2x2 rule
}

// ---- slot stores of one element's tiles (lane = (m, q4) holds C[k][bb] for b = 2 q4 + bb) -----------------------------
__device__ __forceinline__ void sts64_if(unsigned a, double v, unsigned cond) {
  asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %2, 0;\n @q st.shared.f64 [%0], %1;\n}" ::"r"(a), "d"(v), "r"(cond) : "memory");
}
__device__ __forceinline__ void lds64_if(double& v, unsigned a, unsigned cond) {
  asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %2, 0;\n @q ld.shared.f64 %0, [%1];\n}" : "+d"(v) : "r"(a), "r"(cond) : "memory");
}
__device__ __forceinline__ void store_slots(const LaneK& lk, unsigned smem32, int ecol, int par, const double (&C)[6][2]) {
  // bottom in-plane pairs join what the element below left for node plane s (region IN[par]); top in-plane pairs open
  // node plane s+1 (region IN[par ^ 1]); cross pairs are this element's alone.  Values a lane holds in duplicate are not
  // stored (predicated instructions, no branches).
  const unsigned rb = lk.kind == 0 ? O_CR : (((lk.kind == 1) == (par == 0)) ? O_IN0 : O_IN1);
  const unsigned ar = smem32 + rb + ecol * 8;
  const unsigned ldm = lk.kind == 1 ? lk.pred : 0u;  // the adding lanes read first
  double old[12];
#pragma unroll
  for (int t = 0; t < 12; ++t) {
    old[t] = 0.;
    lds64_if(old[t], ar + lk.off[t], (ldm >> t) & 1u);
  }
#pragma unroll
  for (int t = 0; t < 12; ++t) sts64_if(ar + lk.off[t], old[t] + C[t % 6][t / 6], (lk.pred >> t) & 1u);
}

// ---- matrix + slot stores: all 32 lanes on one element -------------------------------------------------------------
__device__ __forceinline__ void element_phase(const LaneK& lk, unsigned smem32, const double* st, int ecol, int par, int lane) {
  const int q4 = lane & 3;
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
  store_slots(lk, smem32, ecol, par, C);
}

// ---- Green-Lagrange strain (GL variant): StructuralElement3D::internal_residual with NONLINEAR_STRAIN ------------------
// (solid_element_3d.cpp:299-409).  With F = I + grad u, g = grad N and the isotropic D the tangent block of a node pair is
//   K_ab = sum_q JxW E [ lam1 (F g_a)(F g_b)^T + mu1 (F g_b)(F g_a)^T + mu1 (g_a . g_b) F F^T + (g_a . S_1 g_b) I ],
// S_1 = lam1 tr(E_GL) I + 2 mu1 E_GL (B_a^T D B_b of the reference's sixteen operator products, plus the initial-stress
// stiffness), i.e. K_ij[a][b] is again a sum of rank-one DMMA products over the quadrature points: 27 products = 54
// mma.sync.m8n8k4.f64 per element for the six tiles i <= j (K_ji[a][b] = K_ij[b][a]), and the slots then hold K itself.
constexpr int QSG = 20;  // doubles staged per (element, qp): inv[3][3], w, H[3][3] = grad u, pad

template <bool SIMP>
__device__ __forceinline__ void geometry_phase_gl(const StructParams& p, const Sm& sm, int s, int base, int q, double* out) {
  const double xi = (q & 1) ? GP : -GP, eta = (q & 2) ? GP : -GP, zeta = (q & 4) ? GP : -GP;
  const double xm = 1. - xi, xp = 1. + xi, ym = 1. - eta, yp = 1. + eta, zm = 1. - zeta, zp = 1. + zeta;
  double Jm[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};  // Jm[r][c] = d x_c / d xi_r
  double Hr[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};  // Hr[i][r] = sum_m u_m,i dN_m/dxi_r
  double r = 0.;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const double fx = (m & 1) ? xp : xm, fy = (m & 2) ? yp : ym, fz = (m & 4) ? zp : zm;
    const double d0 = ((m & 1) ? 0.125 : -0.125) * fy * fz, d1 = ((m & 2) ? 0.125 : -0.125) * fx * fz,
                 d2 = ((m & 4) ? 0.125 : -0.125) * fx * fy;
    const int z = (m & 4) ? s + 1 : s, col = base + ((m >> 1) & 1) * TCX + (m & 1);
    const double* X = sm.xyz + (z % 3) * 3 * NS + col;
    const double x0 = X[0], x1 = X[NS], x2 = X[2 * NS];
    Jm[0][0] = fma(d0, x0, Jm[0][0]); Jm[0][1] = fma(d0, x1, Jm[0][1]); Jm[0][2] = fma(d0, x2, Jm[0][2]);
    Jm[1][0] = fma(d1, x0, Jm[1][0]); Jm[1][1] = fma(d1, x1, Jm[1][1]); Jm[1][2] = fma(d1, x2, Jm[1][2]);
    Jm[2][0] = fma(d2, x0, Jm[2][0]); Jm[2][1] = fma(d2, x1, Jm[2][1]); Jm[2][2] = fma(d2, x2, Jm[2][2]);
    const double* U = sm.u + (z & 3) * 3 * NS + col;
    const double u0 = U[0], u1 = U[NS], u2 = U[2 * NS];
    Hr[0][0] = fma(u0, d0, Hr[0][0]); Hr[0][1] = fma(u0, d1, Hr[0][1]); Hr[0][2] = fma(u0, d2, Hr[0][2]);
    Hr[1][0] = fma(u1, d0, Hr[1][0]); Hr[1][1] = fma(u1, d1, Hr[1][1]); Hr[1][2] = fma(u1, d2, Hr[1][2]);
    Hr[2][0] = fma(u2, d0, Hr[2][0]); Hr[2][1] = fma(u2, d1, Hr[2][1]); Hr[2][2] = fma(u2, d2, Hr[2][2]);
    if (SIMP) r = fma(0.125 * fx * fy * fz, sm.rho[(z % 3) * NS + col], r);
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
  double E = p.E0;
  if (SIMP) E = p.E0 * (p.rho_min + (1. - p.rho_min) * simp_pow(r, p.penalty));
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) out[c * 3 + rr] = inv[c][rr];
  out[9] = det * E;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) out[10 + i * 3 + j] = Hr[i][0] * inv[j][0] + Hr[i][1] * inv[j][1] + Hr[i][2] * inv[j][2];
  out[19] = 0.;
}

__device__ __forceinline__ void element_phase_gl(const StructParams& p, const LaneK& lk, unsigned smem32, const double* st, int ecol,
                                                 int par, int lane) {
  const int q4 = lane & 3;
  double C[6][2];
#pragma unroll
  for (int k = 0; k < 6; ++k) { C[k][0] = 0.; C[k][1] = 0.; }
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    const double2* s2 = reinterpret_cast<const double2*>(st + (q4 + 4 * h) * QSG);
    double v[20];
#pragma unroll
    for (int t = 0; t < 10; ++t) { const double2 x = s2[t]; v[2 * t] = x.x; v[2 * t + 1] = x.y; }
    const double w = v[9];
    const double* H = v + 10;  // H[i * 3 + j] = d u_i / d x_j
    double g[3], gh[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c] = v[c * 3] * lk.d0[h] + v[c * 3 + 1] * lk.d1[h] + v[c * 3 + 2] * lk.d2;
#pragma unroll
    for (int i = 0; i < 3; ++i) gh[i] = g[i] + (H[i * 3] * g[0] + H[i * 3 + 1] * g[1] + H[i * 3 + 2] * g[2]);  // F g
    // F F^T and the Green-Lagrange strain E = 1/2 (H + H^T + H^T H), symmetric: index (i, j) -> KK
    double Bm[6], Eg[6];
    {
      const int KI[6] = {0, 1, 2, 0, 0, 1}, KJ[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int i = KI[k], j = KJ[k];
        Bm[k] = (i == j ? 1. : 0.) + H[i * 3 + j] + H[j * 3 + i] + (H[i * 3] * H[j * 3] + H[i * 3 + 1] * H[j * 3 + 1] + H[i * 3 + 2] * H[j * 3 + 2]);
        Eg[k] = 0.5 * (H[i * 3 + j] + H[j * 3 + i] + (H[i] * H[j] + H[3 + i] * H[3 + j] + H[6 + i] * H[6 + j]));
      }
    }
    const double tr = p.lam1 * (Eg[0] + Eg[1] + Eg[2]), m2 = 2. * p.mu1;
    const double S00 = tr + m2 * Eg[0], S11 = tr + m2 * Eg[1], S22 = tr + m2 * Eg[2], S01 = m2 * Eg[3], S02 = m2 * Eg[4], S12 = m2 * Eg[5];
    const double Sg[3] = {S00 * g[0] + S01 * g[1] + S02 * g[2], S01 * g[0] + S11 * g[1] + S12 * g[2],
                          S02 * g[0] + S12 * g[1] + S22 * g[2]};
    const double wl = w * p.lam1, wm = w * p.mu1, wlm = w * (p.lam1 + p.mu1);
    // diagonal tiles (i, i): (lam1 + mu1) (F g_a)_i (F g_b)_i + sum_J g_a[J] (mu1 (F F^T)_ii g_b[J] + (S_1 g_b)[J])
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      dmma8x8x4(C[i][0], C[i][1], gh[i], wlm * gh[i]);
#pragma unroll
      for (int J = 0; J < 3; ++J) dmma8x8x4(C[i][0], C[i][1], g[J], wm * Bm[i] * g[J] + w * Sg[J]);
    }
    // off-diagonal tiles (i < j): lam1 (F g_a)_i (F g_b)_j + mu1 (F g_a)_j (F g_b)_i + mu1 (F F^T)_ij sum_J g_a[J] g_b[J]
    {
      const int KI[3] = {0, 0, 1}, KJ[3] = {1, 2, 2};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int i = KI[k], j = KJ[k];
        dmma8x8x4(C[3 + k][0], C[3 + k][1], gh[i], wl * gh[j]);
        dmma8x8x4(C[3 + k][0], C[3 + k][1], gh[j], wm * gh[i]);
        const double wb = wm * Bm[3 + k];
#pragma unroll
        for (int J = 0; J < 3; ++J) dmma8x8x4(C[3 + k][0], C[3 + k][1], g[J], wb * g[J]);
      }
    }
  }
  store_slots(lk, smem32, ecol, par, C);
}

// ---- gather: warp = (group of ten owned node columns, dz), lane = (node column, dx); the three dy in turn -------------
// At time s (element plane s just computed): node plane s with dz = 0 (in-plane slots, complete) and dz = +1, node plane
// s + 1 with dz = -1 (cross slots of element plane s).  (dy, dz) are compile-time, dx is the lane's.  The slot address of
// a pair is linear in the corner (cx, cy) of the element column relative to the node: one base register per item, the
// four corners and the components are immediates.
constexpr int DXC = (-10 * NEP + 1) * 8, DYC = (-5 * NEP + 8) * 8;   // pm-indexed arrays (cross, in-plane i != j)
constexpr int DXD = (-1 * NEP + 1) * 8, DYD = (-2 * NEP + 8) * 8;    // in-plane i == j (unordered pairs by diagonals)
constexpr int CST = 16 * NEP * 8, DST = 10 * NEP * 8;                // component strides

struct Item {        // what the three (dy) items of a lane share
  int lx, ly, xyflags, di, killx;
  int64_t nb0;        // CSR row pointer (in node blocks) of the item's node, fetched before the first item needs it
  bool tv, vx0, vx1;
  int e00;
};

template <int DJ, int DK, bool KMODE>
__device__ __forceinline__ void gather_item(const StructParams& p, const Sm& sm, unsigned smem32, const StructUnit& u, int s,
                                            const Item& it, int64_t plane_nodes, double& f0, double& f1, double& f2) {
  const int pl = DK < 0 ? s + 1 : s;  // node plane of the item's row
  const int flags = it.xyflags | (pl == 0 ? 16 : 0) | (pl == p.nz ? 32 : 0);  // x0, xN, y0, yN, z0, zN
  const int kill = it.killx | (DJ < 0 ? 4 : 0) | (DJ > 0 ? 8 : 0) | (DK < 0 ? 16 : 0) | (DK > 0 ? 32 : 0);
  if (!it.tv || (flags & kill)) return;
  const int di = it.di;
  // element columns (corner cx, cy) that hold the pair: x by the lane's dx, y by dy, both inside the tile
  const bool vy0 = DJ <= 0 && it.ly >= 1 && it.ly - 1 < u.ney, vy1 = DJ >= 0 && it.ly < u.ney;
  double S[9];
#pragma unroll
  for (int c = 0; c < 9; ++c) S[c] = 0.;
  if (DK == 0) {
    const unsigned inb = smem32 + ((s & 1) ? O_IN1 : O_IN0) + it.e00 * 8;
    // corner (0, 0): m4 = 3, b4 = 3 + dx + 2 dy
    const unsigned pu = inb + ((1 + di) * 8 + 6 + (1 + DJ)) * (NEP * 8);            // S_(A,B)[i][j], i < j
    const unsigned pl2 = inb + (8 + 2 * (1 + di) + 4 * (1 + DJ) + 1) * (NEP * 8);   // S_(B,A)[i][j], i < j
    const int dlt = di + 2 * DJ;                                                    // b4 - m4
    const unsigned pd = inb + (48 + min(3, 3 + dlt) + updiag(dlt < 0 ? -dlt : dlt)) * (NEP * 8);
#define MASTB200_IN_CORNER(CX, CY)                                                                  \
    {                                                                                               \
      const unsigned o = (unsigned)((CX) * DXC + (CY) * DYC), od = (unsigned)((CX) * DXD + (CY) * DYD); \
      S[1] += lds64(pu + o); S[2] += lds64(pu + o + CST); S[5] += lds64(pu + o + 2 * CST);          \
      S[3] += lds64(pl2 + o); S[6] += lds64(pl2 + o + CST); S[7] += lds64(pl2 + o + 2 * CST);       \
      S[0] += lds64(pd + od); S[4] += lds64(pd + od + DST); S[8] += lds64(pd + od + 2 * DST);       \
    }
    if (vy0) {
      if (it.vx0) MASTB200_IN_CORNER(0, 0)
      if (it.vx1) MASTB200_IN_CORNER(1, 0)
    }
    if (vy1) {
      if (it.vx0) MASTB200_IN_CORNER(0, 1)
      if (it.vx1) MASTB200_IN_CORNER(1, 1)
    }
#undef MASTB200_IN_CORNER
  } else {
    // dz = +1: S_(A,B) with A the bottom node; dz = -1: the slot holds S_(B,A) (B bottom), i.e. the transpose
    const int idx00 = DK > 0 ? (1 + di) * 8 + 6 + (1 + DJ) : 8 + 2 * (1 + di) + 4 * (1 + DJ) + 1;
    const unsigned a = smem32 + O_CR + (idx00 * NEP + it.e00) * 8;
#define MASTB200_CR_CORNER(CX, CY)                                                  \
    {                                                                               \
      const unsigned o = (unsigned)((CX) * DXC + (CY) * DYC);                       \
      _Pragma("unroll") for (int c = 0; c < 9; ++c) S[c] += lds64(a + o + c * CST); \
    }
    if (vy0) {
      if (it.vx0) MASTB200_CR_CORNER(0, 0)
      if (it.vx1) MASTB200_CR_CORNER(1, 0)
    }
    if (vy1) {
      if (it.vx0) MASTB200_CR_CORNER(0, 1)
      if (it.vx1) MASTB200_CR_CORNER(1, 1)
    }
#undef MASTB200_CR_CORNER
  }
  // the block holds S_(row,nbr) or its transpose: K = lam*S + mu*S^T + mu*tr(S)*I either way, with lam/mu swapped;
  // KMODE (Green-Lagrange variant): the slots hold the tangent blocks themselves
  const double l1 = DK < 0 ? p.mu1 : p.lam1, l2 = DK < 0 ? p.lam1 : p.mu1;
  const double trm = p.mu1 * (S[0] + S[4] + S[8]);
  double K[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      K[i][j] = KMODE ? (DK < 0 ? S[j * 3 + i] : S[i * 3 + j]) : l1 * S[i * 3 + j] + l2 * S[j * 3 + i] + (i == j ? trm : 0.);
  const int slotA = it.ly * TCX + it.lx, slotB = slotA + DJ * TCX + di;
  if (p.R) {  // internal force of linear elasticity: f_a += K_ab u_b (unconstrained element matrices)
    const double* U = sm.u + ((pl + DK) & 3) * 3 * NS + slotB;
    const double u0 = U[0], u1 = U[NS], u2 = U[2 * NS];
    f0 += K[0][0] * u0 + K[0][1] * u1 + K[0][2] * u2;
    f1 += K[1][0] * u0 + K[1][1] * u1 + K[1][2] * u2;
    f2 += K[2][0] * u0 + K[2][1] * u1 + K[2][2] * u2;
  }
  if (p.J) {
    // CSR numbering of the node's neighbours: interior nodes have all 27 (column block = direction)
    int c = (DK + 1) * 9 + (DJ + 1) * 3 + di + 1, rs = 81;
    double nadj = 8.;
    if (flags) {
      const int x0 = flags & 1, y0 = (flags >> 2) & 1, z0 = (flags >> 4) & 1;
      const int xcnt = 3 - x0 - ((flags >> 1) & 1), ycnt = 3 - y0 - ((flags >> 3) & 1);
      const int zcnt = 3 - z0 - ((flags >> 5) & 1);
      rs = 3 * xcnt * ycnt * zcnt;
      nadj = double((xcnt - 1) * (ycnt - 1) * (zcnt - 1));
      c = ((DK + 1 - z0) * ycnt + (DJ + 1 - y0)) * xcnt + (di + 1 - x0);
    }
    if (p.cmask) {
      const unsigned cmn = sm.cm[(pl & 3) * NSB + slotA], cmm = sm.cm[((pl + DK) & 3) * NSB + slotB];
      if (cmn | cmm) {  // libMesh constrain_element_matrix_and_vector: row/column zeroed, diagonal 1 per adjacent element
        const bool self = DJ == 0 && DK == 0 && di == 0;
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
    }
    double* dst = p.J + 9 * it.nb0 + 3 * c;
    if (p.j_aligned32) {
      // a row piece starts at an even or an odd double (9 nb0 + i rs + 3 c): one 16-byte store of the aligned pair and one
      // 8-byte store of the third value, chosen without a branch (as in the lock-step kernel's flush)
      const int par0 = ((int)it.nb0 + c) & 1;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double* d = dst + i * rs;
        const bool odd = ((par0 + i * rs) & 1) != 0;
        double* d2 = odd ? d + 1 : d;
        const double a = odd ? K[i][1] : K[i][0], b = odd ? K[i][2] : K[i][1];
        asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(d2), "d"(a), "d"(b) : "memory");
        *(odd ? d : d + 2) = odd ? K[i][0] : K[i][2];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) dst[i * rs + j] = K[i][j];
    }
  }
}

template <int DK, bool KMODE>
__device__ __forceinline__ void gather_dz(const StructParams& p, const Sm& sm, unsigned smem32, const StructUnit& u, int s,
                                          int n_own, int group, int64_t plane_nodes, int lane) {
  const int l3 = (lane * 11) >> 5;  // lane / 3
  Item it;
  it.di = lane - 3 * l3 - 1;
  it.killx = it.di < 0 ? 1 : (it.di > 0 ? 2 : 0);
  const int col = group * 10 + l3;
  // node plane s is flushed by this unit from z0 on; node plane s + 1 gets its dz = -1 blocks if the unit owns it
  const bool plane_on = DK < 0 ? (s + 1 < u.z1 && s < p.nz) : s >= u.z0;
  it.tv = lane < 30 && col < n_own && plane_on;
  const int2 ni = sm.node[col < n_own ? col : 0];
  it.nb0 = it.tv && p.J ? p.nbr_ptr[(int64_t)((DK < 0 ? s + 1 : s) - p.oz0) * plane_nodes + ni.x] : 0;
  it.lx = ni.y & 15; it.ly = (ni.y >> 4) & 15; it.xyflags = ni.y >> 8;
  it.vx0 = it.di <= 0 && it.lx >= 1 && it.lx - 1 < u.nex;
  it.vx1 = it.di >= 0 && it.lx < u.nex;
  it.e00 = (it.ly - 1) * 8 + it.lx - 1;
  double f0 = 0., f1 = 0., f2 = 0.;
  gather_item<0, DK, KMODE>(p, sm, smem32, u, s, it, plane_nodes, f0, f1, f2);
  gather_item<-1, DK, KMODE>(p, sm, smem32, u, s, it, plane_nodes, f0, f1, f2);
  gather_item<1, DK, KMODE>(p, sm, smem32, u, s, it, plane_nodes, f0, f1, f2);
  if (p.R) {
    // fixed order over the triple: (dx = -1) + (dx = 0), then + (dx = +1); the lead lane keeps the dz partial of the node
    const double a0 = __shfl_down_sync(0xffffffffu, f0, 1), a1 = __shfl_down_sync(0xffffffffu, f1, 1), a2 = __shfl_down_sync(0xffffffffu, f2, 1);
    const double b0 = __shfl_down_sync(0xffffffffu, f0, 2), b1 = __shfl_down_sync(0xffffffffu, f1, 2), b2 = __shfl_down_sync(0xffffffffu, f2, 2);
    if (lane < 30 && col < n_own && it.di == -1) {
      double* fp = sm.fp + ((DK + 1) * MAXOWN + col) * 3;
      fp[0] = (f0 + a0) + b0; fp[1] = (f1 + a1) + b1; fp[2] = (f2 + a2) + b2;
    }
  }
}

template <bool SIMP, bool GL>
__global__ void __launch_bounds__(ST, 1) hex8_slots_kernel(const StructParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Sm sm;
  sm.xyz = reinterpret_cast<double*>(smem_raw + O_XYZ);
  sm.u = reinterpret_cast<double*>(smem_raw + O_U);
  sm.stage = reinterpret_cast<double*>(smem_raw + O_STAGE);
  sm.cm = smem_raw + O_CM;
  sm.rho = reinterpret_cast<double*>(smem_raw + O_RHO);
  sm.node = reinterpret_cast<int2*>(smem_raw + O_NODE);
  sm.fp = reinterpret_cast<double*>(smem_raw + O_FP);
  sm.carry = reinterpret_cast<double*>(smem_raw + O_CARRY);
  sm.misc = reinterpret_cast<int*>(smem_raw + O_MISC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  LaneK lk;
  make_lane(lane, lk);
  const unsigned smem32 = (unsigned)__cvta_generic_to_shared(smem_raw);
  double* my_stage = sm.stage + warp * (4 * 8 * QS);
  const int bx2 = (warp & 3) * 2, by2 = (warp >> 2) * 2;  // origin of the warp's 2x2 element block

  for (;;) {
    if (tid == 0) sm.misc[0] = atomicAdd(p.counter, 1);
    __syncthreads();
    const int unit = sm.misc[0];
    if (unit >= p.n_units) break;
    const StructUnit u = p.units[unit];
    const int s_begin = u.z0 > 0 ? u.z0 - 1 : 0;
    const int64_t npx = p.nx + 1, plane_nodes = npx * (p.ny + 1);
    const int ncx = u.nex + 1;
    const int n_own = u.nox * u.noy;
    // per-thread constants of the unit: the state item this thread stages every plane.  The thread -> (node column,
    // component) map is the tile's fixed 9 x 9 shape (constant divisors), so recomputing it per plane costs nothing.
    bool ld_on, ld_u;
    int ld_dst;
    int64_t ld_src;
    {
      const int col = tid / 6, c = tid - col * 6;
      const int ly = col / TCX, lx = col - ly * TCX;
      ld_on = col < NS && lx < ncx && ly <= u.ney;
      const int64_t gid0 = (u.e0x + lx) + (u.e0y + ly) * npx;
      ld_u = c >= 3;
      ld_dst = (ld_u ? c - 3 : c) * NS + col;
      ld_src = gid0 * 3 + (ld_u ? c - 3 : c);
    }
    bool cm_on;
    int64_t cm_src;
    const int cm_dst = tid;
    {
      const int ly = tid / TCX, lx = tid - ly * TCX;
      cm_on = tid < NS && lx < ncx && ly <= u.ney;
      cm_src = (u.e0x + lx) + (u.e0y + ly) * npx;
    }
    const bool cmk_on = cm_on && p.cmask != nullptr;
    auto load_plane = [&](int z) {  // coordinates, displacements (and density) of node plane z, asynchronously
      if (SIMP && cm_on) cp_async8(sm.rho + (z % 3) * NS + cm_dst, p.rho + (int64_t)z * plane_nodes + cm_src);
      if (ld_on) {
        if (ld_u) cp_async8(sm.u + (z & 3) * 3 * NS + ld_dst, p.X + (int64_t)z * plane_nodes * 3 + ld_src);
        else cp_async8(sm.xyz + (z % 3) * 3 * NS + ld_dst, p.xyz + (int64_t)z * plane_nodes * 3 + ld_src);
      }
    };
    if (tid < n_own) {
      const int oyi = tid / u.nox, oxi = tid - oyi * u.nox;
      const int ax = u.o0x + oxi, ay = u.o0y + oyi;
      const int flags = (ax == 0 ? 1 : 0) | (ax == p.nx ? 2 : 0) | (ay == 0 ? 4 : 0) | (ay == p.ny ? 8 : 0);
      sm.node[tid] = make_int2((int)(ax + ay * npx), (ax - u.e0x) | ((ay - u.e0y) << 4) | (flags << 8));
    }
    // prologue: node plane s_begin starts from zero (its element plane below is not visited), no force carried in
    {
      double* zi = reinterpret_cast<double*>(smem_raw + ((s_begin & 1) ? O_IN1 : O_IN0));
      for (int i = tid; i < N_IN * NEP; i += ST) zi[i] = 0.;
      if (tid < MAXOWN * 3) sm.carry[tid] = 0.;
    }
    load_plane(s_begin);
    if (s_begin + 1 <= p.nz) load_plane(s_begin + 1);
    if (cmk_on) {
      sm.cm[(s_begin & 3) * NSB + cm_dst] = p.cmask[(int64_t)s_begin * plane_nodes + cm_src];
      if (s_begin + 1 <= p.nz) sm.cm[((s_begin + 1) & 3) * NSB + cm_dst] = p.cmask[(int64_t)(s_begin + 1) * plane_nodes + cm_src];
    }
    cp_async_wait_all();
    __syncthreads();

    for (int s = s_begin; s < u.z1; ++s) {
      unsigned cmv = 0u;
      if (s + 2 <= p.nz) {  // fetched while this plane is processed
        load_plane(s + 2);
        if (cmk_on) cmv = p.cmask[(int64_t)(s + 2) * plane_nodes + cm_src];
      }
      if (s < p.nz) {
        // ---- element phase: the warp's 2x2 block of elements of plane s ----
        if (!GL) {
          {
            const int el = lane >> 3, q = lane & 7;
            const int lex = bx2 + (el & 1), ley = by2 + (el >> 1);
            if (lex < u.nex && ley < u.ney) geometry_phase<SIMP>(p, sm, s, ley * TCX + lex, q, my_stage + lane * QS);
          }
          __syncwarp();
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            const int lex = bx2 + (c & 1), ley = by2 + (c >> 1);
            if (lex < u.nex && ley < u.ney) element_phase(lk, smem32, my_stage + c * 8 * QS, ley * 8 + lex, s & 1, lane);
          }
        } else {
          // Green-Lagrange: twice as much staged per quadrature point, so two elements at a time (lanes 0-15 stage them)
#pragma unroll 1
          for (int round = 0; round < 2; ++round) {
            if (lane < 16) {
              const int el = 2 * round + (lane >> 3), q = lane & 7;
              const int lex = bx2 + (el & 1), ley = by2 + (el >> 1);
              if (lex < u.nex && ley < u.ney) geometry_phase_gl<SIMP>(p, sm, s, ley * TCX + lex, q, my_stage + lane * QSG);
            }
            __syncwarp();
#pragma unroll 1
            for (int c = 2 * round; c < 2 * round + 2; ++c) {
              const int lex = bx2 + (c & 1), ley = by2 + (c >> 1);
              if (lex < u.nex && ley < u.ney) element_phase_gl(p, lk, smem32, my_stage + (c & 1) * 8 * QSG, ley * 8 + lex, s & 1, lane);
            }
            __syncwarp();
          }
        }
      }
      __syncthreads();
      if (warp < 15) {  // warp = (group of ten owned node columns, dz)
        const int group = warp % 5, dzi = warp / 5;
        if (group * 10 < n_own) {
          if (dzi == 0) gather_dz<-1, GL>(p, sm, smem32, u, s, n_own, group, plane_nodes, lane);
          else if (dzi == 1) gather_dz<0, GL>(p, sm, smem32, u, s, n_own, group, plane_nodes, lane);
          else gather_dz<1, GL>(p, sm, smem32, u, s, n_own, group, plane_nodes, lane);
        }
      }
      if (cmk_on && s + 2 <= p.nz) sm.cm[((s + 2) & 3) * NSB + cm_dst] = (uint8_t)cmv;
      cp_async_wait_all();
      __syncthreads();
      // residual of node plane s: the dz = -1 partial carried from the step before, then dz = 0, +1
      if (p.R && tid < 3 * n_own) {
        const int col = (tid * 21846) >> 16, i = tid - 3 * col;
        const double* fp = sm.fp + col * 3 + i;
        constexpr int CS = MAXOWN * 3;
        double v = sm.carry[tid];
        v += fp[1 * CS]; v += fp[2 * CS];
        sm.carry[tid] = fp[0];
        if (s >= u.z0) {
          const int2 ni = sm.node[col];
          const int slot = ((ni.y >> 4) & 15) * TCX + (ni.y & 15);
          const unsigned cmn = p.cmask ? sm.cm[(s & 3) * NSB + slot] : 0u;
          p.R[((int64_t)s * plane_nodes + ni.x) * 3 + i] = ((cmn >> i) & 1u) ? 0. : v;
        }
      }
    }
  }
}

}  // namespace

size_t hex8_slots_smem_bytes() { return SMEM_BYTES; }
int hex8_slots_threads() { return ST; }

cudaError_t launch_hex8_slots(const StructParams& p, int n_ctas, cudaStream_t stream) {
  if (p.nv != 3 || p.G) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(p.counter, 0, sizeof(int32_t), stream);
  if (e != cudaSuccess) return e;
  static bool attr_set[4][64] = {};  // per kernel variant and device: the attribute is set once, not per launch
  int dev = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  auto go = [&](auto kern, int variant) -> cudaError_t {
    if (dev >= 64 || !attr_set[variant][dev]) {
      cudaError_t er = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
      if (er != cudaSuccess) return er;
      if ((er = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return er;
      if (dev < 64) attr_set[variant][dev] = true;
    }
    kern<<<n_ctas, ST, SMEM_BYTES, stream>>>(p);
    return cudaSuccess;
  };
  if (p.gl) {  // Green-Lagrange strain: tangent only (the residual comes from the matrix-free force kernels)
    if (p.R) return cudaErrorInvalidValue;
    e = p.rho ? go(hex8_slots_kernel<true, true>, 3) : go(hex8_slots_kernel<false, true>, 2);
  } else {
    e = p.rho ? go(hex8_slots_kernel<true, false>, 1) : go(hex8_slots_kernel<false, false>, 0);
  }
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

}  // namespace mastb200
