// sweep_hex8_march.cu -- HEX8 linear-elastic residual + Jacobian sweep for ijk-numbered (build_cube) meshes, 3-dof
// system: register-accumulating column march (sm_100a, fp64).
//
// Same device counterpart as sweep_hex8_struct.cu -- the element loop of MAST::NonlinearImplicitAssembly::
// residual_and_jacobian (/root/reference/src/base/nonlinear_implicit_assembly.cpp:111-186) with
// StructuralElement3D::internal_residual (src/elasticity/solid_element_3d.cpp:135-423, linear strain) and the
// isotropic D of src/property_cards/isotropic_material_property_card.cpp:512-528 -- with a different division of
// labour: the pair blocks S_AB = sum_e sum_q JxW E grad N_A (x) grad N_B never pass through shared-memory accumulators.
//
// A work unit is a column of <= 16 x 4 owned nodes over a z-range; the CTA (16 warps) sweeps it element plane by element
// plane (e = node planes e and e+1).  Per plane:
//   geometry : item = (element, qp) of the <= 17 x 5 elements touching the owned nodes: J^-1, JxW E (constant or SIMP
//              modulus of the interpolated nodal density) and the eight physical gradients scaled by sqrt(JxW E) go to a
//              shared-memory table tab[qp][xy-corner][component][element][z-corner] (element innermost: lanes =
//              consecutive nodes read consecutive elements, conflict-free 16-byte loads of a (bottom, top) corner pair).
//              Elements outside the mesh get zeros, so the march needs no tests.
//   march    : thread = (owned node column A, type, half).  The 9 in-plane directions A -> B of a column are dealt to four
//              warp-uniform thread types so that each has four (element column, direction) visits per plane: centre |
//              the three directions towards column x-1 | towards x+1 | the two y-edges.  Half h pairs {A_e, A_e+1} with
//              B_e+h; a visit adds, for the eight qps, two 3x3 blocks into 18 REGISTER accumulators (6 shared loads per
//              18 DFMA).  When the last element column of a direction is done the finished blocks -- row A_e: dz = 0
//              (half 0), +1 (half 1); row A_e+1: dz = -1 (half 0) -- become K_AB = lam S + mu S^T + mu tr(S) I with
//              libMesh's per-element Dirichlet semantics and are stored, every CSR value exactly once; the half-finished
//              block (A_e+1, B_e+1) of half 1 is carried to the next plane through shared memory.  Force partials
//              f_A += K_AB u_B of the two halves are joined by one shuffle, kept per (node, direction) in shared memory and
//              summed over the nine directions in fixed order by the centre threads after the next barrier.  K_e is never
//              formed.  CSR row pointers of a box have a closed form, so nothing but nodal state is read.
// Summation order of a CSR value: element planes ascending, element columns in the visit order of the direction
// (y-major), qps ascending -- fixed, so runs are bitwise reproducible.  No atomics, no counters; two CTA barriers per plane.
// Nodal coordinates / displacements / density of plane e+2 arrive by cp.async meanwhile.
// Status: parity-green, slower than sweep_hex8_struct.cu (profiles/r01h_hex8_march_ncu.md); opt-in with
// MASTB200_STRUCT_KERNEL=march.
#include <cuda_runtime.h>

#include "kernels.h"

namespace mastb200 {
namespace {

constexpr int MX = kMarchTileX, MY = kMarchTileY;  // owned node columns per tile
constexpr int EXS = MX + 1, EYS = MY + 1;          // element columns touching them
constexpr int ES = EXS * EYS;                      // element slots of the gradient table
constexpr int NCX = MX + 2, NCY = MY + 2, NC = NCX * NCY;  // node columns incl. the halo
constexpr int NN = MX * MY;                        // owned node columns
constexpr int NT = NN * 8;                         // threads: four types x two halves per owned column
constexpr int NCB = (NC + 15) & ~15;
constexpr double GP = 0.57735026918962576451;
static_assert(MX == 16, "a half-warp is one row of owned nodes");

constexpr int O_TAB = 0;                                // [8 qp][4 xy-corner][3][ES][2 z-corner]
constexpr int O_CARRY = O_TAB + 8 * 8 * 3 * ES * 8;     // [9 dir][9][NN]
constexpr int O_F = O_CARRY + 9 * 9 * NN * 8;           // [2][9 dir][3][NN]
constexpr int O_XYZ = O_F + 2 * 9 * 3 * NN * 8;         // [3 planes][3][NC]
constexpr int O_U = O_XYZ + 3 * 3 * NC * 8;             // [3 planes][3][NC]
constexpr int O_RHO = O_U + 3 * 3 * NC * 8;             // [3 planes][NC]
constexpr int O_CM = O_RHO + 3 * NC * 8;                // [3 planes][NCB] uint8
constexpr int O_MISC = O_CM + 3 * NCB;
constexpr int SMEM_BYTES = O_MISC + 16;
static_assert(SMEM_BYTES <= 232448, "shared-memory layout exceeds 227 KB");

__device__ __forceinline__ double lds64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void lds128(unsigned a, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct Sm {
  double* tab;
  double* carry;
  double* f;
  double* xyz;
  double* u;
  double* rho;
  uint8_t* cm;
  int* misc;
};

// per-thread constants of a unit
struct NodeK {
  int t;            // owned column index b*16 + a
  int ncol;         // node column of A in the halo grid
  int64_t inplane;  // in-plane node id of A
  int x0f, y0f, xcnt, ycnt;
  int ax, ay;
  bool active;
};

// ---- geometry: item = (qp, element slot) ------------------------------------------------------------------------------
template <bool SIMP>
__device__ __forceinline__ void geometry_item(const StructParams& p, const Sm& sm, int e, int q, int slot, bool exists) {
  double2* out = reinterpret_cast<double2*>(sm.tab) + (q * 12) * ES + slot;  // (bottom, top) corner pairs
  if (!exists) {
#pragma unroll
    for (int i = 0; i < 12; ++i) out[i * ES] = make_double2(0., 0.);
    return;
  }
  const int ex = slot % EXS, ey = slot / EXS, base = ey * NCX + ex;
  const double xi = (q & 1) ? GP : -GP, eta = (q & 2) ? GP : -GP, zeta = (q & 4) ? GP : -GP;
  const double xm = 1. - xi, xp = 1. + xi, ym = 1. - eta, yp = 1. + eta, zm = 1. - zeta, zp = 1. + zeta;
  // reference gradients factorised: dN_m/dxi = sx t_yz[my][mz], dN_m/deta = sy t_xz[mx][mz], dN_m/dzeta = sz t_xy[mx][my]
  const double fxv[2] = {xm, xp}, fyv[2] = {ym, yp}, fzv[2] = {zm, zp};
  double t_yz[2][2], t_xz[2][2], t_xy[2][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      t_yz[a][b] = 0.125 * fyv[a] * fzv[b]; t_xz[a][b] = 0.125 * fxv[a] * fzv[b]; t_xy[a][b] = 0.125 * fxv[a] * fyv[b];
    }
  double xn[8][3];
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const double* X = sm.xyz + (((m & 4) ? e + 1 : e) % 3) * 3 * NC + base + ((m >> 1) & 1) * NCX + (m & 1);
    xn[m][0] = X[0]; xn[m][1] = X[NC]; xn[m][2] = X[2 * NC];
  }
  double Jm[3][3];  // Jm[r][c] = d x_c / d xi_r = sum over the four edges along r of t * (x_far - x_near)
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double j0 = 0., j1 = 0., j2 = 0.;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        j0 = fma(t_yz[a][b], xn[1 + 2 * a + 4 * b][c] - xn[2 * a + 4 * b][c], j0);
        j1 = fma(t_xz[a][b], xn[a + 2 + 4 * b][c] - xn[a + 4 * b][c], j1);
        j2 = fma(t_xy[a][b], xn[a + 2 * b + 4][c] - xn[a + 2 * b][c], j2);
      }
    Jm[0][c] = j0; Jm[1][c] = j1; Jm[2][c] = j2;
  }
  const double c00 = Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1];
  const double c01 = Jm[1][2] * Jm[2][0] - Jm[1][0] * Jm[2][2];
  const double c02 = Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0];
  const double det = Jm[0][0] * c00 + Jm[0][1] * c01 + Jm[0][2] * c02;
  double E = p.E0;
  if (SIMP) {  // modulus of the interpolated nodal density (examples/structural/example_6/example_6.cpp:100-137)
    double r = 0.;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const double fx = (m & 1) ? xp : xm, fy = (m & 2) ? yp : ym, fz = (m & 4) ? zp : zm;
      r = fma(0.125 * fx * fy * fz, sm.rho[(((m & 4) ? e + 1 : e) % 3) * NC + base + ((m >> 1) & 1) * NCX + (m & 1)], r);
    }
    E = p.E0 * (p.rho_min + (1. - p.rho_min) * simp_pow(r, p.penalty));
  }
  // sqrt(JxW E) / det folded into the inverse: the table holds h_m = sqrt(JxW E) grad N_m, so S_AB = sum_q h_A (x) h_B
  const double sc = sqrt(det * E) * __drcp_rn(det);  // Gauss weights are 1 for the 2x2x2 rule
  double inv[3][3];  // inv[c][r] = d xi_r / d x_c (scaled)
  inv[0][0] = c00 * sc; inv[1][0] = c01 * sc; inv[2][0] = c02 * sc;
  inv[0][1] = (Jm[0][2] * Jm[2][1] - Jm[0][1] * Jm[2][2]) * sc;
  inv[1][1] = (Jm[0][0] * Jm[2][2] - Jm[0][2] * Jm[2][0]) * sc;
  inv[2][1] = (Jm[0][1] * Jm[2][0] - Jm[0][0] * Jm[2][1]) * sc;
  inv[0][2] = (Jm[0][1] * Jm[1][2] - Jm[0][2] * Jm[1][1]) * sc;
  inv[1][2] = (Jm[0][2] * Jm[1][0] - Jm[0][0] * Jm[1][2]) * sc;
  inv[2][2] = (Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0]) * sc;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double P0[2][2], P1[2][2], P2[2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        P0[a][b] = inv[c][0] * t_yz[a][b]; P1[a][b] = inv[c][1] * t_xz[a][b]; P2[a][b] = inv[c][2] * t_xy[a][b];
      }
#pragma unroll
    for (int m = 0; m < 4; ++m) {  // xy-corner m, (bottom, top)
      const int mx = m & 1, my = m >> 1;
      const double g2 = P2[mx][my];
      const double lo = (mx ? P0[my][0] : -P0[my][0]) + (my ? P1[mx][0] : -P1[mx][0]) - g2;
      const double hi = (mx ? P0[my][1] : -P0[my][1]) + (my ? P1[mx][1] : -P1[mx][1]) + g2;
      out[(m * 3 + c) * ES] = make_double2(lo, hi);
    }
  }
}

// ---- one finished 3x3 block -> K_AB, force partial, CSR store -----------------------------------------------------------
// row: node A on plane k; neighbour B = A + (ddx, ddy, dz).  c = rank of B among A's existing neighbours.
__device__ __forceinline__ void flush_block(const StructParams& p, const Sm& sm, const NodeK& nk, const double* S, int ddx, int ddy,
                                            int dz, int k, double* Jrow, double* f, double lam1, double mu1) {
  const int z0f = k == 0 ? 1 : 0, zcnt = 3 - z0f - (k == p.nz ? 1 : 0);
  const int rs = 3 * nk.xcnt * nk.ycnt * zcnt;
  const double trm = mu1 * (S[0] + S[4] + S[8]);
  double K[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) K[i][j] = lam1 * S[i * 3 + j] + mu1 * S[j * 3 + i] + (i == j ? trm : 0.);
  const int bcol = nk.ncol + ddy * NCX + ddx;
  if (p.R) {  // internal force of linear elasticity: f_a += K_ab u_b (unconstrained element matrices)
    const double* U = sm.u + ((k + dz) % 3) * 3 * NC + bcol;
    const double u0 = U[0], u1 = U[NC], u2 = U[2 * NC];
    f[0] += K[0][0] * u0 + K[0][1] * u1 + K[0][2] * u2;
    f[1] += K[1][0] * u0 + K[1][1] * u1 + K[1][2] * u2;
    f[2] += K[2][0] * u0 + K[2][1] * u1 + K[2][2] * u2;
  }
  if (p.J) {
    if (p.cmask) {
      const unsigned cmn = sm.cm[(k % 3) * NCB + nk.ncol], cmm = sm.cm[((k + dz) % 3) * NCB + bcol];
      if (cmn | cmm) {  // libMesh constrain_element_matrix_and_vector: row/column zeroed, diagonal 1 per adjacent element
        const bool self = ddx == 0 && ddy == 0 && dz == 0;
        const double nadj = double((nk.xcnt - 1) * (nk.ycnt - 1) * (zcnt - 1));
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            double v = K[i][j];
            if (((cmn >> i) | (cmm >> j)) & 1u) v = 0.;
            if (self && i == j && ((cmn >> i) & 1u)) v = nadj;
            K[i][j] = v;
          }
      }
    }
    const int c = ((dz + 1 - z0f) * nk.ycnt + (ddy + 1 - nk.y0f)) * nk.xcnt + (ddx + 1 - nk.x0f);
    double* dst = Jrow + 3 * c;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) dst[i * rs + j] = K[i][j];
  }
}

template <bool SIMP>
__global__ void __launch_bounds__(NT, 1) hex8_march_kernel(const StructParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Sm sm;
  sm.tab = reinterpret_cast<double*>(smem_raw + O_TAB);
  sm.carry = reinterpret_cast<double*>(smem_raw + O_CARRY);
  sm.f = reinterpret_cast<double*>(smem_raw + O_F);
  sm.xyz = reinterpret_cast<double*>(smem_raw + O_XYZ);
  sm.u = reinterpret_cast<double*>(smem_raw + O_U);
  sm.rho = reinterpret_cast<double*>(smem_raw + O_RHO);
  sm.cm = smem_raw + O_CM;
  sm.misc = reinterpret_cast<int*>(smem_raw + O_MISC);
  const int tid = threadIdx.x;
  const int type = tid >> 7, half = tid & 1;  // type warp-uniform: 0 centre, 1 / 2 the directions towards column x-1 / x+1, 3 y-edges
  const unsigned tab32 = (unsigned)__cvta_generic_to_shared(sm.tab);
  const double lam1 = p.lam1, mu1 = p.mu1;
  const int64_t npx = p.nx + 1, plane_nodes = npx * (p.ny + 1);

  for (;;) {
    __syncthreads();
    if (tid == 0) sm.misc[0] = atomicAdd(p.counter, 1);
    __syncthreads();
    const int unit = sm.misc[0];
    if (unit >= p.n_units) break;
    const StructUnit u = p.units[unit];
    const int gx0 = u.o0x - 1, gy0 = u.o0y - 1;  // global node coordinates of halo column (0, 0); may be -1
    NodeK nk;
    nk.t = (tid & 127) >> 1;
    {
      const int ai = nk.t & 15, bi = nk.t >> 4;
      nk.active = ai < u.nox && bi < u.noy;
      nk.ax = u.o0x + ai; nk.ay = u.o0y + bi;
      nk.ncol = (bi + 1) * NCX + ai + 1;
      nk.inplane = nk.ax + nk.ay * npx;
      nk.x0f = nk.ax == 0 ? 1 : 0; nk.y0f = nk.ay == 0 ? 1 : 0;
      nk.xcnt = 3 - nk.x0f - (nk.ax == p.nx ? 1 : 0); nk.ycnt = 3 - nk.y0f - (nk.ay == p.ny ? 1 : 0);
    }
    const int e_begin = u.z0 > 0 ? u.z0 - 1 : 0;
    const int e_end = (u.z1 - 1 < p.nz - 1) ? u.z1 - 1 : p.nz - 1;  // inclusive

    auto load_plane = [&](int z) {  // coordinates, displacements (and density) of node plane z, asynchronously
      for (int i = tid; i < NC * 7; i += NT) {
        const int col = i % NC, c = i / NC;
        const int lx = col % NCX, ly = col / NCX;
        const int gx = gx0 + lx, gy = gy0 + ly;
        if (gx < 0 || gx > p.nx || gy < 0 || gy > p.ny || lx > u.nox + 1 || ly > u.noy + 1) continue;
        const int64_t gid = (int64_t)z * plane_nodes + gx + gy * npx;
        if (c < 3) cp_async8(sm.xyz + (z % 3) * 3 * NC + c * NC + col, p.xyz + gid * 3 + c);
        else if (c < 6) cp_async8(sm.u + (z % 3) * 3 * NC + (c - 3) * NC + col, p.X + gid * 3 + (c - 3));
        else if (SIMP) cp_async8(sm.rho + (z % 3) * NC + col, p.rho + gid);
      }
    };
    auto load_cm = [&](int z) {
      if (p.cmask && tid < NC) {
        const int lx = tid % NCX, ly = tid / NCX;
        const int gx = gx0 + lx, gy = gy0 + ly;
        unsigned v = 0u;
        if (gx >= 0 && gx <= p.nx && gy >= 0 && gy <= p.ny) v = p.cmask[(int64_t)z * plane_nodes + gx + gy * npx];
        sm.cm[(z % 3) * NCB + tid] = (uint8_t)v;
      }
    };
    load_plane(e_begin);
    load_plane(e_begin + 1);
    load_cm(e_begin);
    load_cm(e_begin + 1);

    // force of the rows of node plane k: the nine direction partials of buffer b summed in fixed order by the centre threads
    auto force_row = [&](int k, int b) {
      if (type == 0 && half == 0 && nk.active) {
        const unsigned cmn = p.cmask ? sm.cm[(k % 3) * NCB + nk.ncol] : 0u;
        const double* fs = sm.f + b * 27 * NN + nk.t;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          double y = 0.;
#pragma unroll
          for (int d = 0; d < 9; ++d) y += fs[(d * 3 + i) * NN];
          p.R[((int64_t)k * plane_nodes + nk.inplane) * 3 + i] = ((cmn >> i) & 1u) ? 0. : y;
        }
      }
    };
    for (int e = e_begin; e <= e_end; ++e) {
      cp_async_wait_all();
      __syncthreads();  // B1: state of planes e, e+1 visible; the table is free; the force partials of plane e-1 are complete
      if (e + 2 <= p.nz) load_plane(e + 2);
      const bool own_lo = e >= u.z0 && e < u.z1, own_hi = e + 1 >= u.z0 && e + 1 < u.z1;
      if (p.R && e > e_begin && e - 1 >= u.z0) force_row(e - 1, (e - 1) & 1);  // nobody writes that buffer before B2
      // ---- geometry ----
      for (int i = tid; i < 8 * ES; i += NT) {
        const int q = i / ES, slot = i - q * ES;
        const int ex = slot % EXS, ey = slot / EXS;
        const int gex = gx0 + ex, gey = gy0 + ey;
        const bool exists = gex >= 0 && gex < p.nx && gey >= 0 && gey < p.ny && ex <= u.nox && ey <= u.noy;
        geometry_item<SIMP>(p, sm, e, q, slot, exists);
      }
      // node-level CSR row pointers of this thread's two rows: the rows of a box are numbered in node order with
      // xcnt*ycnt*zcnt neighbours each, so the prefix sums have a closed form (== nbr_ptr of the pattern)
      auto row_ptr = [&](int k) -> int64_t {
        const int64_t SX = 3 * (int64_t)p.nx + 1, SY = 3 * (int64_t)p.ny + 1;
        const int64_t pz = 3 * (int64_t)k - (k > 0 ? 1 : 0), pz0 = 3 * (int64_t)p.oz0 - (p.oz0 > 0 ? 1 : 0);
        const int64_t zc = 3 - (k == 0 ? 1 : 0) - (k == p.nz ? 1 : 0);
        const int64_t px = 3 * nk.ax - (nk.ax > 0 ? 1 : 0), py = 3 * nk.ay - (nk.ay > 0 ? 1 : 0);
        return SX * SY * (pz - pz0) + zc * (SX * py + nk.ycnt * px);
      };
      __syncthreads();  // B2: table of plane e complete
      if (e + 2 <= p.nz) load_cm(e + 2);
      // ---- march ----
      // half 0 pairs {A_e, A_e+1} with B_e (blocks: row e dz = 0, row e+1 dz = -1), half 1 with B_e+1 (row e dz = +1 and
      // the carried block (A_e+1, B_e+1)); the two halves of a node are neighbouring lanes
      double alo[9], ahi[9];
      const int cur = e & 1;
      const bool top = e + 1 == p.nz;
#pragma unroll 1
      for (int ui = 0; ui < 4; ++ui) {
        int ddx, ddy, rx, ry;
        bool first, last;
        if (type == 0) { ddx = 0; ddy = 0; rx = ui & 1; ry = ui >> 1; first = ui == 0; last = ui == 3; }
        else if (type == 3) { ddx = 0; ddy = ui < 2 ? -1 : 1; ry = ui < 2 ? 0 : 1; rx = ui & 1; first = !(ui & 1); last = ui & 1; }
        else {  // the three directions towards the x-1 (type 1) / x+1 (type 2) column: corner, edge (two elements), corner
          ddx = type == 1 ? -1 : 1; rx = type == 1 ? 0 : 1;
          ddy = ui == 0 ? -1 : (ui == 3 ? 1 : 0); ry = ui >> 1; first = ui != 2; last = ui != 1;
        }
        const int dir = (ddy + 1) * 3 + ddx + 1;
        double* cr = sm.carry + dir * 9 * NN + nk.t;
        if (first) {
          const bool take = half == 0 && e > e_begin;
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            alo[i] = take ? cr[i * NN] : 0.;
            ahi[i] = 0.;
          }
        }
        {
          const int slot = ((nk.t >> 4) + ry) * EXS + (nk.t & 15) + rx;
          const int mA = (1 - rx) + 2 * (1 - ry), mB = (ddx + 1 - rx) + 2 * (ddy + 1 - ry);
          const unsigned aA = tab32 + (mA * 3 * ES + slot) * 16, aB = tab32 + (mB * 3 * ES + slot) * 16 + half * 8;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            double al[3], ah[3], b[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) lds128(aA + ((q * 12 + c) * ES) * 16, al[c], ah[c]);
#pragma unroll
            for (int c = 0; c < 3; ++c) b[c] = lds64(aB + ((q * 12 + c) * ES) * 16);
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                alo[i * 3 + j] = fma(al[i], b[j], alo[i * 3 + j]);
                ahi[i * 3 + j] = fma(ah[i], b[j], ahi[i * 3 + j]);
              }
          }
        }
        if (last) {
          const int bx = nk.ax + ddx, by = nk.ay + ddy;
          const bool nbr = nk.active && bx >= 0 && bx <= p.nx && by >= 0 && by <= p.ny;
          if (own_lo) {
            double f[3] = {0., 0., 0.};
            if (nbr) flush_block(p, sm, nk, alo, ddx, ddy, half, e, p.J + 9 * row_ptr(e), f, lam1, mu1);
            if (p.R) {
              double* fl = sm.f + (cur * 9 + dir) * 3 * NN + nk.t;
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                const double o = __shfl_xor_sync(0xffffffffu, f[i], 1);
                if (half == 0) fl[i * NN] = (e > e_begin ? fl[i * NN] : 0.) + (f[i] + o);
              }
            }
          }
          if (own_hi) {
            double f[3] = {0., 0., 0.};
            if (nbr && (half == 0 || top)) flush_block(p, sm, nk, ahi, ddx, ddy, half - 1, e + 1, p.J + 9 * row_ptr(e + 1), f, lam1, mu1);
            if (p.R) {
              double* fh = sm.f + ((cur ^ 1) * 9 + dir) * 3 * NN + nk.t;
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                const double o = __shfl_xor_sync(0xffffffffu, f[i], 1);
                if (half == 0) fh[i * NN] = f[i] + o;
              }
            }
          }
          if (half == 1) {
#pragma unroll
            for (int i = 0; i < 9; ++i) cr[i * NN] = ahi[i];
          }
        }
      }
    }
    if (p.R) {  // rows of the last plane (and of the top plane of the mesh)
      __syncthreads();
      const bool own_lo = e_end >= u.z0 && e_end < u.z1, own_top = e_end + 1 == p.nz && e_end + 1 >= u.z0 && e_end + 1 < u.z1;
      if (own_lo) force_row(e_end, e_end & 1);
      if (own_top) force_row(e_end + 1, (e_end & 1) ^ 1);
    }
    cp_async_wait_all();
  }
}

}  // namespace

size_t hex8_march_smem_bytes() { return SMEM_BYTES; }
int hex8_march_threads() { return NT; }

cudaError_t launch_hex8_march(const StructParams& p, int n_ctas, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(p.counter, 0, sizeof(int32_t), stream);
  if (e != cudaSuccess) return e;
  static bool attr_set[4][64] = {};  // per kernel variant and device: the attribute is set once, not per launch
  int dev = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  auto go = [&](auto kern, int variant) -> cudaError_t {
    if (dev >= 64 || !attr_set[variant][dev]) {
      cudaError_t er = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
      if (er != cudaSuccess) return er;
      if (dev < 64) attr_set[variant][dev] = true;
    }
    kern<<<n_ctas, NT, SMEM_BYTES, stream>>>(p);
    return cudaSuccess;
  };
  e = p.rho ? go(hex8_march_kernel<true>, 1) : go(hex8_march_kernel<false>, 0);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

}  // namespace mastb200
