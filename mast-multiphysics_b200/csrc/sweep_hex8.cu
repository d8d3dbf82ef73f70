// sweep_hex8.cu -- HEX8 linear-elastic residual + Jacobian + density-sensitivity sweep (sm_100a, fp64).
//
// Device counterpart of the element loop of MAST::NonlinearImplicitAssembly::residual_and_jacobian
// (/root/reference/src/base/nonlinear_implicit_assembly.cpp:111-186) with
// StructuralElement3D::internal_residual (src/elasticity/solid_element_3d.cpp:135-423, linear strain),
// ::thermal_residual (:950-1089) and the sensitivity sweep of SURVEY 8a row S
// (sensitivity_assemble, nonlinear_implicit_assembly.cpp:515-686, for all nodal-density DVs at once).
//
// One CTA owns one node tile at a time (plan.h).  Per element visit:
//   geometry phase : one lane per (element, quadrature point) -- 4 elements x 8 points per warp -- builds
//                    J, det, the physical gradients g_b of the 8 shape functions, the weight
//                    w = JxW*E(rho_q), the weighted stress P = JxW*sigma and the sensitivity density s.
//   matrix phase   : S_ij[a][b] = sum_q w_q g_a[i] g_b[j] is a 24x24x8 product done with 18
//                    mma.sync.m8n8k4.f64 (DMMA); lane (a, q) holds exactly the (a, q) operand entry, so
//                    the A and B fragments are the same registers (B = w*A) and each lane ends with the
//                    complete 3x3 blocks S[a][b], b in {2(lane%4), 2(lane%4)+1}.
//   accumulate     : blocks are added to the tile's CSR rows in shared memory at plan slots; visits of
//                    one step touch disjoint owned nodes, so no atomics and a fixed summation order.
// At tile end the accumulated S is turned into K = lam*S + mu*S^T + mu*tr(S)*I (isotropic D,
// src/property_cards/isotropic_material_property_card.cpp:512-528, modulus folded into w), the per-element
// Dirichlet semantics of libMesh's constrain_element_matrix_and_vector are applied (row/column zeroed,
// diagonal = number of adjacent elements, rhs zeroed) and every CSR value is written to HBM exactly once
// with contiguous, coalesced stores.  K_e is never stored.
#include <cuda_runtime.h>

#include "device_types.h"
#include "kernels.h"

namespace mastb200 {

namespace {

constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;
constexpr int MAXN = kMaxTileNodes;  // staged tile nodes
constexpr int EL_STRIDE = 280;       // doubles staged per element: V 216 | w 8 | P 48 | s 8   (== 8 mod 16: conflict-free)
constexpr int STAGE_EL = 2;          // elements staged per warp at a time
constexpr int V_HALF = 36;           // stride of one (component, qp-half) 8x4 operand panel (32 + 4 pad)
constexpr int OFF_W = 216, OFF_P = 224, OFF_S = 272;
constexpr int EXTRA = kAccPerNode;   // per owned node: f[3], G, rot-diag

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// HEX8 reference-node signs (libMesh node order)
__device__ __forceinline__ constexpr int sgx(int b) { return (b == 1 || b == 2 || b == 5 || b == 6) ? 1 : -1; }
__device__ __forceinline__ constexpr int sgy(int b) { return (b == 2 || b == 3 || b == 6 || b == 7) ? 1 : -1; }
__device__ __forceinline__ constexpr int sgz(int b) { return (b >= 4) ? 1 : -1; }

struct Smem {
  double* arena;
  double* xyz;
  double* u;
  double* adj;
  double* rho;
  double* T;
  double* stage;
  int32_t* accoff;
  int32_t* gid;
  int32_t* cnt;     // per owned node: number of element contributions accumulated so far
  uint32_t* rec;    // per warp: the 4 visit records of the current chunk
  uint8_t* nbrcnt;
  uint8_t* cm;
};

constexpr int REC_WORDS = 24;  // sizeof(VisitRec8) / 4

__device__ __forceinline__ int ld_volatile_s32(const int32_t* p) { return *reinterpret_cast<const volatile int32_t*>(p); }
__device__ __forceinline__ void st_volatile_s32(int32_t* p, int v) { *reinterpret_cast<volatile int32_t*>(p) = v; }

struct Flags {
  bool J, R, G, simp, thermal, nv6;
};

// ---- geometry phase: lane = (element slot el, quadrature point q) ---------------------------------------
struct QpData {
  double g[8][3];  // physical shape-function gradients
  double w;        // JxW * E(rho_q)
  double P[6];     // JxW * stress (xx,yy,zz,xy,yz,zx), thermal part subtracted
  double s;        // sensitivity density  JxW * dE/drho * eps(adj)^T Dhat eps(u)
  double N[8];
};

__device__ __forceinline__ void geometry_phase(const SweepParams& p, const Smem& sm, const Flags& fl, const int ln[8],
                                               int q, QpData& d) {
  constexpr double gp = 0.57735026918962576451;
  const double xi = (q & 1) ? gp : -gp, eta = (q & 2) ? gp : -gp, zeta = (q & 4) ? gp : -gp;
  const double xm = 1. - xi, xp = 1. + xi, ym = 1. - eta, yp = 1. + eta, zm = 1. - zeta, zp = 1. + zeta;
  double Jm[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};  // Jm[r][c] = d x_c / d xi_r
  double dref[8][3];
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const double fx = sgx(b) > 0 ? xp : xm, fy = sgy(b) > 0 ? yp : ym, fz = sgz(b) > 0 ? zp : zm;
    dref[b][0] = 0.125 * sgx(b) * fy * fz;
    dref[b][1] = 0.125 * sgy(b) * fx * fz;
    dref[b][2] = 0.125 * sgz(b) * fx * fy;
    d.N[b] = 0.125 * fx * fy * fz;
    const double* X = sm.xyz + ln[b] * 3;
    const double x0 = X[0], x1 = X[1], x2 = X[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      Jm[r][0] = fma(dref[b][r], x0, Jm[r][0]);
      Jm[r][1] = fma(dref[b][r], x1, Jm[r][1]);
      Jm[r][2] = fma(dref[b][r], x2, Jm[r][2]);
    }
  }
  // cofactors: inv[c][r] = d xi_r / d x_c
  const double c00 = Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1];
  const double c01 = Jm[1][2] * Jm[2][0] - Jm[1][0] * Jm[2][2];
  const double c02 = Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0];
  const double det = Jm[0][0] * c00 + Jm[0][1] * c01 + Jm[0][2] * c02;
  const double idet = 1. / det;
  double inv[3][3];
  inv[0][0] = c00 * idet; inv[1][0] = c01 * idet; inv[2][0] = c02 * idet;
  inv[0][1] = (Jm[0][2] * Jm[2][1] - Jm[0][1] * Jm[2][2]) * idet;
  inv[1][1] = (Jm[0][0] * Jm[2][2] - Jm[0][2] * Jm[2][0]) * idet;
  inv[2][1] = (Jm[0][1] * Jm[2][0] - Jm[0][0] * Jm[2][1]) * idet;
  inv[0][2] = (Jm[0][1] * Jm[1][2] - Jm[0][2] * Jm[1][1]) * idet;
  inv[1][2] = (Jm[0][2] * Jm[1][0] - Jm[0][0] * Jm[1][2]) * idet;
  inv[2][2] = (Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0]) * idet;
#pragma unroll
  for (int b = 0; b < 8; ++b)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      d.g[b][c] = dref[b][0] * inv[c][0] + dref[b][1] * inv[c][1] + dref[b][2] * inv[c][2];

  // modulus at the point: constant or SIMP of the interpolated nodal density (example_6.cpp:100-137)
  double E = p.E0, dE = 0.;
  if (fl.simp) {
    double r = 0.;
#pragma unroll
    for (int b = 0; b < 8; ++b) r = fma(d.N[b], sm.rho[ln[b]], r);
    E = p.E0 * (p.rho_min + (1. - p.rho_min) * pow(r, p.penalty));
    if (fl.G) dE = p.E0 * (1. - p.rho_min) * p.penalty * pow(r, p.penalty - 1.);
  }
  const double JxW = det;  // Gauss weights are 1 for the 2x2x2 rule
  d.w = JxW * E;
  d.s = 0.;
#pragma unroll
  for (int k = 0; k < 6; ++k) d.P[k] = 0.;
  if (fl.R || fl.G) {
    double H[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};  // H[i][j] = d u_i / d x_j
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const double* U = sm.u + ln[b] * 3;
      const double u0 = U[0], u1 = U[1], u2 = U[2];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        H[0][j] = fma(u0, d.g[b][j], H[0][j]);
        H[1][j] = fma(u1, d.g[b][j], H[1][j]);
        H[2][j] = fma(u2, d.g[b][j], H[2][j]);
      }
    }
    const double e0 = H[0][0], e1 = H[1][1], e2 = H[2][2];
    const double e3 = H[0][1] + H[1][0], e4 = H[1][2] + H[2][1], e5 = H[0][2] + H[2][0];
    const double tr = e0 + e1 + e2;
    if (fl.R) {
      const double lw = p.lam1 * d.w, mw = p.mu1 * d.w;
      double th = 0.;
      if (fl.thermal) {  // thermal_residual: f -= JxW B^T (D alpha dT)
        double t = 0.;
#pragma unroll
        for (int b = 0; b < 8; ++b) t = fma(d.N[b], sm.T[ln[b]], t);
        th = d.w * p.alpha_k * (t - p.T_ref);
      }
      d.P[0] = lw * tr + 2. * mw * e0 - th;
      d.P[1] = lw * tr + 2. * mw * e1 - th;
      d.P[2] = lw * tr + 2. * mw * e2 - th;
      d.P[3] = mw * e3;
      d.P[4] = mw * e4;
      d.P[5] = mw * e5;
    }
    if (fl.G) {
      double A[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const double* U = sm.adj + ln[b] * 3;
        const double u0 = U[0], u1 = U[1], u2 = U[2];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          A[0][j] = fma(u0, d.g[b][j], A[0][j]);
          A[1][j] = fma(u1, d.g[b][j], A[1][j]);
          A[2][j] = fma(u2, d.g[b][j], A[2][j]);
        }
      }
      const double a0 = A[0][0], a1 = A[1][1], a2 = A[2][2];
      const double a3 = A[0][1] + A[1][0], a4 = A[1][2] + A[2][1], a5 = A[0][2] + A[2][0];
      const double dens = p.lam1 * (a0 + a1 + a2) * tr +
                          p.mu1 * (2. * (a0 * e0 + a1 * e1 + a2 * e2) + a3 * e3 + a4 * e4 + a5 * e5);
      d.s = JxW * dE * dens;
    }
  }
}

__device__ __forceinline__ void stage_write(double* st, int q, const QpData& d) {
  const int h = q >> 2, q4 = q & 3;
#pragma unroll
  for (int b = 0; b < 8; ++b)
#pragma unroll
    for (int i = 0; i < 3; ++i) st[(i * 2 + h) * V_HALF + b * 4 + q4] = d.g[b][i];
  st[OFF_W + q] = d.w;
#pragma unroll
  for (int k = 0; k < 6; ++k) st[OFF_P + q * 6 + k] = d.P[k];
  st[OFF_S + q] = d.s;
}

// ---- matrix phase: all 32 lanes on one element; lane = (row node a, qp q4 | q4+4) ------------------------------
// rec: the visit record in shared memory.  All arithmetic (DMMA, lane reductions) happens first; the adds into the
// tile's rows wait until the per-node counters say that every earlier visit of the plan order has landed.
__device__ __forceinline__ void element_phase(const SweepParams& p, const Smem& sm, const Flags& fl, const double* st,
                                              const uint32_t* rec, int extras_base, int lane) {
  const int a = lane >> 2, q4 = lane & 3;
  const uint16_t* rec16 = reinterpret_cast<const uint16_t*>(rec);
  const uint8_t* rec8 = reinterpret_cast<const uint8_t*>(rec);
  const int la = rec16[a];
  const int rank = rec8[16 + a];
  const unsigned slots = rec16[16 + a * 4 + q4];  // bytes 32 + a*8 + 2*q4
  const int accoff = sm.accoff[la];
  const int nnb = sm.nbrcnt[la];
  double A[3][2];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    A[i][0] = st[(i * 2 + 0) * V_HALF + lane];
    A[i][1] = st[(i * 2 + 1) * V_HALF + lane];
  }
  double C[3][3][2];
  double mx = 0., t0 = 0., t1 = 0., t2 = 0., tg = 0.;
  if (fl.J) {
    const double w0 = st[OFF_W + q4], w1 = st[OFF_W + 4 + q4];
    double B[3][2];
#pragma unroll
    for (int j = 0; j < 3; ++j) { B[j][0] = A[j][0] * w0; B[j][1] = A[j][1] * w1; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) { C[i][j][0] = 0.; C[i][j][1] = 0.; }
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) dmma8x8x4(C[i][j][0], C[i][j][1], A[i][h], B[j][h]);
    if (fl.nv6) {
      // 1e-20*max(diag K_e) on the rotational diagonal (solid_element_3d.cpp:413-417); diag of K_e from S_aa
      const int bself = a - 2 * q4;  // the rotational diagonal itself is zero, so the max is at least 0
      if (bself == 0 || bself == 1) {
        const double s0 = C[0][0][bself & 1], s1 = C[1][1][bself & 1], s2 = C[2][2][bself & 1];
        const double trm = p.mu1 * (s0 + s1 + s2);
        mx = fmax(mx, fmax((p.lam1 + p.mu1) * s0 + trm, fmax((p.lam1 + p.mu1) * s1 + trm, (p.lam1 + p.mu1) * s2 + trm)));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
  }
  if (fl.R) {
    const double* P0 = st + OFF_P + q4 * 6;
    const double* P1 = P0 + 24;
    t0 = P0[0] * A[0][0] + P0[3] * A[1][0] + P0[5] * A[2][0] + P1[0] * A[0][1] + P1[3] * A[1][1] + P1[5] * A[2][1];
    t1 = P0[3] * A[0][0] + P0[1] * A[1][0] + P0[4] * A[2][0] + P1[3] * A[0][1] + P1[1] * A[1][1] + P1[4] * A[2][1];
    t2 = P0[5] * A[0][0] + P0[4] * A[1][0] + P0[2] * A[2][0] + P1[5] * A[0][1] + P1[4] * A[1][1] + P1[2] * A[2][1];
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
      t0 += __shfl_xor_sync(0xffffffffu, t0, o);
      t1 += __shfl_xor_sync(0xffffffffu, t1, o);
      t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    }
  }
  if (fl.G) {
    // N_a at qp q4 and q4+4
    constexpr double gp = 0.57735026918962576451;
    const double xi = (q4 & 1) ? gp : -gp, eta = (q4 & 2) ? gp : -gp;
    const double fx = ((a == 1 || a == 2 || a == 5 || a == 6) ? 1. + xi : 1. - xi);
    const double fy = ((a == 2 || a == 3 || a == 6 || a == 7) ? 1. + eta : 1. - eta);
    const double nlo = 0.125 * fx * fy * (a >= 4 ? 1. - gp : 1. + gp);  // zeta = -gp for q < 4
    const double nhi = 0.125 * fx * fy * (a >= 4 ? 1. + gp : 1. - gp);  // zeta = +gp for q >= 4
    tg = nlo * st[OFF_S + q4] + nhi * st[OFF_S + 4 + q4];
    tg += __shfl_xor_sync(0xffffffffu, tg, 1);
    tg += __shfl_xor_sync(0xffffffffu, tg, 2);
  }

  // ---- ordered accumulation ----
  const bool own = accoff >= 0;
  if (own && q4 == 0)
    while (ld_volatile_s32(sm.cnt + la) != rank) {}
  __threadfence_block();
  __syncwarp();
  if (own) {
    if (fl.J) {
      const int rs = 3 * nnb;
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        double* dst = sm.arena + accoff + 3 * ((slots >> (8 * bb)) & 255u);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) dst[i * rs + j] += C[i][j][bb];
      }
    }
    if (q4 == 0) {
      double* ex = sm.arena + extras_base + la * EXTRA;
      if (fl.R) { ex[0] += t0; ex[1] += t1; ex[2] += t2; }
      if (fl.G) ex[3] += tg;
      if (fl.J && fl.nv6) ex[4] += 1.0e-20 * mx;
    }
  }
  __threadfence_block();
  __syncwarp();
  if (own && q4 == 0) st_volatile_s32(sm.cnt + la, rank + 1);
}

// ---- tile flush: S -> K, Dirichlet semantics, coalesced write of the tile's CSR rows -----------------------
__device__ __forceinline__ void flush_tile(const SweepParams& p, const Smem& sm, const Flags& fl, const TileDesc& td,
                                           int extras_base, int warp, int lane) {
  for (int l = warp; l < td.n_owned; l += WARPS) {
    const int n = sm.gid[l];
    const int nnb = sm.nbrcnt[l];
    const int base = sm.accoff[l];
    const int rs = 3 * nnb;
    const unsigned cmn = sm.cm[l];
    const int nadj = p.n_adj_elem[n], nadjd = p.n_adj_diag[n];
    const int64_t row = p.node_row[n];
    const double* ex = sm.arena + extras_base + l * EXTRA;
    int self_c = -1;
    if (fl.J) {
      const uint16_t* nl = p.nbr_local + td.nbrl_begin + base / 9;
      for (int c0 = 0; c0 < nnb; c0 += 32) {
        const int c = c0 + lane;
        bool self = false;
        if (c < nnb) {
          double* blk = sm.arena + base + 3 * c;
          double S[3][3];
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) S[i][j] = blk[i * rs + j];
          const double trm = p.mu1 * (S[0][0] + S[1][1] + S[2][2]);
          const int ml = nl[c];
          self = (ml == l);
          const unsigned cmm = (ml != 0xFFFF) ? sm.cm[ml] : 0u;
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              double k = p.lam1 * S[i][j] + p.mu1 * S[j][i] + (i == j ? trm : 0.);
              if (((cmn >> i) | (cmm >> j)) & 1u) k = 0.;
              if (self && i == j) {
                if ((cmn >> i) & 1u) k = double(nadj + nadjd);  // each adjacent element sets the diagonal to 1
                else k += 1.0e-24 * nadjd;                      // diagonal_elem_subdomain_id elements
              }
              blk[i * rs + j] = k;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, self);
        if (bal) self_c = c0 + __ffs(bal) - 1;
      }
      __syncwarp();
      if (!fl.nv6) {
        double* dst = p.J + 9 * p.nbr_ptr[row];
        const double* src = sm.arena + base;
        for (int k = lane; k < 9 * nnb; k += 32) dst[k] = src[k];
      } else {
        double* dst = p.J + 36 * p.nbr_ptr[row];
        const int rl = 6 * nnb;
        for (int k = lane; k < 36 * nnb; k += 32) {
          const int i = k / rl, r = k - i * rl, c = r / 6, j = r - 6 * c;
          double val = 0.;
          if (i < 3 && j < 3) val = sm.arena[base + i * rs + 3 * c + j];
          else if (i == j && c == self_c) val = ((cmn >> i) & 1u) ? double(nadj + nadjd) : ex[4] + 1.0e-24 * nadjd;
          dst[k] = val;
        }
      }
    }
    if (fl.R && lane < p.nv) p.R[(int64_t)n * p.nv + lane] = (lane < 3 && !((cmn >> lane) & 1u)) ? ex[lane] : 0.;
    if (fl.G && lane == 0) p.G[n] = ex[3];
  }
}

__global__ void __launch_bounds__(THREADS, 1) hex8_sweep_kernel(const SweepParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem sm;
  sm.arena = reinterpret_cast<double*>(smem_raw);
  sm.xyz = sm.arena + p.arena_doubles;
  sm.u = sm.xyz + MAXN * 3;
  sm.adj = sm.u + MAXN * 3;
  sm.rho = sm.adj + MAXN * 3;
  sm.T = sm.rho + MAXN;
  sm.stage = sm.T + MAXN;
  sm.accoff = reinterpret_cast<int32_t*>(sm.stage + WARPS * STAGE_EL * EL_STRIDE);
  sm.gid = sm.accoff + MAXN;
  sm.cnt = sm.gid + MAXN;
  sm.rec = reinterpret_cast<uint32_t*>(sm.cnt + MAXN);
  sm.nbrcnt = reinterpret_cast<uint8_t*>(sm.rec + WARPS * 4 * REC_WORDS);
  sm.cm = sm.nbrcnt + MAXN;
  int& s_next = *reinterpret_cast<int*>(sm.cm + MAXN);  // chunk dispenser of the current tile

  Flags fl;
  fl.J = p.J != nullptr; fl.R = p.R != nullptr; fl.G = p.G != nullptr;
  fl.simp = p.rho != nullptr; fl.thermal = p.thermal != 0; fl.nv6 = p.nv == 6;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int el = lane >> 3, q = lane & 7;
  double* my_stage = sm.stage + warp * STAGE_EL * EL_STRIDE;
  uint32_t* my_rec = sm.rec + warp * 4 * REC_WORDS;

  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const TileDesc td = p.tiles[tile];
    // stage node state of the tile (owned + halo), clear accumulators and counters
    for (int l = tid; l < td.n_nodes; l += THREADS) {
      const int n = p.tile_nodes[td.node_begin + l];
      const unsigned cm = p.cmask ? p.cmask[n] : 0u;
      sm.gid[l] = n;
      sm.accoff[l] = p.tile_acc_off[td.node_begin + l];
      sm.nbrcnt[l] = (uint8_t)p.tile_nbr_cnt[td.node_begin + l];
      sm.cm[l] = (uint8_t)cm;
      sm.cnt[l] = 0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        sm.xyz[l * 3 + c] = p.xyz[(int64_t)n * 3 + c];
        sm.u[l * 3 + c] = p.X[(int64_t)n * p.nv + c];
        // constrain_element_vector zeroes constrained rows of dR/dp, i.e. the adjoint never sees them
        sm.adj[l * 3 + c] = (p.adj && !((cm >> c) & 1u)) ? p.adj[(int64_t)n * p.nv + c] : 0.;
      }
      sm.rho[l] = p.rho ? p.rho[n] : 1.;
      sm.T[l] = p.temperature ? p.temperature[n] : p.T_const;
    }
    for (int k = tid; k < td.arena_doubles; k += THREADS) sm.arena[k] = 0.;
    if (tid == 0) s_next = 0;
    __syncthreads();
    const int extras_base = td.arena_doubles - td.n_owned * EXTRA;
    const int n_chunks = (td.n_visits + 3) >> 2;
    const uint32_t* recs = reinterpret_cast<const uint32_t*>(p.visits + td.visit_begin);

    // warps take chunks of 4 consecutive visits in plan order; the records of the next chunk are prefetched into
    // registers while the current one is processed
    int c = 0;
    if (lane == 0) c = atomicAdd(&s_next, 1);
    c = __shfl_sync(0xffffffffu, c, 0);
    uint32_t r0 = 0, r1 = 0, r2 = 0;
    if (c < n_chunks) {
      const int nw = min(4, td.n_visits - 4 * c) * REC_WORDS;
      const uint32_t* src = recs + (int64_t)c * 4 * REC_WORDS;
      if (lane < nw) r0 = src[lane];
      if (lane + 32 < nw) r1 = src[lane + 32];
      if (lane + 64 < nw) r2 = src[lane + 64];
    }
    while (c < n_chunks) {
      const int nvis = min(4, td.n_visits - 4 * c);
      my_rec[lane] = r0; my_rec[lane + 32] = r1; my_rec[lane + 64] = r2;
      __syncwarp();
      int cn = 0;
      if (lane == 0) cn = atomicAdd(&s_next, 1);
      cn = __shfl_sync(0xffffffffu, cn, 0);
      if (cn < n_chunks) {
        const int nw = min(4, td.n_visits - 4 * cn) * REC_WORDS;
        const uint32_t* src = recs + (int64_t)cn * 4 * REC_WORDS;
        r0 = r1 = r2 = 0;
        if (lane < nw) r0 = src[lane];
        if (lane + 32 < nw) r1 = src[lane + 32];
        if (lane + 64 < nw) r2 = src[lane + 64];
      }
      const bool active = el < nvis;
      QpData d;
      if (active) {
        const uint4 pk = *reinterpret_cast<const uint4*>(my_rec + el * REC_WORDS);
        int ln[8] = {int(pk.x & 0xffffu), int(pk.x >> 16), int(pk.y & 0xffffu), int(pk.y >> 16),
                     int(pk.z & 0xffffu), int(pk.z >> 16), int(pk.w & 0xffffu), int(pk.w >> 16)};
        geometry_phase(p, sm, fl, ln, q, d);
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        __syncwarp();
        if (active && (el >> 1) == half) stage_write(my_stage + (el & 1) * EL_STRIDE, q, d);
        __syncwarp();
#pragma unroll
        for (int e2 = 0; e2 < 2; ++e2) {
          const int ecur = half * 2 + e2;
          if (ecur < nvis) element_phase(p, sm, fl, my_stage + e2 * EL_STRIDE, my_rec + ecur * REC_WORDS, extras_base, lane);
        }
      }
      __syncwarp();
      c = cn;
    }
    __syncthreads();
    flush_tile(p, sm, fl, td, extras_base, warp, lane);
    __syncthreads();
  }
}

}  // namespace

size_t hex8_sweep_smem_bytes(int arena_doubles) {
  return size_t(arena_doubles) * 8 + size_t(MAXN) * 11 * 8 + size_t(WARPS) * STAGE_EL * EL_STRIDE * 8 +
         size_t(MAXN) * (4 + 4 + 4 + 1 + 1) + size_t(WARPS) * 4 * REC_WORDS * 4 + 64;  // +64: chunk dispenser, padding
}

cudaError_t launch_hex8_sweep(const SweepParams& p, int n_ctas, cudaStream_t stream) {
  const size_t smem = hex8_sweep_smem_bytes(p.arena_doubles);
  cudaError_t e = cudaFuncSetAttribute(hex8_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  hex8_sweep_kernel<<<n_ctas, THREADS, smem, stream>>>(p);
  return cudaGetLastError();
}

int hex8_sweep_threads() { return THREADS; }

}  // namespace mastb200
