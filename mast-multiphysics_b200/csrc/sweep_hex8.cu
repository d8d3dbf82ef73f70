// sweep_hex8.cu -- HEX8 linear-elastic residual + Jacobian + density-sensitivity sweep (sm_100a, fp64).
//
// Device counterpart of the element loop of MAST::NonlinearImplicitAssembly::residual_and_jacobian
// (/root/reference/src/base/nonlinear_implicit_assembly.cpp:111-186) with
// StructuralElement3D::internal_residual (src/elasticity/solid_element_3d.cpp:135-423, linear strain),
// ::thermal_residual (:950-1089) and the sensitivity sweep of SURVEY 8a row S
// (sensitivity_assemble, nonlinear_implicit_assembly.cpp:515-686, for all nodal-density DVs at once).
//
// One persistent CTA per SM walks strips of the mesh (plan.h) front by front.  Per step:
//   visits   : warps take chunks of 4 element visits in plan order.
//     geometry phase : one lane per (element, quadrature point) builds J, det, the physical gradients g_b of the
//                      8 shape functions and the weight w = JxW*E(rho_q); gradients are staged in shared memory.
//     matrix phase   : S_ij[a][b] = sum_q w_q g_a[i] g_b[j] is a 24x24x8 product.  Only the six component pairs
//                      i<=j are formed (S_ji[a][b] = S_ij[b][a]) with 12 mma.sync.m8n8k4.f64 (DMMA); lane (a, q)
//                      holds exactly the (a, q) operand entry, so the A and B fragments are the same registers
//                      (B = w*A).
//     accumulate     : the 3x3 block S_ab of every node pair {a,b} with an owned end is added to the pair's
//                      shared-memory block (one block per unordered pair; S_ba = S_ab^T).  Per-node counters
//                      enforce the plan's visit order, so there are no atomics and the summation order into every
//                      CSR entry is fixed.
//   flush    : every node whose last element has been visited gets its rows written: K_ab = lam*S + mu*S^T +
//              mu*tr(S)*I (isotropic D, src/property_cards/isotropic_material_property_card.cpp:512-528, modulus
//              folded into w), the per-element Dirichlet semantics of libMesh's constrain_element_matrix_and_vector
//              (row/column zeroed, diagonal = number of adjacent elements, rhs zeroed), the internal force
//              f_a = sum_b K_ab u_b (linear strain) and one store of every CSR value to HBM.  K_e is never stored.
//   Warp 0 stages the nodal state of the next step and the flush records of the current one while the others compute.
#include <cuda_runtime.h>

#include "device_types.h"
#include "kernels.h"

namespace mastb200 {

namespace {

#ifndef MASTB200_WARPS
#define MASTB200_WARPS 12
#endif
constexpr int WARPS = MASTB200_WARPS;  // 12 warps x 168 registers fill the register file of one SM
constexpr int THREADS = WARPS * 32;
constexpr int STAGE_EL = 2;          // elements staged per warp at a time
constexpr int V_HALF = 36;           // stride of one (component, qp-half) 8x4 operand panel (32 + 4 pad)
constexpr int OFF_W = 216, OFF_PTH = 224, OFF_S = 232;
constexpr int STRIDE_PLAIN = 232;    // V 216 | w 8 | pad      (== 8 mod 16: the two staged elements hit disjoint banks)
constexpr int STRIDE_AUX = 248;      // V 216 | w 8 | pth 8 | s 8 | pad
constexpr int EXTRA = 5;             // per slot (aux layout): thermal f[3], G, rot-diag
constexpr int REC_WORDS = 24;        // sizeof(VisitRec8) / 4

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// HEX8 reference-node signs (libMesh node order)
__device__ __forceinline__ constexpr int sgx(int b) { return (b == 1 || b == 2 || b == 5 || b == 6) ? 1 : -1; }
__device__ __forceinline__ constexpr int sgy(int b) { return (b == 2 || b == 3 || b == 6 || b == 7) ? 1 : -1; }
__device__ __forceinline__ constexpr int sgz(int b) { return (b >= 4) ? 1 : -1; }

// Shared-memory layout: every offset is a compile-time constant, so all staged data is addressed as base + immediate.
template <bool AUX>
struct Lay {
  static constexpr int MS = kMaxSlots;
  static constexpr int STRIDE = AUX ? STRIDE_AUX : STRIDE_PLAIN;
  static constexpr int SMEM_TOTAL = kCtasPerSm == 1 ? 232448 : 233472 / kCtasPerSm - 1024;  // sm_100: 227 KB opt-in per CTA, 228 KB per SM
  static constexpr int FIXED = MS * 48 + (AUX ? MS * (24 + 8 + 8 + 8 * EXTRA) : 0) + WARPS * STAGE_EL * STRIDE * 8 +
                               kFlushBufBytes + WARPS * 4 * REC_WORDS * 4 + MS * 8 + 2 * 32 + 16 + MS;
  static constexpr int ARENA_RAW = ((SMEM_TOTAL - FIXED) / 72) & ~1;
  static constexpr int ARENA_BLOCKS = ARENA_RAW > 4094 ? 4094 : ARENA_RAW;
  static constexpr int O_XYZ = ARENA_BLOCKS * 72;
  static constexpr int O_U = O_XYZ + MS * 24;
  static constexpr int O_ADJ = O_U + MS * 24;
  static constexpr int O_RHO = O_ADJ + MS * 24;
  static constexpr int O_T = O_RHO + MS * 8;
  static constexpr int O_EX = O_T + MS * 8;
  static constexpr int O_STAGE = AUX ? O_EX + MS * 8 * EXTRA : O_ADJ;
  static constexpr int O_FBUF = O_STAGE + WARPS * STAGE_EL * STRIDE * 8;
  static constexpr int O_REC = O_FBUF + kFlushBufBytes;
  static constexpr int O_GID = O_REC + WARPS * 4 * REC_WORDS * 4;
  static constexpr int O_CNT = O_GID + MS * 4;
  static constexpr int O_SD = O_CNT + MS * 4;
  static constexpr int O_MISC = O_SD + 2 * 32;
  static constexpr int O_CM = O_MISC + 16;
  static constexpr int TOTAL = (O_CM + MS + 15) & ~15;
  static_assert(TOTAL <= SMEM_TOTAL, "shared-memory layout exceeds 227 KB");
  static_assert(O_FBUF % 16 == 0 && O_REC % 16 == 0 && O_STAGE % 16 == 0, "16-byte alignment of vector-accessed regions");
};

struct Smem {
  double* arena;   // pair blocks
  double* xyz;
  double* u;
  double* adj;     // aux
  double* rho;     // aux
  double* T;       // aux
  double* extras;  // aux
  double* stage;
  unsigned char* fbuf;
  uint32_t* rec;   // per warp: the 4 visit records of the current chunk
  int32_t* gid;
  int32_t* cnt;    // per slot: number of element contributions accumulated so far
  uint8_t* cm;
  StepDesc* sd;    // [2]
  int* misc;       // [0] chunk dispenser, [1] current tile
};

template <bool AUX>
__device__ __forceinline__ Smem make_smem(unsigned char* base) {
  using L = Lay<AUX>;
  Smem sm;
  sm.arena = reinterpret_cast<double*>(base);
  sm.xyz = reinterpret_cast<double*>(base + L::O_XYZ);
  sm.u = reinterpret_cast<double*>(base + L::O_U);
  sm.adj = reinterpret_cast<double*>(base + L::O_ADJ);
  sm.rho = reinterpret_cast<double*>(base + L::O_RHO);
  sm.T = reinterpret_cast<double*>(base + L::O_T);
  sm.extras = reinterpret_cast<double*>(base + L::O_EX);
  sm.stage = reinterpret_cast<double*>(base + L::O_STAGE);
  sm.fbuf = base + L::O_FBUF;
  sm.rec = reinterpret_cast<uint32_t*>(base + L::O_REC);
  sm.gid = reinterpret_cast<int32_t*>(base + L::O_GID);
  sm.cnt = reinterpret_cast<int32_t*>(base + L::O_CNT);
  sm.sd = reinterpret_cast<StepDesc*>(base + L::O_SD);
  sm.misc = reinterpret_cast<int*>(base + L::O_MISC);
  sm.cm = base + L::O_CM;
  return sm;
}

// per-node accumulation counters in shared memory: acquire load / release store at CTA scope order the pair-block
// updates of successive visits without full fences
__device__ __forceinline__ int ld_acquire_s32(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_s32(int32_t* p, int v) {
  asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

struct Flags {
  bool J, R, G, simp, thermal, nv6, S;  // S: pair blocks are accumulated (J or R requested)
};

// ---- geometry phase: lane = (element slot el, quadrature point q) ---------------------------------------
template <bool AUX>
struct QpData {
  double g[8][3];  // physical shape-function gradients
  double w;        // JxW * E(rho_q)
  double pth;      // AUX: -JxW * E * alpha_k * (T - Tref)   (thermal stress, isotropic)
  double s;        // AUX: sensitivity density  JxW * dE/drho * eps(adj)^T Dhat eps(u)
};

template <bool AUX>
__device__ __forceinline__ void geometry_phase(const SweepParams& p, const Smem& sm, const Flags& fl, const int ln[8],
                                               int q, QpData<AUX>& d) {
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
  const double idet = __drcp_rn(det);
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

  const double JxW = det;  // Gauss weights are 1 for the 2x2x2 rule
  double E = p.E0;
  d.pth = 0.; d.s = 0.;
  if (AUX) {
    double N[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const double fx = sgx(b) > 0 ? xp : xm, fy = sgy(b) > 0 ? yp : ym, fz = sgz(b) > 0 ? zp : zm;
      N[b] = 0.125 * fx * fy * fz;
    }
    // modulus at the point: constant or SIMP of the interpolated nodal density (example_6.cpp:100-137)
    double dE = 0.;
    if (fl.simp) {
      double r = 0.;
#pragma unroll
      for (int b = 0; b < 8; ++b) r = fma(N[b], sm.rho[ln[b]], r);
      E = p.E0 * (p.rho_min + (1. - p.rho_min) * pow(r, p.penalty));
      if (fl.G) dE = p.E0 * (1. - p.rho_min) * p.penalty * pow(r, p.penalty - 1.);
    }
    if (fl.thermal && fl.R) {  // thermal_residual: f -= JxW B^T (D alpha dT)
      double t = 0.;
#pragma unroll
      for (int b = 0; b < 8; ++b) t = fma(N[b], sm.T[ln[b]], t);
      d.pth = -(JxW * E) * p.alpha_k * (t - p.T_ref);
    }
    if (fl.G) {
      double H[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};  // H[i][j] = d u_i / d x_j
      double A[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const double* U = sm.u + ln[b] * 3;
        const double* L = sm.adj + ln[b] * 3;
        const double u0 = U[0], u1 = U[1], u2 = U[2], l0 = L[0], l1 = L[1], l2 = L[2];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          H[0][j] = fma(u0, d.g[b][j], H[0][j]);
          H[1][j] = fma(u1, d.g[b][j], H[1][j]);
          H[2][j] = fma(u2, d.g[b][j], H[2][j]);
          A[0][j] = fma(l0, d.g[b][j], A[0][j]);
          A[1][j] = fma(l1, d.g[b][j], A[1][j]);
          A[2][j] = fma(l2, d.g[b][j], A[2][j]);
        }
      }
      const double e0 = H[0][0], e1 = H[1][1], e2 = H[2][2];
      const double e3 = H[0][1] + H[1][0], e4 = H[1][2] + H[2][1], e5 = H[0][2] + H[2][0];
      const double a0 = A[0][0], a1 = A[1][1], a2 = A[2][2];
      const double a3 = A[0][1] + A[1][0], a4 = A[1][2] + A[2][1], a5 = A[0][2] + A[2][0];
      const double dens = p.lam1 * (a0 + a1 + a2) * (e0 + e1 + e2) +
                          p.mu1 * (2. * (a0 * e0 + a1 * e1 + a2 * e2) + a3 * e3 + a4 * e4 + a5 * e5);
      d.s = JxW * dE * dens;
    }
  }
  d.w = JxW * E;
}

template <bool AUX>
__device__ __forceinline__ void stage_write(double* st, int q, const QpData<AUX>& d) {
  const int h = q >> 2, q4 = q & 3;
#pragma unroll
  for (int b = 0; b < 8; ++b)
#pragma unroll
    for (int i = 0; i < 3; ++i) st[(i * 2 + h) * V_HALF + b * 4 + q4] = d.g[b][i];
  st[OFF_W + q] = d.w;
  if (AUX) { st[OFF_PTH + q] = d.pth; st[OFF_S + q] = d.s; }
}

// ---- matrix phase: all 32 lanes on one element; lane = (row node a, qp q4 | q4+4) ------------------------------
// rec: the visit record in shared memory.  All arithmetic (DMMA, lane reductions) happens first; the adds into the
// pair blocks wait until the per-node counters say that every earlier visit of the plan order has landed.
struct LaneConst {     // matrix-phase constants of lane (a = lane/4, q4 = lane%4), columns b = 2*q4 + bb
  int pidx[2];         // index of pair {a,b} in VisitRec8::pair
  unsigned swap[2];    // a > b: the lane holds the transposed orientation of the pair's (lo,hi) order
  bool le[2];          // a <= b: the lane owns the symmetric components (i,i) of the pair
};

__device__ __forceinline__ LaneConst make_lane_const(int lane) {
  LaneConst lc;
  const int a = lane >> 2, q4 = lane & 3;
#pragma unroll
  for (int bb = 0; bb < 2; ++bb) {
    const int b = 2 * q4 + bb;
    const int lo = min(a, b), hi = max(a, b);
    lc.pidx[bb] = lo * (15 - lo) / 2 + hi;
    lc.swap[bb] = a > b ? 1u : 0u;
    lc.le[bb] = a <= b;
  }
  return lc;
}

template <bool AUX>
__device__ __forceinline__ void element_phase(const SweepParams& p, const Smem& sm, const Flags& fl, const LaneConst& lc,
                                              const double* st, const uint32_t* rec, int lane) {
  const int a = lane >> 2, q4 = lane & 3;
  const uint16_t* rec16 = reinterpret_cast<const uint16_t*>(rec);
  const uint8_t* rec8 = reinterpret_cast<const uint8_t*>(rec);
  const int la = rec16[a];
  const int rank = rec8[16 + a];
  const bool own = rank != 0xFF;
  double A[3][2];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    A[i][0] = st[(i * 2 + 0) * V_HALF + lane];
    A[i][1] = st[(i * 2 + 1) * V_HALF + lane];
  }
  // C[k][bb]: component pairs k = 0..5 <-> (0,0),(1,1),(2,2),(0,1),(0,2),(1,2); columns b = 2*q4 + bb
  double C[6][2];
  double mx = 0., t0 = 0., t1 = 0., t2 = 0., tg = 0.;
  if (fl.S) {
    const double w0 = st[OFF_W + q4], w1 = st[OFF_W + 4 + q4];
    double B[3][2];
#pragma unroll
    for (int j = 0; j < 3; ++j) { B[j][0] = A[j][0] * w0; B[j][1] = A[j][1] * w1; }
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
    if (AUX && fl.nv6) {
      // 1e-20*max(diag K_e) on the rotational diagonal (solid_element_3d.cpp:413-417); diag of K_e from S_aa
      const int bself = a - 2 * q4;  // the rotational diagonal itself is zero, so the max is at least 0
      if (bself == 0 || bself == 1) {
        const double s0 = C[0][bself & 1], s1 = C[1][bself & 1], s2 = C[2][bself & 1];
        const double trm = p.mu1 * (s0 + s1 + s2);
        mx = fmax(mx, fmax((p.lam1 + p.mu1) * s0 + trm, fmax((p.lam1 + p.mu1) * s1 + trm, (p.lam1 + p.mu1) * s2 + trm)));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
  }
  if (AUX && fl.thermal && fl.R) {
    const double p0 = st[OFF_PTH + q4], p1 = st[OFF_PTH + 4 + q4];
    t0 = p0 * A[0][0] + p1 * A[0][1];
    t1 = p0 * A[1][0] + p1 * A[1][1];
    t2 = p0 * A[2][0] + p1 * A[2][1];
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
      t0 += __shfl_xor_sync(0xffffffffu, t0, o);
      t1 += __shfl_xor_sync(0xffffffffu, t1, o);
      t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    }
  }
  if (AUX && fl.G) {
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
  if (own && q4 == 0)
    while (ld_acquire_s32(sm.cnt + la) != rank) __nanosleep(20);
  __syncwarp();
  if (fl.S) {
    // both pair entries of the lane: columns b = 2*q4 and 2*q4+1 of row a
    const unsigned e2[2] = {rec16[12 + lc.pidx[0]], rec16[12 + lc.pidx[1]]};
#pragma unroll
    for (int bb = 0; bb < 2; ++bb) {
      const unsigned e = e2[bb];
      if (e != kPairNone) {
        double* blk = sm.arena + (e & kPairSlotMask) * 9;
        // stored block oriented as S_(a,b) (tr = 0) or S_(b,a) (tr = 1): off-diagonal components (i<j) of C_ij[a][b]
        // go to [i][j] = {1, 2, 5} or to [j][i] = {3, 6, 7}
        const unsigned tr = ((e >> 12) & 1u) ^ lc.swap[bb];
        double* p1 = blk + 2 * tr;  // [1]|[3] and [5]|[7]
        double* p2 = blk + 4 * tr;  // [2]|[6]
        const bool add = (e & kPairInit) == 0;
        const bool dg = lc.le[bb];
        double v01 = 0., v02 = 0., v12 = 0., v00 = 0., v11 = 0., v22 = 0.;
        if (add) {
          v01 = p1[1]; v02 = p2[2]; v12 = p1[5];
          if (dg) { v00 = blk[0]; v11 = blk[4]; v22 = blk[8]; }
        }
        p1[1] = v01 + C[3][bb]; p2[2] = v02 + C[4][bb]; p1[5] = v12 + C[5][bb];
        if (dg) { blk[0] = v00 + C[0][bb]; blk[4] = v11 + C[1][bb]; blk[8] = v22 + C[2][bb]; }
      }
    }
  }
  if (AUX && own && q4 == 0) {
    double* ex = sm.extras + la * EXTRA;
    if (fl.thermal && fl.R) { ex[0] += t0; ex[1] += t1; ex[2] += t2; }
    if (fl.G) ex[3] += tg;
    if (fl.S && fl.nv6) ex[4] += 1.0e-20 * mx;
  }
  __syncwarp();
  if (own && q4 == 0) st_release_s32(sm.cnt + la, rank + 1);
}

// ---- staging of nodal state (loads of one step) by `nthr` threads, this thread being number `t` ---------------
template <bool AUX>
__device__ __forceinline__ void load_states(const SweepParams& p, const Smem& sm, int begin, int n, int t, int nthr) {
  constexpr int UN = 4;
  for (int i0 = t; i0 < n; i0 += nthr * UN) {
    LoadRec lr[UN];
#pragma unroll
    for (int k = 0; k < UN; ++k) {
      const int i = i0 + k * nthr;
      if (i < n) lr[k] = p.loads[begin + i];
    }
#pragma unroll
    for (int k = 0; k < UN; ++k) {
      const int i = i0 + k * nthr;
      if (i >= n) continue;
      const int n_ = lr[k].gid, sl = lr[k].slot;
      const unsigned cm = p.cmask ? p.cmask[n_] : 0u;
      sm.gid[sl] = n_;
      sm.cnt[sl] = 0;
      sm.cm[sl] = (uint8_t)cm;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        sm.xyz[sl * 3 + c] = p.xyz[(int64_t)n_ * 3 + c];
        sm.u[sl * 3 + c] = p.X[(int64_t)n_ * p.nv + c];
      }
      if (AUX) {
        // constrain_element_vector zeroes constrained rows of dR/dp, i.e. the adjoint never sees them
#pragma unroll
        for (int c = 0; c < 3; ++c) sm.adj[sl * 3 + c] = (p.adj && !((cm >> c) & 1u)) ? p.adj[(int64_t)n_ * p.nv + c] : 0.;
        sm.rho[sl] = p.rho ? p.rho[n_] : 1.;
        sm.T[sl] = p.temperature ? p.temperature[n_] : p.T_const;
#pragma unroll
        for (int c = 0; c < EXTRA; ++c) sm.extras[sl * EXTRA + c] = 0.;
      }
    }
  }
}

// ---- flush: S -> K, Dirichlet semantics, f = K u, one store of every CSR value of the finished rows --------------
template <bool AUX>
__device__ __forceinline__ void flush_rows(const SweepParams& p, const Smem& sm, const Flags& fl, const StepDesc& sd,
                                           int warp, int lane) {
  const FlushRow* rows = reinterpret_cast<const FlushRow*>(sm.fbuf);
  const uint32_t* ents = reinterpret_cast<const uint32_t*>(sm.fbuf + (size_t)sd.n_flush * sizeof(FlushRow));
  for (int r = warp; r < sd.n_flush; r += WARPS) {
    const FlushRow fr = rows[r];
    const int nnb = fr.nnb;
    const int sa = fr.slot;
    const unsigned cmn = sm.cm[sa];
    const int n = fr.gid;
    const double nadj = double(fr.nadj + fr.nadjd), nadjd = double(fr.nadjd);
    const double* ex = AUX ? sm.extras + sa * EXTRA : nullptr;
    double r0 = 0., r1 = 0., r2 = 0.;
    for (int c0 = 0; c0 < nnb; c0 += 32) {
      const int c = c0 + lane;
      if (c < nnb) {
        const uint32_t e = ents[fr.ent_rel + c];
        const unsigned bs = e & kFlushNoBlock;
        const bool self = (e & kFlushSelf) != 0;
        const unsigned ns = e >> 16;
        double S[3][3];
        if (bs != kFlushNoBlock) {
          const double* blk = sm.arena + bs * 9;
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) S[i][j] = blk[i * 3 + j];
          if (self) { S[1][0] = S[0][1]; S[2][0] = S[0][2]; S[2][1] = S[1][2]; }
        } else {
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) S[i][j] = 0.;
        }
        // the block holds S_(row,nbr) or its transpose: K = lam*S + mu*S^T + mu*tr(S)*I either way, with lam/mu swapped
        const bool tr = (e & kFlushTransposed) != 0;
        const double l1 = tr ? p.mu1 : p.lam1, l2 = tr ? p.lam1 : p.mu1;
        const double trm = p.mu1 * (S[0][0] + S[1][1] + S[2][2]);
        double K[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) K[i][j] = l1 * S[i][j] + l2 * S[j][i] + (i == j ? trm : 0.);
        unsigned cmm = 0u;
        if (ns != 0xFFFFu) {
          cmm = sm.cm[ns];
          if (fl.R) {  // internal force of linear elasticity: f_a += K_ab u_b (unconstrained element matrices)
            const double u0 = sm.u[ns * 3], u1 = sm.u[ns * 3 + 1], u2 = sm.u[ns * 3 + 2];
            r0 += K[0][0] * u0 + K[0][1] * u1 + K[0][2] * u2;
            r1 += K[1][0] * u0 + K[1][1] * u1 + K[1][2] * u2;
            r2 += K[2][0] * u0 + K[2][1] * u1 + K[2][2] * u2;
          }
        }
        if (fl.J) {
          if (cmn | cmm | (unsigned)fr.nadjd) {  // Dirichlet rows/columns and diagonal_elem_subdomain_id terms (rare)
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                double k = K[i][j];
                if (((cmn >> i) | (cmm >> j)) & 1u) k = 0.;
                if (self && i == j) {
                  if ((cmn >> i) & 1u) k = nadj;       // each adjacent element sets the diagonal to 1
                  else k += 1.0e-24 * nadjd;           // diagonal_elem_subdomain_id elements
                }
                K[i][j] = k;
              }
          }
          if (!fl.nv6) {
            double* dst = p.J + 9 * fr.dst_block + 3 * c;
            const int rs = 3 * nnb;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j) dst[i * rs + j] = K[i][j];
          } else {
            double* dst = p.J + 36 * fr.dst_block + 6 * c;
            const int rl = 6 * nnb;
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
              for (int j = 0; j < 6; ++j) {
                double val = 0.;
                if (i < 3 && j < 3) val = K[i][j];
                else if (i == j && self) val = ((cmn >> i) & 1u) ? nadj : (AUX ? ex[4] : 0.) + 1.0e-24 * nadjd;
                dst[i * rl + j] = val;
              }
          }
        }
      }
    }
    if (fl.R) {
      // deterministic butterfly over the 32 lanes for the three sums at once: lanes 0, 8 and 16 end with r0, r1, r2
      const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
      const double g0 = __shfl_xor_sync(0xffffffffu, hi16 ? r0 : r2, 16);
      const double g1 = __shfl_xor_sync(0xffffffffu, hi16 ? r1 : 0., 16);
      const double x0 = (hi16 ? r2 : r0) + g0, x1 = (hi16 ? 0. : r1) + g1;
      double y = (hi8 ? x1 : x0) + __shfl_xor_sync(0xffffffffu, hi8 ? x0 : x1, 8);
      y += __shfl_xor_sync(0xffffffffu, y, 4);
      y += __shfl_xor_sync(0xffffffffu, y, 2);
      y += __shfl_xor_sync(0xffffffffu, y, 1);
      if ((lane & 7) == 0 && lane < 24) {
        const int i = lane >> 3;
        double v = y;
        if (AUX) v += ex[i];
        p.R[(int64_t)n * p.nv + i] = ((cmn >> i) & 1u) ? 0. : v;
      } else if (fl.nv6 && lane >= 25 && lane < 28) {
        p.R[(int64_t)n * p.nv + (lane - 22)] = 0.;  // rotational dofs of the 6-variable system carry no residual
      }
    }
    if (AUX && fl.G && lane == 0) p.G[n] = ex[3];
  }
}

template <bool AUX>
__global__ void __launch_bounds__(THREADS, kCtasPerSm) hex8_sweep_kernel(const SweepParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int STRIDE = Lay<AUX>::STRIDE;
  const Smem sm = make_smem<AUX>(smem_raw);

  Flags fl;
  fl.J = p.J != nullptr; fl.R = p.R != nullptr; fl.G = p.G != nullptr;
  fl.simp = p.rho != nullptr; fl.thermal = p.thermal != 0; fl.nv6 = p.nv == 6;
  fl.S = fl.J || fl.R;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int el = lane >> 3, q = lane & 7;
  const LaneConst lc = make_lane_const(lane);
  double* my_stage = sm.stage + warp * STAGE_EL * STRIDE;
  uint32_t* my_rec = sm.rec + warp * 4 * REC_WORDS;
  int& s_next = sm.misc[0];
  int& s_tile = sm.misc[1];

  for (;;) {
    if (tid == 0) s_tile = atomicAdd(p.tile_counter, 1);
    __syncthreads();
    const int tile = s_tile;
    if (tile >= p.n_tiles) break;
    const TileDesc td = p.tiles[tile];
    if (tid < 8) reinterpret_cast<int*>(&sm.sd[0])[tid] = reinterpret_cast<const int*>(p.steps + td.step_begin)[tid];
    if (tid == 0) s_next = 0;
    __syncthreads();
    load_states<AUX>(p, sm, sm.sd[0].load_begin, sm.sd[0].n_loads, tid, THREADS);
    __syncthreads();

    for (int s = 0; s < td.n_steps; ++s) {
      const StepDesc sd = sm.sd[s & 1];
      if (warp == 0) {
        // stage the flush records of this step and the nodal state of the next one
        if (s + 1 < td.n_steps) {
          StepDesc* nsd = &sm.sd[(s + 1) & 1];
          if (lane < 8) reinterpret_cast<int*>(nsd)[lane] = reinterpret_cast<const int*>(p.steps + td.step_begin + s + 1)[lane];
          __syncwarp();
          load_states<AUX>(p, sm, nsd->load_begin, nsd->n_loads, lane, 32);
        }
        const uint4* src_r = reinterpret_cast<const uint4*>(p.flush_rows + sd.flush_begin);
        uint4* dst_r = reinterpret_cast<uint4*>(sm.fbuf);
        for (int k = lane; k < sd.n_flush * 2; k += 32) dst_r[k] = src_r[k];
        const uint4* src_e = reinterpret_cast<const uint4*>(p.flush_ent + sd.fent_begin);
        uint4* dst_e = reinterpret_cast<uint4*>(sm.fbuf + (size_t)sd.n_flush * sizeof(FlushRow));
        for (int k = lane; k < (sd.n_fent + 3) / 4; k += 32) dst_e[k] = src_e[k];
      }
      // warps take chunks of consecutive visits in plan order: 4 visits while more than a round of work is left, then
      // the remainder split evenly over the warps so that the step ends without a tail; the records of the next chunk
      // are prefetched into registers while the current one is processed
      const int full = sd.n_visits / (4 * WARPS) * WARPS;          // chunks of 4
      const int rem = sd.n_visits - 4 * full;
      const int base = rem / WARPS, extra = rem - base * WARPS;     // last round: `extra` chunks of base+1, then base
      const int n_chunks = full + (base ? WARPS : extra);
      const uint32_t* recs = reinterpret_cast<const uint32_t*>(p.visits + sd.visit_begin);
      auto chunk_range = [&](int c, int& begin, int& size) {
        if (c < full) { begin = 4 * c; size = 4; }
        else { const int j = c - full; begin = 4 * full + j * base + min(j, extra); size = base + (j < extra ? 1 : 0); }
      };
      int c = 0;
      if (lane == 0) c = atomicAdd(&s_next, 1);
      c = __shfl_sync(0xffffffffu, c, 0);
      uint32_t r0 = 0, r1 = 0, r2 = 0;
      int nvis = 0, nvis_next = 0;
      if (c < n_chunks) {
        int b0;
        chunk_range(c, b0, nvis);
        const int nw = nvis * REC_WORDS;
        const uint32_t* src = recs + (int64_t)b0 * REC_WORDS;
        if (lane < nw) r0 = src[lane];
        if (lane + 32 < nw) r1 = src[lane + 32];
        if (lane + 64 < nw) r2 = src[lane + 64];
      }
      while (c < n_chunks) {
        my_rec[lane] = r0; my_rec[lane + 32] = r1; my_rec[lane + 64] = r2;
        __syncwarp();
        int cn = 0;
        if (lane == 0) cn = atomicAdd(&s_next, 1);
        cn = __shfl_sync(0xffffffffu, cn, 0);
        if (cn < n_chunks) {
          int b0;
          chunk_range(cn, b0, nvis_next);
          const int nw = nvis_next * REC_WORDS;
          const uint32_t* src = recs + (int64_t)b0 * REC_WORDS;
          r0 = r1 = r2 = 0;
          if (lane < nw) r0 = src[lane];
          if (lane + 32 < nw) r1 = src[lane + 32];
          if (lane + 64 < nw) r2 = src[lane + 64];
        }
        const bool active = el < nvis;
        QpData<AUX> d;
        if (active) {
          const uint4 pk = *reinterpret_cast<const uint4*>(my_rec + el * REC_WORDS);
          int ln[8] = {int(pk.x & 0xffffu), int(pk.x >> 16), int(pk.y & 0xffffu), int(pk.y >> 16),
                       int(pk.z & 0xffffu), int(pk.z >> 16), int(pk.w & 0xffffu), int(pk.w >> 16)};
          geometry_phase<AUX>(p, sm, fl, ln, q, d);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          __syncwarp();
          if (active && (el >> 1) == half) stage_write<AUX>(my_stage + (el & 1) * STRIDE, q, d);
          __syncwarp();
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            const int ecur = half * 2 + e2;
            if (ecur < nvis) element_phase<AUX>(p, sm, fl, lc, my_stage + e2 * STRIDE, my_rec + ecur * REC_WORDS, lane);
          }
        }
        __syncwarp();
        c = cn; nvis = nvis_next;
      }
      __syncthreads();
      if (tid == 0) s_next = 0;
      flush_rows<AUX>(p, sm, fl, sd, warp, lane);
      __syncthreads();
    }
  }
}

}  // namespace

size_t hex8_sweep_smem_bytes(bool aux) { return aux ? Lay<true>::TOTAL : Lay<false>::TOTAL; }
int hex8_sweep_arena_blocks(bool aux) { return aux ? Lay<true>::ARENA_BLOCKS : Lay<false>::ARENA_BLOCKS; }

cudaError_t launch_hex8_sweep(const SweepParams& p, int n_ctas, cudaStream_t stream) {
  const size_t smem = hex8_sweep_smem_bytes(p.aux != 0);
  cudaError_t e = cudaMemsetAsync(p.tile_counter, 0, sizeof(int32_t), stream);
  if (e != cudaSuccess) return e;
  if (p.aux) {
    e = cudaFuncSetAttribute(hex8_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hex8_sweep_kernel<true><<<n_ctas, THREADS, smem, stream>>>(p);
  } else {
    e = cudaFuncSetAttribute(hex8_sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hex8_sweep_kernel<false><<<n_ctas, THREADS, smem, stream>>>(p);
  }
  return cudaGetLastError();
}

int hex8_sweep_threads() { return THREADS; }

}  // namespace mastb200
