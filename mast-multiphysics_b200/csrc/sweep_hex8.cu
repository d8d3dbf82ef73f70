// sweep_hex8.cu -- HEX8 linear-elastic residual + Jacobian + density-sensitivity sweep (sm_100a, fp64).
//
// Device counterpart of the element loop of MAST::NonlinearImplicitAssembly::residual_and_jacobian
// (/root/reference/src/base/nonlinear_implicit_assembly.cpp:111-186) with
// StructuralElement3D::internal_residual (src/elasticity/solid_element_3d.cpp:135-423, linear strain),
// ::thermal_residual (:950-1089) and the sensitivity sweep of SURVEY 8a row S
// (sensitivity_assemble, nonlinear_implicit_assembly.cpp:515-686, for all nodal-density DVs at once).
//
// One persistent CTA per SM walks strips of the mesh (plan.h) front by front.  The CTA is warp-specialised:
//   12 compute warps each walk their own row of the plan's per-step visit grid, four visits at a time.
//     geometry phase : one lane per (element, quadrature point) builds J, its inverse and the weight
//                      w = JxW*E(rho_q) and stages them (10 doubles) in shared memory.
//     matrix phase   : all 32 lanes on one element, lane = (node a, quadrature points q4 and q4+4).  The lane forms
//                      its own operand entries g_a = J^-T dN_a (the reference gradients are lane constants);
//                      S_ij[a][b] = sum_q w_q g_a[i] g_b[j] is a 24x24x8 product of which only the six component
//                      pairs i<=j are formed (S_ji[a][b] = S_ij[b][a]) with 12 mma.sync.m8n8k4.f64 (DMMA); the A and
//                      B fragments are the same registers (B = w*A).
//     accumulate     : the 3x3 block S_ab of every node pair {a,b} with an owned end is added to the pair's
//                      shared-memory block (one block per unordered pair; S_ba = S_ab^T).  Per-node counters
//                      (acquire/release) enforce the plan's visit order: no atomics, fixed summation order.
//   4 flush warps run one step behind: every node whose last element has been visited gets its rows written:
//              K_ab = lam*S + mu*S^T + mu*tr(S)*I (isotropic D, src/property_cards/isotropic_material_property_card.cpp:
//              512-528, modulus folded into w), the per-element Dirichlet semantics of libMesh's
//              constrain_element_matrix_and_vector (row/column zeroed, diagonal = number of adjacent elements, rhs
//              zeroed), the internal force f_a = sum_b K_ab u_b (linear strain) and one store of every CSR value to
//              HBM.  K_e is never stored.
//   1 staging warp fetches flush records, nodal state and step descriptors one to three steps ahead.
//   The groups meet only at named barriers (visits of step s done -> flush s; flush s done -> visits s+2).
#include <cuda_runtime.h>

#include "device_types.h"
#include "kernels.h"

namespace mastb200 {

namespace {

__device__ __forceinline__ void stg256(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void stg128(double* p, double a, double b) {
  asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}


#ifndef MASTB200_NC
#define MASTB200_NC 11
#endif
constexpr int NC = MASTB200_NC;      // compute warps (11 + 4 + 1 = 16 warps x 128 registers fill the register file)
constexpr int NF = 4;                // flush warps
constexpr int WARPS = NC + NF + 1;   // + one staging warp
constexpr int THREADS = WARPS * 32;
constexpr int QS_PLAIN = 10;         // doubles staged per (element, qp): inv[3][3], w
constexpr int QS_AUX = 14;           // + pth, s, pad (stride of 28 words keeps the 128-bit stores conflict-free)
constexpr int EXTRA = 5;             // per slot (aux layout): thermal f[3], G, rot-diag
constexpr int REC_WORDS = 24;        // sizeof(VisitRec8) / 4
constexpr double GP = 0.57735026918962576451;

// named barriers (0 is __syncthreads)
constexpr int BAR_VISITS = 1;        // +parity: visits of step s done, its flush records staged (compute, stager arrive; flush sync)
constexpr int BAR_FLUSH = 3;         // +parity: flush of step s done, state of step s+2 staged (flush, stager arrive; compute sync)
constexpr int BAR_FDONE = 5;         // +parity: flush of step s done (flush arrive; stager sync)

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

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
  static constexpr int QS = AUX ? QS_AUX : QS_PLAIN;
  static constexpr int STAGE_WARP = 4 * 8 * QS * 8;   // bytes staged per compute warp: 4 elements x 8 qps
  static constexpr int SMEM_TOTAL = 232448;            // sm_100: 227 KB opt-in per CTA
  static constexpr int FIXED = MS * 48 + (AUX ? MS * (24 + 8 + 8 + 8 * EXTRA) : 0) + NC * STAGE_WARP + 2 * kFlushBufBytes +
                               NC * 4 * REC_WORDS * 4 + MS * 8 + 4 * 48 + 16 + MS;
  static constexpr int ARENA_RAW = ((SMEM_TOTAL - FIXED) / 72) & ~1;
  static constexpr int ARENA_BLOCKS = ARENA_RAW > 4094 ? 4094 : ARENA_RAW;
  static constexpr int O_XYZ = ARENA_BLOCKS * 72;
  static constexpr int O_U = O_XYZ + MS * 24;
  static constexpr int O_ADJ = O_U + MS * 24;
  static constexpr int O_RHO = O_ADJ + MS * 24;
  static constexpr int O_T = O_RHO + MS * 8;
  static constexpr int O_EX = O_T + MS * 8;
  static constexpr int O_STAGE = AUX ? O_EX + MS * 8 * EXTRA : O_ADJ;
  static constexpr int O_FBUF = O_STAGE + NC * STAGE_WARP;
  static constexpr int O_REC = O_FBUF + 2 * kFlushBufBytes;
  static constexpr int O_GID = O_REC + NC * 4 * REC_WORDS * 4;
  static constexpr int O_CNT = O_GID + MS * 4;
  static constexpr int O_SD = O_CNT + MS * 4;
  static constexpr int O_MISC = O_SD + 4 * 48;
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
  unsigned char* fbuf;   // two buffers of kFlushBufBytes (step parity)
  uint32_t* rec;   // per compute warp: the 4 visit records of the current chunk
  int32_t* gid;
  int32_t* cnt;    // per slot: number of element contributions accumulated so far
  uint8_t* cm;
  StepDesc* sd;    // ring of 4
  int* misc;       // [0] current tile
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

// ---- geometry phase: lane = (element el, quadrature point q); stages inv = J^-1 (inv[c][r] = d xi_r / d x_c) and w ---
template <bool AUX>
__device__ __forceinline__ void geometry_phase(const SweepParams& p, const Smem& sm, const Flags& fl, const int ln[8],
                                               int q, double* out) {
  const double xi = (q & 1) ? GP : -GP, eta = (q & 2) ? GP : -GP, zeta = (q & 4) ? GP : -GP;
  const double xm = 1. - xi, xp = 1. + xi, ym = 1. - eta, yp = 1. + eta, zm = 1. - zeta, zp = 1. + zeta;
  double Jm[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};  // Jm[r][c] = d x_c / d xi_r
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const double fx = sgx(b) > 0 ? xp : xm, fy = sgy(b) > 0 ? yp : ym, fz = sgz(b) > 0 ? zp : zm;
    const double d0 = 0.125 * sgx(b) * fy * fz, d1 = 0.125 * sgy(b) * fx * fz, d2 = 0.125 * sgz(b) * fx * fy;
    const double* X = sm.xyz + ln[b] * 3;
    const double x0 = X[0], x1 = X[1], x2 = X[2];
    Jm[0][0] = fma(d0, x0, Jm[0][0]); Jm[0][1] = fma(d0, x1, Jm[0][1]); Jm[0][2] = fma(d0, x2, Jm[0][2]);
    Jm[1][0] = fma(d1, x0, Jm[1][0]); Jm[1][1] = fma(d1, x1, Jm[1][1]); Jm[1][2] = fma(d1, x2, Jm[1][2]);
    Jm[2][0] = fma(d2, x0, Jm[2][0]); Jm[2][1] = fma(d2, x1, Jm[2][1]); Jm[2][2] = fma(d2, x2, Jm[2][2]);
  }
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

  const double JxW = det;  // Gauss weights are 1 for the 2x2x2 rule
  double E = p.E0, pth = 0., s = 0.;
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
      E = p.E0 * (p.rho_min + (1. - p.rho_min) * simp_pow(r, p.penalty));
      if (fl.G) dE = p.E0 * (1. - p.rho_min) * p.penalty * pow(r, p.penalty - 1.);
    }
    if (fl.thermal && fl.R) {  // thermal_residual: f -= JxW B^T (D alpha dT)
      double t = 0.;
#pragma unroll
      for (int b = 0; b < 8; ++b) t = fma(N[b], sm.T[ln[b]], t);
      pth = -(JxW * E) * p.alpha_k * (t - p.T_ref);
    }
    if (fl.G) {
      double H[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};  // H[i][j] = d u_i / d x_j
      double A[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const double fx = sgx(b) > 0 ? xp : xm, fy = sgy(b) > 0 ? yp : ym, fz = sgz(b) > 0 ? zp : zm;
        const double d0 = 0.125 * sgx(b) * fy * fz, d1 = 0.125 * sgy(b) * fx * fz, d2 = 0.125 * sgz(b) * fx * fy;
        double g[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) g[c] = d0 * inv[c][0] + d1 * inv[c][1] + d2 * inv[c][2];
        const double* U = sm.u + ln[b] * 3;
        const double* L = sm.adj + ln[b] * 3;
        const double u0 = U[0], u1 = U[1], u2 = U[2], l0 = L[0], l1 = L[1], l2 = L[2];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          H[0][j] = fma(u0, g[j], H[0][j]); H[1][j] = fma(u1, g[j], H[1][j]); H[2][j] = fma(u2, g[j], H[2][j]);
          A[0][j] = fma(l0, g[j], A[0][j]); A[1][j] = fma(l1, g[j], A[1][j]); A[2][j] = fma(l2, g[j], A[2][j]);
        }
      }
      const double e0 = H[0][0], e1 = H[1][1], e2 = H[2][2];
      const double e3 = H[0][1] + H[1][0], e4 = H[1][2] + H[2][1], e5 = H[0][2] + H[2][0];
      const double a0 = A[0][0], a1 = A[1][1], a2 = A[2][2];
      const double a3 = A[0][1] + A[1][0], a4 = A[1][2] + A[2][1], a5 = A[0][2] + A[2][0];
      const double dens = p.lam1 * (a0 + a1 + a2) * (e0 + e1 + e2) +
                          p.mu1 * (2. * (a0 * e0 + a1 * e1 + a2 * e2) + a3 * e3 + a4 * e4 + a5 * e5);
      s = JxW * dE * dens;
    }
  }
  double2* o = reinterpret_cast<double2*>(out);
  o[0] = make_double2(inv[0][0], inv[0][1]);
  o[1] = make_double2(inv[0][2], inv[1][0]);
  o[2] = make_double2(inv[1][1], inv[1][2]);
  o[3] = make_double2(inv[2][0], inv[2][1]);
  o[4] = make_double2(inv[2][2], JxW * E);
  if (AUX) o[5] = make_double2(pth, s);
}

// ---- matrix phase: all 32 lanes on one element; lane = (row node a, qp q4 | q4+4) ------------------------------
struct LaneConst {     // constants of lane (a = lane/4, q4 = lane%4); columns b = 2*q4 + bb
  double d0[2], d1[2], d2;  // reference gradients of N_a at qps q4 (h = 0) and q4+4 (h = 1)
  double nlo, nhi;          // N_a at those qps (aux)
  unsigned bits;            // pair indices (bits 0-5, 8-13), a > b flags (16, 17), a <= b flags (18, 19)
};

__device__ __forceinline__ LaneConst make_lane_const(int lane) {
  LaneConst lc;
  const int a = lane >> 2, q4 = lane & 3;
  const double sx = (((a + 1) >> 1) & 1) ? 1. : -1., sy = ((a >> 1) & 1) ? 1. : -1., sz = (a >> 2) ? 1. : -1.;
  const double fx = 1. + sx * ((q4 & 1) ? GP : -GP), fy = 1. + sy * ((q4 & 2) ? GP : -GP);
  const double fz0 = 1. - sz * GP, fz1 = 1. + sz * GP;
  lc.d0[0] = 0.125 * sx * fy * fz0; lc.d0[1] = 0.125 * sx * fy * fz1;
  lc.d1[0] = 0.125 * sy * fx * fz0; lc.d1[1] = 0.125 * sy * fx * fz1;
  lc.d2 = 0.125 * sz * fx * fy;
  lc.nlo = 0.125 * fx * fy * fz0; lc.nhi = 0.125 * fx * fy * fz1;
  unsigned bits = 0;
#pragma unroll
  for (int bb = 0; bb < 2; ++bb) {
    const int b = 2 * q4 + bb;
    const int lo = min(a, b), hi = max(a, b);
    bits |= (unsigned)(lo * (15 - lo) / 2 + hi) << (8 * bb);
    bits |= (a > b ? 1u : 0u) << (16 + bb);
    bits |= (a <= b ? 1u : 0u) << (18 + bb);
  }
  lc.bits = bits;
  return lc;
}

// st: the element's staged qp data; rec: the visit record in shared memory.  All arithmetic (DMMA, lane reductions)
// happens first; the adds into the pair blocks wait until the per-node counters say that every earlier visit of the
// plan order has landed.
template <bool AUX>
__device__ __forceinline__ void element_phase(const SweepParams& p, const Smem& sm, const Flags& fl, const LaneConst& lc,
                                              const double* st, const uint32_t* rec, int lane) {
  constexpr int QS = Lay<AUX>::QS;
  const int a = lane >> 2, q4 = lane & 3;
  const uint16_t* rec16 = reinterpret_cast<const uint16_t*>(rec);
  const uint8_t* rec8 = reinterpret_cast<const uint8_t*>(rec);
  const int la = rec16[a];
  const int rank = rec8[16 + a];
  const bool own = rank != 0xFF;
  // operand entries of the lane: g_a[i] at its two quadrature points, A[i][h] = sum_r inv_h[i][r] * dN_a/dxi_r
  double A[3][2], w[2], pth[2], ss[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const double2* s2 = reinterpret_cast<const double2*>(st + (q4 + 4 * h) * QS);
    const double2 v0 = s2[0], v1 = s2[1], v2 = s2[2], v3 = s2[3], v4 = s2[4];
    A[0][h] = v0.x * lc.d0[h] + v0.y * lc.d1[h] + v1.x * lc.d2;
    A[1][h] = v1.y * lc.d0[h] + v2.x * lc.d1[h] + v2.y * lc.d2;
    A[2][h] = v3.x * lc.d0[h] + v3.y * lc.d1[h] + v4.x * lc.d2;
    w[h] = v4.y;
    if (AUX) { const double2 v5 = s2[5]; pth[h] = v5.x; ss[h] = v5.y; }
  }
  // C[k][bb]: component pairs k = 0..5 <-> (0,0),(1,1),(2,2),(0,1),(0,2),(1,2); columns b = 2*q4 + bb
  double C[6][2];
  double mx = 0., t0 = 0., t1 = 0., t2 = 0., tg = 0.;
  if (fl.S) {
    double B[3][2];
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
    t0 = pth[0] * A[0][0] + pth[1] * A[0][1];
    t1 = pth[0] * A[1][0] + pth[1] * A[1][1];
    t2 = pth[0] * A[2][0] + pth[1] * A[2][1];
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
      t0 += __shfl_xor_sync(0xffffffffu, t0, o);
      t1 += __shfl_xor_sync(0xffffffffu, t1, o);
      t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    }
  }
  if (AUX && fl.G) {
    tg = lc.nlo * ss[0] + lc.nhi * ss[1];
    tg += __shfl_xor_sync(0xffffffffu, tg, 1);
    tg += __shfl_xor_sync(0xffffffffu, tg, 2);
  }

  // ---- ordered accumulation ----
  if (own && q4 == 0)
    while (ld_acquire_s32(sm.cnt + la) != rank) __nanosleep(20);
  __syncwarp();
  if (fl.S) {
#pragma unroll
    for (int bb = 0; bb < 2; ++bb) {
      const unsigned e = rec16[12 + ((lc.bits >> (8 * bb)) & 63u)];  // pair {a, b = 2*q4 + bb}
      if (e != kPairNone) {
        double* blk = sm.arena + (e & kPairSlotMask) * 9;
        // stored block oriented as S_(a,b) (tr = 0) or S_(b,a) (tr = 1): off-diagonal components (i<j) of C_ij[a][b]
        // go to [i][j] = {1, 2, 5} or to [j][i] = {3, 6, 7}
        const unsigned tr = ((e >> 12) ^ (lc.bits >> (16 + bb))) & 1u;
        double* p1 = blk + 2 * tr;  // [1]|[3] and [5]|[7]
        double* p2 = blk + 4 * tr;  // [2]|[6]
        const bool add = (e & kPairInit) == 0;
        const bool dg = ((lc.bits >> (18 + bb)) & 1u) != 0;
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
// The records and then the nodal data of up to four nodes per thread are fetched together so that the two dependent
// global-memory latencies are paid once per call, not once per node.
template <bool AUX>
__device__ __forceinline__ void load_states(const SweepParams& p, const Smem& sm, int begin, int n, int t, int nthr) {
  constexpr int UN = 4;
  for (int i0 = t; i0 < n; i0 += nthr * UN) {
    LoadRec lr[UN];
#pragma unroll
    for (int k = 0; k < UN; ++k) {
      const int i = i0 + k * nthr;
      lr[k] = p.loads[begin + min(i, n - 1)];
    }
    double x[UN][3], u[UN][3], ad[UN][3], rho[UN], T[UN];
    unsigned cm[UN];
#pragma unroll
    for (int k = 0; k < UN; ++k) {
      const int64_t g = lr[k].gid;
      cm[k] = p.cmask ? p.cmask[g] : 0u;
#pragma unroll
      for (int c = 0; c < 3; ++c) { x[k][c] = p.xyz[g * 3 + c]; u[k][c] = p.X[g * p.nv + c]; }
      if (AUX) {
#pragma unroll
        for (int c = 0; c < 3; ++c) ad[k][c] = p.adj ? p.adj[g * p.nv + c] : 0.;
        rho[k] = p.rho ? p.rho[g] : 1.;
        T[k] = p.temperature ? p.temperature[g] : p.T_const;
      }
    }
#pragma unroll
    for (int k = 0; k < UN; ++k) {
      if (i0 + k * nthr >= n) continue;
      const int sl = lr[k].slot;
      sm.gid[sl] = lr[k].gid;
      sm.cnt[sl] = 0;
      sm.cm[sl] = (uint8_t)cm[k];
#pragma unroll
      for (int c = 0; c < 3; ++c) { sm.xyz[sl * 3 + c] = x[k][c]; sm.u[sl * 3 + c] = u[k][c]; }
      if (AUX) {
        // constrain_element_vector zeroes constrained rows of dR/dp, i.e. the adjoint never sees them
#pragma unroll
        for (int c = 0; c < 3; ++c) sm.adj[sl * 3 + c] = ((cm[k] >> c) & 1u) ? 0. : ad[k][c];
        sm.rho[sl] = rho[k];
        sm.T[sl] = T[k];
#pragma unroll
        for (int c = 0; c < EXTRA; ++c) sm.extras[sl * EXTRA + c] = 0.;
      }
    }
  }
}

// flush records of one step -> shared memory buffer `buf` (by `nthr` threads), as asynchronous 16-byte copies: all of
// them are in flight at once; the caller waits with cp_async_wait_all() before publishing the buffer
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void load_flush_records(const SweepParams& p, unsigned char* buf, const StepDesc& sd, int t, int nthr) {
  const uint4* src_r = reinterpret_cast<const uint4*>(p.flush_rows + sd.flush_begin);
  uint4* dst_r = reinterpret_cast<uint4*>(buf);
  for (int k = t; k < sd.n_flush * 2; k += nthr) cp_async16(dst_r + k, src_r + k);
  const uint4* src_e = reinterpret_cast<const uint4*>(p.flush_ent + sd.fent_begin);
  uint4* dst_e = reinterpret_cast<uint4*>(buf + (size_t)sd.n_flush * sizeof(FlushRow));
  for (int k = t; k < (sd.n_fent + 3) / 4; k += nthr) cp_async16(dst_e + k, src_e + k);
}

// ---- flush: S -> K, Dirichlet semantics, f = K u, one store of every CSR value of the finished rows --------------
// four rows per warp at a time, eight lanes per row; a lane walks the row's neighbour blocks c = l8, l8+8, ...
template <bool AUX>
__device__ __forceinline__ void flush_rows(const SweepParams& p, const Smem& sm, const Flags& fl, const StepDesc& sd,
                                           const unsigned char* fbuf, int fwarp, int lane) {
  const FlushRow* rows = reinterpret_cast<const FlushRow*>(fbuf);
  const uint32_t* ents = reinterpret_cast<const uint32_t*>(fbuf + (size_t)sd.n_flush * sizeof(FlushRow));
  const int sub = lane >> 3, l8 = lane & 7;
  for (int rbase = fwarp * 4; rbase < sd.n_flush; rbase += NF * 4) {
    const int r = rbase + sub;
    const bool valid = r < sd.n_flush;
    const FlushRow fr = rows[valid ? r : rbase];
    const int nnb = valid ? fr.nnb : 0;
    const int sa = fr.slot;
    const unsigned cmn = sm.cm[sa];
    const int n = fr.gid;
    const double nadj = double(fr.nadj + fr.nadjd), nadjd = double(fr.nadjd);
    const double* ex = AUX ? sm.extras + sa * EXTRA : nullptr;
    const uint32_t* ent = ents + fr.ent_rel;
    double r0 = 0., r1 = 0., r2 = 0.;
    for (int c = l8; __any_sync(0xffffffffu, c < nnb); c += 8) {
      if (c < nnb) {
        const uint32_t e = ent[c];
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
          const bool wide = (reinterpret_cast<uintptr_t>(p.J) & 31) == 0;  // 32-byte aligned J: vector stores
          if (!fl.nv6) {
            double* dst = p.J + 9 * fr.dst_block + 3 * c;
            const int rs = 3 * nnb;
            // (16-byte + 8-byte stores of a row piece, which pay in sweep_hex8_struct.cu, measured 1.6 % slower here)
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j) dst[i * rs + j] = K[i][j];
          } else {
            double* dst = p.J + 36 * fr.dst_block + 6 * c;
            const int rl = 6 * nnb;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
              double v[6];
#pragma unroll
              for (int j = 0; j < 6; ++j) {
                double val = 0.;
                if (i < 3 && j < 3) val = K[i][j];
                else if (i == j && self) val = ((cmn >> i) & 1u) ? nadj : (AUX ? ex[4] : 0.) + 1.0e-24 * nadjd;
                v[j] = val;
              }
              double* d = dst + i * rl;
              if (wide) {
                // six values = 48 bytes starting at a multiple of 48 bytes: on a 32-byte boundary for every other piece --
                // one 32-byte and one 16-byte store instead of six 8-byte ones, chosen without a branch
                const bool odd = ((c + i * nnb) & 1) != 0;
                stg128(odd ? d : d + 4, odd ? v[0] : v[4], odd ? v[1] : v[5]);
                stg256(odd ? d + 2 : d, odd ? v[2] : v[0], odd ? v[3] : v[1], odd ? v[4] : v[2], odd ? v[5] : v[3]);
              } else {
#pragma unroll
                for (int j = 0; j < 6; ++j) d[j] = v[j];
              }
            }
          }
        }
      }
    }
    if (fl.R) {
      // deterministic butterfly over the 8 lanes of a row for the three sums at once: lanes 0, 2 and 4 of the group end
      // with r0, r1, r2
      const bool hi4 = (lane & 4) != 0, hi2 = (lane & 2) != 0;
      const double g0 = __shfl_xor_sync(0xffffffffu, hi4 ? r0 : r2, 4);
      const double g1 = __shfl_xor_sync(0xffffffffu, hi4 ? r1 : 0., 4);
      const double x0 = (hi4 ? r2 : r0) + g0, x1 = (hi4 ? 0. : r1) + g1;
      double y = (hi2 ? x1 : x0) + __shfl_xor_sync(0xffffffffu, hi2 ? x0 : x1, 2);
      y += __shfl_xor_sync(0xffffffffu, y, 1);
      if (valid) {
        if ((l8 & 1) == 0 && l8 < 6) {
          const int i = l8 >> 1;
          double v = y;
          if (AUX) v += ex[i];
          p.R[(int64_t)n * p.nv + i] = ((cmn >> i) & 1u) ? 0. : v;
        } else if (fl.nv6 && (l8 & 1)) {
          if (l8 < 6) p.R[(int64_t)n * p.nv + 3 + (l8 >> 1)] = 0.;  // rotational dofs of the 6-variable system: no residual
        }
      }
    }
    if (AUX && fl.G && valid && l8 == 0) p.G[n] = ex[3];
  }
}

template <bool AUX>
#ifdef MASTB200_MAXNREG
__global__ void __maxnreg__(MASTB200_MAXNREG) hex8_sweep_kernel(const SweepParams p) {
#else
__global__ void __launch_bounds__(THREADS, 1) hex8_sweep_kernel(const SweepParams p) {
#endif
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int QS = Lay<AUX>::QS;
  const Smem sm = make_smem<AUX>(smem_raw);

  Flags fl;
  fl.J = p.J != nullptr; fl.R = p.R != nullptr; fl.G = p.G != nullptr;
  fl.simp = p.rho != nullptr; fl.thermal = p.thermal != 0; fl.nv6 = p.nv == 6;
  fl.S = fl.J || fl.R;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int& s_tile = sm.misc[0];
  constexpr int SDW = sizeof(StepDesc) / 4;

  for (;;) {
    if (tid == 0) s_tile = atomicAdd(p.tile_counter, 1);
    __syncthreads();
    const int tile = s_tile;
    if (tile >= p.n_tiles) break;
    const TileDesc td = p.tiles[tile];
    // prologue: descriptors of the first three steps, state of the first two, flush records of the first
    if (tid < 3 * SDW && tid / SDW < td.n_steps)
      reinterpret_cast<int*>(sm.sd)[tid] = reinterpret_cast<const int*>(p.steps + td.step_begin)[tid];
    __syncthreads();
    load_states<AUX>(p, sm, sm.sd[0].load_begin, sm.sd[0].n_loads, tid, THREADS);
    if (td.n_steps > 1) load_states<AUX>(p, sm, sm.sd[1].load_begin, sm.sd[1].n_loads, tid, THREADS);
    load_flush_records(p, sm.fbuf, sm.sd[0], tid, THREADS);
    cp_async_wait_all();
    __syncthreads();

    if (warp < NC) {
      // ================================ compute warps ================================
      // Warp w walks row w of every step's visit grid, four positions at a time; the records of the next group (of this
      // or the next step) are prefetched into registers while the current one is processed.
      const int el = lane >> 3, q = lane & 7;
      const LaneConst lc = make_lane_const(lane);
      double* my_stage = sm.stage + warp * (4 * 8 * QS);
      uint32_t* my_rec = sm.rec + warp * 4 * REC_WORDS;
      uint32_t r0 = 0, r1 = 0, r2 = 0;
      int pf_s = -1, pf_g = -1;
      auto fetch = [&](const StepDesc& d, int s_, int g) {
        const int nw = min(4, d.n_pos - g) * REC_WORDS;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.visits + d.visit_begin + (int64_t)warp * d.n_pos + g);
        r0 = r1 = r2 = 0;
        if (lane < nw) r0 = src[lane];
        if (lane + 32 < nw) r1 = src[lane + 32];
        if (lane + 64 < nw) r2 = src[lane + 64];
        pf_s = s_; pf_g = g;
      };
      for (int s = 0; s < td.n_steps; ++s) {
        if (s >= 2) bar_sync(BAR_FLUSH + (s & 1), THREADS);  // flush of step s-2 done, state of step s staged
        const StepDesc sd = sm.sd[s & 3];
        for (int g0 = 0; g0 < sd.n_pos; g0 += 4) {
          const int nvis = min(4, sd.n_pos - g0);
          if (pf_s != s || pf_g != g0) fetch(sd, s, g0);
          my_rec[lane] = r0; my_rec[lane + 32] = r1; my_rec[lane + 64] = r2;
          __syncwarp();
          if (g0 + 4 < sd.n_pos) fetch(sd, s, g0 + 4);
          else if (s + 1 < td.n_steps && sm.sd[(s + 1) & 3].n_pos > 0) fetch(sm.sd[(s + 1) & 3], s + 1, 0);
          if (el < nvis) {
            const uint4 pk = *reinterpret_cast<const uint4*>(my_rec + el * REC_WORDS);
            if ((pk.x & 0xffffu) != 0xffffu) {
              int ln[8] = {int(pk.x & 0xffffu), int(pk.x >> 16), int(pk.y & 0xffffu), int(pk.y >> 16),
                           int(pk.z & 0xffffu), int(pk.z >> 16), int(pk.w & 0xffffu), int(pk.w >> 16)};
              geometry_phase<AUX>(p, sm, fl, ln, q, my_stage + lane * QS);
            }
          }
          __syncwarp();
          for (int e = 0; e < nvis; ++e) {
            if ((my_rec[e * REC_WORDS] & 0xffffu) == 0xffffu) continue;  // empty grid position
            element_phase<AUX>(p, sm, fl, lc, my_stage + e * 8 * QS, my_rec + e * REC_WORDS, lane);
          }
          __syncwarp();
        }
        bar_arrive(BAR_VISITS + (s & 1), THREADS);
      }
    } else if (warp < NC + NF) {
      // ================================ flush warps ================================
      const int fw = warp - NC;
      for (int s = 0; s < td.n_steps; ++s) {
        bar_sync(BAR_VISITS + (s & 1), THREADS);   // every visit of step s has landed; its flush records are staged
        const StepDesc sd = sm.sd[s & 3];
        flush_rows<AUX>(p, sm, fl, sd, sm.fbuf + (s & 1) * kFlushBufBytes, fw, lane);
        if (s + 1 < td.n_steps) bar_arrive(BAR_FDONE + (s & 1), (NF + 1) * 32);
        if (s + 2 < td.n_steps) bar_arrive(BAR_FLUSH + (s & 1), THREADS);
      }
    } else {
      // ================================ staging warp ================================
      // everything later steps need from global memory, fetched in the shadow of the visits: at iteration s the flush
      // records of step s+1, the nodal state of step s+2 and the descriptor of step s+3
      bar_arrive(BAR_VISITS + 0, THREADS);         // step 0 was staged by the prologue
      for (int s = 0; s < td.n_steps; ++s) {
        if (s >= 1) bar_sync(BAR_FDONE + ((s - 1) & 1), (NF + 1) * 32);  // flush of step s-1 done: buffers and slots are free
        if (s + 1 < td.n_steps) load_flush_records(p, sm.fbuf + ((s + 1) & 1) * kFlushBufBytes, sm.sd[(s + 1) & 3], lane, 32);
        if (s + 2 < td.n_steps) {
          const StepDesc nsd = sm.sd[(s + 2) & 3];
          load_states<AUX>(p, sm, nsd.load_begin, nsd.n_loads, lane, 32);
        }
        if (s + 3 < td.n_steps && lane < SDW)
          reinterpret_cast<int*>(&sm.sd[(s + 3) & 3])[lane] = reinterpret_cast<const int*>(p.steps + td.step_begin + s + 3)[lane];
        cp_async_wait_all();
        __syncwarp();
        if (s + 1 < td.n_steps) bar_arrive(BAR_VISITS + ((s + 1) & 1), THREADS);
        if (s + 2 < td.n_steps) bar_arrive(BAR_FLUSH + (s & 1), THREADS);
      }
    }
    __syncthreads();
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
int hex8_sweep_workers() { return NC; }

}  // namespace mastb200
