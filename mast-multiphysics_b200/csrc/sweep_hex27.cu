// sweep_hex27.cu -- HEX27 (tri-quadratic) linear-elastic residual + Jacobian sweep (sm_100a, fp64, DMMA).
//
// Device counterpart of the element loop of MAST::NonlinearImplicitAssembly::residual_and_jacobian
// (/root/reference/src/base/nonlinear_implicit_assembly.cpp:111-186) with StructuralElement3D::internal_residual
// (src/elasticity/solid_element_3d.cpp:135-423, linear strain; 3x3x3 Gauss rule = libMesh's default for SECOND order,
// src/mesh/fe_base.cpp:103-105) and ::thermal_residual (:950-1089).
//
// The path is fp64-compute-bound (K_e is 81x81 from 27 quadrature points), so the contraction runs on the fp64 tensor
// pipe.  With an isotropic D the element matrix is K_ab = lam*S_ab + mu*S_ab^T + mu*tr(S_ab)*I with
//   S_ij[a][b] = sum_q w_q (dN_a/dx_i)(q) (dN_b/dx_j)(q),   w_q = JxW_q * E(x_q),
// i.e. nine 27x27x27 products (padded to 32x32x28) instead of B^T D B with K = 162: 1008 mma.sync.m8n8k4.f64 per
// element (516 kflop) against 2.1 Mflop for the dense form.
//
// One CTA (4 warps) per element, five CTAs per SM; elements are processed colour by colour (no two elements of a colour
// share a node), so the adds into the CSR values and the residual are plain read-modify-writes without atomics and the
// summation order into every entry is the colour order -- fixed by the plan, independent of scheduling.  K_e is never
// stored: the accumulator fragments go from registers straight to their CSR slots through the per-element position map
// (a precomputed element -> nnz map, one byte per node pair).
// Opt-in (MASTB200_H27_ORDERED=1): ONE launch in the plan's order (spatial bricks, colour-major inside a brick:
// plan.cpp build_ordered_scatter) where an element starts when all node-sharing elements before it have finished (done
// flags, acquire / release at gpu scope), the first element to touch a CSR block stores and the later ones add (position
// map, bit 7: no memset of J).  Parity-green and bitwise reproducible, but slower on B200: with 740 elements in flight the
// L2 holds only ~4 element-times of CSR rows while an 8-colour dependency chain needs 8 (profiles/r02c_hex27_ordered.md).
//   phase 1  one thread per quadrature point: J, J^-1, w_q (SIMP modulus of the interpolated nodal density), thermal
//            stress factor
//   phase 2  gradients g_i[q][a] = J^-T dN_a, staged in shared memory in three arrays with a q-stride of 36 doubles so
//            that both DMMA operand fetches are bank-conflict-free
//   phase 3  warp w owns node rows 8w..8w+7: two passes over column tiles, 9 component pairs x 2 tiles x 7 k-steps of
//            DMMA each, B operand = w_q * g formed on the fly
//   phase 4  epilogue per pass: K from S, f_a += K_ab u_b, libMesh's per-element Dirichlet semantics, scatter.
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"

namespace mastb200 {
namespace {

constexpr int TB = 128;
#ifndef MASTB200_H27_CTAS
#define MASTB200_H27_CTAS 5
#endif
constexpr int QSTR = 36;              // q-stride of the gradient arrays (doubles): 36 = 4 mod 16 -> conflict-free fragments
constexpr int GSZ = 28 * QSTR;        // one component array: 28 (27 + zero pad) quadrature points x 36 columns

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

struct Smem27 {
  double tabW[28];            // Gauss weights (the shape tables N_b(q), dN_b/dxi_r(q) stay in global memory / L1)
  double g[3][GSZ];           // physical gradients [i][q*36 + a]
  double wq[28];              // JxW * E at the point (0 for the pad point)
  double pth[28];             // thermal stress factor at the point
  double inv[27 * 9];         // J^-1 per point: inv[q][c*3+r] = d xi_r / d x_c
  double xs[27 * 3], us[27 * 3], rho[27], T[27], fth[27 * 3];
  double red[4];
  int64_t base[27];           // nbr_ptr[row] of each node (node-level CSR offset)
  int32_t gid[27], row[27], nnb[27];
  uint8_t cm[32];
  uint8_t pos[27 * 27 + 7];   // position of node b in the neighbour list of node a
};

// read of a value another CTA of the SAME launch may have added to: past L1 in the ordered sweep
__device__ __forceinline__ double ld_acc(const double* a, bool ordered) { return ordered ? __ldcg(a) : *a; }

__global__ void __launch_bounds__(TB, MASTB200_H27_CTAS) hex27_sweep_kernel(const Hex27Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem27& s = *reinterpret_cast<Smem27*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nv = p.nv;
  const bool wantJ = p.J != nullptr || p.tmp_K != nullptr, wantR = p.R != nullptr || p.tmp_f != nullptr;
  const bool ordered = p.pred_ptr != nullptr;
  const unsigned pos_mask = p.pos_flagged ? 0x7fu : 0xffu;  // bit 7 of a position-map entry: first writer of the block

  // tables and the zero padding of the operand arrays: once per CTA
  const double* __restrict__ tabN = p.tab;             // [q][b]
  const double* __restrict__ tabD = p.tab + 27 * 27;   // [q][b][r]
  if (tid < 27) s.tabW[tid] = p.tab[27 * 27 * 4 + tid];
  for (int k = tid; k < 3 * GSZ; k += TB) (&s.g[0][0])[k] = 0.;
  if (tid < 28) { s.wq[tid] = 0.; s.pth[tid] = 0.; }

  for (int idx = blockIdx.x; idx < p.n_elems; idx += gridDim.x) {
    const int e = p.elems ? p.elems[idx] : p.elem_first + idx;
    __syncthreads();  // previous element done with the staging arrays
    if (p.pred_ptr) {  // ordered sweep: wait for the node-sharing elements before this one in the order
      const int pb = p.pred_ptr[idx], pn = p.pred_ptr[idx + 1] - pb;
      if (tid < pn) {
        const int32_t* flag = p.done + p.pred[pb + tid];
        int v;
        do {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
          if (v != p.epoch) __nanosleep(64);
        } while (v != p.epoch);
      }
      __syncthreads();
    }
    if (tid < 27) {
      const int n = p.conn[(int64_t)e * 27 + tid];
      s.gid[tid] = n;
      const int r = p.node_row[n];
      s.row[tid] = r;
      s.base[tid] = r >= 0 ? p.nbr_ptr[r] : 0;
      s.nnb[tid] = r >= 0 ? (int)(p.nbr_ptr[r + 1] - p.nbr_ptr[r]) : 0;
      s.cm[tid] = p.cmask ? p.cmask[n] : 0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        s.xs[tid * 3 + c] = p.xyz[(int64_t)n * 3 + c];
        s.us[tid * 3 + c] = p.X[(int64_t)n * nv + c];
      }
      s.rho[tid] = p.rho ? p.rho[n] : 1.;
      s.T[tid] = p.temperature ? p.temperature[n] : p.T_const;
    }
    for (int k = tid; k < 27 * 27; k += TB) s.pos[k] = p.posmap[(int64_t)e * 729 + k];
    __syncthreads();

    if (p.diag_elem && p.diag_elem[e]) {  // diagonal_elem_subdomain_id (nonlinear_implicit_assembly.cpp:117-130)
      if (wantJ && tid < 27 && s.row[tid] >= 0) {
        const int64_t rl = (int64_t)s.nnb[tid] * nv;
        double* blk = p.J + s.base[tid] * nv * nv + (int64_t)(s.pos[tid * 27 + tid] & pos_mask) * nv;
        for (int i = 0; i < nv; ++i) blk[i * rl + i] = ld_acc(blk + i * rl + i, ordered) + (((s.cm[tid] >> i) & 1u) ? 1. : 1.0e-24);
      }
      if (p.pred_ptr) {
        __threadfence();
        __syncthreads();
        if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.done + e), "r"(p.epoch) : "memory");
      }
      continue;
    }

    // ---- phase 1: geometry at the quadrature points ----
    if (tid < 27) {
      const int q = tid;
      double Jm[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};  // Jm[r][c] = d x_c / d xi_r
      double rq = 0., tq = 0.;
      for (int b = 0; b < 27; ++b) {
        const double* d = tabD + (q * 27 + b) * 3;
        const double x0 = s.xs[b * 3], x1 = s.xs[b * 3 + 1], x2 = s.xs[b * 3 + 2];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          Jm[r][0] = fma(d[r], x0, Jm[r][0]); Jm[r][1] = fma(d[r], x1, Jm[r][1]); Jm[r][2] = fma(d[r], x2, Jm[r][2]);
        }
        const double N = __ldg(tabN + q * 27 + b);
        rq = fma(N, s.rho[b], rq);
        tq = fma(N, s.T[b], tq);
      }
      const double c00 = Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1];
      const double c01 = Jm[1][2] * Jm[2][0] - Jm[1][0] * Jm[2][2];
      const double c02 = Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0];
      const double det = Jm[0][0] * c00 + Jm[0][1] * c01 + Jm[0][2] * c02;
      const double idet = 1. / det;
      double* inv = s.inv + q * 9;
      inv[0] = c00 * idet; inv[3] = c01 * idet; inv[6] = c02 * idet;
      inv[1] = (Jm[0][2] * Jm[2][1] - Jm[0][1] * Jm[2][2]) * idet;
      inv[4] = (Jm[0][0] * Jm[2][2] - Jm[0][2] * Jm[2][0]) * idet;
      inv[7] = (Jm[0][1] * Jm[2][0] - Jm[0][0] * Jm[2][1]) * idet;
      inv[2] = (Jm[0][1] * Jm[1][2] - Jm[0][2] * Jm[1][1]) * idet;
      inv[5] = (Jm[0][2] * Jm[1][0] - Jm[0][0] * Jm[1][2]) * idet;
      inv[8] = (Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0]) * idet;
      const double JxW = det * s.tabW[q];
      double E = p.E0;
      if (p.rho) E = p.E0 * (p.rho_min + (1. - p.rho_min) * simp_pow(rq, p.penalty));  // example_6.cpp:100-137
      s.wq[q] = JxW * E;
      s.pth[q] = p.thermal ? -(JxW * E) * p.alpha_k * (tq - p.T_ref) : 0.;  // f -= JxW B^T (D alpha dT)
    }
    __syncthreads();

    // ---- phase 2: physical gradients g_i[q][a] ----
    for (int k = tid; k < 27 * 27; k += TB) {
      const int q = k / 27, a = k - q * 27;
      const double* d = tabD + k * 3;
      const double* inv = s.inv + q * 9;
#pragma unroll
      for (int c = 0; c < 3; ++c) s.g[c][q * QSTR + a] = d[0] * inv[c * 3] + d[1] * inv[c * 3 + 1] + d[2] * inv[c * 3 + 2];
    }
    __syncthreads();
    if (p.thermal && wantR && tid < 81) {
      const int a = tid / 3, i = tid - a * 3;
      double t = 0.;
      for (int q = 0; q < 27; ++q) t = fma(s.pth[q], s.g[i][q * QSTR + a], t);
      s.fth[tid] = t;
    }

    // ---- phases 3 + 4 ----
    const int al = lane >> 2, kq = lane & 3;
    const int a = 8 * warp + al;
    const bool a_ok = a < 27;
    const int rowa = a_ok ? s.row[a] : -1;
    const unsigned cma = a_ok ? s.cm[a] : 0u;
    const int64_t rl = a_ok ? (int64_t)s.nnb[a] * nv : 0;
    double* rowJ = (wantJ && rowa >= 0) ? p.J + s.base[a] * nv * nv : nullptr;
    double* tmpK = (p.tmp_K && rowa >= 0) ? p.tmp_K + ((int64_t)idx * 27 + a) * (27 * 9) : nullptr;  // two-pass: element-major
    double f0 = 0., f1 = 0., f2 = 0., mx = 0.;
    for (int pass = 0; pass < 2; ++pass) {
      double C[9][2][2];
#pragma unroll
      for (int k = 0; k < 9; ++k) { C[k][0][0] = C[k][0][1] = C[k][1][0] = C[k][1][1] = 0.; }
#pragma unroll 1
      for (int ks = 0; ks < 7; ++ks) {
        const int q = 4 * ks + kq;
        const double w = s.wq[q];
        const double A0 = s.g[0][q * QSTR + a], A1 = s.g[1][q * QSTR + a], A2 = s.g[2][q * QSTR + a];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int bc = 8 * (2 * pass + t) + al;
          const double B0 = w * s.g[0][q * QSTR + bc], B1 = w * s.g[1][q * QSTR + bc], B2 = w * s.g[2][q * QSTR + bc];
          dmma8x8x4(C[0][t][0], C[0][t][1], A0, B0);
          dmma8x8x4(C[1][t][0], C[1][t][1], A0, B1);
          dmma8x8x4(C[2][t][0], C[2][t][1], A0, B2);
          dmma8x8x4(C[3][t][0], C[3][t][1], A1, B0);
          dmma8x8x4(C[4][t][0], C[4][t][1], A1, B1);
          dmma8x8x4(C[5][t][0], C[5][t][1], A1, B2);
          dmma8x8x4(C[6][t][0], C[6][t][1], A2, B0);
          dmma8x8x4(C[7][t][0], C[7][t][1], A2, B1);
          dmma8x8x4(C[8][t][0], C[8][t][1], A2, B2);
        }
      }
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int bb = 0; bb < 2; ++bb) {
          const int b = 8 * (2 * pass + t) + 2 * kq + bb;
          if (!a_ok || b >= 27) continue;
          double K[3][3];
          const double trm = p.mu1 * (C[0][t][bb] + C[4][t][bb] + C[8][t][bb]);
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
              K[i][j] = p.lam1 * C[i * 3 + j][t][bb] + p.mu1 * C[j * 3 + i][t][bb] + (i == j ? trm : 0.);
          if (wantR) {  // internal force of linear elasticity: f_a += K_ab u_b (unconstrained element matrix)
            const double u0 = s.us[b * 3], u1 = s.us[b * 3 + 1], u2 = s.us[b * 3 + 2];
            f0 += K[0][0] * u0 + K[0][1] * u1 + K[0][2] * u2;
            f1 += K[1][0] * u0 + K[1][1] * u1 + K[1][2] * u2;
            f2 += K[2][0] * u0 + K[2][1] * u1 + K[2][2] * u2;
          }
          if (a == b) mx = fmax(mx, fmax(K[0][0], fmax(K[1][1], K[2][2])));
          if (tmpK) {  // two-pass assembly: the block is stored, the row gather sums it (same per-element Dirichlet semantics)
            const unsigned cmb = s.cm[b];
            double* dst = tmpK + b * 9;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                double k = K[i][j];
                if (((cma >> i) | (cmb >> j)) & 1u) k = (a == b && i == j && ((cma >> i) & 1u)) ? 1. : 0.;
                dst[i * 3 + j] = k;
              }
          } else if (rowJ) {
            const unsigned cmb = s.cm[b];
            const unsigned pm = s.pos[a * 27 + b];
            const bool first = p.first_writer && (pm & 0x80u);  // no element before this one has touched the block
            double* dst = rowJ + (int64_t)(pm & pos_mask) * nv;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                double k = K[i][j];
                if (((cma >> i) | (cmb >> j)) & 1u) k = (a == b && i == j && ((cma >> i) & 1u)) ? 1. : 0.;
                dst[i * rl + j] = first ? k : ld_acc(dst + i * rl + j, ordered) + k;
              }
          }
        }
    }
    if (nv == 6 && wantJ) {  // 1e-20*max(diag K_e) on the rotational diagonal (solid_element_3d.cpp:413-417)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (lane == 0) s.red[warp] = mx;
    }
    __syncthreads();  // red and fth complete
    if (wantR) {
      f0 += __shfl_xor_sync(0xffffffffu, f0, 1); f0 += __shfl_xor_sync(0xffffffffu, f0, 2);
      f1 += __shfl_xor_sync(0xffffffffu, f1, 1); f1 += __shfl_xor_sync(0xffffffffu, f1, 2);
      f2 += __shfl_xor_sync(0xffffffffu, f2, 1); f2 += __shfl_xor_sync(0xffffffffu, f2, 2);
      if (kq == 0 && rowa >= 0) {
        if (p.thermal) { f0 += s.fth[a * 3]; f1 += s.fth[a * 3 + 1]; f2 += s.fth[a * 3 + 2]; }
        if (p.tmp_f) {
          double* r = p.tmp_f + ((int64_t)idx * 27 + a) * 3;
          r[0] = (cma & 1u) ? 0. : f0; r[1] = (cma & 2u) ? 0. : f1; r[2] = (cma & 4u) ? 0. : f2;
        } else {
          double* r = p.R + (int64_t)s.gid[a] * nv;
          if (!(cma & 1u)) r[0] = ld_acc(r, ordered) + f0;
          if (!(cma & 2u)) r[1] = ld_acc(r + 1, ordered) + f1;
          if (!(cma & 4u)) r[2] = ld_acc(r + 2, ordered) + f2;
        }
      }
    }
    if (nv == 6 && wantJ && tid < 27 && s.row[tid] >= 0) {
      const double m = fmax(fmax(s.red[0], s.red[1]), fmax(s.red[2], s.red[3]));
      const int64_t rl6 = (int64_t)s.nnb[tid] * 6;
      double* blk = p.J + s.base[tid] * 36 + (int64_t)(s.pos[tid * 27 + tid] & pos_mask) * 6;
      for (int i = 3; i < 6; ++i) blk[i * rl6 + i] = ld_acc(blk + i * rl6 + i, ordered) + (((s.cm[tid] >> i) & 1u) ? 1. : 1.0e-20 * m);
    }
    if (p.pred_ptr) {  // this element's adds are visible before its flag
      __threadfence();
      __syncthreads();
      if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.done + e), "r"(p.epoch) : "memory");
    }
  }
}

// ---- pass 1 of the two-pass assembly with the symmetric half of the contraction ---------------------------------------
// K_ba = K_ab^T element-wise, so only the ten tile pairs (row tile <= column tile) of the 4 x 4 tile grid are contracted: 630
// DMMA per element instead of 1008, five warps with two tile pairs each; a lane stores its block (a, b) and, off the tile
// diagonal, the mirrored block (b, a).  The element force is not K u any more but what the reference computes,
// f_a = sum_q JxW B_a^T sigma_q (solid_element_3d.cpp:289-297): stress per point, then one thread per (node, component).
// 3-dof system, no diagonal elements (the host keeps the other kernel for those).
constexpr int TB5 = 160;

// a 3x3 block = 72 contiguous bytes at an 8-byte aligned address: four 16-byte stores and one 8-byte store (three sectors per
// lane instead of nine partial-sector writes)
__device__ __forceinline__ void store_block72(double* dst, const double (&v)[9]) {
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    double2* d2 = reinterpret_cast<double2*>(dst);
    d2[0] = make_double2(v[0], v[1]); d2[1] = make_double2(v[2], v[3]); d2[2] = make_double2(v[4], v[5]); d2[3] = make_double2(v[6], v[7]);
    dst[8] = v[8];
  } else {
    dst[0] = v[0];
    double2* d2 = reinterpret_cast<double2*>(dst + 1);
    d2[0] = make_double2(v[1], v[2]); d2[1] = make_double2(v[3], v[4]); d2[2] = make_double2(v[5], v[6]); d2[3] = make_double2(v[7], v[8]);
  }
}

#ifndef MASTB200_H27_SYM_CTAS
#define MASTB200_H27_SYM_CTAS 4
#endif
__global__ void __launch_bounds__(TB5, MASTB200_H27_SYM_CTAS) hex27_sym_pass1_kernel(const Hex27Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem27& s = *reinterpret_cast<Smem27*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool wantR = p.tmp_f != nullptr;
  const double* __restrict__ tabN = p.tab;             // [q][b]
  const double* __restrict__ tabD = p.tab + 27 * 27;   // [q][b][r]
  if (tid < 27) s.tabW[tid] = p.tab[27 * 27 * 4 + tid];
  for (int k = tid; k < 3 * GSZ; k += TB5) (&s.g[0][0])[k] = 0.;
  if (tid < 28) { s.wq[tid] = 0.; s.pth[tid] = 0.; }

  for (int idx = blockIdx.x; idx < p.n_elems; idx += gridDim.x) {
    const int e = p.elem_first + idx;
    __syncthreads();  // previous element done with the staging arrays
    if (tid < 27) {
      const int n = p.conn[(int64_t)e * 27 + tid];
      s.gid[tid] = n;
      s.row[tid] = p.node_row[n];
      s.cm[tid] = p.cmask ? p.cmask[n] : 0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        s.xs[tid * 3 + c] = p.xyz[(int64_t)n * 3 + c];
        s.us[tid * 3 + c] = p.X[(int64_t)n * 3 + c];
      }
      s.rho[tid] = p.rho ? p.rho[n] : 1.;
      s.T[tid] = p.temperature ? p.temperature[n] : p.T_const;
    }
    __syncthreads();
    // ---- phase 1: geometry at the quadrature points; 81 threads form the rows of J, 27 more interpolate rho and T ----
    if (tid < 81) {
      const int q = tid / 3, r = tid - q * 3;
      double j0 = 0., j1 = 0., j2 = 0.;  // row r of Jm[r][c] = d x_c / d xi_r
      for (int b = 0; b < 27; ++b) {
        const double d = tabD[(q * 27 + b) * 3 + r];
        j0 = fma(d, s.xs[b * 3], j0); j1 = fma(d, s.xs[b * 3 + 1], j1); j2 = fma(d, s.xs[b * 3 + 2], j2);
      }
      double* Jr = s.inv + q * 9 + r * 3;
      Jr[0] = j0; Jr[1] = j1; Jr[2] = j2;
    } else if (tid < 108) {
      const int q = tid - 81;
      double rq = 0., tq = 0.;
      for (int b = 0; b < 27; ++b) {
        const double N = __ldg(tabN + q * 27 + b);
        rq = fma(N, s.rho[b], rq);
        tq = fma(N, s.T[b], tq);
      }
      s.fth[q] = rq; s.fth[27 + q] = tq;
    }
    __syncthreads();
    if (tid < 27) {
      const int q = tid;
      double Jm[3][3];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Jm[r][c] = s.inv[q * 9 + r * 3 + c];
      const double rq = s.fth[q], tq = s.fth[27 + q];
      const double c00 = Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1];
      const double c01 = Jm[1][2] * Jm[2][0] - Jm[1][0] * Jm[2][2];
      const double c02 = Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0];
      const double det = Jm[0][0] * c00 + Jm[0][1] * c01 + Jm[0][2] * c02;
      const double idet = 1. / det;
      double* inv = s.inv + q * 9;
      inv[0] = c00 * idet; inv[3] = c01 * idet; inv[6] = c02 * idet;
      inv[1] = (Jm[0][2] * Jm[2][1] - Jm[0][1] * Jm[2][2]) * idet;
      inv[4] = (Jm[0][0] * Jm[2][2] - Jm[0][2] * Jm[2][0]) * idet;
      inv[7] = (Jm[0][1] * Jm[2][0] - Jm[0][0] * Jm[2][1]) * idet;
      inv[2] = (Jm[0][1] * Jm[1][2] - Jm[0][2] * Jm[1][1]) * idet;
      inv[5] = (Jm[0][2] * Jm[1][0] - Jm[0][0] * Jm[1][2]) * idet;
      inv[8] = (Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0]) * idet;
      const double JxW = det * s.tabW[q];
      double E = p.E0;
      if (p.rho) E = p.E0 * (p.rho_min + (1. - p.rho_min) * simp_pow(rq, p.penalty));  // example_6.cpp:100-137
      s.wq[q] = JxW * E;
      s.pth[q] = p.thermal ? -(JxW * E) * p.alpha_k * (tq - p.T_ref) : 0.;  // f -= JxW B^T (D alpha dT)
    }
    __syncthreads();
    // ---- phase 2: physical gradients g_i[q][a] ----
    for (int k = tid; k < 27 * 27; k += TB5) {
      const int q = k / 27, a = k - q * 27;
      const double* d = tabD + k * 3;
      const double* inv = s.inv + q * 9;
#pragma unroll
      for (int c = 0; c < 3; ++c) s.g[c][q * QSTR + a] = d[0] * inv[c * 3] + d[1] * inv[c * 3 + 1] + d[2] * inv[c * 3 + 2];
    }
    __syncthreads();
    // ---- phase 2b: stress per point (w_q E-scaled; the thermal part rides on the diagonal), into the J^-1 storage:
    //      81 threads form the rows of H = grad u, 27 turn them into the stress ----
    if (wantR && tid < 81) {
      const int q = tid / 3, i = tid - q * 3;
      double h0 = 0., h1 = 0., h2 = 0.;  // H[i][j] = d u_i / d x_j
      for (int b = 0; b < 27; ++b) {
        const double u = s.us[b * 3 + i];
        h0 = fma(u, s.g[0][q * QSTR + b], h0); h1 = fma(u, s.g[1][q * QSTR + b], h1); h2 = fma(u, s.g[2][q * QSTR + b], h2);
      }
      double* Hr = s.inv + q * 9 + i * 3;
      Hr[0] = h0; Hr[1] = h1; Hr[2] = h2;
    }
    __syncthreads();
    if (wantR && tid < 27) {
      const int q = tid;
      double H[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) H[i][j] = s.inv[q * 9 + i * 3 + j];
      const double w = s.wq[q], tr = p.lam1 * (H[0][0] + H[1][1] + H[2][2]), m2 = 2. * p.mu1;
      double* sg = s.inv + q * 9;  // xx, yy, zz, xy, yz, zx
      sg[0] = w * (tr + m2 * H[0][0]) + s.pth[q]; sg[1] = w * (tr + m2 * H[1][1]) + s.pth[q]; sg[2] = w * (tr + m2 * H[2][2]) + s.pth[q];
      sg[3] = w * p.mu1 * (H[0][1] + H[1][0]); sg[4] = w * p.mu1 * (H[1][2] + H[2][1]); sg[5] = w * p.mu1 * (H[0][2] + H[2][0]);
    }
    __syncthreads();
    if (wantR && tid < 81) {  // f_a[i] = sum_q sum_j sigma_q[i][j] g_j[q][a]
      const int a = tid / 3, i = tid - a * 3;
      double f = 0.;
      for (int q = 0; q < 27; ++q) {
        const double* sg = s.inv + q * 9;
        const double g0 = s.g[0][q * QSTR + a], g1 = s.g[1][q * QSTR + a], g2 = s.g[2][q * QSTR + a];
        const double r0 = i == 0 ? sg[0] : (i == 1 ? sg[3] : sg[5]);
        const double r1 = i == 0 ? sg[3] : (i == 1 ? sg[1] : sg[4]);
        const double r2 = i == 0 ? sg[5] : (i == 1 ? sg[4] : sg[2]);
        f += r0 * g0 + r1 * g1 + r2 * g2;
      }
      if (s.row[a] >= 0) p.tmp_f[((int64_t)idx * 27 + a) * 3 + i] = ((s.cm[a] >> i) & 1u) ? 0. : f;
    }
    if (!p.tmp_K) continue;

    // ---- phases 3 + 4: two tile pairs per warp ----
    const int al = lane >> 2, kq = lane & 3;
#pragma unroll 1
    for (int pi = 0; pi < 2; ++pi) {
      // pairs in the order (0,0) (0,1) (0,2) (0,3) (1,1) (1,2) (1,3) (2,2) (2,3) (3,3)
      const int pr = 2 * warp + pi;
      const int rt = pr < 4 ? 0 : (pr < 7 ? 1 : (pr < 9 ? 2 : 3));
      const int ct = pr < 4 ? pr : (pr < 7 ? pr - 3 : (pr < 9 ? pr - 5 : 3));
      const int a = 8 * rt + al, bc = 8 * ct + al;
      double C[9][2];
#pragma unroll
      for (int k = 0; k < 9; ++k) { C[k][0] = 0.; C[k][1] = 0.; }
#pragma unroll 1
      for (int ks = 0; ks < 7; ++ks) {
        const int q = 4 * ks + kq;
        const double w = s.wq[q];
        const double A0 = s.g[0][q * QSTR + a], A1 = s.g[1][q * QSTR + a], A2 = s.g[2][q * QSTR + a];
        const double B0 = w * s.g[0][q * QSTR + bc], B1 = w * s.g[1][q * QSTR + bc], B2 = w * s.g[2][q * QSTR + bc];
        dmma8x8x4(C[0][0], C[0][1], A0, B0);
        dmma8x8x4(C[1][0], C[1][1], A0, B1);
        dmma8x8x4(C[2][0], C[2][1], A0, B2);
        dmma8x8x4(C[3][0], C[3][1], A1, B0);
        dmma8x8x4(C[4][0], C[4][1], A1, B1);
        dmma8x8x4(C[5][0], C[5][1], A1, B2);
        dmma8x8x4(C[6][0], C[6][1], A2, B0);
        dmma8x8x4(C[7][0], C[7][1], A2, B1);
        dmma8x8x4(C[8][0], C[8][1], A2, B2);
      }
      if (a >= 27) continue;
      const unsigned cma = s.cm[a];
      const bool own_a = s.row[a] >= 0;
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        const int b = 8 * ct + 2 * kq + bb;
        if (b >= 27) continue;
        double K[3][3];
        const double trm = p.mu1 * (C[0][bb] + C[4][bb] + C[8][bb]);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) K[i][j] = p.lam1 * C[i * 3 + j][bb] + p.mu1 * C[j * 3 + i][bb] + (i == j ? trm : 0.);
        const unsigned cmb = s.cm[b];
        if (cma | cmb) {  // per-element Dirichlet semantics (symmetric: the mirrored block sees the same zeros)
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
              if (((cma >> i) | (cmb >> j)) & 1u) K[i][j] = (a == b && i == j && ((cma >> i) & 1u)) ? 1. : 0.;
        }
        if (own_a) {
          const double v[9] = {K[0][0], K[0][1], K[0][2], K[1][0], K[1][1], K[1][2], K[2][0], K[2][1], K[2][2]};
          store_block72(p.tmp_K + (((int64_t)idx * 27 + a) * 27 + b) * 9, v);
        }
        if (rt != ct && s.row[b] >= 0) {  // the mirrored block K_ba = K_ab^T
          const double v[9] = {K[0][0], K[1][0], K[2][0], K[0][1], K[1][1], K[2][1], K[0][2], K[1][2], K[2][2]};
          store_block72(p.tmp_K + (((int64_t)idx * 27 + b) * 27 + a) * 9, v);
        }
      }
    }
  }
}

// ---- pass 2 of the two-pass assembly: row gather -------------------------------------------------------------------
// One warp per owned row of the chunk.  The node's blocks are summed in shared memory (<= 125 neighbours x 9 values) over
// its elements inside the chunk in ascending element id -- the order of the reference's element loop -- from the
// element-major buffer of pass 1 (243 contiguous doubles per (element, node): coalesced reads), then the three CSR rows
// leave as whole runs (coalesced 8-byte stores, every CSR value written once per chunk that touches the row).
constexpr int GW2 = 4;         // warps per CTA
constexpr int GMAXNB = 128;    // neighbours per row (position map: 7 bits)

__global__ void __launch_bounds__(GW2 * 32) hex27_row_gather_kernel(const Hex27GatherParams p) {
  __shared__ double acc_all[GW2][GMAXNB * 9];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ir = (int64_t)blockIdx.x * GW2 + warp;
  if (ir >= p.n_rows) return;
  double* acc = acc_all[warp];
  const int r = p.rows[ir];
  const int n = p.owned_nodes[r];
  const int64_t nb0 = p.nbr_ptr[r];
  const int nnb = (int)(p.nbr_ptr[r + 1] - nb0);
  const unsigned pmask = p.pos_flagged ? 0x7fu : 0xffu;
  const int64_t k0 = p.ea_ptr[ir], k1 = p.ea_ptr[ir + 1];
  const bool first_chunk = p.row_first[ir] != 0;  // no element of the row lies before the chunk: stored, otherwise added to
  for (int t = lane; t < nnb * 9; t += 32) acc[t] = 0.;
  double f = 0.;  // lanes 0-2: the node's force components
  // the row's (element, local node) codes come from the plan; the loads of element k + 1 (eight doubles per lane and the
  // positions of its 27 blocks) are in flight while element k is added
  auto fetch = [&](int code, double (&v)[8], unsigned& pb) {
    const double* src = p.tmp_K + (int64_t)code * (27 * 9);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int t = lane + 32 * u;
      v[u] = t < 27 * 9 ? __ldcs(src + t) : 0.;  // streamed once
    }
    const int el = (int)__umulhi((unsigned)code, 159072863u);  // code / 27 (exact below 8e8)
    pb = lane < 27 ? (unsigned)(p.posmap[((int64_t)(p.elem_first + el) * 27 + (code - el * 27)) * 27 + lane] & pmask) : 0u;
  };
  for (int64_t kb = k0; kb < k1; kb += 32) {
    const int cnt = (int)(k1 - kb < 32 ? k1 - kb : 32);
    const int mine = lane < cnt ? p.ea[kb + lane] : 0;
    double va[8], vb[8];
    unsigned pa = 0u, pbn = 0u;
    fetch(__shfl_sync(0xffffffffu, mine, 0), va, pa);
    for (int el = 0; el < cnt; ++el) {
      const int code = __shfl_sync(0xffffffffu, mine, el);
      if (el + 1 < cnt) fetch(__shfl_sync(0xffffffffu, mine, el + 1), vb, pbn);
      __syncwarp();
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int t = lane + 32 * u;
        const int b = (t * 7282) >> 16;  // t / 9 for t < 243 (and 27.. for the padding lanes: masked below)
        const unsigned pos = __shfl_sync(0xffffffffu, pa, b < 27 ? b : 0);
        if (t < 27 * 9) acc[pos * 9 + (t - b * 9)] += va[u];
      }
      if (p.R && lane < 3) f += p.tmp_f[(int64_t)code * 3 + lane];
#pragma unroll
      for (int u = 0; u < 8; ++u) va[u] = vb[u];
      pa = pbn;
    }
  }
  __syncwarp();
  if (p.J) {
    double* row = p.J + 9 * nb0;
    const int rl = 3 * nnb, total = 9 * nnb;
    int i = 0, rem = lane;  // output index o = i * rl + rem
    for (int o = lane; o < total; o += 32) {
      while (rem >= rl) { rem -= rl; ++i; }
      const int c = (rem * 43691) >> 17;  // rem / 3 for rem < 384
      const double v = acc[c * 9 + i * 3 + (rem - 3 * c)];
      row[o] = first_chunk ? v : row[o] + v;
      rem += 32;
    }
  }
  if (p.R && lane < 3) {
    double* d = p.R + (int64_t)n * 3 + lane;
    *d = first_chunk ? f : *d + f;
  }
}

}  // namespace

cudaError_t launch_hex27_sym_pass1(const Hex27Params& p, int sm_count, cudaStream_t stream) {
  if (p.n_elems <= 0) return cudaSuccess;
  if (p.nv != 3 || p.diag_elem || p.elems) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(hex27_sym_pass1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem27));
  if (e != cudaSuccess) return e;
  const int grid = p.n_elems < MASTB200_H27_SYM_CTAS * sm_count ? p.n_elems : MASTB200_H27_SYM_CTAS * sm_count;
  hex27_sym_pass1_kernel<<<grid, TB5, sizeof(Smem27), stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_hex27_row_gather(const Hex27GatherParams& p, cudaStream_t stream) {
  if (p.n_rows <= 0) return cudaSuccess;
  hex27_row_gather_kernel<<<(unsigned)((p.n_rows + GW2 - 1) / GW2), GW2 * 32, 0, stream>>>(p);
  return cudaGetLastError();
}

size_t hex27_sweep_smem_bytes() { return sizeof(Smem27); }

cudaError_t launch_hex27_sweep(const Hex27Params& p, int sm_count, cudaStream_t stream) {
  if (p.n_elems <= 0) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(hex27_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem27));
  if (e != cudaSuccess) return e;
  // the ordered sweep waits on other CTAs: every CTA of the grid must be resident at once
  int per_sm = MASTB200_H27_CTAS;
  if (p.pred_ptr) {
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hex27_sweep_kernel, TB, sizeof(Smem27))) != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    per_sm = per_sm < MASTB200_H27_CTAS ? per_sm : MASTB200_H27_CTAS;
  }
  const int grid = p.n_elems < per_sm * sm_count ? p.n_elems : per_sm * sm_count;
  hex27_sweep_kernel<<<grid, TB, sizeof(Smem27), stream>>>(p);
  return cudaGetLastError();
}

}  // namespace mastb200
