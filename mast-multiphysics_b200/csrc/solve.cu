// solve.cu -- device-resident linear solve on the assembled CSR: block-Jacobi preconditioned conjugate gradients (sm_100a, fp64).
//
// Device counterpart of the Krylov solve that follows every assembly in the reference: NonlinearSystem::solve ->
// PETSc SNES/KSP, and the explicit linear_solver->solve(*matrix, pc, dsol, rhs, tol, max_it) of sensitivity_solve /
// adjoint_solve (/root/reference/src/base/nonlinear_system.cpp:777-782, 834-839).  With J already in HBM in the PETSc row
// layout, solving there removes the 15.6 GB device-to-host copy of the CSR values (C3) from every Newton step.
//
// Matrix layout (mast_b200.h): the nv = 3 rows of owned node r are contiguous, row i holds nnb*3 values in neighbour
// order: J[9 nb0 + i 3 nnb + 3 c + j], nb0 = nbr_ptr[r], c = position of the neighbour in nbr[nb0 .. nb0 + nnb).
//   spmv        one warp per owned node: the 9 nnb values of its three rows are one contiguous run, read with fully
//               coalesced 8-byte loads; x of a neighbour block comes from L2 (27 / 125 neighbours, mostly consecutive).
//               The warp also leaves p_A . (A p)_A for the step length: block partials, no atomics.
//   precond     3x3 diagonal block of every owned node inverted once per solve (the block-Jacobi of PETSc's
//               -pc_type bjacobi with one node per block; rows of constrained dofs are identity rows there).
//   vector ops  fused: x += a p, r -= a Ap, z = M^-1 r, partial sums of r.z and r.r in one pass; p = z + b p in another.
// Scalars (alpha, beta, the norms) live on the device; the host reads the residual norm every `check` iterations only.
// All reductions are two-stage with a fixed grid, so a solve is bitwise reproducible.  Multi-rank: the caller's hooks
// refresh the ghost values of p before every product and add the scalars over the ranks (context.cu, NCCL).
#include <cuda_runtime.h>

#include <cstdlib>

#include "kernels.h"

namespace mastb200 {
namespace {

constexpr int RB = 1024;  // blocks of the two-stage reductions

__global__ void __launch_bounds__(128) bj_setup_kernel(int64_t n_owned, const int32_t* __restrict__ owned, const int64_t* __restrict__ nbr_ptr,
                                                       const int32_t* __restrict__ nbr, const double* __restrict__ J,
                                                       double* __restrict__ minv) {
  const int64_t r = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (r >= n_owned) return;
  const int64_t b = nbr_ptr[r], nnb = nbr_ptr[r + 1] - b;
  const int32_t self = owned[r];
  int64_t lo = 0, hi = nnb - 1;
  while (lo < hi) {  // neighbours ascending
    const int64_t m = (lo + hi) >> 1;
    if (nbr[b + m] < self) lo = m + 1; else hi = m;
  }
  double a[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) a[i][j] = J[9 * b + i * 3 * nnb + 3 * lo + j];
  const double c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1], c01 = a[1][2] * a[2][0] - a[1][0] * a[2][2],
               c02 = a[1][0] * a[2][1] - a[1][1] * a[2][0];
  const double det = a[0][0] * c00 + a[0][1] * c01 + a[0][2] * c02;
  const double id = det != 0. ? 1. / det : 0.;
  double* m = minv + r * 9;
  m[0] = c00 * id; m[3] = c01 * id; m[6] = c02 * id;
  m[1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * id;
  m[4] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id;
  m[7] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id;
  m[2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * id;
  m[5] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id;
  m[8] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id;
}

// y_A = sum_c K_Ac x_c for the owned nodes; partial[block] = sum over the block's nodes of x_A . y_A (if partial).
// Lane = position p = 3 c + j of a scalar row (three positions per lane and chunk of 96): the x value of a position is
// loaded once and serves the node's three rows.  The product is latency-bound unless the dependent round trips
// row pointer -> neighbour ids -> x are taken out of the node's own iteration: the row pointers arrive two iterations
// ahead and the neighbour ids of the first chunk one iteration ahead, so the nine streaming 256-byte runs of J, the x
// values and the next node's ids are all issued before the first dependent instruction (the one-value-per-iteration loop
// this replaces ran at 1.95 TB/s).  J is read once per product: evict-first (ld.global.cs), so that the stream does not
// push x out of L2.  The grid is one resident wave (launch_spmv_node).
__global__ void __launch_bounds__(256) spmv_node_kernel(int64_t n_owned, const int32_t* __restrict__ owned,
                                                        const int64_t* __restrict__ nbr_ptr, const int32_t* __restrict__ nbr,
                                                        const double* __restrict__ J, const double* __restrict__ x,
                                                        double* __restrict__ y, double* __restrict__ partial) {
  __shared__ double red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double blocksum = 0.;
  const int64_t stride = (int64_t)gridDim.x * 8;
  int64_t r = (int64_t)blockIdx.x * 8 + warp;
  int jj[3], cc[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { cc[k] = (32 * k + lane) / 3; jj[k] = 32 * k + lane - 3 * cc[k]; }
  // pipeline: (b, e) of this node, (b1, e1) of the next one, ids nb[] of this node's first chunk
  int64_t b = 0, e = 0, b1 = 0, e1 = 0;
  int32_t nb[3] = {-1, -1, -1};
  if (r < n_owned) { b = nbr_ptr[r]; e = nbr_ptr[r + 1]; }
  if (r + stride < n_owned) { b1 = nbr_ptr[r + stride]; e1 = nbr_ptr[r + stride + 1]; }
#pragma unroll
  for (int k = 0; k < 3; ++k) nb[k] = 32 * k + lane < 3 * (int)(e - b) ? __ldg(nbr + b + cc[k]) : -1;
  for (; r < n_owned; r += stride) {
    const int rl = 3 * (int)(e - b);
    const double* v = J + 9 * b;
    double j0[3], j1[3], j2[3], xv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int p = 32 * k + lane;
      const bool ok = p < rl;
      j0[k] = ok ? __ldcs(v + p) : 0.;
      j1[k] = ok ? __ldcs(v + rl + p) : 0.;
      j2[k] = ok ? __ldcs(v + 2 * rl + p) : 0.;
      xv[k] = nb[k] >= 0 ? x[(int64_t)nb[k] * 3 + jj[k]] : 0.;
    }
    // next node: ids of its first chunk, row pointers of the one after
    const int64_t bc = b;
    b = b1; e = e1;
#pragma unroll
    for (int k = 0; k < 3; ++k) nb[k] = 32 * k + lane < 3 * (int)(e - b) ? __ldg(nbr + b + cc[k]) : -1;
    if (r + 2 * stride < n_owned) { b1 = nbr_ptr[r + 2 * stride]; e1 = nbr_ptr[r + 2 * stride + 1]; } else { b1 = e1 = 0; }
    double a0 = 0., a1 = 0., a2 = 0.;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      a0 += j0[k] * xv[k];
      a1 += j1[k] * xv[k];
      a2 += j2[k] * xv[k];
    }
    for (int p0 = 96; p0 < rl; p0 += 96) {  // rows with more than 32 neighbours (HEX27): the remaining chunks
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int p = p0 + 32 * k + lane;
        const bool ok = p < rl;
        const int c = p / 3;
        const int32_t id = ok ? __ldg(nbr + bc + c) : -1;
        j0[k] = ok ? __ldcs(v + p) : 0.;
        j1[k] = ok ? __ldcs(v + rl + p) : 0.;
        j2[k] = ok ? __ldcs(v + 2 * rl + p) : 0.;
        xv[k] = ok ? x[(int64_t)id * 3 + (p - 3 * c)] : 0.;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        a0 += j0[k] * xv[k];
        a1 += j1[k] * xv[k];
        a2 += j2[k] * xv[k];
      }
    }
    // three sums over the warp in six 64-bit shuffles instead of fifteen: the first stage leaves (a0, a1) in the lower and
    // (a2, 0) in the upper half-warp, the second one value per lane; lanes 0, 8, 16 end with the three components
    {
      const bool h4 = (lane & 16) != 0, h3 = (lane & 8) != 0;
      const double r0 = __shfl_xor_sync(0xffffffffu, h4 ? a0 : a2, 16);
      const double r1 = __shfl_xor_sync(0xffffffffu, h4 ? a1 : 0., 16);
      const double v0 = (h4 ? a2 : a0) + r0, v1 = (h4 ? 0. : a1) + r1;
      double w = (h3 ? v1 : v0) + __shfl_xor_sync(0xffffffffu, h3 ? v0 : v1, 8);
      w += __shfl_xor_sync(0xffffffffu, w, 4);
      w += __shfl_xor_sync(0xffffffffu, w, 2);
      w += __shfl_xor_sync(0xffffffffu, w, 1);
      if ((lane & 7) == 0 && lane < 24) {
        const int comp = lane >> 3;
        const int64_t n = owned[r];
        y[n * 3 + comp] = w;
        blocksum += x[n * 3 + comp] * w;
      }
    }
  }
  // lanes 0, 8, 16 hold the warp's x . y partials
  blocksum += __shfl_xor_sync(0xffffffffu, blocksum, 8);
  blocksum += __shfl_xor_sync(0xffffffffu, blocksum, 16);
  if (partial) {
    if (lane == 0) red[warp] = blocksum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.;
      for (int w = 0; w < 8; ++w) s += red[w];
      partial[blockIdx.x] = s;
    }
  }
}

// out[k] = sum of partial[k * stride .. k * stride + n) for k < n_out (one block, fixed order)
__global__ void __launch_bounds__(256) reduce_multi_kernel(const double* __restrict__ partial, int n, int n_out, int stride, double* out) {
  __shared__ double red[256];
  for (int k = 0; k < n_out; ++k) {
    double s = 0.;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[k * stride + i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
      if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[k] = red[0];
    __syncthreads();
  }
}

// scal[]: 0 rz, 1 pAp, 2 alpha, 3 rz_new, 4 rr, 5 beta, 6 rhs norm^2
__global__ void alpha_kernel(double* scal) { scal[2] = scal[1] != 0. ? scal[0] / scal[1] : 0.; }
__global__ void beta_kernel(double* scal) {
  scal[5] = scal[0] != 0. ? scal[3] / scal[0] : 0.;
  scal[0] = scal[3];
}

// r = rhs (x = 0), z = M^-1 r, p = z over the owned nodes; partial sums of r.z and r.r
__global__ void __launch_bounds__(256) pcg_init_kernel(int64_t n_owned, const int32_t* __restrict__ owned, const double* __restrict__ rhs,
                                                       const double* __restrict__ minv, double* x, double* r, double* z, double* p,
                                                       double* partial) {
  __shared__ double ra[256], rb[256];
  double s0 = 0., s1 = 0.;
  for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < n_owned; k += (int64_t)gridDim.x * 256) {
    const int64_t n = owned[k];
    const double r0 = rhs[n * 3], r1 = rhs[n * 3 + 1], r2 = rhs[n * 3 + 2];
    const double* m = minv + k * 9;
    const double z0 = m[0] * r0 + m[1] * r1 + m[2] * r2, z1 = m[3] * r0 + m[4] * r1 + m[5] * r2, z2 = m[6] * r0 + m[7] * r1 + m[8] * r2;
    x[n * 3] = 0.; x[n * 3 + 1] = 0.; x[n * 3 + 2] = 0.;
    r[n * 3] = r0; r[n * 3 + 1] = r1; r[n * 3 + 2] = r2;
    z[n * 3] = z0; z[n * 3 + 1] = z1; z[n * 3 + 2] = z2;
    p[n * 3] = z0; p[n * 3 + 1] = z1; p[n * 3 + 2] = z2;
    s0 += r0 * z0 + r1 * z1 + r2 * z2;
    s1 += r0 * r0 + r1 * r1 + r2 * r2;
  }
  ra[threadIdx.x] = s0; rb[threadIdx.x] = s1;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) { ra[threadIdx.x] += ra[threadIdx.x + w]; rb[threadIdx.x] += rb[threadIdx.x + w]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { partial[blockIdx.x] = ra[0]; partial[RB + blockIdx.x] = rb[0]; }
}

// x += alpha p, r -= alpha Ap, z = M^-1 r; partial sums of r.z and r.r
__global__ void __launch_bounds__(256) pcg_update_kernel(int64_t n_owned, const int32_t* __restrict__ owned, const double* __restrict__ scal,
                                                         const double* __restrict__ minv, const double* __restrict__ p,
                                                         const double* __restrict__ Ap, double* x, double* r, double* z,
                                                         double* partial) {
  __shared__ double ra[256], rb[256];
  const double alpha = scal[2];
  double s0 = 0., s1 = 0.;
  for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < n_owned; k += (int64_t)gridDim.x * 256) {
    const int64_t n = owned[k];
    double rr[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      x[n * 3 + i] += alpha * p[n * 3 + i];
      rr[i] = r[n * 3 + i] - alpha * Ap[n * 3 + i];
      r[n * 3 + i] = rr[i];
    }
    const double* m = minv + k * 9;
    const double z0 = m[0] * rr[0] + m[1] * rr[1] + m[2] * rr[2], z1 = m[3] * rr[0] + m[4] * rr[1] + m[5] * rr[2],
                 z2 = m[6] * rr[0] + m[7] * rr[1] + m[8] * rr[2];
    z[n * 3] = z0; z[n * 3 + 1] = z1; z[n * 3 + 2] = z2;
    s0 += rr[0] * z0 + rr[1] * z1 + rr[2] * z2;
    s1 += rr[0] * rr[0] + rr[1] * rr[1] + rr[2] * rr[2];
  }
  ra[threadIdx.x] = s0; rb[threadIdx.x] = s1;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) { ra[threadIdx.x] += ra[threadIdx.x + w]; rb[threadIdx.x] += rb[threadIdx.x + w]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { partial[blockIdx.x] = ra[0]; partial[RB + blockIdx.x] = rb[0]; }
}

__global__ void __launch_bounds__(256) pcg_direction_kernel(int64_t n_owned, const int32_t* __restrict__ owned, const double* __restrict__ scal,
                                                            const double* __restrict__ z, double* p) {
  const double beta = scal[5];
  for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < n_owned; k += (int64_t)gridDim.x * 256) {
    const int64_t n = owned[k];
#pragma unroll
    for (int i = 0; i < 3; ++i) p[n * 3 + i] = z[n * 3 + i] + beta * p[n * 3 + i];
  }
}

// rigid-body modes of StructuralNearNullVectorSpace (structural_near_null_vector_space.cpp:32-110): three translations,
// three rotations by th = 1e-5 about the origin (with the rotation variable itself = th in the six-variable system)
__global__ void __launch_bounds__(256) near_null_kernel(int64_t n_nodes, int nv, const double* __restrict__ xyz, double* modes) {
  const int64_t n = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (n >= n_nodes) return;
  const double px = xyz[n * 3], py = xyz[n * 3 + 1], pz = xyz[n * 3 + 2];
  const double th = 1.e-5, c = cos(th), s = sin(th);
  const int64_t nd = n_nodes * nv;
  for (int m = 0; m < 6; ++m) {
    double o[6] = {0., 0., 0., 0., 0., 0.};
    if (m < 3) o[m] = 1.;
    else if (m == 3) { o[3] = th; o[1] = (c * py - s * pz) - py; o[2] = (s * py + c * pz) - pz; }
    else if (m == 4) { o[4] = th; o[0] = (c * px + s * pz) - px; o[2] = (-s * px + c * pz) - pz; }
    else { o[5] = th; o[0] = (c * px - s * py) - px; o[1] = (s * px + c * py) - py; }
    for (int v = 0; v < nv; ++v) modes[m * nd + n * nv + v] = o[v];
  }
}

}  // namespace

int pcg_reduction_blocks() { return RB; }

cudaError_t launch_bj_setup(int64_t n_owned, const int32_t* owned, const int64_t* nbr_ptr, const int32_t* nbr, const double* J,
                            double* minv, cudaStream_t s) {
  bj_setup_kernel<<<(unsigned)((n_owned + 127) / 128), 128, 0, s>>>(n_owned, owned, nbr_ptr, nbr, J, minv);
  return cudaGetLastError();
}
cudaError_t launch_spmv_node(int64_t n_owned, const int32_t* owned, const int64_t* nbr_ptr, const int32_t* nbr, const double* J,
                             const double* x, double* y, double* partial, int sm_count, cudaStream_t s) {
  // partial (if any) has RB entries: the grid is RB blocks or fewer (unused entries are zeroed by the caller once)
  const int64_t want = (n_owned + 7) / 8;
  static int per_sm = 0;
  if (!per_sm) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spmv_node_kernel, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
  }
  int64_t wave = (int64_t)(sm_count > 0 ? sm_count : 148) * per_sm;  // one resident wave: no second, partly filled one
  if (wave > RB) wave = RB;
  const int grid = (int)(want < wave ? want : wave);
  spmv_node_kernel<<<grid, 256, 0, s>>>(n_owned, owned, nbr_ptr, nbr, J, x, y, partial);
  return cudaGetLastError();
}
cudaError_t launch_reduce_multi(const double* partial, int n, int n_out, int stride, double* out, cudaStream_t s) {
  reduce_multi_kernel<<<1, 256, 0, s>>>(partial, n, n_out, stride, out);
  return cudaGetLastError();
}
cudaError_t launch_pcg_alpha(double* scal, cudaStream_t s) { alpha_kernel<<<1, 1, 0, s>>>(scal); return cudaGetLastError(); }
cudaError_t launch_pcg_beta(double* scal, cudaStream_t s) { beta_kernel<<<1, 1, 0, s>>>(scal); return cudaGetLastError(); }
cudaError_t launch_pcg_init(int64_t n_owned, const int32_t* owned, const double* rhs, const double* minv, double* x, double* r,
                            double* z, double* p, double* partial, cudaStream_t s) {
  pcg_init_kernel<<<RB, 256, 0, s>>>(n_owned, owned, rhs, minv, x, r, z, p, partial);
  return cudaGetLastError();
}
cudaError_t launch_pcg_update(int64_t n_owned, const int32_t* owned, const double* scal, const double* minv, const double* p,
                              const double* Ap, double* x, double* r, double* z, double* partial, cudaStream_t s) {
  pcg_update_kernel<<<RB, 256, 0, s>>>(n_owned, owned, scal, minv, p, Ap, x, r, z, partial);
  return cudaGetLastError();
}
cudaError_t launch_pcg_direction(int64_t n_owned, const int32_t* owned, const double* scal, const double* z, double* p, cudaStream_t s) {
  pcg_direction_kernel<<<RB, 256, 0, s>>>(n_owned, owned, scal, z, p);
  return cudaGetLastError();
}
cudaError_t launch_near_null(int64_t n_nodes, int nv, const double* xyz, double* modes, cudaStream_t s) {
  near_null_kernel<<<(unsigned)((n_nodes + 255) / 256), 256, 0, s>>>(n_nodes, nv, xyz, modes);
  return cudaGetLastError();
}

}  // namespace mastb200
