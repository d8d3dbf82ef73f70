// loads.cu -- side loads, compliance and small deterministic reductions (sm_100a, fp64).
//
// SURFACE_PRESSURE: StructuralElement3D::surface_pressure_residual
// (/root/reference/src/elasticity/solid_element_3d.cpp:815-876):  f -= sum_qp JxW N^T (p n) on a HEX8 face
// with the 2x2 Gauss rule.  Gathered per loaded node (entries ordered by (element, side), the order in which the
// reference's element loop meets them) so that the sum into each residual entry has a fixed order.
// ComplianceOutput::evaluate / output_derivative_for_elem (src/elasticity/compliance_output.cpp:71-102, 265-296):
// q = -sum_e vec_e . u_e with vec_e = -F_e  =>  q = sum_n F_n . u_n ; dq/dX = F (constrained rows zeroed).
#include "kernels.h"

namespace mastb200 {
namespace {

constexpr int LB = 128;

// force on face node k from a uniform pressure on a bilinear face: sum_qp N_k (p * t1 x t2)
__device__ void face_node_force(const double* __restrict__ xyz, const int32_t* fn, int k, double press, double F[3]) {
  constexpr double gp = 0.57735026918962576451;
  double X[4][3];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int c = 0; c < 3; ++c) X[m][c] = xyz[(int64_t)fn[m] * 3 + c];
  F[0] = F[1] = F[2] = 0.;
  const int si[4] = {-1, 1, 1, -1}, sj[4] = {-1, -1, 1, 1};
#pragma unroll
  for (int jq = 0; jq < 2; ++jq)
#pragma unroll
    for (int iq = 0; iq < 2; ++iq) {
      const double xi = iq ? gp : -gp, eta = jq ? gp : -gp;
      double t1[3] = {0., 0., 0.}, t2[3] = {0., 0., 0.};
      double Nk = 0.;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const double fx = 1. + si[m] * xi, fy = 1. + sj[m] * eta;
        const double dxi = 0.25 * si[m] * fy, deta = 0.25 * sj[m] * fx;
        if (m == k) Nk = 0.25 * fx * fy;
#pragma unroll
        for (int c = 0; c < 3; ++c) { t1[c] += dxi * X[m][c]; t2[c] += deta * X[m][c]; }
      }
      const double n0 = t1[1] * t2[2] - t1[2] * t2[1], n1 = t1[2] * t2[0] - t1[0] * t2[2], n2 = t1[0] * t2[1] - t1[1] * t2[0];
      F[0] += Nk * press * n0; F[1] += Nk * press * n1; F[2] += Nk * press * n2;
    }
}

// the same for a bi-quadratic (QUAD9) face of a HEX27 with the 3x3 Gauss rule; libMesh QUAD9 node order: corners,
// edge midpoints (bottom, right, top, left), centre
__device__ void face_node_force9(const double* __restrict__ xyz, const int32_t* fn, int k, double press, double F[3]) {
  const double gp[3] = {-0.77459666924148337704, 0., 0.77459666924148337704};
  const double gw[3] = {5. / 9., 8. / 9., 5. / 9.};
  const int FI[9] = {0, 1, 1, 0, 2, 1, 2, 0, 2}, FJ[9] = {0, 0, 1, 1, 0, 2, 1, 2, 2};  // 0: -1, 1: +1, 2: mid
  auto l = [](int c, double x) { return c == 0 ? 0.5 * x * (x - 1.) : c == 1 ? 0.5 * x * (x + 1.) : (1. - x * x); };
  auto dl = [](int c, double x) { return c == 0 ? x - 0.5 : c == 1 ? x + 0.5 : -2. * x; };
  double X[9][3];
  for (int m = 0; m < 9; ++m)
    for (int c = 0; c < 3; ++c) X[m][c] = xyz[(int64_t)fn[m] * 3 + c];
  F[0] = F[1] = F[2] = 0.;
  for (int jq = 0; jq < 3; ++jq)
    for (int iq = 0; iq < 3; ++iq) {
      const double xi = gp[iq], eta = gp[jq];
      double t1[3] = {0., 0., 0.}, t2[3] = {0., 0., 0.};
      double Nk = 0.;
      for (int m = 0; m < 9; ++m) {
        const double li = l(FI[m], xi), lj = l(FJ[m], eta);
        const double dxi = dl(FI[m], xi) * lj, deta = li * dl(FJ[m], eta);
        if (m == k) Nk = li * lj;
        for (int c = 0; c < 3; ++c) { t1[c] += dxi * X[m][c]; t2[c] += deta * X[m][c]; }
      }
      const double n0 = t1[1] * t2[2] - t1[2] * t2[1], n1 = t1[2] * t2[0] - t1[0] * t2[2], n2 = t1[0] * t2[1] - t1[1] * t2[0];
      const double w = gw[iq] * gw[jq] * Nk * press;
      F[0] += w * n0; F[1] += w * n1; F[2] += w * n2;
    }
}

__global__ void __launch_bounds__(LB) face_load_kernel(FaceLoadPlan fp, const double* __restrict__ xyz,
                                                       const double* __restrict__ X, const uint8_t* __restrict__ cmask,
                                                       int nv, double* R, double* dq, double* partial) {
  __shared__ double red[LB];
  const int t = blockIdx.x * LB + threadIdx.x;
  double dot = 0.;
  if (t < fp.n_loaded_nodes && fp.node[t] >= fp.node_lo && fp.node[t] < fp.node_hi) {
    const int n = fp.node[t];
    double F[3] = {0., 0., 0.};
    for (int e = fp.ptr[t]; e < fp.ptr[t + 1]; ++e) {
      double f[3];
      if (fp.nfn == 9) face_node_force9(xyz, fp.face_nodes + (int64_t)e * 9, fp.face_k[e], fp.pressure[e], f);
      else face_node_force(xyz, fp.face_nodes + (int64_t)e * 4, fp.face_k[e], fp.pressure[e], f);
      F[0] += f[0]; F[1] += f[1]; F[2] += f[2];
    }
    const unsigned cm = cmask ? cmask[n] : 0u;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const bool con = (cm >> i) & 1u;
      if (R && !con) R[(int64_t)n * nv + i] -= F[i];
      if (dq && !con) dq[(int64_t)n * nv + i] = F[i];
      if (X) dot += F[i] * X[(int64_t)n * nv + i];
    }
  }
  if (partial) {
    red[threadIdx.x] = dot;
    __syncthreads();
    for (int s = LB / 2; s > 0; s >>= 1) {
      if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
  }
}

__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* __restrict__ partial, int n, double* out) {
  __shared__ double red[256];
  double s = 0.;
  for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

__global__ void __launch_bounds__(256) sumsq_owned_kernel(const double* __restrict__ R, const int32_t* __restrict__ owned,
                                                          int64_t n_owned, int nv, double* partial) {
  __shared__ double red[256];
  double s = 0.;
  const int64_t total = n_owned * nv;
  for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < total; k += (int64_t)gridDim.x * 256) {
    const int64_t r = k / nv;
    const double v = R[(int64_t)owned[r] * nv + (k - r * nv)];
    s += v * v;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// point loads (nonlinear_implicit_assembly.cpp:190-247): R[dofs of node] -= load, constrained dofs untouched; the host has
// summed the loads of a node in the order they were given and dropped the nodes this rank does not own
__global__ void __launch_bounds__(128) point_load_kernel(int n, const int32_t* __restrict__ node, const double* __restrict__ load,
                                                         const uint8_t* __restrict__ cmask, int nv, double* R, int node_lo, int node_hi) {
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t >= n) return;
  const int nd = node[t];
  if (nd < node_lo || nd >= node_hi) return;
  const unsigned cm = cmask ? cmask[nd] : 0u;
  for (int i = 0; i < nv; ++i)
    if (!((cm >> i) & 1u)) R[(int64_t)nd * nv + i] -= load[(int64_t)t * nv + i];
}

__global__ void __launch_bounds__(128) scatter_field_kernel(int64_t n, const int32_t* __restrict__ node,
                                                            const double* __restrict__ w, double* field) {
  const int64_t t = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (t < n) field[node[t]] = w[t];
}

// sum over the owned nodes of a . b (fixed grid: deterministic two-stage reduction)
__global__ void __launch_bounds__(256) dot_owned_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                        const int32_t* __restrict__ owned, int64_t n_owned, int nv, double* partial) {
  __shared__ double red[256];
  double s = 0.;
  const int64_t total = n_owned * nv;
  for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < total; k += (int64_t)gridDim.x * 256) {
    const int64_t r = k / nv;
    const int64_t i = (int64_t)owned[r] * nv + (k - r * nv);
    s += a[i] * b[i];
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(128) dv_map_transpose_kernel(const double* __restrict__ G, int64_t n_dv,
                                                               const int64_t* __restrict__ dv_ptr,
                                                               const int32_t* __restrict__ dv_node,
                                                               const double* __restrict__ dv_weight, double* sens) {
  const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (i >= n_dv) return;
  double s = 0.;
  for (int64_t k = dv_ptr[i]; k < dv_ptr[i + 1]; ++k) s += dv_weight[k] * G[dv_node[k]];
  sens[i] = s;
}

}  // namespace

int face_loads_blocks(int n_loaded_nodes) { return (n_loaded_nodes + LB - 1) / LB; }

cudaError_t launch_face_loads(const FaceLoadPlan& fp, const double* xyz, const double* X, const uint8_t* cmask, int nv,
                              double* R, double* dq, double* partial, int* n_blocks, cudaStream_t stream) {
  const int nb = face_loads_blocks(fp.n_loaded_nodes);
  if (n_blocks) *n_blocks = nb;
  if (nb == 0) return cudaSuccess;
  face_load_kernel<<<nb, LB, 0, stream>>>(fp, xyz, X, cmask, nv, R, dq, partial);
  return cudaGetLastError();
}

cudaError_t launch_reduce_partials(const double* partial, int n, double* out, cudaStream_t stream) {
  reduce_partials_kernel<<<1, 256, 0, stream>>>(partial, n, out);
  return cudaGetLastError();
}

cudaError_t launch_sumsq_owned(const double* R, const int32_t* owned_nodes, int64_t n_owned, int nv, double* partial,
                               int n_blocks, cudaStream_t stream) {
  sumsq_owned_kernel<<<n_blocks, 256, 0, stream>>>(R, owned_nodes, n_owned, nv, partial);
  return cudaGetLastError();
}

cudaError_t launch_point_loads(int n, const int32_t* node, const double* load, const uint8_t* cmask, int nv, double* R,
                              cudaStream_t stream, int node_lo, int node_hi) {
  if (n == 0) return cudaSuccess;
  point_load_kernel<<<(n + 127) / 128, 128, 0, stream>>>(n, node, load, cmask, nv, R, node_lo, node_hi);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) sub_vector_kernel(int64_t n, double* __restrict__ a, const double* __restrict__ b) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t < n) a[t] -= b[t];
}
cudaError_t launch_sub_vector(int64_t n, double* a, const double* b, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  sub_vector_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, a, b);
  return cudaGetLastError();
}

cudaError_t launch_scatter_field(int64_t n, const int32_t* node, const double* w, double* field, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  scatter_field_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(n, node, w, field);
  return cudaGetLastError();
}

cudaError_t launch_dot_owned(const double* a, const double* b, const int32_t* owned_nodes, int64_t n_owned, int nv,
                             double* partial, int n_blocks, cudaStream_t stream) {
  dot_owned_kernel<<<n_blocks, 256, 0, stream>>>(a, b, owned_nodes, n_owned, nv, partial);
  return cudaGetLastError();
}

cudaError_t launch_dv_map_transpose(const double* G, int64_t n_dv, const int64_t* dv_ptr, const int32_t* dv_node,
                                    const double* dv_weight, double* sens, cudaStream_t stream) {
  if (n_dv == 0) return cudaSuccess;
  dv_map_transpose_kernel<<<(unsigned)((n_dv + 127) / 128), 128, 0, stream>>>(G, n_dv, dv_ptr, dv_node, dv_weight, sens);
  return cudaGetLastError();
}

}  // namespace mastb200
