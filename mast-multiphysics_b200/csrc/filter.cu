// filter.cu -- the density filter of the SIMP loop on the device: radius search at setup, rho = W d, grad = W^T G, material
// volume and its gradient (sm_100a, fp64).
//
// Device counterpart of MAST::FilterBase (/root/reference/src/level_set/filter_base.cpp): the geometric filter map of
// ::_init / ::_init2 (:275-372, 378-447: every node within `radius` of node 1 with weight (radius - d_12) / sum, the
// neighbours in ascending node id as the linear search of ::_init meets them), ::compute_filtered_values (:59-97: filtered
// value for the design-variable nodes, the input itself for the others) and, for the gradient recipe of
// examples/structural/example_6/example_6.cpp:877-939, its transpose; plus the material volume of example_6.cpp:714-791
// (sum over the elements of volume x mean nodal density) and its derivative w.r.t. the nodal densities.
//
//   setup   nodes are binned into a uniform grid of cell size `radius` (cell ids sorted with thrust -- setup, not the hot
//           path); one thread per node counts, then collects the neighbours within the radius from the 27 cells around it,
//           orders them by node id and keeps the raw weights (radius - d) and their sum.
//   apply   one thread per node: rho_n = sum_k raw_nk d_k / sum_n (fixed order: bitwise reproducible); the transpose uses
//           the symmetry of the neighbour relation: grad_n = sum_k raw_nk G_k / sum_k over the design-variable rows k.
//   volume  V = v . rho with the nodal volumes v_n = sum_{e ∋ n} vol(e) / nn computed once per mesh by a row-owner gather.
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/scan.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>

#include "context.h"

namespace {

struct Grid {
  double lo[3];
  double h;
  int n[3];
};

__device__ __forceinline__ int cell_coord(double x, double lo, double h, int n) {
  int c = (int)floor((x - lo) / h);
  return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

__global__ void __launch_bounds__(256) cell_id_kernel(int64_t n_nodes, const double* __restrict__ xyz, Grid g, int64_t* cell) {
  const int64_t n = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (n >= n_nodes) return;
  const int cx = cell_coord(xyz[n * 3], g.lo[0], g.h, g.n[0]), cy = cell_coord(xyz[n * 3 + 1], g.lo[1], g.h, g.n[1]),
            cz = cell_coord(xyz[n * 3 + 2], g.lo[2], g.h, g.n[2]);
  cell[n] = ((int64_t)cz * g.n[1] + cy) * g.n[0] + cx;
}

// first position of every cell in the sorted (cell, node) list; cell_start has n_cells + 1 entries
__global__ void __launch_bounds__(256) cell_start_kernel(int64_t n_nodes, const int64_t* __restrict__ sorted_cell, int64_t n_cells,
                                                         int64_t* cell_start) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i > n_nodes) return;
  const int64_t cur = i < n_nodes ? sorted_cell[i] : n_cells, prev = i > 0 ? sorted_cell[i - 1] : -1;
  for (int64_t c = prev + 1; c <= cur; ++c) cell_start[c] = i;
}

// pass 0: count the neighbours of every node; pass 1: collect them (ascending node id), raw weights and their sum
template <int PASS>
__global__ void __launch_bounds__(128) neighbours_kernel(int64_t n_nodes, const double* __restrict__ xyz, Grid g, double radius,
                                                         const int64_t* __restrict__ cell_start, const int32_t* __restrict__ sorted_node,
                                                         int64_t* count, const int64_t* __restrict__ row_ptr, int32_t* col, double* raw,
                                                         double* sum) {
  const int64_t n = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (n >= n_nodes) return;
  const double x = xyz[n * 3], y = xyz[n * 3 + 1], z = xyz[n * 3 + 2];
  const int cx = cell_coord(x, g.lo[0], g.h, g.n[0]), cy = cell_coord(y, g.lo[1], g.h, g.n[1]), cz = cell_coord(z, g.lo[2], g.h, g.n[2]);
  int64_t k = PASS ? row_ptr[n] : 0;
  const int64_t k0 = k;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int ax = cx + dx, ay = cy + dy, az = cz + dz;
        if (ax < 0 || ay < 0 || az < 0 || ax >= g.n[0] || ay >= g.n[1] || az >= g.n[2]) continue;
        const int64_t c = ((int64_t)az * g.n[1] + ay) * g.n[0] + ax;
        for (int64_t i = cell_start[c]; i < cell_start[c + 1]; ++i) {
          const int32_t m = sorted_node[i];
          const double ex = x - xyz[(int64_t)m * 3], ey = y - xyz[(int64_t)m * 3 + 1], ez = z - xyz[(int64_t)m * 3 + 2];
          const double d = sqrt(ex * ex + ey * ey + ez * ez);  // (node_1 - node_2).norm(), filter_base.cpp:415-416
          if (d <= radius) {
            if (PASS) {
              // insertion by node id: the list stays ascending, the order of the reference's linear search
              int64_t p = k;
              while (p > k0 && col[p - 1] > m) { col[p] = col[p - 1]; raw[p] = raw[p - 1]; --p; }
              col[p] = m; raw[p] = radius - d;
            }
            ++k;
          }
        }
      }
  if (!PASS) {
    count[n] = k;
  } else {
    double s = 0.;
    for (int64_t p = k0; p < k; ++p) s += raw[p];  // ascending node id, as the reference accumulates `sum`
    sum[n] = s;
  }
}

// rho_n = sum_k (raw_nk / sum_n) d_k for design-variable nodes, d_n otherwise (compute_filtered_values, :59-97)
__global__ void __launch_bounds__(128) filter_apply_kernel(int64_t n_nodes, const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                                                           const double* __restrict__ raw, const double* __restrict__ sum,
                                                           const uint8_t* __restrict__ dv, const double* __restrict__ in, double* out) {
  const int64_t n = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (n >= n_nodes) return;
  if (dv && !dv[n]) { out[n] = in[n]; return; }
  const double s = sum[n];
  double acc = 0.;
  for (int64_t p = row_ptr[n]; p < row_ptr[n + 1]; ++p) acc += in[col[p]] * (raw[p] / s);
  out[n] = acc;
}

// grad_n = sum over design-variable rows k within the radius of W[k,n] G_k = raw_nk / sum_k G_k  (+ G_n for a node that is
// not a design variable: its row of W is the identity)
__global__ void __launch_bounds__(128) filter_transpose_kernel(int64_t n_nodes, const int64_t* __restrict__ row_ptr,
                                                               const int32_t* __restrict__ col, const double* __restrict__ raw,
                                                               const double* __restrict__ sum, const uint8_t* __restrict__ dv,
                                                               const double* __restrict__ G, double* grad) {
  const int64_t n = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (n >= n_nodes) return;
  double acc = (dv && !dv[n]) ? G[n] : 0.;
  for (int64_t p = row_ptr[n]; p < row_ptr[n + 1]; ++p) {
    const int32_t k = col[p];
    if (dv && !dv[k]) continue;
    acc += G[k] * (raw[p] / sum[k]);
  }
  grad[n] = acc;
}

// element volumes (sum of JxW over the quadrature points; tab: N[q][b], dN[q][b][3], w[q]) and nodal volumes
__global__ void __launch_bounds__(128) elem_volume_kernel(int64_t n_elem, int nn, int nq, const int32_t* __restrict__ conn,
                                                          const double* __restrict__ tab, const double* __restrict__ xyz, double* vol) {
  const int64_t e = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (e >= n_elem) return;
  const int32_t* en = conn + e * nn;
  double v = 0.;
  for (int q = 0; q < nq; ++q) {
    const double* dNt = tab + (size_t)nq * nn + (size_t)q * nn * 3;
    double J[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};
    for (int b = 0; b < nn; ++b) {
      const double* X = xyz + (int64_t)en[b] * 3;
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) J[r][c] = fma(dNt[b * 3 + r], X[c], J[r][c]);
    }
    const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) + J[0][1] * (J[1][2] * J[2][0] - J[1][0] * J[2][2]) +
                       J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    v += det * tab[(size_t)nq * nn * 4 + q];
  }
  vol[e] = v;
}
__global__ void __launch_bounds__(128) nodal_volume_kernel(int64_t n_nodes, int nn, const int64_t* __restrict__ n2e_ptr,
                                                           const int32_t* __restrict__ n2e, const double* __restrict__ vol,
                                                           const uint8_t* __restrict__ diag_elem, double* nodal) {
  const int64_t n = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (n >= n_nodes) return;
  double acc = 0.;
  for (int64_t k = n2e_ptr[n]; k < n2e_ptr[n + 1]; ++k) {
    const int32_t e = n2e[k];
    if (diag_elem && diag_elem[e]) continue;
    acc += vol[e] / nn;
  }
  nodal[n] = acc;
}

}  // namespace

extern "C" {

int mastb200_filter_init(mastb200_context* ctx, double radius, const uint8_t* dv_node, int64_t* nnz_out) {
  if (!ctx || !(radius > 0.)) return MASTB200_ERR_ARG;
  if (!ctx->have_mesh) return ctx->fail(MASTB200_ERR_ARG, "set_mesh first");
  if (ctx->planning_only) return ctx->fail(MASTB200_ERR_CUDA, "planning-only context (device -1): the filter runs on the device; there is no CPU fallback");
  CU(cudaSetDevice(ctx->device));
  const int64_t nn = ctx->n_nodes;
  Grid g;
  double hi[3];
  for (int d = 0; d < 3; ++d) { g.lo[d] = 1e300; hi[d] = -1e300; }
  for (int64_t n = 0; n < nn; ++n)
    for (int d = 0; d < 3; ++d) { g.lo[d] = std::min(g.lo[d], ctx->xyz[n * 3 + d]); hi[d] = std::max(hi[d], ctx->xyz[n * 3 + d]); }
  g.h = radius * (1. + 1e-12);  // a neighbour within the radius is never more than one cell away
  int64_t n_cells = 0;
  for (;;) {  // a radius far below the node spacing would make more cells than nodes: coarser cells keep the search exact
    n_cells = 1;
    bool too_many = false;
    for (int d = 0; d < 3; ++d) {
      const double c = std::floor((hi[d] - g.lo[d]) / g.h) + 1.;
      if (c > 1e6) { too_many = true; break; }
      g.n[d] = (int)c;
      n_cells *= g.n[d];
    }
    if (!too_many && n_cells <= 4 * nn + 1024) break;
    g.h *= 2.;
  }
  DevBuf<int64_t> d_cell, d_start, d_count;
  DevBuf<int32_t> d_node;
  CU(d_cell.ensure(nn)); CU(d_node.ensure(nn)); CU(d_start.ensure(n_cells + 1)); CU(d_count.ensure(nn + 1));
  cell_id_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, ctx->stream>>>(nn, ctx->d_xyz.p, g, d_cell.p);
  auto pol = thrust::cuda::par.on(ctx->stream);
  thrust::sequence(pol, thrust::device_ptr<int32_t>(d_node.p), thrust::device_ptr<int32_t>(d_node.p + nn));
  thrust::stable_sort_by_key(pol, thrust::device_ptr<int64_t>(d_cell.p), thrust::device_ptr<int64_t>(d_cell.p + nn),
                             thrust::device_ptr<int32_t>(d_node.p));
  cell_start_kernel<<<(unsigned)((nn + 256) / 256), 256, 0, ctx->stream>>>(nn, d_cell.p, n_cells, d_start.p);
  neighbours_kernel<0><<<(unsigned)((nn + 127) / 128), 128, 0, ctx->stream>>>(nn, ctx->d_xyz.p, g, radius, d_start.p, d_node.p, d_count.p,
                                                                               nullptr, nullptr, nullptr, nullptr);
  CU(ctx->f_ptr.ensure(nn + 1));
  CU(cudaMemsetAsync(d_count.p + nn, 0, sizeof(int64_t), ctx->stream));
  thrust::exclusive_scan(pol, thrust::device_ptr<int64_t>(d_count.p), thrust::device_ptr<int64_t>(d_count.p + nn + 1),
                         thrust::device_ptr<int64_t>(ctx->f_ptr.p));
  int64_t nnz = 0;
  CU(cudaMemcpyAsync(&nnz, ctx->f_ptr.p + nn, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(ctx->f_col.ensure(nnz)); CU(ctx->f_raw.ensure(nnz)); CU(ctx->f_sum.ensure(nn));
  neighbours_kernel<1><<<(unsigned)((nn + 127) / 128), 128, 0, ctx->stream>>>(nn, ctx->d_xyz.p, g, radius, d_start.p, d_node.p, nullptr,
                                                                               ctx->f_ptr.p, ctx->f_col.p, ctx->f_raw.p, ctx->f_sum.p);
  ctx->f_has_dv = dv_node != nullptr;
  if (dv_node) {
    CU(ctx->f_dv.ensure(nn));
    CU(cudaMemcpyAsync(ctx->f_dv.p, dv_node, nn, cudaMemcpyHostToDevice, ctx->stream));
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->launches += 4;
  ctx->f_nnz = nnz; ctx->f_radius = radius; ctx->have_filter = true;
  if (nnz_out) *nnz_out = nnz;
  return MASTB200_OK;
}

int mastb200_filter_get(mastb200_context* ctx, int64_t* row_ptr, int32_t* col, double* weight) {
  if (!ctx || !row_ptr) return MASTB200_ERR_ARG;
  if (!ctx->have_filter) return ctx->fail(MASTB200_ERR_ARG, "mastb200_filter_init first");
  CU(cudaSetDevice(ctx->device));
  const int64_t nn = ctx->n_nodes;
  CU(cudaMemcpyAsync(row_ptr, ctx->f_ptr.p, (nn + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (col) CU(cudaMemcpyAsync(col, ctx->f_col.p, ctx->f_nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<double> sum;
  if (weight) {
    sum.resize(nn);
    CU(cudaMemcpyAsync(weight, ctx->f_raw.p, ctx->f_nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(sum.data(), ctx->f_sum.p, nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  if (weight)
    for (int64_t n = 0; n < nn; ++n)
      for (int64_t p = row_ptr[n]; p < row_ptr[n + 1]; ++p) weight[p] /= sum[n];  // vec[i].second /= sum  (:434-438)
  return MASTB200_OK;
}

static int filter_run(mastb200_context* ctx, const double* in, double* out, bool transpose) {
  if (!ctx || !in || !out) return MASTB200_ERR_ARG;
  if (!ctx->have_filter) return ctx->fail(MASTB200_ERR_ARG, "mastb200_filter_init first");
  CU(cudaSetDevice(ctx->device));
  const int64_t nn = ctx->n_nodes;
  const double* din = nullptr;
  double* dout = nullptr;
  int rc;
  if ((rc = resolve_in(ctx, in, nn, ctx->d_G, &din))) return rc;
  if ((rc = resolve_out(ctx, out, nn, ctx->d_dfield, &dout))) return rc;
  if (din == dout) return ctx->fail(MASTB200_ERR_ARG, "filter: input and output must be different arrays");
  const uint8_t* dv = ctx->f_has_dv ? ctx->f_dv.p : nullptr;
  if (transpose)
    filter_transpose_kernel<<<(unsigned)((nn + 127) / 128), 128, 0, ctx->stream>>>(nn, ctx->f_ptr.p, ctx->f_col.p, ctx->f_raw.p,
                                                                                   ctx->f_sum.p, dv, din, dout);
  else
    filter_apply_kernel<<<(unsigned)((nn + 127) / 128), 128, 0, ctx->stream>>>(nn, ctx->f_ptr.p, ctx->f_col.p, ctx->f_raw.p, ctx->f_sum.p,
                                                                               dv, din, dout);
  CU(cudaGetLastError());
  ctx->launches++;
  if ((rc = finish_out(ctx, out, nn, dout))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}

int mastb200_filter_apply(mastb200_context* ctx, const double* d, double* rho) { return filter_run(ctx, d, rho, false); }
int mastb200_filter_apply_transpose(mastb200_context* ctx, const double* G, double* grad) { return filter_run(ctx, G, grad, true); }

int mastb200_volume(mastb200_context* ctx, const double* rho, double* volume, double* dV_drho) {
  if (!ctx || (!volume && !dV_drho)) return MASTB200_ERR_ARG;
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (ctx->elem_type != MASTB200_HEX8 && ctx->elem_type != MASTB200_HEX27)
    return ctx->fail(MASTB200_ERR_UNSUPPORTED, "volume: 3D solids (HEX8, HEX27) only");
  CU(cudaSetDevice(ctx->device));
  if ((rc = ensure_generic_plan(ctx))) return rc;
  mastb200_context::GenericPlan& G = ctx->gplan;
  const int64_t nn = ctx->n_nodes;
  if (!ctx->have_nodal_vol) {  // once per mesh
    DevBuf<double> vol;
    CU(vol.ensure(ctx->n_elem));
    CU(ctx->nodal_vol.ensure(nn));
    elem_volume_kernel<<<(unsigned)((ctx->n_elem + 127) / 128), 128, 0, ctx->stream>>>(ctx->n_elem, ctx->nn, ctx->nn, G.conn.p, G.tab.p,
                                                                                      ctx->d_xyz.p, vol.p);
    nodal_volume_kernel<<<(unsigned)((nn + 127) / 128), 128, 0, ctx->stream>>>(nn, ctx->nn, G.n2e_ptr.p, G.n2e.p, vol.p,
                                                                              ctx->has_diag ? G.diag.p : nullptr, ctx->nodal_vol.p);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->launches += 2;
    ctx->have_nodal_vol = true;
  }
  if (volume) {
    if (!rho) return ctx->fail(MASTB200_ERR_ARG, "volume needs the nodal density");
    const double* drho = nullptr;
    if ((rc = resolve_in(ctx, rho, nn, ctx->d_G, &drho))) return rc;
    const int nb = 1024;
    CU(ctx->d_partial.ensure(4096)); CU(ctx->d_scalar.ensure(8));
    CU(launch_dot_owned(ctx->nodal_vol.p, drho, ctx->d_owned_nodes.p, ctx->pat.n_owned, 1, ctx->d_partial.p, nb, ctx->stream));
    CU(launch_reduce_partials(ctx->d_partial.p, nb, ctx->d_scalar.p + 2, ctx->stream));
    ctx->launches += 2;
    CU(cudaMemcpyAsync(volume, ctx->d_scalar.p + 2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (dV_drho) {
    if (is_device_ptr(dV_drho)) CU(cudaMemcpyAsync(dV_drho, ctx->nodal_vol.p, nn * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    else CU(cudaMemcpyAsync(dV_drho, ctx->nodal_vol.p, nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}

}  // extern "C"
