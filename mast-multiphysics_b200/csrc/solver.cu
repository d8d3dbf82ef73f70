// solver.cu -- host side of the communicator (ghost exchange over NCCL) and of the device linear solve.
//
//   communicator  <- AssemblyBase::build_localized_vector (/root/reference/src/base/assembly_base.cpp:209-228): the reference
//                    localizes X inside every assembly call (nonlinear_implicit_assembly.cpp:93-95) and sums norms /
//                    outputs over the ranks (:258-265, compliance_output.cpp:239-262).  Here the context owns an
//                    ncclComm_t (created from a unique id the host broadcasts, or attached) and per-neighbour send / recv
//                    node lists; the exchange is enqueued on the context stream -- no host wait between halo and sweep.
//                    NCCL is reached through dlopen("libnccl.so.2"), so the library loads on boxes without it.
//   solve         <- the Krylov solve after assembly (src/base/nonlinear_system.cpp:777-782, 834-839); kernels in solve.cu.
#include <dlfcn.h>
#include <nccl.h>

#include "context.h"

namespace {

struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
  bool load() {
    if (h) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) { err = std::string("dlopen(libnccl.so.2): ") + (dlerror() ? dlerror() : "not found"); return false; }
    auto sym = [&](const char* n) { return dlsym(h, n); };
    GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
    Send = (decltype(Send))sym("ncclSend");
    Recv = (decltype(Recv))sym("ncclRecv");
    AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
    GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
    GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !Send || !Recv || !AllReduce || !GroupStart || !GroupEnd) {
      err = "libnccl.so.2 lacks a required symbol";
      h = nullptr;
      return false;
    }
    return true;
  }
};
NcclApi g_nccl;

#define NC(call)                                                                                              \
  do {                                                                                                        \
    ncclResult_t r_ = (call);                                                                                 \
    if (r_ != ncclSuccess)                                                                                    \
      return ctx->fail(MASTB200_ERR_CUDA, "%s: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error"); \
  } while (0)

__global__ void __launch_bounds__(256) pack_nodes_kernel(int64_t n, const int32_t* __restrict__ nodes, int width,
                                                         const double* __restrict__ V, double* __restrict__ buf) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= n * width) return;
  const int64_t k = t / width;
  buf[t] = V[(int64_t)nodes[k] * width + (t - k * width)];
}
__global__ void __launch_bounds__(256) unpack_nodes_kernel(int64_t n, const int32_t* __restrict__ nodes, int width,
                                                           const double* __restrict__ buf, double* __restrict__ V) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= n * width) return;
  const int64_t k = t / width;
  V[(int64_t)nodes[k] * width + (t - k * width)] = buf[t];
}

// >= 0 when nodes[0..n) is the ascending contiguous range [first, first + n)
int64_t contiguous_first(const int32_t* nodes, int64_t n) {
  if (n == 0) return -1;
  for (int64_t i = 1; i < n; ++i)
    if (nodes[i] != nodes[0] + i) return -1;
  return nodes[0];
}

}  // namespace

void comm_release(mastb200_context* ctx) {
  for (auto* nb : ctx->neighbours) delete nb;
  ctx->neighbours.clear();
  if (ctx->nccl_comm && ctx->nccl_owned && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
  ctx->nccl_owned = false;
}

// ghost refresh of the device vector dV (width values per node) on the context stream
int comm_exchange(mastb200_context* ctx, double* dV, int width) {
  if (ctx->neighbours.empty()) return MASTB200_OK;
  if (!ctx->nccl_comm) return ctx->fail(MASTB200_ERR_ARG, "ghost exchange configured without a communicator");
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  for (auto* nb : ctx->neighbours) {
    if (nb->n_send && nb->send_first < 0) {
      CU(nb->send_buf.ensure(nb->n_send * width));
      pack_nodes_kernel<<<(unsigned)((nb->n_send * width + 255) / 256), 256, 0, ctx->stream>>>(nb->n_send, nb->d_send.p, width, dV,
                                                                                              nb->send_buf.p);
      ctx->launches++;
    }
    if (nb->n_recv && nb->recv_first < 0) CU(nb->recv_buf.ensure(nb->n_recv * width));
  }
  NC(g_nccl.GroupStart());
  for (auto* nb : ctx->neighbours) {
    if (nb->n_send) {
      const double* src = nb->send_first >= 0 ? dV + nb->send_first * width : nb->send_buf.p;
      NC(g_nccl.Send(src, (size_t)(nb->n_send * width), ncclDouble, nb->rank, comm, ctx->stream));
    }
    if (nb->n_recv) {
      double* dst = nb->recv_first >= 0 ? dV + nb->recv_first * width : nb->recv_buf.p;
      NC(g_nccl.Recv(dst, (size_t)(nb->n_recv * width), ncclDouble, nb->rank, comm, ctx->stream));
    }
  }
  NC(g_nccl.GroupEnd());
  for (auto* nb : ctx->neighbours)
    if (nb->n_recv && nb->recv_first < 0) {
      unpack_nodes_kernel<<<(unsigned)((nb->n_recv * width + 255) / 256), 256, 0, ctx->stream>>>(nb->n_recv, nb->d_recv.p, width,
                                                                                                nb->recv_buf.p, dV);
      ctx->launches++;
    }
  CU(cudaGetLastError());
  return MASTB200_OK;
}

// in-place sum over the ranks of n device doubles (no-op for one rank)
int comm_all_reduce(mastb200_context* ctx, double* d_vals, int n) {
  if (ctx->comm_world <= 1 || !ctx->nccl_comm) return MASTB200_OK;
  NC(g_nccl.AllReduce(d_vals, d_vals, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  return MASTB200_OK;
}

extern "C" {

int mastb200_comm_unique_id(void* id, int32_t bytes) {
  if (!id || bytes < (int32_t)sizeof(ncclUniqueId)) return MASTB200_ERR_ARG;
  if (!g_nccl.load()) { g_create_error = g_nccl.err; return MASTB200_ERR_CUDA; }
  ncclUniqueId u;
  if (g_nccl.GetUniqueId(&u) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return MASTB200_ERR_CUDA; }
  std::memcpy(id, &u, sizeof u);
  return MASTB200_OK;
}

int mastb200_comm_init(mastb200_context* ctx, int32_t rank, int32_t world, const void* id) {
  if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return MASTB200_ERR_ARG;
  if (ctx->planning_only) return ctx->fail(MASTB200_ERR_CUDA, "planning-only context");
  if (!g_nccl.load()) return ctx->fail(MASTB200_ERR_CUDA, "%s", g_nccl.err.c_str());
  CU(cudaSetDevice(ctx->device));
  if (ctx->nccl_comm && ctx->nccl_owned) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof u);
  ncclComm_t comm = nullptr;
  NC(g_nccl.CommInitRank(&comm, world, u, rank));
  ctx->nccl_comm = comm; ctx->nccl_owned = true; ctx->comm_rank = rank; ctx->comm_world = world;
  return MASTB200_OK;
}

int mastb200_comm_attach(mastb200_context* ctx, void* nccl_comm, int32_t rank, int32_t world) {
  if (!ctx || world < 1 || rank < 0 || rank >= world) return MASTB200_ERR_ARG;
  if (nccl_comm && !g_nccl.load()) return ctx->fail(MASTB200_ERR_CUDA, "%s", g_nccl.err.c_str());
  if (ctx->nccl_comm && ctx->nccl_owned) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = nccl_comm; ctx->nccl_owned = false; ctx->comm_rank = rank; ctx->comm_world = world;
  return MASTB200_OK;
}

int mastb200_set_ghost_exchange(mastb200_context* ctx, int32_t n_nbr, const int32_t* nbr_rank, const int64_t* send_ptr,
                                const int32_t* send_nodes, const int64_t* recv_ptr, const int32_t* recv_nodes, int32_t auto_localize) {
  if (!ctx || n_nbr < 0) return MASTB200_ERR_ARG;
  if (!ctx->have_mesh) return ctx->fail(MASTB200_ERR_ARG, "set_mesh first");
  if (ctx->planning_only) return ctx->fail(MASTB200_ERR_CUDA, "planning-only context");
  CU(cudaSetDevice(ctx->device));
  for (auto* nb : ctx->neighbours) delete nb;
  ctx->neighbours.clear();
  ctx->auto_localize = auto_localize != 0;
  if (n_nbr == 0) return MASTB200_OK;
  if (!nbr_rank || !send_ptr || !recv_ptr) return MASTB200_ERR_ARG;
  for (int k = 0; k < n_nbr; ++k) {
    auto* nb = new mastb200_context::Neighbour();
    ctx->neighbours.push_back(nb);
    nb->rank = nbr_rank[k];
    nb->n_send = send_ptr[k + 1] - send_ptr[k];
    nb->n_recv = recv_ptr[k + 1] - recv_ptr[k];
    const int32_t* sn = send_nodes + send_ptr[k];
    const int32_t* rn = recv_nodes + recv_ptr[k];
    for (int64_t i = 0; i < nb->n_send; ++i)
      if (sn[i] < 0 || sn[i] >= ctx->n_nodes) return ctx->fail(MASTB200_ERR_ARG, "bad send node");
    for (int64_t i = 0; i < nb->n_recv; ++i)
      if (rn[i] < 0 || rn[i] >= ctx->n_nodes) return ctx->fail(MASTB200_ERR_ARG, "bad recv node");
    nb->send_first = contiguous_first(sn, nb->n_send);
    nb->recv_first = contiguous_first(rn, nb->n_recv);
    if (nb->n_send && nb->send_first < 0) {
      CU(nb->d_send.ensure(nb->n_send));
      CU(cudaMemcpyAsync(nb->d_send.p, sn, nb->n_send * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (nb->n_recv && nb->recv_first < 0) {
      CU(nb->d_recv.ensure(nb->n_recv));
      CU(cudaMemcpyAsync(nb->d_recv.p, rn, nb->n_recv * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}

int mastb200_build_localized_vector(mastb200_context* ctx, double* V, int32_t width) {
  if (!ctx || !V || width < 1) return MASTB200_ERR_ARG;
  if (ctx->planning_only) return ctx->fail(MASTB200_ERR_CUDA, "planning-only context");
  CU(cudaSetDevice(ctx->device));
  const size_t n = size_t(ctx->n_nodes) * width;
  if (is_device_ptr(V)) {
    if (int rc = comm_exchange(ctx, V, width)) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return MASTB200_OK;
  }
  CU(ctx->d_vec2.ensure(n));
  CU(cudaMemcpyAsync(ctx->d_vec2.p, V, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (int rc = comm_exchange(ctx, ctx->d_vec2.p, width)) return rc;
  CU(cudaMemcpyAsync(V, ctx->d_vec2.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}

int mastb200_all_reduce_sum(mastb200_context* ctx, double* vals, int32_t n) {
  if (!ctx || !vals || n < 1 || n > 64) return MASTB200_ERR_ARG;
  if (ctx->comm_world <= 1 || !ctx->nccl_comm) return MASTB200_OK;
  CU(cudaSetDevice(ctx->device));
  CU(ctx->d_pscal.ensure(128));
  CU(cudaMemcpyAsync(ctx->d_pscal.p + 64, vals, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (int rc = comm_all_reduce(ctx, ctx->d_pscal.p + 64, n)) return rc;
  CU(cudaMemcpyAsync(vals, ctx->d_pscal.p + 64, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}

int mastb200_near_null_space(mastb200_context* ctx, double* modes) {
  if (!ctx || !modes) return MASTB200_ERR_ARG;
  if (!ctx->have_mesh || !ctx->have_dofs) return ctx->fail(MASTB200_ERR_ARG, "set_mesh and set_dof_map first");
  if (ctx->planning_only) return ctx->fail(MASTB200_ERR_CUDA, "planning-only context");
  if (ctx->nv != 3 && ctx->nv != 6) return ctx->fail(MASTB200_ERR_ARG, "rigid-body modes need a structural system (nv = 3 or 6)");
  CU(cudaSetDevice(ctx->device));
  const size_t n = size_t(ctx->n_nodes) * ctx->nv * 6;
  double* d = nullptr;
  if (int rc = resolve_out(ctx, modes, n, ctx->d_scratch, &d)) return rc;
  CU(launch_near_null(ctx->n_nodes, ctx->nv, ctx->d_xyz.p, d, ctx->stream));
  ctx->launches++;
  if (int rc = finish_out(ctx, modes, n, d)) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}

int mastb200_solve(mastb200_context* ctx, const double* J_vals, const double* rhs, double* x, double rtol, int32_t max_it,
                   mastb200_solve_info* info) {
  if (!ctx || !J_vals || !rhs || !x) return MASTB200_ERR_ARG;
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (ctx->nv != 3) return ctx->fail(MASTB200_ERR_UNSUPPORTED, "solve: the 3-dof structural system only (the rotational rows of nv = 6 are 1e-20)");
  CU(cudaSetDevice(ctx->device));
  if ((rc = ensure_generic_plan(ctx))) return rc;
  mastb200_context::GenericPlan& G = ctx->gplan;
  if (!G.nbr.p) {  // the HEX27 sweep works from position maps: the neighbour lists go up on the first solve
    CU(G.nbr.upload(ctx->pat.nbr, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  const int64_t n_owned = ctx->pat.n_owned;
  const size_t n_dof = size_t(ctx->n_nodes) * 3, nnz = (size_t)plan_nnz(ctx);
  const double *dJ = nullptr, *drhs = nullptr;
  double* dx = nullptr;
  if ((rc = resolve_in(ctx, J_vals, nnz, ctx->d_J, &dJ))) return rc;
  if ((rc = resolve_in(ctx, rhs, n_dof, ctx->d_R, &drhs))) return rc;
  if ((rc = resolve_out(ctx, x, n_dof, ctx->d_px, &dx))) return rc;
  const int RB = pcg_reduction_blocks();
  CU(ctx->d_minv.ensure(size_t(n_owned) * 9));
  CU(ctx->d_pr.ensure(n_dof)); CU(ctx->d_pz.ensure(n_dof)); CU(ctx->d_pp.ensure(n_dof)); CU(ctx->d_pAp.ensure(n_dof));
  CU(ctx->d_ppart.ensure(size_t(3) * RB)); CU(ctx->d_pscal.ensure(128));
  double *r = ctx->d_pr.p, *z = ctx->d_pz.p, *p = ctx->d_pp.p, *Ap = ctx->d_pAp.p, *part = ctx->d_ppart.p, *scal = ctx->d_pscal.p;
  cudaEvent_t e0 = ctx->ev[0], e1 = ctx->ev[3];
  CU(cudaEventRecord(e0, ctx->stream));
  CU(cudaMemsetAsync(part, 0, size_t(3) * RB * sizeof(double), ctx->stream));
  CU(cudaMemsetAsync(scal, 0, 16 * sizeof(double), ctx->stream));
  CU(cudaMemsetAsync(dx, 0, n_dof * sizeof(double), ctx->stream));
  CU(cudaMemsetAsync(p, 0, n_dof * sizeof(double), ctx->stream));
  const int32_t* owned = ctx->d_owned_nodes.p;
  CU(launch_bj_setup(n_owned, owned, G.nbr_ptr.p, G.nbr.p, dJ, ctx->d_minv.p, ctx->stream));
  CU(launch_pcg_init(n_owned, owned, drhs, ctx->d_minv.p, dx, r, z, p, part, ctx->stream));
  CU(launch_reduce_multi(part, RB, 2, RB, scal + 3, ctx->stream));  // scal[3] = r.z, scal[4] = r.r = |rhs|^2
  if ((rc = comm_all_reduce(ctx, scal + 3, 2))) return rc;
  CU(launch_pcg_beta(scal, ctx->stream));                            // scal[0] = r.z (beta = 0: p = z already)
  ctx->launches += 4;
  double h[2] = {0., 0.};
  CU(cudaMemcpyAsync(h, scal + 3, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  const double rhs2 = h[1];
  int it = 0;
  bool converged = rhs2 == 0.;
  double rel = rhs2 == 0. ? 0. : 1.;
  const int check = 10;
  const int grid_spmv = (int)std::min<int64_t>((n_owned + 7) / 8, RB);
  while (!converged && it < max_it) {
    const int burst = std::min(check, (int)max_it - it);
    for (int k = 0; k < burst; ++k) {
      if ((rc = comm_exchange(ctx, p, 3))) return rc;
      CU(launch_spmv_node(n_owned, owned, G.nbr_ptr.p, G.nbr.p, dJ, p, Ap, part + 2 * RB, ctx->sm_count, ctx->stream));
      CU(launch_reduce_multi(part + 2 * RB, grid_spmv, 1, 0, scal + 1, ctx->stream));
      if ((rc = comm_all_reduce(ctx, scal + 1, 1))) return rc;
      CU(launch_pcg_alpha(scal, ctx->stream));
      CU(launch_pcg_update(n_owned, owned, scal, ctx->d_minv.p, p, Ap, dx, r, z, part, ctx->stream));
      CU(launch_reduce_multi(part, RB, 2, RB, scal + 3, ctx->stream));
      if ((rc = comm_all_reduce(ctx, scal + 3, 2))) return rc;
      CU(launch_pcg_beta(scal, ctx->stream));
      CU(launch_pcg_direction(n_owned, owned, scal, z, p, ctx->stream));
      ctx->launches += 7;
    }
    it += burst;
    CU(cudaMemcpyAsync(h, scal + 3, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    rel = std::sqrt(h[1] / rhs2);
    if (!(rel == rel)) return ctx->fail(MASTB200_ERR_ARG, "solve: the iteration broke down (matrix not symmetric positive definite?)");
    converged = rel <= rtol;
  }
  CU(cudaEventRecord(e1, ctx->stream));
  if ((rc = finish_out(ctx, x, n_dof, dx))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  if (info) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e0, e1) != cudaSuccess) ms = 0.f;
    info->iterations = it; info->converged = converged ? 1 : 0; info->rel_residual = rel; info->solve_ms = ms;
  }
  ctx->timed = false;
  return MASTB200_OK;
}

}  // extern "C"
