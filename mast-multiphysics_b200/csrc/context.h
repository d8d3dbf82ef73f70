// context.h -- the host-side context behind the C-ABI of include/mast_b200.h: device copies of mesh, dof map, plans and
// state, the stream all work is ordered on, and the helpers shared by context.cu (sweeps) and solver.cu (communicator,
// linear solve).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mast_b200.h"
#include "device_types.h"
#include "kernels.h"
#include "plan.h"

using namespace mastb200;

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr; n = 0;
  }
  cudaError_t ensure(size_t count) {
    if (count <= n && p) return cudaSuccess;
    release();
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  cudaError_t upload(const std::vector<T>& v, cudaStream_t s) {
    cudaError_t e = ensure(v.size());
    if (e != cudaSuccess || v.empty()) return e;
    return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
  }
};

inline bool is_device_ptr(const void* ptr) {
  if (!ptr) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

static const int H8_SIDE[6][4] = {{0, 3, 2, 1}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};
static const int H27_SIDE[6][9] = {{0, 3, 2, 1, 11, 10, 9, 8, 20},  {0, 1, 5, 4, 8, 13, 16, 12, 21}, {1, 2, 6, 5, 9, 14, 17, 13, 22},
                            {2, 3, 7, 6, 10, 15, 18, 14, 23}, {3, 0, 4, 7, 11, 12, 19, 15, 24}, {4, 5, 6, 7, 16, 17, 18, 19, 25}};

extern std::string g_create_error;

struct mastb200_context {
  int device = 0;
  bool planning_only = false;  // device == -1: pattern/plan construction on the host only, no sweeps
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  bool timed = false;
  std::string err;
  int sm_count = 0;

  // mesh / dof map (host)
  int32_t elem_type = 0, nn = 0, nv = 0;
  int64_t n_elem = 0, n_nodes = 0;
  std::vector<double> xyz;
  std::vector<int32_t> conn;
  std::vector<uint8_t> diag_elem, owned, cmask;
  bool has_constraints = false, has_ghosts = false, has_diag = false;
  mastb200_material mat{};
  bool have_mesh = false, have_dofs = false, have_mat = false, have_plan = false;

  Pattern pat;
  // streaming plans: [0] plain layout (R/J of the 3-dof system), [1] aux layout (adjoint, density, temperature, nv=6)
  struct DevicePlan {
    bool built = false;
    PlanLimits lim{};
    size_t smem = 0;
    int n_ctas = 0;
    int64_t n_tiles = 0, n_visits = 0, bytes = 0;
    int32_t max_live_blocks = 0, max_live_slots = 0;
    StreamPlan host;  // kept only by planning-only contexts (tests inspect it)
    DevBuf<TileDesc> tiles;
    DevBuf<StepDesc> steps;
    DevBuf<VisitRec8> visits;
    DevBuf<FlushRow> frows;
    DevBuf<uint32_t> fent;
    DevBuf<LoadRec> loads;
    void reset() {
      built = false; host = StreamPlan();
      tiles.release(); steps.release(); visits.release(); frows.release(); fent.release(); loads.release();
    }
  } dplan[2];
  int last_plan = 0;
  bool last_struct = false;
  double carry_ms = 0.;  // time of the first part of a two-part sweep (R/J, then sensitivities), added to the stats
  // device plan of the gather (QUAD4 heat) and coloured-scatter (HEX27) sweeps
  struct GenericPlan {
    bool built = false;
    DevBuf<int32_t> conn, n2e, nbr, node_row, color_elems, ent_row;  // ent_row: owned row of every neighbour entry (plate)
    // ordered single-launch HEX27 scatter: processing order, predecessor lists, done flags (value = sweep counter)
    DevBuf<int32_t> order, pred_ptr, pred, done;
    int32_t n_order = 0, epoch = 0;
    bool ordered = false, first_writer = false;
    DevBuf<int64_t> n2e_ptr, nbr_ptr;
    DevBuf<uint8_t> diag, posmap;
    DevBuf<double> tab;
    std::vector<int32_t> color_ptr;
    // two-pass HEX27 assembly: chunks of consecutive element ids, the owned rows each chunk touches, the element-major buffers
    bool tp_built = false;
    int32_t tp_chunk = 0;
    std::vector<int64_t> tp_row_ptr;     // per chunk: range in tp_rows
    DevBuf<int32_t> tp_rows, tp_ea;      // rows per chunk; per row the (element - chunk start) * 27 + local node codes
    DevBuf<int64_t> tp_ea_ptr;           // [rows + chunks]: per chunk, rows + 1 offsets into tp_ea
    DevBuf<uint8_t> tp_first;
    DevBuf<double> tp_K, tp_f;
    void reset() {
      built = false; color_ptr.clear(); ordered = false; first_writer = false; n_order = 0; epoch = 0;
      order.release(); pred_ptr.release(); pred.release(); done.release();
      conn.release(); n2e.release(); nbr.release(); node_row.release(); color_elems.release(); ent_row.release();
      n2e_ptr.release(); nbr_ptr.release(); diag.release(); posmap.release(); tab.release();
      tp_built = false; tp_chunk = 0; tp_row_ptr.clear(); tp_rows.release(); tp_ea.release(); tp_ea_ptr.release(); tp_first.release(); tp_K.release(); tp_f.release();
    }
  } gplan;
  // implicit-connectivity sweep of ijk-numbered (build_cube) HEX8 meshes (sweep_hex8_struct.cu)
  struct StructPlan {
    bool tried = false, ok = false;
    int32_t nx = 0, ny = 0, nz = 0, oz0 = 0, oz1 = 0, n_units = 0, n_ctas = 0;
    int64_t n_visits = 0;
    DevBuf<StructUnit> units;
    DevBuf<int32_t> arr_off;
    DevBuf<int64_t> nbr_ptr;
    DevBuf<StructUnit> s_units;  // units of the structured sensitivity sweep (sweep_hex8_sens_struct.cu: larger tiles)
    int32_t s_n_units = 0;
    DevBuf<double> zeros;     // source of the bulk copies (TMA) that clear the lock-step kernel's accumulators
    // units of the register-accumulating march (sweep_hex8_march.cu; 3-dof system)
    DevBuf<StructUnit> m_units;
    int32_t m_n_units = 0, m_n_ctas = 0;
    int64_t m_n_visits = 0;
    bool march = false;       // the march kernel serves nv = 3
    bool last_march = false;  // ... and ran the last sweep
    bool slots = false;       // the slot kernel (sweep_hex8_slots.cu) serves nv = 3 without fused sensitivities
    // the same units grouped by z-segment (ascending z, cost-sorted inside a group): a sweep with host buffers runs group by
    // group so that the copies of X (in) and R (out) overlap the kernels
    DevBuf<StructUnit> g_units;
    std::vector<int32_t> g_ptr, g_z0, g_z1;
    bool last_slots = false;
    void reset() { tried = ok = false; units.release(); arr_off.release(); nbr_ptr.release(); zeros.release(); s_units.release(); m_units.release(); g_units.release(); g_ptr.clear(); g_z0.clear(); g_z1.clear(); }
  } splan;
  cudaStream_t copy_stream = nullptr;  // second stream + events of the segmented host-buffer sweep (created on first use)
  cudaEvent_t seg_ev[66] = {};  // [g] planes of X for segment g are in, [32 + g] segment g is swept, [64], [65] fork / join
  bool seg_ready = false;
  bool async_device = false;  // mastb200_set_async: sweeps whose buffers all live on the device return without a stream sync
  int max_optin = 232448;  // sm_100: 227 KB opt-in shared memory per CTA
  DevBuf<int32_t> d_tile_counter, d_owned_nodes;
  DevBuf<uint8_t> d_cmask;
  // device: mesh + state + staging for host pointers
  DevBuf<double> d_xyz, d_X, d_adj, d_rho, d_T, d_R, d_J, d_G, d_dq, d_partial, d_scalar, d_scratch, d_st[4];
  DevBuf<uint8_t> d_elem_mask;
  bool rho_set = false, T_nodal = false, thermal = false;
  const double* rho_ext = nullptr;  // caller-owned device density
  const double* T_ext = nullptr;
  double T_const = 0., T_ref = 0.;
  const double* last_R = nullptr;   // device pointer of the last residual
  // face loads: raw (element, side, pressure) entries and the per-owned-node gather lists built from them
  std::vector<int32_t> sp_elem, sp_side;
  std::vector<double> sp_p;
  std::vector<int32_t> fl_node, fl_ptr, fl_face_nodes;
  std::vector<uint8_t> fl_k;
  std::vector<double> fl_p;
  DevBuf<int32_t> d_fl_node, d_fl_ptr, d_fl_face_nodes;
  DevBuf<uint8_t> d_fl_k;
  DevBuf<double> d_fl_p;
  // point loads: loads of the OWNED nodes, summed per node in the order given (nonlinear_implicit_assembly.cpp:190-247)
  std::vector<int32_t> pl_node;
  std::vector<double> pl_load;
  DevBuf<int32_t> d_pl_node, d_dvn;
  DevBuf<double> d_pl_load, d_dfield, d_dvw, d_vec2;
  std::vector<int32_t> pl_raw_node;
  std::vector<double> pl_raw_load;
  // heat side loads (CSR over elements)
  DevBuf<int32_t> d_hs_ptr, d_hs_side;
  DevBuf<double> d_hs_val;
  bool have_heat_sides = false;
  // build_square numbering of a QUAD4 heat mesh recognised (quad4_heat_struct_kernel)
  bool q4s_tried = false, q4s_ok = false;
  int32_t q4s_nx = 0, q4s_ny = 0;
  bool last_q4s = false;
  double rad_sbc = 0., rad_T0 = 0.;  // SURFACE_RADIATION_HEAT_FLUX parameters of the heat side loads
  // dv map staging
  DevBuf<int64_t> d_dv_ptr;
  DevBuf<int32_t> d_dv_node;
  DevBuf<double> d_dv_w, d_sens;

  // communicator (solver.cu): NCCL through dlopen, ghost exchange lists per neighbouring rank
  void* nccl_comm = nullptr;       // ncclComm_t
  bool nccl_owned = false;         // created by mastb200_comm_init (destroyed with the context)
  int comm_rank = 0, comm_world = 1;
  struct Neighbour {
    int rank = 0;
    int64_t n_send = 0, n_recv = 0;
    int64_t send_first = -1, recv_first = -1;  // >= 0: the list is the contiguous node range [first, first + n)
    DevBuf<int32_t> d_send, d_recv;            // node lists (only when not contiguous)
    DevBuf<double> send_buf, recv_buf;
  };
  std::vector<Neighbour*> neighbours;
  bool auto_localize = false;      // sweeps refresh the ghost values of X themselves (device X, in place)
  // density filter and material volume (filter.cu)
  DevBuf<int64_t> f_ptr;
  DevBuf<int32_t> f_col;
  DevBuf<double> f_raw, f_sum, nodal_vol;
  DevBuf<uint8_t> f_dv;
  bool have_filter = false, f_has_dv = false, have_nodal_vol = false;
  int64_t f_nnz = 0;
  double f_radius = 0.;
  // linear solve (solver.cu)
  DevBuf<double> d_minv, d_pr, d_pz, d_pp, d_pAp, d_px, d_ppart, d_pscal;

  int64_t launches = 0;

  int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
  int cuda_fail(cudaError_t e, const char* what) {
    return fail(MASTB200_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  }
};

#define CU(call)                                                   \
  do {                                                             \
    cudaError_t e_ = (call);                                       \
    if (e_ != cudaSuccess) return ctx->cuda_fail(e_, #call);       \
  } while (0)

// ---- helpers shared between the translation units (context.cu) ----
// device view of an input array: device pointers pass through, host pointers are copied into `stage`
int resolve_in(mastb200_context* ctx, const double* ptr, size_t n, DevBuf<double>& stage, const double** out);
int resolve_out(mastb200_context* ctx, double* ptr, size_t n, DevBuf<double>& stage, double** out);
int finish_out(mastb200_context* ctx, double* ptr, size_t n, const double* dev);
int check_ready(mastb200_context* ctx);
int ensure_generic_plan(mastb200_context* ctx);
int64_t plan_nnz(const mastb200_context* ctx);
int localize_if_needed(mastb200_context* ctx, const double* dV, int width);
// ghost refresh of a device vector on the context stream (no-op without a configured exchange); solver.cu
int comm_exchange(mastb200_context* ctx, double* dV, int width);
int comm_all_reduce(mastb200_context* ctx, double* d_vals, int n);
void comm_release(mastb200_context* ctx);
