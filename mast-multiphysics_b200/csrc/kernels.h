// kernels.h -- launchers of the sm_100a kernels (implemented in sweep_*.cu, loads.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "device_types.h"

namespace mastb200 {

#ifndef MASTB200_CTAS
#define MASTB200_CTAS 1                // resident CTAs per SM of the sweep kernel (tuning knob; shared memory is split)
#endif
constexpr int kCtasPerSm = MASTB200_CTAS;
constexpr int kMaxSlots = kCtasPerSm == 1 ? 448 : 288;           // node state slots staged in shared memory per strip
constexpr int kFlushBufBytes = kCtasPerSm == 1 ? 8192 : 6144;    // flush records of one step staged in shared memory (x2)

// HEX8 linear-elastic sweep (sweep_hex8.cu).  aux: adjoint / density / temperature state and per-node extras staged
size_t hex8_sweep_smem_bytes(bool aux);
int hex8_sweep_arena_blocks(bool aux);  // pair blocks of the shared-memory arena (compile-time layout)
int hex8_sweep_threads();
int hex8_sweep_workers();  // compute warps = rows of the plan's per-step visit grid
cudaError_t launch_hex8_sweep(const SweepParams& p, int n_ctas, cudaStream_t stream);

// SURFACE_PRESSURE side loads gathered per loaded node (loads.cu)
struct FaceLoadPlan {
  int32_t n_loaded_nodes = 0;
  const int32_t* node = nullptr;      // [n_loaded_nodes]
  const int32_t* ptr = nullptr;       // [n_loaded_nodes+1]
  const int32_t* face_nodes = nullptr;  // [n_entries][4] nodes of the loaded face (libMesh side ordering)
  const uint8_t* face_k = nullptr;    // [n_entries] which face node this entry's node is
  const double* pressure = nullptr;   // [n_entries]
};
// R[nv*n+i] -= F_i (unconstrained dofs) if R; dq[nv*n+i] = F_i (unconstrained) if dq;
// partial[block] = sum_n F_n . u_n if partial (then reduce_partials).
cudaError_t launch_face_loads(const FaceLoadPlan& fp, const double* xyz, const double* X, const uint8_t* cmask, int nv,
                              double* R, double* dq, double* partial, int* n_blocks, cudaStream_t stream);
int face_loads_blocks(int n_loaded_nodes);

// deterministic reductions: out[0] = sum(partial[0..n)) ; sum of squares of R over owned nodes
cudaError_t launch_reduce_partials(const double* partial, int n, double* out, cudaStream_t stream);
cudaError_t launch_sumsq_owned(const double* R, const int32_t* owned_nodes, int64_t n_owned, int nv, double* partial,
                               int n_blocks, cudaStream_t stream);

// G^T W : sens[i] = sum_k w[k] * G[node[k]]
cudaError_t launch_dv_map_transpose(const double* G, int64_t n_dv, const int64_t* dv_ptr, const int32_t* dv_node,
                                    const double* dv_weight, double* sens, cudaStream_t stream);

}  // namespace mastb200
