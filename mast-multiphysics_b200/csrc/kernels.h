// kernels.h -- launchers of the sm_100a kernels (implemented in sweep_*.cu, loads.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "device_types.h"
#include "struct_layout.h"

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

#ifdef __CUDACC__
// rho^penalty of the SIMP law (examples/structural/example_6/example_6.cpp:100-137): the usual small integer penalties by
// multiplication (a few ulp from pow, which costs ~150 instructions per quadrature point)
__device__ __forceinline__ double simp_pow(double rho, double penalty) {
  if (penalty == 3.) return rho * rho * rho;
  if (penalty == 4.) { const double r2 = rho * rho; return r2 * r2; }
  if (penalty == 2.) return rho * rho;
  if (penalty == 1.) return rho;
  return pow(rho, penalty);
}
#endif

// ---- HEX8 sweep for ijk-numbered (build_cube) meshes with implicit connectivity (sweep_hex8_struct.cu) --------------
struct StructUnit {          // a column of owned nodes over a z-range, with the elements that touch it
  int16_t e0x, nex;          // first visited element / number of visited elements in x (<= 8)
  int16_t o0x, nox;          // first owned node / number of owned nodes in x (<= 7)
  int16_t e0y, ney, o0y, noy;
  int32_t z0, z1;            // owned node planes [z0, z1)
};
struct StructParams {
  const StructUnit* units = nullptr;
  int32_t n_units = 0;
  int32_t* counter = nullptr;            // dynamic unit scheduler (zeroed before every launch)
  int32_t nx = 0, ny = 0, nz = 0;        // elements per direction of the (local) box
  int32_t oz0 = 0;                       // first owned node plane: CSR rows start there
  int32_t nv = 3;                        // 3 (compact) or 6 (MAST's structural system: rotational rows kept, see flush)
  const int64_t* nbr_ptr = nullptr;      // node-level CSR row pointer of the owned rows
  const int32_t* arr_off = nullptr;      // [123] offsets of the pair arrays within an IN (42) / CR (81) region
  int32_t in1_off = 0, cr0_off = 0, cr1_off = 0;
  const double* zeros = nullptr;         // >= kStructZeroDoubles zeros, 16-byte aligned (lock-step kernel: bulk-copy clear), or null
  const double* xyz = nullptr;
  const double* X = nullptr;
  const uint8_t* cmask = nullptr;
  const double* rho = nullptr;           // nodal density (SIMP modulus) or null
  const double* adj = nullptr;           // adjoint [n_nodes*3] (== X for the compliance) when G is requested
  double* G = nullptr;                   // [n_nodes] density sensitivities fused into the sweep, or null
  double* J = nullptr;
  double* R = nullptr;
  int32_t j_aligned32 = 0;               // J is 32-byte aligned: the nv = 6 flush uses 16- and 32-byte stores
  int32_t gl = 0;                        // slot kernel: Green-Lagrange strain, tangent only (R must be null)
  double E0 = 0., lam1 = 0., mu1 = 0., rho_min = 0., penalty = 1.;
};
constexpr int kStructZeroDoubles = 81 * 81 + 8;
size_t hex8_struct_smem_bytes();
size_t hex8_struct_sens_smem_bytes();   // of the variant with fused density sensitivities (0: not built for this CTA shape)
int hex8_struct_threads();
cudaError_t launch_hex8_struct(const StructParams& p, int n_ctas, cudaStream_t stream);
// register-accumulating column march for the 3-dof system (sweep_hex8_march.cu): units of <= 16 x 4 owned node columns
constexpr int kMarchTileX = 16, kMarchTileY = 4;
size_t hex8_march_smem_bytes();
int hex8_march_threads();
cudaError_t launch_hex8_march(const StructParams& p, int n_ctas, cudaStream_t stream);
// density sensitivities alone on the same units (sweep_hex8_sens_struct.cu; 3-dof system, SIMP, linear strain)
constexpr int kSensTileX = 15, kSensTileY = 11;   // owned node columns per unit (16 x 12 elements)
size_t hex8_sens_struct_smem_bytes();
cudaError_t launch_hex8_sens_struct(const StructParams& p, int sm_count, cudaStream_t stream);
// element results in shared-memory slots, row-owner gather straight to the CSR (sweep_hex8_slots.cu; 3-dof system, same units
// as the lock-step kernel)
size_t hex8_slots_smem_bytes();
int hex8_slots_threads();
cudaError_t launch_hex8_slots(const StructParams& p, int n_ctas, cudaStream_t stream);

// ---- row-owner gather kernels for small elements (sweep_quad4_heat.cu) ---------------------------------------------
constexpr int kGatherMaxNbr = 24;      // neighbours per row the gather kernels keep in shared memory
struct GatherParams {
  int64_t n_owned = 0;
  const int32_t* owned_nodes = nullptr;  // [n_owned]
  const int64_t* nbr_ptr = nullptr;      // node-level CSR of the owned rows
  const int32_t* nbr = nullptr;
  const int64_t* n2e_ptr = nullptr;      // node -> adjacent elements (ascending id)
  const int32_t* n2e = nullptr;
  const int32_t* conn = nullptr;         // [n_elem][nn]
  const uint8_t* diag_elem = nullptr;    // [n_elem] or null
  const uint8_t* cmask = nullptr;        // [n_nodes] or null
  const double* xyz = nullptr;
  const double* X = nullptr;
  double* J = nullptr;
  double* R = nullptr;
  double k0 = 0., k_beta = 0., thickness = 1., heat_source = 0.;
  int32_t k_of_T = 0;
  // heat side loads per element (CSR over elements): side, HEAT_FLUX value, CONVECTION_HEAT_FLUX coefficient, T_ambient
  const int32_t* sl_ptr = nullptr;       // [n_elem+1] or null (no side loads)
  const int32_t* sl_side = nullptr;
  const double* sl_val = nullptr;        // [n_entries][4]: flux, h_c, T_amb, emissivity
  double rad_sbc = 0., rad_T0 = 0.;      // Stefan-Boltzmann constant, reference-zero temperature (radiation)
};
cudaError_t launch_quad4_heat(const GatherParams& p, cudaStream_t stream);
// the same sweep for build_square-numbered meshes (node i + j(nx+1), element i + j nx, every node owned, no
// diagonal elements): implicit connectivity, every element evaluated once per CTA tile (quad4_heat_struct_kernel)
cudaError_t launch_quad4_heat_struct(const GatherParams& p, int32_t nx, int32_t ny, cudaStream_t stream);

// ---- HEX27 coloured DMMA sweep (sweep_hex27.cu) -----------------------------------------------------------------------
struct Hex27Params {
  const int32_t* elems = nullptr;        // element ids in processing order (of one colour, or the whole ordered sweep)
  int32_t n_elems = 0;
  // ordered single-launch sweep: predecessor lists per position, done flags per element, the value that marks 'done'
  const int32_t* pred_ptr = nullptr;
  const int32_t* pred = nullptr;
  int32_t* done = nullptr;
  int32_t epoch = 0;
  int32_t pos_flagged = 0;               // posmap entries carry the first-writer flag in bit 7 (positions are 7 bits)
  int32_t first_writer = 0;              // use the flag: the first element to touch a block stores (no memset of J)
  const int32_t* conn = nullptr;         // [n_elem][27]
  const uint8_t* posmap = nullptr;       // [n_elem][27][27] position of node b in the neighbour list of node a
  const int32_t* node_row = nullptr;     // [n_nodes] owned-row index or -1
  const int64_t* nbr_ptr = nullptr;
  const double* tab = nullptr;           // N[27][27] (q,b), dN[27][27][3] (q,b,r), w[27]
  const uint8_t* diag_elem = nullptr;
  const uint8_t* cmask = nullptr;
  const double* xyz = nullptr;
  const double* X = nullptr;
  const double* rho = nullptr;
  const double* temperature = nullptr;
  double* J = nullptr;
  double* R = nullptr;
  double E0 = 0., lam1 = 0., mu1 = 0., rho_min = 0., penalty = 1., alpha_k = 0., T_const = 0., T_ref = 0.;
  int32_t nv = 3, thermal = 0;
  // two-pass assembly (3-dof system): pass 1 = this kernel over a chunk of consecutive element ids [elem_first, elem_first +
  // n_elems) with elems == nullptr, K_e blocks and f_e STORED element-major (no colours, no read-modify-write); pass 2 = the
  // row gather below
  int32_t elem_first = 0;
  double* tmp_K = nullptr;               // [n_elems][27][27][9] block (a, b) of the chunk's element, constrained as the element is
  double* tmp_f = nullptr;               // [n_elems][27][3]
};
size_t hex27_sweep_smem_bytes();
cudaError_t launch_hex27_sweep(const Hex27Params& p, int sm_count, cudaStream_t stream);
// pass 2: one warp per owned row touched by the chunk sums the blocks of the row's elements inside the chunk (ascending id)
// in shared memory and writes the node's three CSR rows as whole runs; rows that earlier chunks have written are added to
struct Hex27GatherParams {
  int64_t n_rows = 0;                    // rows of this chunk
  const int32_t* rows = nullptr;         // owned-row indices
  const int64_t* ea_ptr = nullptr;       // [n_rows + 1] range of the row's (element, local node) list
  const int32_t* ea = nullptr;           // (element - elem_first) * 27 + local node, ascending element id
  const uint8_t* row_first = nullptr;    // [n_rows] 1: no element of the row lies before the chunk (store), 0: add
  const int32_t* owned_nodes = nullptr;
  const int64_t* nbr_ptr = nullptr;
  const int64_t* n2e_ptr = nullptr;
  const int32_t* n2e = nullptr;
  const int32_t* conn = nullptr;
  const uint8_t* posmap = nullptr;
  int32_t pos_flagged = 0;
  int32_t elem_first = 0, n_elems = 0;
  const double* tmp_K = nullptr;
  const double* tmp_f = nullptr;
  double* J = nullptr;
  double* R = nullptr;
};
cudaError_t launch_hex27_row_gather(const Hex27GatherParams& p, cudaStream_t stream);
// pass 1 with the symmetric half of the tile pairs (630 instead of 1008 DMMA per element; 3-dof system, no diagonal elements)
cudaError_t launch_hex27_sym_pass1(const Hex27Params& p, int sm_count, cudaStream_t stream);


// ---- QUAD4 Mindlin plate in the xy-plane (sweep_quad4_plate.cu): one thread per (owned node, neighbour) block ------
struct PlateParams {
  int64_t n_owned = 0, n_entries = 0;    // n_entries = nbr_ptr[n_owned]
  const int32_t* owned_nodes = nullptr;
  const int64_t* nbr_ptr = nullptr;
  const int32_t* nbr = nullptr;
  const int32_t* ent_row = nullptr;      // [n_entries] owned row of every neighbour entry
  const int64_t* n2e_ptr = nullptr;
  const int32_t* n2e = nullptr;
  const int32_t* conn = nullptr;         // [n_elem][4]
  const uint8_t* diag_elem = nullptr;
  const uint8_t* cmask = nullptr;
  const double* xyz = nullptr;
  const double* X = nullptr;             // [n_nodes*6]
  double* J = nullptr;
  double* R = nullptr;
  double c11 = 0., c12 = 0., c33 = 0.;   // plane-stress / plane-strain material matrix C
  double shear_G = 0., thickness = 1., offset = 0., kappa = 5. / 6.;
  int32_t local_frames = 0;              // 1: some element's nodes do not share one z -> per-element local frames (T K T^T)
  int32_t nonlinear = 0;                 // 1: NONLINEAR_STRAIN -- Green-Lagrange membrane strain + von Karman strain
};
cudaError_t launch_quad4_plate(const PlateParams& p, cudaStream_t stream);

// ---- HEX8 row-owner gather with Green-Lagrange (or linear) strain (sweep_hex8_gather.cu) ----------------------------
struct SolidGatherParams {
  int64_t n_owned = 0;
  const int32_t* owned_nodes = nullptr;
  const int64_t* nbr_ptr = nullptr;      // node-level CSR of the owned rows (at most 32 neighbours per row)
  const int32_t* nbr = nullptr;
  const int64_t* n2e_ptr = nullptr;
  const int32_t* n2e = nullptr;
  const int32_t* conn = nullptr;         // [n_elem][8] (HEX27: [n_elem][27])
  const double* tab = nullptr;           // HEX27: N[27][27], dN[27][27][3], w[27]
  const uint8_t* diag_elem = nullptr;
  const uint8_t* cmask = nullptr;
  const double* xyz = nullptr;
  const double* X = nullptr;
  const double* rho = nullptr;
  const double* temperature = nullptr;
  double* J = nullptr;
  double* R = nullptr;
  double E0 = 0., lam1 = 0., mu1 = 0., rho_min = 0., penalty = 1., alpha_k = 0., T_const = 0., T_ref = 0.;
  int32_t nv = 3, thermal = 0, nonlinear = 0;
};
cudaError_t launch_hex8_gather(const SolidGatherParams& p, cudaStream_t stream);
cudaError_t launch_hex27_gather(const SolidGatherParams& p, cudaStream_t stream);  // sweep_hex27_gather.cu (<= 128 neighbours per row)

// ---- stress output: per-qp stress/strain/von Mises, p-norm aggregate and its state derivative (stress_output.cu) ----
struct StressParams {
  int64_t n_elem = 0, n_owned = 0;
  int32_t nn = 8, nq = 8, nv = 3, nonlinear = 0;
  int32_t scratch_nodal = 0;             // scratch holds [element][node] sums instead of [element][qp] point values
  const int32_t* conn = nullptr;
  const double* tab = nullptr;           // N[nq][nn], dN[nq][nn][3], w[nq]
  const double* xyz = nullptr;
  const double* X = nullptr;
  const double* rho = nullptr;
  const double* temperature = nullptr;   // nodal temperature or null (T_const): thermal stress of the matrix-free force, mode 4
  double alpha_k = 0., T_const = 0., T_ref = 0.;
  const uint8_t* elem_mask = nullptr;    // participating elements or null (all)
  const uint8_t* diag_elem = nullptr;
  const uint8_t* cmask = nullptr;
  const int32_t* owned_nodes = nullptr;
  const int64_t* n2e_ptr = nullptr;
  const int32_t* n2e = nullptr;
  double E0 = 0., lam1 = 0., mu1 = 0., rho_min = 0., penalty = 1.;
  double p_norm = 2., rho_exp = 0., sigma0 = 1., sum_num = 0., sum_den = 0.;
  double* out_stress = nullptr;          // [n_elem*nq][6] or null
  double* out_strain = nullptr;
  double* out_vm = nullptr;              // [n_elem*nq]
  double* out_jxw = nullptr;
  double* partial = nullptr;             // [2*blocks] block sums of (numerator, denominator) or null
  double* scratch = nullptr;             // [n_elem*nq][9] (derivative passes)
  double* dq_dX = nullptr;               // [n_nodes*nv]
};
int stress_blocks(int64_t n_elem, int nq);
cudaError_t launch_stress_qp(const StressParams& p, cudaStream_t stream);
cudaError_t launch_reduce_pairs(const double* partial, int n, double* out, cudaStream_t stream);
cudaError_t launch_stress_derivative(const StressParams& p, cudaStream_t stream);  // two launches
// matrix-free internal force with a per-point modulus into p.dq_dX (two launches); scratch: [n_elem*nq][9] doubles.
// mode 0 E(rho), 1 density design variable (dfield = nodal filter column), 2 d/dE0, 3 lam1 / mu1 are d/dnu,
// 4 thermal stress E(rho) alpha_k (T - T_ref) I instead of D eps (the TEMPERATURE volume load as a force vector)
cudaError_t launch_internal_force(const StressParams& p, int mode, const double* dfield, cudaStream_t stream);
// density sensitivities alone (no R/J): G[n] = adj . dR/d(rho_n); scratch: [n_elem*nq] doubles (two launches)
cudaError_t launch_density_sensitivities(const StressParams& p, const double* adj, double* G, cudaStream_t stream);
// dq/d(rho_n) of the stress aggregate at fixed state (two launches); scratch: [n_elem*nq] doubles
cudaError_t launch_stress_density_derivative(const StressParams& p, double* G, cudaStream_t stream);

// SURFACE_PRESSURE side loads gathered per loaded node (loads.cu)
struct FaceLoadPlan {
  int32_t n_loaded_nodes = 0;
  const int32_t* node = nullptr;      // [n_loaded_nodes]
  const int32_t* ptr = nullptr;       // [n_loaded_nodes+1]
  int32_t nfn = 4;                    // nodes per face: 4 (HEX8 face, 2x2 Gauss) or 9 (HEX27 face, 3x3 Gauss)
  const int32_t* face_nodes = nullptr;  // [n_entries][nfn] nodes of the loaded face (libMesh side ordering)
  const uint8_t* face_k = nullptr;    // [n_entries] which face node this entry's node is
  const double* pressure = nullptr;   // [n_entries]
  int32_t node_lo = 0, node_hi = 0x7fffffff;  // only loaded nodes in [node_lo, node_hi) are applied (segmented sweeps)
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

// point loads: R[nv*node[t]+i] -= load[t*nv+i] on unconstrained dofs; field[node[t]] = w[t]; sum over owned nodes of a.b
cudaError_t launch_sub_vector(int64_t n, double* a, const double* b, cudaStream_t stream);  // a -= b
cudaError_t launch_point_loads(int n, const int32_t* node, const double* load, const uint8_t* cmask, int nv, double* R,
                               cudaStream_t stream, int node_lo = 0, int node_hi = 0x7fffffff);
cudaError_t launch_scatter_field(int64_t n, const int32_t* node, const double* w, double* field, cudaStream_t stream);
cudaError_t launch_dot_owned(const double* a, const double* b, const int32_t* owned_nodes, int64_t n_owned, int nv,
                             double* partial, int n_blocks, cudaStream_t stream);

// ---- device linear solve (solve.cu): block-Jacobi preconditioned CG on the node-block CSR of the 3-dof system ---------
int pcg_reduction_blocks();
cudaError_t launch_bj_setup(int64_t n_owned, const int32_t* owned, const int64_t* nbr_ptr, const int32_t* nbr, const double* J,
                            double* minv, cudaStream_t s);
cudaError_t launch_spmv_node(int64_t n_owned, const int32_t* owned, const int64_t* nbr_ptr, const int32_t* nbr, const double* J,
                             const double* x, double* y, double* partial, int sm_count, cudaStream_t s);
cudaError_t launch_reduce_multi(const double* partial, int n, int n_out, int stride, double* out, cudaStream_t s);
cudaError_t launch_pcg_alpha(double* scal, cudaStream_t s);
cudaError_t launch_pcg_beta(double* scal, cudaStream_t s);
cudaError_t launch_pcg_init(int64_t n_owned, const int32_t* owned, const double* rhs, const double* minv, double* x, double* r,
                            double* z, double* p, double* partial, cudaStream_t s);
cudaError_t launch_pcg_update(int64_t n_owned, const int32_t* owned, const double* scal, const double* minv, const double* p,
                              const double* Ap, double* x, double* r, double* z, double* partial, cudaStream_t s);
cudaError_t launch_pcg_direction(int64_t n_owned, const int32_t* owned, const double* scal, const double* z, double* p, cudaStream_t s);
cudaError_t launch_near_null(int64_t n_nodes, int nv, const double* xyz, double* modes, cudaStream_t s);

// G^T W : sens[i] = sum_k w[k] * G[node[k]]
cudaError_t launch_dv_map_transpose(const double* G, int64_t n_dv, const int64_t* dv_ptr, const int32_t* dv_node,
                                    const double* dv_weight, double* sens, cudaStream_t stream);

}  // namespace mastb200
