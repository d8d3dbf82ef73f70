/* mast_b200.h -- C-ABI of the B200-native element-assembly + sensitivity path.
 *
 * This is the drop-in boundary: MAST's host classes (MAST::NonlinearImplicitAssembly & co., mirrored
 * in mast-multiphysics_b200/host/) pull plain arrays out of libMesh/PETSc objects once per mesh and
 * call these entry points once per sweep.  Plain pointers and sizes only, int status returns, no
 * exceptions cross the boundary.  One context per rank/GPU; calls on one context are stream-ordered
 * and not thread-safe (the reference is single-threaded per MPI rank and not re-entrant,
 * /root/reference/src/base/assembly_elem_operation.cpp:48-78).
 *
 * Every data pointer marked [h|d] may be a host pointer or a CUDA device pointer; the library detects
 * which (cudaPointerGetAttributes).  Host buffers are copied inside the call; device buffers are used
 * in place.  There is no CPU fallback: every compute entry point fails with MASTB200_ERR_CUDA if no
 * sm_100 device is usable.
 *
 * Reference interfaces replaced (paths relative to /root/reference):
 *   mastb200_residual_and_jacobian   <- MAST::NonlinearImplicitAssembly::residual_and_jacobian
 *                                       src/base/nonlinear_implicit_assembly.cpp:64-266 (assembly_base.h:226-232)
 *   mastb200_sensitivity_sweep       <- MAST::NonlinearImplicitAssembly::sensitivity_assemble
 *                                       src/base/nonlinear_implicit_assembly.cpp:515-686, folded with
 *                                       AssemblyBase::calculate_output_adjoint_sensitivity_multiple_parameters_no_direct
 *                                       src/base/assembly_base.cpp:594-639 (all DVs in one sweep)
 *   mastb200_compliance / _derivative<- AssemblyBase::calculate_output / calculate_output_derivative with
 *                                       MAST::ComplianceOutput, src/base/assembly_base.cpp:253-430,
 *                                       src/elasticity/compliance_output.cpp:71-102, 265-296
 *   mastb200_apply_dv_map_transpose  <- the per-DV loop of assembly_base.cpp:618-638 with the filter columns of
 *                                       src/level_set/filter_base.cpp:59-97 as the nodal-density sensitivities
 *   mastb200_res_l2_norm             <- NonlinearImplicitAssembly::res_l2_norm (nonlinear_implicit_assembly.cpp:258-264)
 *   mastb200_build_pattern/_get_*    <- libMesh DofMap sparsity + PETSc AIJ preallocation consumed at
 *                                       nonlinear_implicit_assembly.cpp:182-183
 */
#ifndef MAST_B200_H
#define MAST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MASTB200_OK 0
#define MASTB200_ERR_ARG 1      /* bad argument / call order */
#define MASTB200_ERR_CUDA 2     /* CUDA runtime error or no usable device */
#define MASTB200_ERR_UNSUPPORTED 3 /* valid MAST configuration not (yet) on the device path */
#define MASTB200_ERR_PLAN 4     /* mesh cannot be tiled within shared-memory limits */

/* element types (libMesh node ordering) */
#define MASTB200_HEX8 8
#define MASTB200_HEX27 27
#define MASTB200_QUAD4 4 /* nv = 1 heat conduction (xy-plane), nv = 6 Mindlin plate (anywhere in space) */

typedef struct mastb200_context mastb200_context;

/* Isotropic property card + strain type.  E(x) = E0, or the SIMP law of
 * examples/structural/example_6/example_6.cpp:100-137 when a nodal density field is set. */
typedef struct mastb200_material {
  double E0, nu, alpha;
  double rho_min, penalty;     /* SIMP: E = E0*(rho_min + (1-rho_min)*rho^penalty) */
  int32_t nonlinear_strain;    /* 0 LINEAR_STRAIN, 1 NONLINEAR_STRAIN (Green-Lagrange) */
  /* heat conduction card (QUAD4): k(T) = k0*(1+k_beta*T) if k_of_T, section thickness */
  double k0, k_beta, thickness;
  int32_t k_of_T;
  double heat_source;
  /* 2D plate section (QUAD4 with nv = 6: StructuralElement2D, Mindlin bending): thickness above, mid-plane offset,
   * shear correction factor, plane_strain = 0 (plane stress) | 1
   * (solid_2d_section_element_property_card.cpp:395-510, isotropic_material_property_card.cpp:395-426) */
  double offset, kappa;
  int32_t plane_strain;
} mastb200_material;

/* ---- lifetime ------------------------------------------------------------------------------ */
int mastb200_create(int device, mastb200_context** ctx);
int mastb200_destroy(mastb200_context* ctx);
const char* mastb200_last_error(const mastb200_context* ctx); /* never NULL; ctx may be NULL for create errors */
const char* mastb200_version(void);

/* ---- problem definition (host pointers; copied) -------------------------------------------- */
/* conn: [n_elem][nn] node ids; xyz: [n_nodes][3]; diag_elem: [n_elem] or NULL
 * (AssemblyBase::diagonal_elem_subdomain_id members). */
int mastb200_set_mesh(mastb200_context* ctx, int32_t elem_type, int64_t n_elem, int64_t n_nodes,
                      const double* xyz, const int32_t* conn, const uint8_t* diag_elem);
/* dof(node,var) = nv*node + var (one libMesh variable group).  nv = 6 (MAST structural system), 3 (compact)
 * or 1 (heat).  constrained: [n_nodes*nv] homogeneous-Dirichlet flags or NULL.
 * owned: [n_nodes] or NULL (all owned): rows are produced only for owned nodes (the rank's row block);
 * the other nodes are ghosts whose values arrive in X (AssemblyBase::build_localized_vector). */
int mastb200_set_dof_map(mastb200_context* ctx, int32_t nv, const uint8_t* constrained, const uint8_t* owned);
int mastb200_set_material(mastb200_context* ctx, const mastb200_material* mat);
/* SURFACE_PRESSURE side loads on (element, libMesh side) pairs; n = 0 clears. */
int mastb200_set_side_pressure(mastb200_context* ctx, int64_t n, const int32_t* elem, const int32_t* side,
                               const double* pressure);
/* Heat-conduction side loads on (element, libMesh side) pairs of a QUAD4 mesh (nv = 1): HEAT_FLUX value and
 * CONVECTION_HEAT_FLUX coefficient / ambient temperature per entry (conv_coeff, T_amb may be NULL); n = 0 clears.
 *   <- HeatConductionElementBase::surface_flux_residual, ::surface_convection_residual,
 *      src/heat_conduction/heat_conduction_elem_base.cpp:876-914, 1100-1155 (dispatch :428-483) */
int mastb200_set_heat_side_loads(mastb200_context* ctx, int64_t n, const int32_t* elem, const int32_t* side,
                                 const double* heat_flux, const double* conv_coeff, const double* T_amb);
/* The same with SURFACE_RADIATION_HEAT_FLUX: emissivity per entry (NULL: none), Stefan-Boltzmann constant and the
 * reference-zero temperature of the boundary condition; the entry's T_amb is the radiation's ambient temperature.
 * f += JxW N th sb eps ((T + T0)^4 - (T_amb + T0)^4), jac += JxW N N^T th sb eps 4 (T + T0)^3.
 *   <- HeatConductionElementBase::surface_radiation_residual, heat_conduction_elem_base.cpp:1404-1458 */
int mastb200_set_heat_side_loads_ex(mastb200_context* ctx, int64_t n, const int32_t* elem, const int32_t* side,
                                    const double* heat_flux, const double* conv_coeff, const double* T_amb,
                                    const double* emissivity, double sb_const, double T_ref_zero);
/* Point loads (discipline.point_loads()): load[n][nv] on node[n]; R[dofs of the node] -= load on the unconstrained dofs
 * of the nodes this rank owns (loads of one node are summed in the order given); n = 0 clears.  Call after set_dof_map.
 *   <- NonlinearImplicitAssembly::residual_and_jacobian, src/base/nonlinear_implicit_assembly.cpp:190-247 */
int mastb200_set_point_loads(mastb200_context* ctx, int64_t n, const int32_t* node, const double* load);
/* TEMPERATURE volume load: nodal field [h|d] (or NULL => constant), reference temperature. enable=0 clears. */
int mastb200_set_temperature(mastb200_context* ctx, int32_t enable, const double* T_nodal, double T_const,
                             double T_ref);
/* nodal density field [h|d] for the SIMP law (NULL => homogeneous E0). Cheap: call every design iteration. */
int mastb200_set_density(mastb200_context* ctx, const double* rho_nodal);

/* Builds the sparsity pattern (every dof pair sharing an element, columns sorted) and the device
 * assembly plan (node tiles, element visit lists, element->nnz slot maps).  Call after mesh/dof map. */
int mastb200_build_pattern(mastb200_context* ctx, int64_t* n_rows, int64_t* nnz);
/* Copies the CSR pattern of the owned rows out (host pointers): row_ptr[n_rows+1], col_idx[nnz]. */
int mastb200_get_pattern(const mastb200_context* ctx, int64_t* row_ptr, int32_t* col_idx);
/* The same for a sample of rows (the full col_idx of a 200^3 mesh is 7.8 GB): rows[n] are CSR row numbers;
 * row_begin[n] = position of the row's first value in J_vals, row_len[n] = its number of values,
 * col_idx (may be NULL) = the rows' column indices concatenated in the order of rows[]. */
int mastb200_get_pattern_rows(const mastb200_context* ctx, int64_t n, const int64_t* rows, int64_t* row_begin,
                              int32_t* row_len, int32_t* col_idx);
/* dof id of each CSR row: row_dof[n_rows] (owned nodes ascending, nv rows per node). */
int mastb200_get_row_dofs(const mastb200_context* ctx, int64_t* row_dof);

/* ---- sweeps --------------------------------------------------------------------------------- */
/* X: [n_nodes*nv] localized solution (owned + ghost values) [h|d].
 * R: [n_nodes*nv] or NULL [h|d] -- entries of owned rows are written, ghost rows are zeroed.
 * J_vals: [nnz] or NULL [h|d], in the order of mastb200_get_pattern.  Outputs are overwritten (the
 * reference zeroes R and J first, nonlinear_implicit_assembly.cpp:81-82) and left "closed". */
int mastb200_residual_and_jacobian(mastb200_context* ctx, const double* X, double* R, double* J_vals);
int mastb200_res_l2_norm(mastb200_context* ctx, double* norm); /* of the last R (owned rows) */

/* One fused sweep for all density design variables:  G[n] = adj . dR/d(rho_n)  for every owned node n
 * (SURVEY 8a row S), optionally together with R/J of the same state (either may be NULL).
 * adj: [n_nodes*nv] localized adjoint [h|d] (compliance: adj == X).  G: [n_nodes] [h|d]. */
int mastb200_sensitivity_sweep(mastb200_context* ctx, const double* X, const double* adj, double* G,
                               double* R, double* J_vals);
/* sens[i] = sum_k dv_weight[k] * G[dv_node[k]], k in [dv_ptr[i], dv_ptr[i+1])   (host CSC arrays; G [h|d]) */
int mastb200_apply_dv_map_transpose(mastb200_context* ctx, const double* G, int64_t n_dv, const int64_t* dv_ptr,
                                    const int32_t* dv_node, const double* dv_weight, double* sens);

/* dR/dp of ONE parameter as a vector, dRdp[n_nodes*nv] [h|d] (rows of owned nodes, constrained dofs zeroed per element):
 *   MASTB200_PARAM_DENSITY_DV  nodal-density design variable under the SIMP law; dv_node[n_k], dv_weight[n_k] (host) =
 *                              the nodes and weights of its filter column (d rho_n / dp)
 *   MASTB200_PARAM_E           Young's modulus E0 of the isotropic card (dv arguments ignored)
 *   MASTB200_PARAM_NU          Poisson's ratio
 *   MASTB200_PARAM_K0          (QUAD4 heat conduction) conductivity k0 of the material card, k(T) = k0 (1 + k_beta T)
 *   MASTB200_PARAM_THICKNESS   (QUAD4 heat conduction) section thickness: conductance, volume source and side loads
 * 3D solids: matrix-free (two launches), K_e is never formed.  Heat conduction: one residual launch with the card swapped
 * (every term is linear in k0 or h)  <- HeatConductionElementBase::internal_residual_sensitivity and the load
 * sensitivities, src/heat_conduction/heat_conduction_elem_base.cpp:552-617, 957-1002, 1217-1275, 1524-1583, 1750-1794.
 *   <- NonlinearImplicitAssembly::sensitivity_assemble, src/base/nonlinear_implicit_assembly.cpp:515-686, with
 *      StructuralElement3D::internal_residual_sensitivity (src/elasticity/solid_element_3d.cpp:566-788, D <- dD/dp at :640)
 *      and IsotropicMaterialProperty::StiffnessMatrix3D::derivative (isotropic_material_property_card.cpp:530-567) */
#define MASTB200_PARAM_DENSITY_DV 0
#define MASTB200_PARAM_E 1
#define MASTB200_PARAM_NU 2
#define MASTB200_PARAM_K0 3
#define MASTB200_PARAM_THICKNESS 4
int mastb200_sensitivity_assemble(mastb200_context* ctx, const double* X, int32_t param, int64_t n_k, const int32_t* dv_node,
                                  const double* dv_weight, double* dRdp);
/* out[n_nodes*nv] [h|d] = J(X) dX without forming J (constrained dofs zeroed per element), 3D solids, linear strain
 *   <- NonlinearImplicitAssembly::linearized_jacobian_solution_product, src/base/nonlinear_implicit_assembly.cpp:270-383
 *      (declared at nonlinear_implicit_assembly.h:115), elem_linearized_jacobian_solution_product */
int mastb200_linearized_jacobian_solution_product(mastb200_context* ctx, const double* X, const double* dX, double* out);

/* ComplianceOutput: q = -sum f_ext . u  (owned part; caller sums over ranks unless skip_comm_sum); f_ext = SURFACE_PRESSURE
 * sides and the TEMPERATURE volume load (3D solids) */
int mastb200_compliance(mastb200_context* ctx, const double* X, double* q);
/* ComplianceOutput::evaluate_sensitivity through AssemblyBase::calculate_output_direct_sensitivity for parameters that do
 * not enter the loads: dq = -f_ext . dX/dp (dXdp [h|d] may be NULL => 0)
 *   <- src/base/assembly_base.cpp:434-550, src/elasticity/compliance_output.cpp:106-150 */
int mastb200_compliance_direct_sensitivity(mastb200_context* ctx, const double* X, const double* dXdp, double* dq);
/* dq_dX[n_nodes*nv] [h|d] = -f_ext with constrained rows zeroed */
int mastb200_compliance_derivative(mastb200_context* ctx, const double* X, double* dq_dX);

/* StressStrainOutputBase (3D solids): stress[n_elem*nq][6], strain[...][6] (Voigt xx,yy,zz,xy,yz,zx; engineering
 * shear), von_mises[n_elem*nq], jxw[n_elem*nq] at the quadrature points, any may be NULL [h|d]
 *   <- StructuralElement3D::calculate_stress, src/elasticity/solid_element_3d.cpp:1146-1265;
 *      Data::von_Mises_stress, src/elasticity/stress_output_base.cpp:182-195 */
int mastb200_stress(mastb200_context* ctx, const double* X, double* stress, double* strain, double* von_mises,
                    double* jxw);
/* p-norm / KS-weighted von Mises aggregate  q = sigma0 * (num/den)^(1/p_norm),
 *   num = sum (vm/sigma0)^p_norm e^{rho_exp (vm-sigma0)/sigma0} JxW,  den = sum e^{...} JxW
 * over the participating elements (elem_mask[n_elem] or NULL = all; OutputAssemblyElemOperations::
 * set_participating_elements).  sums[2] = {num, den} of this rank (the caller adds them over ranks, the reference's
 * comm().sum at stress_output_base.cpp:903-904); q (may be NULL) from this rank's sums.
 *   <- AssemblyBase::calculate_output + StressStrainOutputBase::functional_for_all_elems,
 *      src/base/assembly_base.cpp:253-333, src/elasticity/stress_output_base.cpp:851-909 */
int mastb200_stress_functional(mastb200_context* ctx, const double* X, double p_norm, double rho_exp, double sigma0,
                               const uint8_t* elem_mask, double* sums, double* q);
/* dq_dX[n_nodes*nv] [h|d] of that aggregate given the global sums {num, den}; rows of owned nodes, constrained
 * dofs zeroed per element
 *   <- AssemblyBase::calculate_output_derivative + StressStrainOutputBase::functional_state_derivartive_for_elem,
 *      src/base/assembly_base.cpp:338-430, src/elasticity/stress_output_base.cpp:1104-1172 */
int mastb200_stress_functional_derivative(mastb200_context* ctx, const double* X, double p_norm, double rho_exp,
                                          double sigma0, const uint8_t* elem_mask, const double* sums, double* dq_dX);

/* dq_drho[n_nodes] [h|d]: partial derivative (fixed state) of that aggregate w.r.t. the nodal densities of the SIMP law,
 * rows of owned nodes; with mastb200_apply_dv_map_transpose it gives output_sensitivity_total(p) of every density DV,
 * the dq/dp term that calculate_output_adjoint_sensitivity adds to adj . dR/dp (include_partial_sens = true)
 *   <- AssemblyBase::calculate_output_direct_sensitivity + StressStrainOutputBase::functional_sensitivity_for_elem,
 *      src/base/assembly_base.cpp:434-550, src/elasticity/stress_output_base.cpp:911-1043, with
 *      dstress/dp = dD/dp strain (pattern of src/elasticity/structural_element_2d.cpp:664-677; the 3D element of the
 *      reference sets no stress sensitivity at all, src/elasticity/solid_element_3d.cpp:1146-1265) */
int mastb200_stress_functional_density_derivative(mastb200_context* ctx, const double* X, double p_norm, double rho_exp,
                                                  double sigma0, const uint8_t* elem_mask, const double* sums,
                                                  double* dq_drho);

/* ---- transient and modal variants of the assembly ------------------------------------------------------------------ */
/* M_vals[nnz] [h|d]: the mass matrix (3D solids: rho_m = density, on the three translations) or the capacitance matrix
 * (QUAD4 heat conduction: rho_m = rho cp, times the section thickness) in the pattern of mastb200_get_pattern, per-element
 * constraint semantics of constrain_element_matrix.  lumped = 1: if_diagonal_mass_matrix (vol / n_shape for the solids,
 * row sums for heat conduction).
 *   <- StructuralElement3D::inertial_residual, src/elasticity/solid_element_3d.cpp:48-130;
 *      HeatConductionElementBase::velocity_residual, src/heat_conduction/heat_conduction_elem_base.cpp:155-243 */
int mastb200_mass_matrix(mastb200_context* ctx, double rho_m, int32_t lumped, double* M_vals);
/* R = f_x(X) + M rate, J = df_x/dX + jac_coeff M, constrained per element as one matrix / vector.  rate [h|d]: the
 * acceleration (second-order structural system, jac_coeff = 1 / beta / dt^2) or the temperature rate (first-order heat
 * conduction, jac_coeff = 1 / beta / dt) the transient solver derives from X; R / J_vals [h|d], either may be NULL.
 *   <- TransientAssembly::residual_and_jacobian, src/base/transient_assembly.cpp:63-180, with
 *      SecondOrderNewmarkTransientSolver::elem_calculations, src/solver/second_order_newmark_transient_solver.cpp:277-380 */
int mastb200_transient_residual_and_jacobian(mastb200_context* ctx, const double* X, const double* rate, double rho_m,
                                             int32_t lumped, double jac_coeff, double* R, double* J_vals);
/* A_vals = tangent stiffness at X, B_vals = mass matrix (both [nnz], [h|d])
 *   <- EigenproblemAssembly::eigenproblem_assemble, src/base/eigenproblem_assembly.cpp:90-180 */
int mastb200_eigenproblem_assemble(mastb200_context* ctx, const double* X, double rho_m, int32_t lumped, double* A_vals,
                                   double* B_vals);

/* ---- density filter and material volume of the SIMP loop ------------------------------------------------------------ */
/* Geometric filter over the nodes of the mesh: W[n,k] = (radius - |x_n - x_k|) / sum_k' (radius - |x_n - x_k'|) for the
 * nodes k within `radius` of n, neighbours in ascending node id.  dv_node[n_nodes] (host) or NULL = every node is a
 * design variable; rows of the other nodes are the identity.  The radius search runs on the device at this call.
 * Every node within the radius must be local (one rank, or a radius inside the ghost layer).
 *   <- MAST::FilterBase::_init / _init2, src/level_set/filter_base.cpp:275-372, 378-447 */
int mastb200_filter_init(mastb200_context* ctx, double radius, const uint8_t* dv_node, int64_t* nnz);
/* the filter as CSR (host arrays; col / weight may be NULL): row_ptr[n_nodes+1], col[nnz], weight[nnz] */
int mastb200_filter_get(mastb200_context* ctx, int64_t* row_ptr, int32_t* col, double* weight);
/* rho[n_nodes] = W d   <- FilterBase::compute_filtered_values, filter_base.cpp:59-97.  [h|d], rho != d */
int mastb200_filter_apply(mastb200_context* ctx, const double* d, double* rho);
/* grad[n_nodes] = W^T G: nodal sensitivities G_n (mastb200_sensitivity_sweep, ..._density_derivative) to design-variable
 * sensitivities, the per-DV loop of examples/structural/example_6/example_6.cpp:877-939 as one product.  [h|d] */
int mastb200_filter_apply_transpose(mastb200_context* ctx, const double* G, double* grad);
/* material volume V = sum_e vol(e) mean_{a in e} rho_a over this rank's owned nodes' share, and dV/d(rho_n) [n_nodes]
 * (either may be NULL; the gradient does not depend on rho)   <- example_6.cpp:714-791 */
int mastb200_volume(mastb200_context* ctx, const double* rho, double* volume, double* dV_drho);

/* ---- communicator: ghost exchange and sums over the ranks of one box (NCCL over NVLink) ------------------------- */
/* One process per GPU.  Rank 0 calls mastb200_comm_unique_id (id: at least 128 bytes), the host broadcasts the bytes by
 * whatever it has (MPI_Bcast in a MAST run, torch.distributed in bench.py), every rank calls mastb200_comm_init; or an
 * existing ncclComm_t is attached (not destroyed with the context).  libnccl.so.2 is loaded on first use (dlopen). */
int mastb200_comm_unique_id(void* id, int32_t bytes);
int mastb200_comm_init(mastb200_context* ctx, int32_t rank, int32_t world, const void* id);
int mastb200_comm_attach(mastb200_context* ctx, void* nccl_comm, int32_t rank, int32_t world);
/* Ghost lists in the context's local node numbering, per neighbouring rank k (CSR over the neighbours):
 * send_nodes[send_ptr[k] .. send_ptr[k+1]) = my owned nodes rank nbr_rank[k] keeps as ghosts, recv_nodes[...] = my ghosts it
 * owns; both sides list the shared nodes in the same order.  Contiguous ranges travel without packing.
 * auto_localize = 1: every sweep refreshes the ghost values of its X (adjoint, ...) itself before the kernels, on the
 * context stream -- X must then be a device vector and is updated in place; n_nbr = 0 clears.
 *   <- AssemblyBase::build_localized_vector, src/base/assembly_base.cpp:209-228 (send_list of the libMesh dof map) */
int mastb200_set_ghost_exchange(mastb200_context* ctx, int32_t n_nbr, const int32_t* nbr_rank, const int64_t* send_ptr,
                                const int32_t* send_nodes, const int64_t* recv_ptr, const int32_t* recv_nodes,
                                int32_t auto_localize);
/* V [h|d]: [n_nodes*width] with valid owned values; the ghost values are fetched from their owners, in place. */
int mastb200_build_localized_vector(mastb200_context* ctx, double* V, int32_t width);
/* vals[n] (host, n <= 64) summed over the ranks in place (res_l2_norm^2, compliance, aggregate sums); no-op for one rank */
int mastb200_all_reduce_sum(mastb200_context* ctx, double* vals, int32_t n);

/* ---- the linear solve after assembly, on the device CSR ----------------------------------------------------------- */
typedef struct mastb200_solve_info {
  int32_t iterations, converged;
  double rel_residual;  /* |rhs - J x| / |rhs| of the recurrence */
  double solve_ms;      /* CUDA-event time of the solve on the context stream */
} mastb200_solve_info;
/* x [h|d]: [n_nodes*nv] solution of J x = rhs over the owned rows by conjugate gradients with the 3x3 node blocks of J as
 * the preconditioner, initial guess zero, stop at |r| <= rtol |rhs| or max_it iterations.  J_vals [h|d] in the order of
 * mastb200_get_pattern (normally the device array the last sweep filled: J never leaves HBM); rhs [h|d].
 * 3-dof structural system (nv = 3; HEX8, HEX27); with a communicator the ranks solve the distributed system together
 * (ghost refresh of the search direction per iteration, scalars summed over the ranks).
 *   <- NonlinearSystem::solve / sensitivity_solve / adjoint_solve: linear_solver->solve(*matrix, pc, dsol, rhs, tol, max_it),
 *      src/base/nonlinear_system.cpp:777-782, 834-839 (PETSc KSP: -ksp_type cg -pc_type bjacobi) */
int mastb200_solve(mastb200_context* ctx, const double* J_vals, const double* rhs, double* x, double rtol, int32_t max_it,
                   mastb200_solve_info* info);
/* modes[6][n_nodes*nv] [h|d]: the rigid-body modes a multigrid preconditioner takes as its near null space: three
 * translations, three rotations by 1e-5 about the origin
 *   <- StructuralNearNullVectorSpace / StructuralModes, src/elasticity/structural_near_null_vector_space.cpp:32-150 */
int mastb200_near_null_space(mastb200_context* ctx, double* modes);

/* ---- introspection for bench / tests -------------------------------------------------------- */
typedef struct mastb200_stats {
  int64_t n_tiles, n_visits, n_owned_nodes, n_elem;
  int64_t kernel_launches;       /* cumulative count of this library's kernel launches */
  double last_sweep_ms;          /* CUDA-event time of the last sweep's kernels on the context stream */
  double last_main_kernel_ms;    /* ... of the dominant (tile sweep) kernel alone */
  int64_t plan_bytes;            /* device bytes of the static plan streamed per sweep */
  int32_t smem_bytes, threads_per_cta, ctas;
  int32_t structured;            /* the last HEX8 sweep used an implicit-connectivity kernel (build_cube numbering): 1 lock-step, 2 column march */
} mastb200_stats;
int mastb200_get_stats(const mastb200_context* ctx, mastb200_stats* s);
/* Streaming assembly plan of a planning-only context (device -1), for host-side verification of the integer
 * bookkeeping (tests/test_host_logic.py replays it on the CPU).  aux: 0 plain, 1 auxiliary shared-memory layout.
 * sizes[9]: number of tiles, steps, visit records, flush rows, flush entries, loads; arena blocks; state slots;
 * compute warps.  Record layouts: mast-multiphysics_b200/csrc/plan.h (TileDesc 16 B, StepDesc 48 B, VisitRec8 96 B, FlushRow 32 B,
 * flush entry 4 B, LoadRec 8 B). */
int mastb200_debug_plan_sizes(mastb200_context* ctx, int32_t aux, int64_t* sizes);
int mastb200_debug_plan_copy(mastb200_context* ctx, int32_t aux, void* tiles, void* steps, void* visits,
                             void* flush_rows, void* flush_ent, void* loads);
/* the CUDA stream all work of this context is ordered on (cudaStream_t as void*) */
void* mastb200_stream(const mastb200_context* ctx);

/* Stream-ordered mode ("one context per rank/GPU, stream-ordered").  Off (default): every call returns after its
 * work has finished, as MAST's assembly does (R and J closed on return, nonlinear_implicit_assembly.cpp:258-265).  On: a
 * residual_and_jacobian / sensitivity_sweep call whose X, R, J, G (and adjoint) are all DEVICE pointers only enqueues its
 * work on mastb200_stream() and returns; the caller orders later work on that stream (the device solve, the next sweep) or
 * synchronises it.  Calls with host buffers still synchronise.  mastb200_get_stats synchronises the stream. */
int mastb200_set_async(mastb200_context* ctx, int32_t on);

#ifdef __cplusplus
}
#endif
#endif
