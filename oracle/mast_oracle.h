/* mast_oracle.h -- CPU ORACLE for the MAST element-assembly + sensitivity hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain, scalar, dependency-free C++ restatement of the
 * reference algorithm (MASTmultiphysics/mast-multiphysics @ 7860f776).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it; the
 * product library (mast_b200/csrc) never links, includes or calls it.
 *
 * PARITY STATUS: "parity unpinned" for element matrices and global assembly -- the reference holds
 * golden vectors only for the material matrices (pinned in tests/test_oracle_golden.py); libMesh,
 * PETSc and Eigen, which own the shape functions, quadrature, dof numbering, constraint handling and
 * ADD_VALUES accumulation, are un-vendored dependencies (libMesh 1.5.1 / PETSc 3.7.7 per
 * /root/reference/.travis.yml, ci/build_dependencies.sh) and are restated here from their published
 * definitions.  The reference cannot be compiled in this image (no libMesh/PETSc/Eigen/MPI), so
 * there is no oracle/_ref.  See DESIGN.md section 3.
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 */
#ifndef MAST_ORACLE_H
#define MAST_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_HEX8 = 8, ORACLE_HEX27 = 27, ORACLE_QUAD4 = 4 };

/* One analysis: mesh + dof layout + property card + loads.  All arrays are caller-owned. */
typedef struct OracleProblem {
  /* mesh (libMesh node ordering of each element) */
  int32_t elem_type;          /* ORACLE_HEX8 | ORACLE_HEX27 (3D solid) | ORACLE_QUAD4 (nv=1: 2D heat conduction; nv=6: Mindlin plate) */
  int64_t n_elem, n_nodes;
  const double* xyz;          /* [n_nodes][3] */
  const int32_t* conn;        /* [n_elem][nn] */
  const uint8_t* diag_elem;   /* [n_elem] or NULL: 1 => element is in AssemblyBase::diagonal_elem_subdomain_id */
  /* dof layout: dof(node, var) = nv*node + var (libMesh variable-group numbering, one group) */
  int32_t nv;                 /* 6 = reference structural system (u,v,w,tx,ty,tz); 3 = compact; 1 = heat */
  const uint8_t* constrained; /* [n_nodes*nv] or NULL: 1 => homogeneous Dirichlet dof */
  /* isotropic material card; E(x) = E0 if rho == NULL, else SIMP E0*(rho_min + (1-rho_min)*rho(x)^penalty) */
  double E0, nu, alpha;
  const double* rho;          /* [n_nodes] nodal density field or NULL */
  double rho_min, penalty;
  int32_t nonlinear_strain;   /* 0 = LINEAR_STRAIN, 1 = NONLINEAR_STRAIN (Green-Lagrange) */
  /* side loads: SURFACE_PRESSURE on (element, side) pairs, constant pressure per entry */
  int64_t n_side_loads;
  const int32_t* side_elem;   /* [n_side_loads] */
  const int32_t* side_id;     /* [n_side_loads] libMesh side number */
  const double* side_pressure;/* [n_side_loads] */
  /* volume load: TEMPERATURE (thermoelastic) if thermal != 0; T nodal field or constant */
  int32_t thermal;
  const double* temperature;  /* [n_nodes] or NULL => temperature_const */
  double temperature_const, ref_temperature;
  /* heat conduction (elem_type QUAD4, nv=1): k(T) = k0*(1+k_beta*T), section thickness h */
  double k0, k_beta, thickness;
  int32_t k_of_T;             /* 1 => attach the active solution function (dk/dT tangent term) */
  double heat_source;         /* volumetric heat source q (HEAT_SOURCE volume load), 0 = none */
  /* 2D plate section (elem_type QUAD4, nv=6: StructuralElement2D with Mindlin bending): thickness above, mid-plane
   * offset, shear correction factor kappa, plane_strain = 0 (plane stress) | 1 */
  double offset, kappa;
  int32_t plane_strain;
  /* heat conduction side loads (QUAD4, nv=1) reuse side_elem / side_id; side_pressure holds the HEAT_FLUX value and
   * these two the CONVECTION_HEAT_FLUX coefficient and ambient temperature per entry (NULL: none) */
  const double* side_conv_coeff;
  const double* side_T_amb;
  /* SURFACE_RADIATION_HEAT_FLUX on the same entries: emissivity per entry (NULL: none; side_T_amb is its ambient
   * temperature), Stefan-Boltzmann constant and reference-zero temperature of the boundary condition */
  const double* side_emissivity;
  double rad_sb_const, rad_T_ref_zero;
} OracleProblem;

/* libMesh-style sparsity: every (row,col) pair of dofs that share an element; columns sorted.
 * Call with col_idx == NULL to get nnz and row_ptr; again with col_idx to fill. Returns nnz. */
int64_t oracle_build_pattern(const OracleProblem* p, int64_t* row_ptr, int32_t* col_idx);

/* NonlinearImplicitAssembly::residual_and_jacobian. R and/or J_vals may be NULL. Returns 0 on success. */
int oracle_residual_and_jacobian(const OracleProblem* p, const double* X, double* R, double* J_vals,
                                 const int64_t* row_ptr, const int32_t* col_idx);

/* elem_calculations of one element (no constraints): vec[nd], mat[nd*nd] row-major, nd = nv*nn,
 * variable-major element dof order. */
int oracle_elem_calculations(const OracleProblem* p, int64_t e, const double* X, int want_jac,
                             double* vec, double* mat);

/* Design variables: rho_n = sum_i W[n,i] d_i given as CSC columns (dv_ptr[n_dv+1], dv_node, dv_weight).
 * NonlinearImplicitAssembly::sensitivity_assemble for DV i -> dres_dp[n_dof]. */
int oracle_sensitivity_assemble(const OracleProblem* p, const double* X, int64_t dv_begin, int64_t dv_end,
                                const int32_t* dv_node, const double* dv_weight, double* dres_dp);

/* sensitivity_assemble for a Parameter of the isotropic card: kind 1 = E0, 2 = nu -> dres_dp[n_dof] */
int oracle_sensitivity_assemble_param(const OracleProblem* p, const double* X, int kind, double* dres_dp);
/* NonlinearImplicitAssembly::linearized_jacobian_solution_product: out[n_dof] = sum_e constrain(K_e(X) dX_e) */
int oracle_linearized_jacobian_solution_product(const OracleProblem* p, const double* X, const double* dX, double* out);
/* point loads of residual_and_jacobian: R[dofs of node[i]] -= load[i][nv] on unconstrained dofs */
int oracle_add_point_loads(const OracleProblem* p, int64_t n, const int32_t* node, const double* load, double* R);

/* mass / capacitance matrix in the pattern (EigenproblemAssembly's B; constrain_element_matrix per element) */
int oracle_mass_matrix(const OracleProblem* p, double rho_m, int lumped, const int64_t* row_ptr, const int32_t* col_idx,
                       double* M_vals);
/* TransientAssembly::residual_and_jacobian with the element combination of the Newmark solvers:
 * R = f_x(X) + M rate, J = df_x/dX + jac_coeff M, constrained per element */
int oracle_transient_residual_and_jacobian(const OracleProblem* p, const double* X, const double* rate, double rho_m, int lumped,
                                           double jac_coeff, double* R, double* J_vals, const int64_t* row_ptr,
                                           const int32_t* col_idx);

/* AssemblyBase::calculate_output_adjoint_sensitivity_multiple_parameters_no_direct: sens[i] = adj . dR/dp_i */
int oracle_adjoint_sensitivity(const OracleProblem* p, const double* X, const double* adj, int64_t n_dv,
                               const int64_t* dv_ptr, const int32_t* dv_node, const double* dv_weight,
                               double* sens);

/* ComplianceOutput through AssemblyBase::calculate_output: returns q = -sum_e f_ext,e . u_e */
double oracle_compliance(const OracleProblem* p, const double* X);
/* AssemblyBase::calculate_output_derivative with ComplianceOutput: dq_dX[n_dof] */
int oracle_compliance_derivative(const OracleProblem* p, const double* X, double* dq_dX);

/* StressStrainOutputBase: per quadrature point stress[6], strain[6], von Mises, JxW  (out arrays
 * sized n_elem*nq*{6,6,1,1}); then the p-norm aggregate functional_for_all_elems. */
int oracle_stress(const OracleProblem* p, const double* X, double* stress, double* strain, double* vm, double* jxw);
double oracle_stress_functional(const OracleProblem* p, const double* X, double p_norm, double rho_exp,
                                double sigma0, double* dq_dX /* [n_dof] or NULL */);

/* partial derivative (fixed state) of that aggregate w.r.t. nodal-density design variables (CSC columns as above):
 * StressStrainOutputBase::functional_sensitivity_for_all_elems with dstress/dp = dD/dp * strain */
int oracle_stress_functional_density_sensitivity(const OracleProblem* p, const double* X, double p_norm, double rho_exp,
                                                 double sigma0, int64_t n_dv, const int64_t* dv_ptr, const int32_t* dv_node,
                                                 const double* dv_weight, double* sens);

/* MAST::FilterBase: geometric density filter (linear radius search), filtered values, and the material volume of
 * example_6 with its derivative w.r.t. the nodal densities */
int64_t oracle_filter_weights(int64_t n_nodes, const double* xyz, double radius, int64_t* row_ptr, int32_t* col, double* weight);
void oracle_filtered_values(int64_t n_nodes, const int64_t* row_ptr, const int32_t* col, const double* weight,
                            const uint8_t* dv, const double* input, double* output);
double oracle_volume(const OracleProblem* p, const double* rho, double* dV_drho);

/* material matrices for the golden-vector tests */
void oracle_stiffness_3d(double E, double nu, double* D36);
void oracle_stiffness_3d_derivative(double E, double nu, double dE, double dnu, double* D36);
void oracle_thermal_expansion_A_3d(double E, double nu, double alpha, double* v6);

/* Solid2DSectionProperty matrices (solid_2d_section_element_property_card.cpp:395-510, :122-134): A = C h,
 * B = C h off, D = C (h^3/12 + h off^2) (3x3 row-major each), transverse shear S = G kappa h (2x2) */
void oracle_plate_section_matrices(double E, double nu, double h, double off, double kappa, int plane_strain,
                                   double* A9, double* B9, double* D9, double* S4);

/* FE tables of one element (libMesh FE restatement): phi[nn*nq], dphi[nn*nq*3], JxW[nq], xyz[nq*3] */
int oracle_fe_tables(int32_t elem_type, const double* coords /* [nn][3] */, double* phi, double* dphi,
                     double* jxw, double* xyz_qp);

/* timing helper for bench.py: one residual_and_jacobian sweep over elements [e0,e1) with `threads`
 * std::threads each owning a contiguous element block and a private row range (interface rows are
 * reduced afterwards, like the PETSc stash). Returns seconds. */
double oracle_time_residual_and_jacobian(const OracleProblem* p, const double* X, double* R, double* J_vals,
                                         const int64_t* row_ptr, const int32_t* col_idx, int threads);
/* the same; *threads_used = the thread count the sweep ran with (halved until blocks of equal parity share no node) */
double oracle_time_residual_and_jacobian_ex(const OracleProblem* p, const double* X, double* R, double* J_vals,
                                            const int64_t* row_ptr, const int32_t* col_idx, int threads, int* threads_used);

#ifdef __cplusplus
}
#endif
#endif
