// mass.cu -- mass / capacitance matrices and the transient (Newmark) and eigenproblem assemblies built on them (sm_100a, fp64).
//
// Device counterparts of
//   StructuralElement3D::inertial_residual      /root/reference/src/elasticity/solid_element_3d.cpp:48-130
//       consistent: jac_xddot += JxW N^T rho N on the three translations; if_diagonal_mass_matrix: vol / n_shape rho
//   HeatConductionElementBase::velocity_residual src/heat_conduction/heat_conduction_elem_base.cpp:155-243
//       jac_xdot += JxW (rho cp) N^T N (times the thickness of the 2D section card), row-sum lumped if diagonal
//   TransientAssembly::residual_and_jacobian    src/base/transient_assembly.cpp:63-180 with the element combination of the
//       Newmark solvers (src/solver/second_order_newmark_transient_solver.cpp:277-380): vec = f_m + f_x,
//       mat = jac_coeff df_m/d(rate) + df_x/dx, constrained together
//   EigenproblemAssembly::eigenproblem_assemble src/base/eigenproblem_assembly.cpp:90-180: A = K, B = M, each with
//       constrain_element_matrix (:167-168)
//
// M_AB is a scalar per node pair (times the identity on the translations), so it lives on the node-level pattern:
//   pass 1  one thread per (element, quadrature point): JxW -> scratch (the Jacobian is evaluated once per point)
//   pass 2  one thread per neighbour entry (A, B) of the node-level CSR: the elements around A in ascending id (the
//           reference's ADD order) that also hold B, m_AB = sum_e sum_q JxW N_a N_b; the value is written to / added into
//           the diagonal of the entry's nv x nv block with libMesh's per-element Dirichlet semantics
//   pass 3  (transient) one thread per owned node: R_A += sum_B m_AB rate_B in neighbour order
// No atomics, fixed summation order: bitwise reproducible.
#include "context.h"

namespace {

constexpr int MAXNN = 27;

__global__ void __launch_bounds__(128) jxw_kernel(int64_t n_elem, int nn, int nq, int dim, const int32_t* __restrict__ conn,
                                                  const double* __restrict__ tab, const double* __restrict__ xyz, double* jxw) {
  const int64_t t = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (t >= n_elem * nq) return;
  const int64_t e = t / nq;
  const int q = (int)(t - e * nq);
  const int32_t* en = conn + e * nn;
  const double* dNt = tab + (size_t)nq * nn + (size_t)q * nn * 3;
  double J[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};
  for (int b = 0; b < nn; ++b) {
    const double* X = xyz + (int64_t)en[b] * 3;
    for (int r = 0; r < dim; ++r)
      for (int c = 0; c < dim; ++c) J[r][c] = fma(dNt[b * 3 + r], X[c], J[r][c]);
  }
  const double det = dim == 3 ? J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) + J[0][1] * (J[1][2] * J[2][0] - J[1][0] * J[2][2]) +
                                    J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0])
                              : J[0][0] * J[1][1] - J[0][1] * J[1][0];
  jxw[t] = det * tab[(size_t)nq * nn * 4 + q];
}

struct MassParams {
  int64_t n_owned, n_entries;
  int nn, nq, nv, ndiag;        // ndiag: dofs of a node that carry mass (3 translations, or the temperature)
  int lumped;                   // 0 consistent, 1 vol / nn (structural), 2 row sums (heat)
  int mode;                     // 0: out = M (constrained diagonal = adjacent elements), 1: out += coeff M (constrained entries skipped)
  double c, coeff;
  const int32_t* owned; const int64_t* nbr_ptr; const int32_t* nbr; const int32_t* ent_row;
  const int64_t* n2e_ptr; const int32_t* n2e; const int32_t* conn; const uint8_t* diag_elem; const uint8_t* cmask;
  const double* tab; const double* jxw;
  double* m_entry;              // [n_entries] unconstrained m_AB
  double* out;                  // CSR values
};

__global__ void __launch_bounds__(128) mass_entry_kernel(const MassParams p) {
  const int64_t k = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (k >= p.n_entries) return;
  const int r = p.ent_row[k];
  const int A = p.owned[r], B = p.nbr[k];
  const int nn = p.nn, nq = p.nq;
  double m = 0.;
  int n_adj = 0;
  for (int64_t ke = p.n2e_ptr[A]; ke < p.n2e_ptr[A + 1]; ++ke) {
    const int64_t e = p.n2e[ke];
    if (p.diag_elem && p.diag_elem[e]) continue;
    ++n_adj;
    const int32_t* en = p.conn + e * nn;
    int a = -1, b = -1;
    for (int i = 0; i < nn; ++i) { if (en[i] == A) a = i; if (en[i] == B) b = i; }
    if (b < 0) continue;
    double me = 0.;
    if (p.lumped == 1) {           // vol / n_shape on the diagonal (solid_element_3d.cpp:84-99)
      if (A == B) { for (int q = 0; q < nq; ++q) me += p.jxw[e * nq + q]; me /= (1. * nn); }
    } else if (p.lumped == 2) {    // row sum of the consistent matrix on the diagonal (heat_conduction_elem_base.cpp:232-236)
      if (A == B)
        for (int bb = 0; bb < nn; ++bb)
          for (int q = 0; q < nq; ++q) me += p.jxw[e * nq + q] * p.tab[(size_t)q * nn + a] * p.tab[(size_t)q * nn + bb];
    } else {
      for (int q = 0; q < nq; ++q) me += p.jxw[e * nq + q] * p.tab[(size_t)q * nn + a] * p.tab[(size_t)q * nn + b];
    }
    m += p.c * me;
  }
  if (p.m_entry) p.m_entry[k] = m;
  if (!p.out) return;
  const int64_t b0 = p.nbr_ptr[r], nnb = p.nbr_ptr[r + 1] - b0;
  const unsigned cmA = p.cmask ? p.cmask[A] : 0u, cmB = p.cmask ? p.cmask[B] : 0u;
  double* blk = p.out + b0 * p.nv * p.nv + (k - b0) * p.nv;
  for (int i = 0; i < p.ndiag; ++i) {
    const bool con = ((cmA >> i) | (cmB >> i)) & 1u;
    double* dst = blk + (int64_t)i * nnb * p.nv + i;
    if (p.mode == 0) *dst = con ? ((A == B && ((cmA >> i) & 1u)) ? (double)n_adj : 0.) : m;
    else if (!con) *dst += p.coeff * m;
  }
  if (p.mode == 0 && A == B)  // dofs without mass (rotations of the six-variable system): constrained rows still get their 1s
    for (int i = p.ndiag; i < p.nv; ++i)
      if ((cmA >> i) & 1u) blk[(int64_t)i * nnb * p.nv + i] = (double)n_adj;
}

__global__ void __launch_bounds__(128) mass_times_rate_kernel(const MassParams p, const double* __restrict__ rate, double* R) {
  const int64_t r = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (r >= p.n_owned) return;
  const int A = p.owned[r];
  const unsigned cmA = p.cmask ? p.cmask[A] : 0u;
  double acc[3] = {0., 0., 0.};
  for (int64_t k = p.nbr_ptr[r]; k < p.nbr_ptr[r + 1]; ++k) {
    const double m = p.m_entry[k];
    const double* v = rate + (int64_t)p.nbr[k] * p.nv;
    for (int i = 0; i < p.ndiag; ++i) acc[i] = fma(m, v[i], acc[i]);
  }
  for (int i = 0; i < p.ndiag; ++i)
    if (!((cmA >> i) & 1u)) R[(int64_t)A * p.nv + i] += acc[i];  // constrain_element_vector: constrained rows stay zero
}

int mass_setup(mastb200_context* ctx, double rho_m, int lumped, MassParams& p) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  const bool heat = ctx->elem_type == MASTB200_QUAD4 && ctx->nv == 1;
  if (!heat && ctx->elem_type != MASTB200_HEX8 && ctx->elem_type != MASTB200_HEX27)
    return ctx->fail(MASTB200_ERR_UNSUPPORTED, "mass matrix: 3D solids (HEX8, HEX27) and QUAD4 heat conduction");
  if (!heat && ctx->nv == 1) return ctx->fail(MASTB200_ERR_ARG, "structural mass matrix needs nv = 3 or 6");
  CU(cudaSetDevice(ctx->device));
  if ((rc = ensure_generic_plan(ctx))) return rc;
  mastb200_context::GenericPlan& G = ctx->gplan;
  if (!G.nbr.p) CU(G.nbr.upload(ctx->pat.nbr, ctx->stream));
  if (!G.ent_row.p) {
    std::vector<int32_t> er((size_t)ctx->pat.nbr_ptr[ctx->pat.n_owned]);
    for (int64_t r = 0; r < ctx->pat.n_owned; ++r)
      for (int64_t k = ctx->pat.nbr_ptr[r]; k < ctx->pat.nbr_ptr[r + 1]; ++k) er[k] = (int32_t)r;
    CU(G.ent_row.upload(er, ctx->stream));
  }
  if (heat && !G.tab.p) {  // QUAD4 tables in the same layout: N[q][b], dN[q][b][3], w[q] (2x2 Gauss, xi fastest, unit weights)
    static const int8_t C[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
    const double gp = 0.57735026918962576451;
    std::vector<double> tab(4 * 4 * 4 + 4, 0.);
    for (int q = 0; q < 4; ++q) {
      const double xi = (q & 1) ? gp : -gp, eta = (q & 2) ? gp : -gp;
      for (int b = 0; b < 4; ++b) {
        tab[q * 4 + b] = 0.25 * (1. + C[b][0] * xi) * (1. + C[b][1] * eta);
        tab[16 + (q * 4 + b) * 3] = 0.25 * C[b][0] * (1. + C[b][1] * eta);
        tab[16 + (q * 4 + b) * 3 + 1] = 0.25 * C[b][1] * (1. + C[b][0] * xi);
      }
      tab[64 + q] = 1.;
    }
    CU(G.tab.upload(tab, ctx->stream));
  }
  const int nq = ctx->nn;
  CU(ctx->d_scratch.ensure(size_t(ctx->n_elem) * nq * 9));  // shared with the other qp kernels: never smaller
  jxw_kernel<<<(unsigned)((ctx->n_elem * nq + 127) / 128), 128, 0, ctx->stream>>>(ctx->n_elem, ctx->nn, nq, heat ? 2 : 3, G.conn.p,
                                                                                 G.tab.p, ctx->d_xyz.p, ctx->d_scratch.p);
  CU(cudaGetLastError());
  ctx->launches++;
  p.n_owned = ctx->pat.n_owned; p.n_entries = ctx->pat.nbr_ptr[ctx->pat.n_owned];
  p.nn = ctx->nn; p.nq = nq; p.nv = ctx->nv; p.ndiag = heat ? 1 : 3;
  p.lumped = lumped ? (heat ? 2 : 1) : 0;
  p.c = heat ? rho_m * ctx->mat.thickness : rho_m;
  p.owned = ctx->d_owned_nodes.p; p.nbr_ptr = G.nbr_ptr.p; p.nbr = G.nbr.p; p.ent_row = G.ent_row.p;
  p.n2e_ptr = G.n2e_ptr.p; p.n2e = G.n2e.p; p.conn = G.conn.p;
  p.diag_elem = ctx->has_diag ? G.diag.p : nullptr;
  p.cmask = ctx->has_constraints ? ctx->d_cmask.p : nullptr;
  p.tab = G.tab.p; p.jxw = ctx->d_scratch.p;
  p.m_entry = nullptr; p.out = nullptr; p.mode = 0; p.coeff = 0.;
  return MASTB200_OK;
}

}  // namespace

extern "C" {

int mastb200_mass_matrix(mastb200_context* ctx, double rho_m, int32_t lumped, double* M_vals) {
  if (!ctx || !M_vals) return MASTB200_ERR_ARG;
  MassParams p;
  int rc = mass_setup(ctx, rho_m, lumped, p);
  if (rc) return rc;
  const size_t nnz = (size_t)plan_nnz(ctx);
  double* dM = nullptr;
  if ((rc = resolve_out(ctx, M_vals, nnz, ctx->d_J, &dM))) return rc;
  CU(cudaMemsetAsync(dM, 0, nnz * sizeof(double), ctx->stream));
  p.out = dM; p.mode = 0;
  mass_entry_kernel<<<(unsigned)((p.n_entries + 127) / 128), 128, 0, ctx->stream>>>(p);
  CU(cudaGetLastError());
  ctx->launches++;
  if ((rc = finish_out(ctx, M_vals, nnz, dM))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}

int mastb200_transient_residual_and_jacobian(mastb200_context* ctx, const double* X, const double* rate, double rho_m,
                                             int32_t lumped, double jac_coeff, double* R, double* J_vals) {
  if (!ctx || !X || !rate) return MASTB200_ERR_ARG;
  int rc = check_ready(ctx);
  if (rc) return rc;
  const size_t n_dof = size_t(ctx->n_nodes) * ctx->nv, nnz = (size_t)plan_nnz(ctx);
  // the static part f_x(X), df_x/dX into device arrays (the caller's if they are device pointers)
  double *dR = nullptr, *dJ = nullptr;
  if ((rc = resolve_out(ctx, R, n_dof, ctx->d_R, &dR))) return rc;
  if ((rc = resolve_out(ctx, J_vals, nnz, ctx->d_J, &dJ))) return rc;
  if ((rc = mastb200_residual_and_jacobian(ctx, X, dR, dJ))) return rc;
  const double* drate = nullptr;
  if ((rc = resolve_in(ctx, rate, n_dof, ctx->d_adj, &drate))) return rc;
  if ((rc = localize_if_needed(ctx, drate, ctx->nv))) return rc;
  MassParams p;
  if ((rc = mass_setup(ctx, rho_m, lumped, p))) return rc;
  CU(ctx->d_dfield.ensure((size_t)p.n_entries));
  p.m_entry = ctx->d_dfield.p; p.out = dJ; p.mode = 1; p.coeff = jac_coeff;
  mass_entry_kernel<<<(unsigned)((p.n_entries + 127) / 128), 128, 0, ctx->stream>>>(p);
  if (dR) mass_times_rate_kernel<<<(unsigned)((p.n_owned + 127) / 128), 128, 0, ctx->stream>>>(p, drate, dR);
  CU(cudaGetLastError());
  ctx->launches += dR ? 2 : 1;
  ctx->last_R = dR;
  if ((rc = finish_out(ctx, R, n_dof, dR))) return rc;
  if ((rc = finish_out(ctx, J_vals, nnz, dJ))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}

int mastb200_eigenproblem_assemble(mastb200_context* ctx, const double* X, double rho_m, int32_t lumped, double* A_vals,
                                   double* B_vals) {
  if (!ctx || !X || !A_vals || !B_vals) return MASTB200_ERR_ARG;
  // A: the tangent stiffness at X (df_x/dX with constrain_element_matrix), B: the mass matrix, both in the pattern
  int rc = mastb200_residual_and_jacobian(ctx, X, nullptr, A_vals);
  if (rc) return rc;
  return mastb200_mass_matrix(ctx, rho_m, lumped, B_vals);
}

}  // extern "C"
