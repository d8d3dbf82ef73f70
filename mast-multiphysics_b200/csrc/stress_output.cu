// stress_output.cu -- stress / strain at the quadrature points, von Mises p-norm aggregate and its state derivative
// for 3D solids (HEX8, HEX27), sm_100a fp64.
//
// Device counterpart of StructuralElement3D::calculate_stress (/root/reference/src/elasticity/solid_element_3d.cpp:
// 1146-1265: sigma = D eps at every quadrature point, stored with JxW), StressStrainOutputBase::Data::von_Mises_stress
// and ::dvon_Mises_stress_dX (src/elasticity/stress_output_base.cpp:182-230), the aggregate of
// ::functional_for_all_elems (:851-909) and ::functional_state_derivartive_for_elem (:1104-1172) as driven by
// AssemblyBase::calculate_output / calculate_output_derivative (src/base/assembly_base.cpp:253-430).
// The 3D dstress/dX is libmesh_error() in the reference (solid_element_3d.cpp:1258-1261); as in the 2D element
// (structural_element_2d.cpp:555-611) it is D B_lin.
//
//   pass 1  one thread per (element, quadrature point): geometry, strain, stress, von Mises; deterministic block sums of
//           (vm/s0)^p e^{rho (vm-s0)/s0} JxW and e^{...} JxW  (two-stage reduction, no atomics).
//   pass 2  same thread layout with the global sums known: the derivative of the aggregate w.r.t. the dofs of node a of
//           the element is sum_q T_q dN_a/dxi(q) with T_q = kappa_q M_q J^-T (M_q the symmetric tensor of D dvm/dsigma),
//           one 3x3 matrix per point, written to a scratch array.
//   pass 3  row-owner gather: one thread per owned node sums its adjacent elements in ascending id (the reference's
//           ADD order), constrained dofs zeroed per element (constrain_element_vector).
#include <cuda_runtime.h>

#include "kernels.h"

namespace mastb200 {
namespace {

constexpr int TB = 128;

struct QpGeom {
  double inv[3][3];  // inv[c][r] = d xi_r / d x_c
  double det, wq;
  double eps[6], sig[6];
  double E, rho;
  double H[3][3];    // displacement gradient d u_i / d x_j (F = I + H for the Green-Lagrange strain)
};

// geometry + linear/Green-Lagrange strain + stress at (e, q); tab: N[q][b], dN[q][b][3], w[q]
__device__ __forceinline__ void eval_qp(const StressParams& p, int64_t e, int q, QpGeom& g) {
  const int nn = p.nn, nq = p.nq;
  const double* Nt = p.tab + (size_t)q * nn;
  const double* dNt = p.tab + (size_t)nq * nn + (size_t)q * nn * 3;
  const int32_t* en = p.conn + e * nn;
  double Jm[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};
  for (int b = 0; b < nn; ++b) {
    const double* X = p.xyz + (int64_t)en[b] * 3;
    const double d0 = dNt[b * 3], d1 = dNt[b * 3 + 1], d2 = dNt[b * 3 + 2];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      Jm[0][c] = fma(d0, X[c], Jm[0][c]); Jm[1][c] = fma(d1, X[c], Jm[1][c]); Jm[2][c] = fma(d2, X[c], Jm[2][c]);
    }
  }
  const double c00 = Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1];
  const double c01 = Jm[1][2] * Jm[2][0] - Jm[1][0] * Jm[2][2];
  const double c02 = Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0];
  const double det = Jm[0][0] * c00 + Jm[0][1] * c01 + Jm[0][2] * c02;
  const double idet = 1. / det;
  g.inv[0][0] = c00 * idet; g.inv[1][0] = c01 * idet; g.inv[2][0] = c02 * idet;
  g.inv[0][1] = (Jm[0][2] * Jm[2][1] - Jm[0][1] * Jm[2][2]) * idet;
  g.inv[1][1] = (Jm[0][0] * Jm[2][2] - Jm[0][2] * Jm[2][0]) * idet;
  g.inv[2][1] = (Jm[0][1] * Jm[2][0] - Jm[0][0] * Jm[2][1]) * idet;
  g.inv[0][2] = (Jm[0][1] * Jm[1][2] - Jm[0][2] * Jm[1][1]) * idet;
  g.inv[1][2] = (Jm[0][2] * Jm[1][0] - Jm[0][0] * Jm[1][2]) * idet;
  g.inv[2][2] = (Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0]) * idet;
  g.det = det;
  g.wq = p.tab[(size_t)nq * nn * 4 + q];
  double H[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};
  double rho = 0.;
  for (int b = 0; b < nn; ++b) {
    const double d0 = dNt[b * 3], d1 = dNt[b * 3 + 1], d2 = dNt[b * 3 + 2];
    double gr[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) gr[c] = d0 * g.inv[c][0] + d1 * g.inv[c][1] + d2 * g.inv[c][2];
    const double* U = p.X + (int64_t)en[b] * p.nv;
    const double u0 = U[0], u1 = U[1], u2 = U[2];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      H[0][j] = fma(u0, gr[j], H[0][j]); H[1][j] = fma(u1, gr[j], H[1][j]); H[2][j] = fma(u2, gr[j], H[2][j]);
    }
    if (p.rho) rho = fma(Nt[b], p.rho[en[b]], rho);
  }
  double E = p.E0;
  if (p.rho) E = p.E0 * (p.rho_min + (1. - p.rho_min) * simp_pow(rho, p.penalty));
  g.E = E; g.rho = rho;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) g.H[i][j] = H[i][j];
  if (p.nonlinear) {
    double Eg[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        Eg[i][j] = 0.5 * (H[i][j] + H[j][i] + (H[0][i] * H[0][j] + H[1][i] * H[1][j] + H[2][i] * H[2][j]));
    g.eps[0] = Eg[0][0]; g.eps[1] = Eg[1][1]; g.eps[2] = Eg[2][2];
    g.eps[3] = Eg[0][1] + Eg[1][0]; g.eps[4] = Eg[1][2] + Eg[2][1]; g.eps[5] = Eg[0][2] + Eg[2][0];
  } else {
    g.eps[0] = H[0][0]; g.eps[1] = H[1][1]; g.eps[2] = H[2][2];
    g.eps[3] = H[0][1] + H[1][0]; g.eps[4] = H[1][2] + H[2][1]; g.eps[5] = H[0][2] + H[2][0];
  }
  const double lam = E * p.lam1, mu = E * p.mu1, tr = g.eps[0] + g.eps[1] + g.eps[2];
  g.sig[0] = lam * tr + 2. * mu * g.eps[0]; g.sig[1] = lam * tr + 2. * mu * g.eps[1]; g.sig[2] = lam * tr + 2. * mu * g.eps[2];
  g.sig[3] = mu * g.eps[3]; g.sig[4] = mu * g.eps[4]; g.sig[5] = mu * g.eps[5];
}

// rho^(penalty-1) of E'(rho): the usual small integer SIMP penalties by multiplication
__device__ __forceinline__ double pow_pm1(double rho, double pm1) {
  if (pm1 == 3.) return rho * rho * rho;
  if (pm1 == 2.) return rho * rho;
  if (pm1 == 1.) return rho;
  if (pm1 == 0.) return 1.;
  return pow(rho, pm1);
}

__device__ __forceinline__ double vm_pp(const double* s) {
  const double a = s[0] - s[1], b = s[1] - s[2], c = s[2] - s[0];
  return 0.5 * (a * a + b * b + c * c) + 3.0 * (s[3] * s[3] + s[4] * s[4] + s[5] * s[5]);
}

// deterministic CTA sum of two values -> partial[2*block], partial[2*block+1]
__device__ __forceinline__ void block_sum2(double a, double b, double* partial) {
  __shared__ double sa[TB / 32], sb[TB / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double x = 0., y = 0.;
    for (int w = 0; w < TB / 32; ++w) { x += sa[w]; y += sb[w]; }
    partial[2 * blockIdx.x] = x;
    partial[2 * blockIdx.x + 1] = y;
  }
}

__global__ void __launch_bounds__(TB) stress_qp_kernel(const StressParams p) {
  const int64_t t = (int64_t)blockIdx.x * TB + threadIdx.x;
  const int64_t nqp = p.n_elem * p.nq;
  double num = 0., den = 0.;
  if (t < nqp) {
    const int64_t e = t / p.nq;
    const int q = (int)(t - e * p.nq);
    QpGeom g;
    eval_qp(p, e, q, g);
    const double vm = sqrt(vm_pp(g.sig));
    const double JxW = g.det * g.wq;
    if (p.out_stress)
#pragma unroll
      for (int i = 0; i < 6; ++i) p.out_stress[t * 6 + i] = g.sig[i];
    if (p.out_strain)
#pragma unroll
      for (int i = 0; i < 6; ++i) p.out_strain[t * 6 + i] = g.eps[i];
    if (p.out_vm) p.out_vm[t] = vm;
    if (p.out_jxw) p.out_jxw[t] = JxW;
    // diagonal_elem_subdomain_id elements take no part in outputs (assembly_base.cpp:299), as in the derivative kernels
    if (p.partial && (!p.elem_mask || p.elem_mask[e]) && !(p.diag_elem && p.diag_elem[e])) {
      const double sp = (vm - p.sigma0) / p.sigma0;  // p_norm_weight = 1 (stress_output_base.cpp:278-281)
      const double ex = (p.rho_exp * sp > 1.e2) ? exp(1.e2) : exp(p.rho_exp * sp);
      num = pow(vm / p.sigma0, p.p_norm) * ex * JxW;
      den = ex * JxW;
    }
  }
  if (p.partial) block_sum2(num, den, p.partial);
}

// chain-rule factor of one quadrature point: d(aggregate) = kappa_q * d(vm_q)
// (functional_state_derivartive_for_elem / functional_sensitivity_for_elem, stress_output_base.cpp:978-1043, 1104-1172)
__device__ __forceinline__ double aggregate_kappa(const StressParams& p, double vm, double JxW) {
  const double Sn = p.sum_num, Sd = p.sum_den;
  const double sp_w = (vm - p.sigma0) / p.sigma0;
  const double sp_s = pow(vm / p.sigma0, p.p_norm), sp1_s = pow(vm / p.sigma0, p.p_norm - 1.);
  double cn, cd;
  if (p.rho_exp * sp_w > 1.e2) {
    cn = p.p_norm * sp1_s * exp(1.e2) / p.sigma0 * JxW;
    cd = 0.;
  } else {
    const double ex = exp(p.rho_exp * sp_w);
    cd = ex * p.rho_exp / p.sigma0 * JxW;
    cn = 1. / p.sigma0 * JxW * ex * (p.p_norm * sp1_s + sp_s * p.rho_exp);
  }
  return p.sigma0 / p.p_norm * pow(Sn / Sd, 1. / p.p_norm - 1.) * (cn / Sd - Sn / (Sd * Sd) * cd);
}

// partial derivative of the aggregate w.r.t. the nodal densities at fixed state: dstress/d(rho_n) = N_n E'(rho_q)/E sigma
// (pattern of structural_element_2d.cpp:664-677), hence d(vm)/d(rho_n) = N_n E'/E vm; point scalar for the nodal gather
__global__ void __launch_bounds__(TB) stress_drho_qp_kernel(const StressParams p) {
  const int64_t t = (int64_t)blockIdx.x * TB + threadIdx.x;
  const int64_t nqp = p.n_elem * p.nq;
  if (t >= nqp) return;
  const int64_t e = t / p.nq;
  const int q = (int)(t - e * p.nq);
  if ((p.elem_mask && !p.elem_mask[e]) || (p.diag_elem && p.diag_elem[e])) { p.scratch[t] = 0.; return; }
  QpGeom g;
  eval_qp(p, e, q, g);
  const double vm = sqrt(vm_pp(g.sig));
  const double kappa = aggregate_kappa(p, vm, g.det * g.wq);
  const double dE = p.E0 * (1. - p.rho_min) * p.penalty * pow(g.rho, p.penalty - 1.);
  p.scratch[t] = kappa * (dE / g.E) * vm;
}

// pass 2: T_q = kappa_q * M_q * J^-T per (element, qp), 9 doubles
__global__ void __launch_bounds__(TB) stress_dq_qp_kernel(const StressParams p) {
  const int64_t t = (int64_t)blockIdx.x * TB + threadIdx.x;
  const int64_t nqp = p.n_elem * p.nq;
  if (t >= nqp) return;
  const int64_t e = t / p.nq;
  const int q = (int)(t - e * p.nq);
  double* T = p.scratch + t * 9;
  if (p.elem_mask && !p.elem_mask[e]) {
#pragma unroll
    for (int i = 0; i < 9; ++i) T[i] = 0.;
    return;
  }
  QpGeom g;
  eval_qp(p, e, q, g);  // the stress follows the strain measure; its derivative uses B_lin (the reference's 2D pattern)
  const double pp = vm_pp(g.sig);
  const double vm = sqrt(pp);
  const double kappa = aggregate_kappa(p, vm, g.det * g.wq);
  double dv[6] = {0., 0., 0., 0., 0., 0.};
  if (fabs(pp) > 0.) {
    const double h = 0.5 / sqrt(pp);
    const double a = g.sig[0] - g.sig[1], b = g.sig[1] - g.sig[2], c = g.sig[2] - g.sig[0];
    dv[0] = (a - c) * h; dv[1] = (b - a) * h; dv[2] = (c - b) * h;
    dv[3] = 6. * g.sig[3] * h; dv[4] = 6. * g.sig[4] * h; dv[5] = 6. * g.sig[5] * h;
  }
  // m = D dv (D symmetric): de_j = m^T B_lin[:, j]
  const double lam = g.E * p.lam1, mu = g.E * p.mu1, tr = dv[0] + dv[1] + dv[2];
  const double m0 = lam * tr + 2. * mu * dv[0], m1 = lam * tr + 2. * mu * dv[1], m2 = lam * tr + 2. * mu * dv[2];
  const double m3 = mu * dv[3], m4 = mu * dv[4], m5 = mu * dv[5];
  const double M[3][3] = {{m0, m3, m5}, {m3, m1, m4}, {m5, m4, m2}};
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int r = 0; r < 3; ++r)
      T[k * 3 + r] = kappa * (M[k][0] * g.inv[0][r] + M[k][1] * g.inv[1][r] + M[k][2] * g.inv[2][r]);
}


// ---- matrix-free internal force with a per-point modulus (linear strain) -------------------------------------------------
// f_a = sum_q JxW B_a^T sigma_q with sigma_q = Ehat_q D_1 eps(u): the element vector of StructuralElement3D::
// internal_residual (solid_element_3d.cpp:135-423) when Ehat = E(rho_q), of ::internal_residual_sensitivity (:566-788, D <-
// dD/dp at :640) when Ehat / D_1 carry the parameter derivative, and of elem_linearized_jacobian_solution_product
// (structural_nonlinear_assembly.cpp:142-166) when u is the perturbation dX.  Same three passes as the aggregate
// derivative: T_q = JxW sigma_q J^-T per point, then the row-owner gather f_a = sum_e sum_q T_q dN_a/dxi(q) in ascending
// element id with constrained dofs zeroed per element (constrain_element_vector).
//   mode 0: Ehat = E(rho_q)                                   (f_int of the state, J dX of a perturbation)
//   mode 1: Ehat = E'(rho_q) sum_b N_b(q) dfield[b]           (nodal-density design variable with filter column dfield)
//   mode 2: Ehat = E(rho_q) / E0                              (parameter E0 of the property card)
//   mode 3: Ehat = E(rho_q), D_1 <- dD_1/dnu (lam1, mu1 of the params are the derivatives)
//   mode 4: sigma_q = E(rho_q) alpha_k (T_q - T_ref) I: the TEMPERATURE volume load B^T D alpha dT of thermal_residual
//           (solid_element_3d.cpp:950-1089, linear strain) as a force vector
//   mode 5: the tangent product K_T(u) du with Green-Lagrange strain (dfield = du as a dof vector): the first Piola-Kirchhoff
//           increment dP = grad(du) S + F dS, dS = E D_1 sym(F^T grad du) -- material tangent + initial-stress stiffness of
//           solid_element_3d.cpp:327-409 applied to a vector; with linear strain it reduces to mode 0 of du
// dP = grad(du) S + F dS with dS = E D_1 sym(F^T grad du) (F = I + H with Green-Lagrange strain; F = I and no initial-stress
// term with linear strain): the increment of the first Piola-Kirchhoff stress along du, written over S
__device__ __forceinline__ void tangent_increment(const StressParams& p, double E, const double (&H)[3][3], const double (&dH)[3][3],
                                                  double (&S)[3][3]) {
  double Dm[3][3];  // F^T grad(du)
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      Dm[i][j] = dH[i][j] + (p.nonlinear ? H[0][i] * dH[0][j] + H[1][i] * dH[1][j] + H[2][i] * dH[2][j] : 0.);
  const double lam = E * p.lam1, mu = E * p.mu1, tr = lam * (Dm[0][0] + Dm[1][1] + Dm[2][2]);
  double dS[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) dS[i][j] = mu * (Dm[i][j] + Dm[j][i]) + (i == j ? tr : 0.);
  double P[3][3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double v = dS[k][c];
      if (p.nonlinear)
        v += H[k][0] * dS[0][c] + H[k][1] * dS[1][c] + H[k][2] * dS[2][c] + dH[k][0] * S[0][c] + dH[k][1] * S[1][c] + dH[k][2] * S[2][c];
      P[k][c] = v;
    }
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int c = 0; c < 3; ++c) S[k][c] = P[k][c];
}

__global__ void __launch_bounds__(TB) force_qp_kernel(const StressParams p, int mode, const double* __restrict__ dfield) {
  const int64_t t = (int64_t)blockIdx.x * TB + threadIdx.x;
  const int64_t nqp = p.n_elem * p.nq;
  if (t >= nqp) return;
  const int64_t e = t / p.nq;
  const int q = (int)(t - e * p.nq);
  double* T = p.scratch + t * 9;
  if (p.diag_elem && p.diag_elem[e]) {
#pragma unroll
    for (int i = 0; i < 9; ++i) T[i] = 0.;
    return;
  }
  QpGeom g;
  eval_qp(p, e, q, g);  // sig = E(rho_q) D_1 eps with the lam1 / mu1 of the params
  double f = g.det * g.wq;
  if (mode == 1) {
    const double* Nt = p.tab + (size_t)q * p.nn;
    const int32_t* en = p.conn + e * p.nn;
    double dr = 0.;
    for (int b = 0; b < p.nn; ++b) dr = fma(Nt[b], dfield[en[b]], dr);
    const double dE = p.E0 * (1. - p.rho_min) * p.penalty * pow_pm1(g.rho, p.penalty - 1.);
    f *= dE * dr / g.E;
  } else if (mode == 2) {
    f /= p.E0;
  }
  if (mode == 4) {
    double Tq = p.T_const;
    if (p.temperature) {
      const double* Nt = p.tab + (size_t)q * p.nn;
      const int32_t* en = p.conn + e * p.nn;
      Tq = 0.;
      for (int b = 0; b < p.nn; ++b) Tq = fma(Nt[b], p.temperature[en[b]], Tq);
    }
    const double st = g.E * p.alpha_k * (Tq - p.T_ref);
    g.sig[0] = g.sig[1] = g.sig[2] = st;
    g.sig[3] = g.sig[4] = g.sig[5] = 0.;
  }
  double S[3][3] = {{g.sig[0], g.sig[3], g.sig[5]}, {g.sig[3], g.sig[1], g.sig[4]}, {g.sig[5], g.sig[4], g.sig[2]}};
  if (mode == 5) {
    double dH[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};  // grad(du)
    const double* dNt = p.tab + (size_t)p.nq * p.nn + (size_t)q * p.nn * 3;
    const int32_t* en = p.conn + e * p.nn;
    for (int b = 0; b < p.nn; ++b) {
      const double d0 = dNt[b * 3], d1 = dNt[b * 3 + 1], d2 = dNt[b * 3 + 2];
      double gr[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) gr[c] = d0 * g.inv[c][0] + d1 * g.inv[c][1] + d2 * g.inv[c][2];
      const double* D = dfield + (int64_t)en[b] * p.nv;
      const double v0 = D[0], v1 = D[1], v2 = D[2];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        dH[0][j] = fma(v0, gr[j], dH[0][j]); dH[1][j] = fma(v1, gr[j], dH[1][j]); dH[2][j] = fma(v2, gr[j], dH[2][j]);
      }
    }
    tangent_increment(p, g.E, g.H, dH, S);  // S <- dP
  } else if (p.nonlinear) {  // Green-Lagrange: f_a = JxW B_a(F)^T S = JxW (F S) grad N_a with F = I + H (modes 1-3: S <- dD/dp E_GL)
    double FS[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) FS[k][c] = S[k][c] + (g.H[k][0] * S[0][c] + g.H[k][1] * S[1][c] + g.H[k][2] * S[2][c]);
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) S[k][c] = FS[k][c];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int r = 0; r < 3; ++r) T[k * 3 + r] = f * (S[k][0] * g.inv[0][r] + S[k][1] * g.inv[1][r] + S[k][2] * g.inv[2][r]);
}

// pass 3: dq_dX rows of the owned nodes
__global__ void __launch_bounds__(TB) stress_dq_gather_kernel(const StressParams p) {
  const int64_t r = (int64_t)blockIdx.x * TB + threadIdx.x;
  if (r >= p.n_owned) return;
  const int n = p.owned_nodes[r];
  const unsigned cm = p.cmask ? p.cmask[n] : 0u;
  const int nn = p.nn, nq = p.nq;
  double acc[3] = {0., 0., 0.};
  for (int64_t ke = p.n2e_ptr[n]; ke < p.n2e_ptr[n + 1]; ++ke) {
    const int64_t e = p.n2e[ke];
    if (p.diag_elem && p.diag_elem[e]) continue;
    const int32_t* en = p.conn + e * nn;
    int a = 0;
    for (int b = 0; b < nn; ++b)
      if (en[b] == n) a = b;
    double o[3] = {0., 0., 0.};
    for (int q = 0; q < nq; ++q) {
      const double* T = p.scratch + (e * nq + q) * 9;
      const double* d = p.tab + (size_t)nq * nn + ((size_t)q * nn + a) * 3;
#pragma unroll
      for (int k = 0; k < 3; ++k) o[k] += T[k * 3] * d[0] + T[k * 3 + 1] * d[1] + T[k * 3 + 2] * d[2];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) acc[k] += o[k];
  }
  for (int k = 0; k < p.nv; ++k) p.dq_dX[(int64_t)n * p.nv + k] = (k < 3 && !((cm >> k) & 1u)) ? acc[k] : 0.;
}

__global__ void reduce_pairs_kernel(const double* partial, int n, double* out) {
  // single warp, fixed order: lane-strided partial sums, then a butterfly
  double a = 0., b = 0.;
  for (int i = threadIdx.x; i < n; i += 32) { a += partial[2 * i]; b += partial[2 * i + 1]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (threadIdx.x == 0) { out[0] = a; out[1] = b; }
}

}  // namespace

int stress_blocks(int64_t n_elem, int nq) { return (int)((n_elem * nq + TB - 1) / TB); }

cudaError_t launch_stress_qp(const StressParams& p, cudaStream_t stream) {
  stress_qp_kernel<<<stress_blocks(p.n_elem, p.nq), TB, 0, stream>>>(p);
  return cudaGetLastError();
}
cudaError_t launch_reduce_pairs(const double* partial, int n, double* out, cudaStream_t stream) {
  reduce_pairs_kernel<<<1, 32, 0, stream>>>(partial, n, out);
  return cudaGetLastError();
}
// ---- HEX8 specialisation of the element force (modes 0-4): closed-form reference gradients, one (element, qp) per thread,
// the eight points of an element in adjacent lanes; the point tensors T_q = f (F) sigma J^-T meet in shared memory and lane a
// leaves the element's nodal force f_a = sum_q T_q dN_a/dxi(q) (q ascending): 24 doubles per element instead of 72 per
// element going through HBM, and the gather reads three values per (node, element).
__global__ void __launch_bounds__(256, 2) force_hex8_nodal_kernel(const StressParams p, int mode, const double* __restrict__ dfield) {
  __shared__ double Ts[256 * 9];
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const bool live = t < p.n_elem * 8;
  const int64_t e = live ? (t >> 3) : 0;
  const int q = (int)(t & 7);
  const bool skip = !live || (p.diag_elem && p.diag_elem[e]);
  const double GPq = 0.57735026918962576451;
  const double xi = (q & 1) ? GPq : -GPq, eta = (q & 2) ? GPq : -GPq, zeta = (q & 4) ? GPq : -GPq;
  const double xm = 1. - xi, xp = 1. + xi, ym = 1. - eta, yp = 1. + eta, zm = 1. - zeta, zp = 1. + zeta;
  int en[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) en[b] = p.conn[e * 8 + b];
  double Jm[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool px = b == 1 || b == 2 || b == 5 || b == 6, py = b == 2 || b == 3 || b == 6 || b == 7, pz = b >= 4;
    const double fx = px ? xp : xm, fy = py ? yp : ym, fz = pz ? zp : zm;
    const double d0 = (px ? 0.125 : -0.125) * fy * fz, d1 = (py ? 0.125 : -0.125) * fx * fz, d2 = (pz ? 0.125 : -0.125) * fx * fy;
    const double* X = p.xyz + (int64_t)en[b] * 3;
    const double x0 = X[0], x1 = X[1], x2 = X[2];
    Jm[0][0] = fma(d0, x0, Jm[0][0]); Jm[0][1] = fma(d0, x1, Jm[0][1]); Jm[0][2] = fma(d0, x2, Jm[0][2]);
    Jm[1][0] = fma(d1, x0, Jm[1][0]); Jm[1][1] = fma(d1, x1, Jm[1][1]); Jm[1][2] = fma(d1, x2, Jm[1][2]);
    Jm[2][0] = fma(d2, x0, Jm[2][0]); Jm[2][1] = fma(d2, x1, Jm[2][1]); Jm[2][2] = fma(d2, x2, Jm[2][2]);
  }
  const double c00 = Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1];
  const double c01 = Jm[1][2] * Jm[2][0] - Jm[1][0] * Jm[2][2];
  const double c02 = Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0];
  const double det = Jm[0][0] * c00 + Jm[0][1] * c01 + Jm[0][2] * c02;
  const double idet = 1. / det;
  double inv[3][3];
  inv[0][0] = c00 * idet; inv[1][0] = c01 * idet; inv[2][0] = c02 * idet;
  inv[0][1] = (Jm[0][2] * Jm[2][1] - Jm[0][1] * Jm[2][2]) * idet;
  inv[1][1] = (Jm[0][0] * Jm[2][2] - Jm[0][2] * Jm[2][0]) * idet;
  inv[2][1] = (Jm[0][1] * Jm[2][0] - Jm[0][0] * Jm[2][1]) * idet;
  inv[0][2] = (Jm[0][1] * Jm[1][2] - Jm[0][2] * Jm[1][1]) * idet;
  inv[1][2] = (Jm[0][2] * Jm[1][0] - Jm[0][0] * Jm[1][2]) * idet;
  inv[2][2] = (Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0]) * idet;
  double H[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};
  double rho = 0., dr = 0., tq = 0.;
  double dH[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};  // grad(du), mode 5
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool px = b == 1 || b == 2 || b == 5 || b == 6, py = b == 2 || b == 3 || b == 6 || b == 7, pz = b >= 4;
    const double fx = px ? xp : xm, fy = py ? yp : ym, fz = pz ? zp : zm;
    const double d0 = (px ? 0.125 : -0.125) * fy * fz, d1 = (py ? 0.125 : -0.125) * fx * fz, d2 = (pz ? 0.125 : -0.125) * fx * fy;
    const double Nb = 0.125 * fx * fy * fz;
    double g[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c] = d0 * inv[c][0] + d1 * inv[c][1] + d2 * inv[c][2];
    const int64_t n = en[b];
    const double* U = p.X + n * p.nv;
    const double u0 = U[0], u1 = U[1], u2 = U[2];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      H[0][j] = fma(u0, g[j], H[0][j]); H[1][j] = fma(u1, g[j], H[1][j]); H[2][j] = fma(u2, g[j], H[2][j]);
    }
    if (p.rho) rho = fma(Nb, p.rho[n], rho);
    if (mode == 1) dr = fma(Nb, dfield[n], dr);
    if (mode == 4 && p.temperature) tq = fma(Nb, p.temperature[n], tq);
    if (mode == 5) {
      const double* D = dfield + n * p.nv;
      const double v0 = D[0], v1 = D[1], v2 = D[2];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        dH[0][j] = fma(v0, g[j], dH[0][j]); dH[1][j] = fma(v1, g[j], dH[1][j]); dH[2][j] = fma(v2, g[j], dH[2][j]);
      }
    }
  }
  double E = p.E0;
  if (p.rho) E = p.E0 * (p.rho_min + (1. - p.rho_min) * simp_pow(rho, p.penalty));
  double eps[6];
  if (p.nonlinear) {
    double Eg[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        Eg[i][j] = 0.5 * (H[i][j] + H[j][i] + (H[0][i] * H[0][j] + H[1][i] * H[1][j] + H[2][i] * H[2][j]));
    eps[0] = Eg[0][0]; eps[1] = Eg[1][1]; eps[2] = Eg[2][2];
    eps[3] = Eg[0][1] + Eg[1][0]; eps[4] = Eg[1][2] + Eg[2][1]; eps[5] = Eg[0][2] + Eg[2][0];
  } else {
    eps[0] = H[0][0]; eps[1] = H[1][1]; eps[2] = H[2][2];
    eps[3] = H[0][1] + H[1][0]; eps[4] = H[1][2] + H[2][1]; eps[5] = H[0][2] + H[2][0];
  }
  const double lam = E * p.lam1, mu = E * p.mu1, tr = eps[0] + eps[1] + eps[2];
  const double s0 = lam * tr + 2. * mu * eps[0], s1 = lam * tr + 2. * mu * eps[1], s2 = lam * tr + 2. * mu * eps[2];
  const double s3 = mu * eps[3], s4 = mu * eps[4], s5 = mu * eps[5];
  double f = det;  // Gauss weights are 1 for the 2x2x2 rule
  if (mode == 1) f *= p.E0 * (1. - p.rho_min) * p.penalty * pow_pm1(rho, p.penalty - 1.) * dr / E;
  else if (mode == 2) f /= p.E0;
  if (skip) f = 0.;
  double S[3][3] = {{s0, s3, s5}, {s3, s1, s4}, {s5, s4, s2}};
  if (mode == 4) {  // the TEMPERATURE volume load as a force vector: sigma = E alpha_k (T - T_ref) I (linear strain)
    const double st = E * p.alpha_k * ((p.temperature ? tq : p.T_const) - p.T_ref);
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) S[k][c] = k == c ? st : 0.;
  } else if (mode == 5) {
    tangent_increment(p, E, H, dH, S);
  } else if (p.nonlinear) {
    double FS[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) FS[k][c] = S[k][c] + (H[k][0] * S[0][c] + H[k][1] * S[1][c] + H[k][2] * S[2][c]);
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) S[k][c] = FS[k][c];
  }
  double* T = Ts + threadIdx.x * 9;
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int r = 0; r < 3; ++r) T[k * 3 + r] = f * (S[k][0] * inv[0][r] + S[k][1] * inv[1][r] + S[k][2] * inv[2][r]);
  __syncthreads();
  // lane a of the element's eight: f_a[k] = sum_q sum_r T_q[k][r] dN_a/dxi_r(q)
  const int a = q;
  const bool ax = a == 1 || a == 2 || a == 5 || a == 6, ay = a == 2 || a == 3 || a == 6 || a == 7, az = a >= 4;
  double o[3] = {0., 0., 0.};
#pragma unroll
  for (int qq = 0; qq < 8; ++qq) {
    const double* Tq = Ts + ((threadIdx.x & ~7) | qq) * 9;
    const double fx = ((qq & 1) != 0) == ax ? 1. + GPq : 1. - GPq;
    const double fy = ((qq & 2) != 0) == ay ? 1. + GPq : 1. - GPq;
    const double fz = ((qq & 4) != 0) == az ? 1. + GPq : 1. - GPq;
    const double d0 = (ax ? 0.125 : -0.125) * fy * fz, d1 = (ay ? 0.125 : -0.125) * fx * fz, d2 = (az ? 0.125 : -0.125) * fx * fy;
#pragma unroll
    for (int k = 0; k < 3; ++k) o[k] += Tq[k * 3] * d0 + Tq[k * 3 + 1] * d1 + Tq[k * 3 + 2] * d2;
  }
  if (live) {
    double* dst = p.scratch + e * 24 + a * 3;
    dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];
  }
}

// row-owner gather of the nodal element forces: adjacent elements in ascending id, constrained dofs zeroed
__global__ void __launch_bounds__(TB) force_hex8_gather_kernel(const StressParams p) {
  const int64_t r = (int64_t)blockIdx.x * TB + threadIdx.x;
  if (r >= p.n_owned) return;
  const int n = p.owned_nodes[r];
  const unsigned cm = p.cmask ? p.cmask[n] : 0u;
  double acc[3] = {0., 0., 0.};
  for (int64_t ke = p.n2e_ptr[n]; ke < p.n2e_ptr[n + 1]; ++ke) {
    const int64_t e = p.n2e[ke];
    const int32_t* en = p.conn + e * 8;
    int a = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b)
      if (en[b] == n) a = b;
    const double* o = p.scratch + e * 24 + a * 3;
    acc[0] += o[0]; acc[1] += o[1]; acc[2] += o[2];
  }
  for (int k = 0; k < p.nv; ++k) p.dq_dX[(int64_t)n * p.nv + k] = (k < 3 && !((cm >> k) & 1u)) ? acc[k] : 0.;
}

cudaError_t launch_internal_force(const StressParams& p, int mode, const double* dfield, cudaStream_t stream) {
  if (p.nn == 8) {  // HEX8: nodal element forces (24 doubles per element) instead of point tensors
    force_hex8_nodal_kernel<<<(int)((p.n_elem * 8 + 255) / 256), 256, 0, stream>>>(p, mode, dfield);
    cudaError_t e8 = cudaGetLastError();
    if (e8 != cudaSuccess) return e8;
    force_hex8_gather_kernel<<<(int)((p.n_owned + TB - 1) / TB), TB, 0, stream>>>(p);
    return cudaGetLastError();
  }
  force_qp_kernel<<<stress_blocks(p.n_elem, p.nq), TB, 0, stream>>>(p, mode, dfield);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  stress_dq_gather_kernel<<<(int)((p.n_owned + TB - 1) / TB), TB, 0, stream>>>(p);
  return cudaGetLastError();
}
cudaError_t launch_stress_derivative(const StressParams& p, cudaStream_t stream) {
  stress_dq_qp_kernel<<<stress_blocks(p.n_elem, p.nq), TB, 0, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  stress_dq_gather_kernel<<<(int)((p.n_owned + TB - 1) / TB), TB, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace mastb200

// ---- density sensitivities alone (SURVEY 8a row S without R/J): G_n = adj . dR/d(rho_n) for every owned node ------------
// Device counterpart of NonlinearImplicitAssembly::sensitivity_assemble (/root/reference/src/base/
// nonlinear_implicit_assembly.cpp:515-686) + StructuralElement3D::internal_residual_sensitivity
// (src/elasticity/solid_element_3d.cpp:566-788, D <- dD/dp at :640) + the dot products of
// AssemblyBase::calculate_output_adjoint_sensitivity_multiple_parameters_no_direct (src/base/assembly_base.cpp:594-639)
// for nodal-density design variables under the SIMP law (examples/structural/example_6/example_6.cpp:100-137), linear
// strain:  adj . dR_e/d(rho_n) = sum_q N_n(q) JxW E'(rho_q) eps(adj)^T D_1 eps(u), adj zeroed on constrained dofs
// (constrain_element_vector).  Pass 1: one thread per (element, qp) -> the scalar s_q; pass 2: row-owner gather,
// one thread per owned node over its adjacent elements in ascending id (the reference's ADD order).
namespace mastb200 {
namespace {

__global__ void __launch_bounds__(128) sens_qp_kernel(const StressParams p, const double* __restrict__ adj) {
  const int64_t t = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (t >= p.n_elem * p.nq) return;
  const int64_t e = t / p.nq;
  const int q = (int)(t - e * p.nq);
  if (p.diag_elem && p.diag_elem[e]) { p.scratch[t] = 0.; return; }
  const int nn = p.nn, nq = p.nq;
  const double* Nt = p.tab + (size_t)q * nn;
  const double* dNt = p.tab + (size_t)nq * nn + (size_t)q * nn * 3;
  const int32_t* en = p.conn + e * nn;
  double Jm[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};
  for (int b = 0; b < nn; ++b) {
    const double* X = p.xyz + (int64_t)en[b] * 3;
    const double d0 = dNt[b * 3], d1 = dNt[b * 3 + 1], d2 = dNt[b * 3 + 2];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      Jm[0][c] = fma(d0, X[c], Jm[0][c]); Jm[1][c] = fma(d1, X[c], Jm[1][c]); Jm[2][c] = fma(d2, X[c], Jm[2][c]);
    }
  }
  const double c00 = Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1];
  const double c01 = Jm[1][2] * Jm[2][0] - Jm[1][0] * Jm[2][2];
  const double c02 = Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0];
  const double det = Jm[0][0] * c00 + Jm[0][1] * c01 + Jm[0][2] * c02;
  const double idet = 1. / det;
  double inv[3][3];
  inv[0][0] = c00 * idet; inv[1][0] = c01 * idet; inv[2][0] = c02 * idet;
  inv[0][1] = (Jm[0][2] * Jm[2][1] - Jm[0][1] * Jm[2][2]) * idet;
  inv[1][1] = (Jm[0][0] * Jm[2][2] - Jm[0][2] * Jm[2][0]) * idet;
  inv[2][1] = (Jm[0][1] * Jm[2][0] - Jm[0][0] * Jm[2][1]) * idet;
  inv[0][2] = (Jm[0][1] * Jm[1][2] - Jm[0][2] * Jm[1][1]) * idet;
  inv[1][2] = (Jm[0][2] * Jm[1][0] - Jm[0][0] * Jm[1][2]) * idet;
  inv[2][2] = (Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0]) * idet;
  double H[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}}, A[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};
  double rho = 0.;
  for (int b = 0; b < nn; ++b) {
    const double d0 = dNt[b * 3], d1 = dNt[b * 3 + 1], d2 = dNt[b * 3 + 2];
    double g[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c] = d0 * inv[c][0] + d1 * inv[c][1] + d2 * inv[c][2];
    const int64_t n = en[b];
    const double* U = p.X + n * p.nv;
    const double* L = adj + n * p.nv;
    const unsigned cm = p.cmask ? p.cmask[n] : 0u;
    const double u0 = U[0], u1 = U[1], u2 = U[2];
    const double l0 = (cm & 1u) ? 0. : L[0], l1 = (cm & 2u) ? 0. : L[1], l2 = (cm & 4u) ? 0. : L[2];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      H[0][j] = fma(u0, g[j], H[0][j]); H[1][j] = fma(u1, g[j], H[1][j]); H[2][j] = fma(u2, g[j], H[2][j]);
      A[0][j] = fma(l0, g[j], A[0][j]); A[1][j] = fma(l1, g[j], A[1][j]); A[2][j] = fma(l2, g[j], A[2][j]);
    }
    rho = fma(Nt[b], p.rho[n], rho);
  }
  double e0 = H[0][0], e1 = H[1][1], e2 = H[2][2];
  double e3 = H[0][1] + H[1][0], e4 = H[1][2] + H[2][1], e5 = H[0][2] + H[2][0];
  double a0 = A[0][0], a1 = A[1][1], a2 = A[2][2];
  double a3 = A[0][1] + A[1][0], a4 = A[1][2] + A[2][1], a5 = A[0][2] + A[2][0];
  if (p.nonlinear) {
    // Green-Lagrange strain E = 1/2 (H + H^T + H^T H) of the state and its variation along the adjoint,
    // sum_a B_a(F) adj_a = sym(F^T grad adj) with F = I + H (solid_element_3d.cpp:566-788 with the operators of :1308-1463)
    double Eg[3][3], Dm[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        Eg[i][j] = 0.5 * (H[i][j] + H[j][i] + (H[0][i] * H[0][j] + H[1][i] * H[1][j] + H[2][i] * H[2][j]));
        Dm[i][j] = A[i][j] + (H[0][i] * A[0][j] + H[1][i] * A[1][j] + H[2][i] * A[2][j]);  // (F^T grad adj)_ij
      }
    e0 = Eg[0][0]; e1 = Eg[1][1]; e2 = Eg[2][2];
    e3 = Eg[0][1] + Eg[1][0]; e4 = Eg[1][2] + Eg[2][1]; e5 = Eg[0][2] + Eg[2][0];
    a0 = Dm[0][0]; a1 = Dm[1][1]; a2 = Dm[2][2];
    a3 = Dm[0][1] + Dm[1][0]; a4 = Dm[1][2] + Dm[2][1]; a5 = Dm[0][2] + Dm[2][0];
  }
  const double dens = p.lam1 * (a0 + a1 + a2) * (e0 + e1 + e2) +
                      p.mu1 * (2. * (a0 * e0 + a1 * e1 + a2 * e2) + a3 * e3 + a4 * e4 + a5 * e5);
  const double dE = p.E0 * (1. - p.rho_min) * p.penalty * pow(rho, p.penalty - 1.);
  p.scratch[t] = det * p.tab[(size_t)nq * nn * 4 + q] * dE * dens;
}

// HEX8 specialisation: reference gradients from the node signs (no table loads), fully unrolled, one (element, qp) per
// thread with the eight qps of an element in adjacent lanes (their node loads coalesce into broadcasts)
__device__ __forceinline__ double pow_penalty_m1(double rho, double pm1) {
  // E'(rho) needs rho^(penalty-1): small integer exponents (the usual SIMP penalties) by multiplication
  if (pm1 == 3.) return rho * rho * rho;
  if (pm1 == 2.) return rho * rho;
  if (pm1 == 1.) return rho;
  if (pm1 == 0.) return 1.;
  return pow(rho, pm1);
}

__global__ void __launch_bounds__(256, 2) sens_qp_hex8_kernel(const StressParams p, const double* __restrict__ adj) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const bool live = t < p.n_elem * 8;
  const int64_t e = live ? (t >> 3) : 0;
  const int q = (int)(t & 7);
  const bool skip = !live || (p.diag_elem && p.diag_elem[e]);
  const double GPq = 0.57735026918962576451;
  const double xi = (q & 1) ? GPq : -GPq, eta = (q & 2) ? GPq : -GPq, zeta = (q & 4) ? GPq : -GPq;
  const double xm = 1. - xi, xp = 1. + xi, ym = 1. - eta, yp = 1. + eta, zm = 1. - zeta, zp = 1. + zeta;
  int en[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) en[b] = p.conn[e * 8 + b];
  double Jm[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool px = b == 1 || b == 2 || b == 5 || b == 6, py = b == 2 || b == 3 || b == 6 || b == 7, pz = b >= 4;
    const double fx = px ? xp : xm, fy = py ? yp : ym, fz = pz ? zp : zm;
    const double d0 = (px ? 0.125 : -0.125) * fy * fz, d1 = (py ? 0.125 : -0.125) * fx * fz, d2 = (pz ? 0.125 : -0.125) * fx * fy;
    const double* X = p.xyz + (int64_t)en[b] * 3;
    const double x0 = X[0], x1 = X[1], x2 = X[2];
    Jm[0][0] = fma(d0, x0, Jm[0][0]); Jm[0][1] = fma(d0, x1, Jm[0][1]); Jm[0][2] = fma(d0, x2, Jm[0][2]);
    Jm[1][0] = fma(d1, x0, Jm[1][0]); Jm[1][1] = fma(d1, x1, Jm[1][1]); Jm[1][2] = fma(d1, x2, Jm[1][2]);
    Jm[2][0] = fma(d2, x0, Jm[2][0]); Jm[2][1] = fma(d2, x1, Jm[2][1]); Jm[2][2] = fma(d2, x2, Jm[2][2]);
  }
  const double c00 = Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1];
  const double c01 = Jm[1][2] * Jm[2][0] - Jm[1][0] * Jm[2][2];
  const double c02 = Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0];
  const double det = Jm[0][0] * c00 + Jm[0][1] * c01 + Jm[0][2] * c02;
  const double idet = 1. / det;
  double inv[3][3];
  inv[0][0] = c00 * idet; inv[1][0] = c01 * idet; inv[2][0] = c02 * idet;
  inv[0][1] = (Jm[0][2] * Jm[2][1] - Jm[0][1] * Jm[2][2]) * idet;
  inv[1][1] = (Jm[0][0] * Jm[2][2] - Jm[0][2] * Jm[2][0]) * idet;
  inv[2][1] = (Jm[0][1] * Jm[2][0] - Jm[0][0] * Jm[2][1]) * idet;
  inv[0][2] = (Jm[0][1] * Jm[1][2] - Jm[0][2] * Jm[1][1]) * idet;
  inv[1][2] = (Jm[0][2] * Jm[1][0] - Jm[0][0] * Jm[1][2]) * idet;
  inv[2][2] = (Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0]) * idet;
  double H[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}}, A[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};
  double rho = 0.;
  const bool same = adj == p.X;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool px = b == 1 || b == 2 || b == 5 || b == 6, py = b == 2 || b == 3 || b == 6 || b == 7, pz = b >= 4;
    const double fx = px ? xp : xm, fy = py ? yp : ym, fz = pz ? zp : zm;
    const double d0 = (px ? 0.125 : -0.125) * fy * fz, d1 = (py ? 0.125 : -0.125) * fx * fz, d2 = (pz ? 0.125 : -0.125) * fx * fy;
    const double Nb = 0.125 * fx * fy * fz;
    double g[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c] = d0 * inv[c][0] + d1 * inv[c][1] + d2 * inv[c][2];
    const int64_t n = en[b];
    const double* U = p.X + n * p.nv;
    const unsigned cm = p.cmask ? p.cmask[n] : 0u;
    const double u0 = U[0], u1 = U[1], u2 = U[2];
    double l0 = u0, l1 = u1, l2 = u2;
    if (!same) { const double* L = adj + n * p.nv; l0 = L[0]; l1 = L[1]; l2 = L[2]; }
    if (cm & 1u) l0 = 0.;
    if (cm & 2u) l1 = 0.;
    if (cm & 4u) l2 = 0.;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      H[0][j] = fma(u0, g[j], H[0][j]); H[1][j] = fma(u1, g[j], H[1][j]); H[2][j] = fma(u2, g[j], H[2][j]);
      A[0][j] = fma(l0, g[j], A[0][j]); A[1][j] = fma(l1, g[j], A[1][j]); A[2][j] = fma(l2, g[j], A[2][j]);
    }
    rho = fma(Nb, p.rho[n], rho);
  }
  const double e0 = H[0][0], e1 = H[1][1], e2 = H[2][2];
  const double e3 = H[0][1] + H[1][0], e4 = H[1][2] + H[2][1], e5 = H[0][2] + H[2][0];
  const double a0 = A[0][0], a1 = A[1][1], a2 = A[2][2];
  const double a3 = A[0][1] + A[1][0], a4 = A[1][2] + A[2][1], a5 = A[0][2] + A[2][0];
  const double dens = p.lam1 * (a0 + a1 + a2) * (e0 + e1 + e2) +
                      p.mu1 * (2. * (a0 * e0 + a1 * e1 + a2 * e2) + a3 * e3 + a4 * e4 + a5 * e5);
  const double dE = p.E0 * (1. - p.rho_min) * p.penalty * pow_penalty_m1(rho, p.penalty - 1.);
  const double sq = skip ? 0. : det * dE * dens;
  // the eight lanes of the element turn the point values into nodal ones: lane a ends with sum_q N_a(q) s_q (q ascending)
  const int a = q, lane = threadIdx.x & 31;
  const bool ax = a == 1 || a == 2 || a == 5 || a == 6, ay = a == 2 || a == 3 || a == 6 || a == 7, az = a >= 4;
  double acc = 0.;
#pragma unroll
  for (int qq = 0; qq < 8; ++qq) {
    const double s = __shfl_sync(0xffffffffu, sq, (lane & ~7) | qq);
    const double fx = ((qq & 1) != 0) == ax ? 1. + GPq : 1. - GPq;
    const double fy = ((qq & 2) != 0) == ay ? 1. + GPq : 1. - GPq;
    const double fz = ((qq & 4) != 0) == az ? 1. + GPq : 1. - GPq;
    acc = fma(0.125 * fx * fy * fz, s, acc);
  }
  if (live) p.scratch[t] = acc;  // [element][node]
}

__global__ void __launch_bounds__(128) sens_gather_kernel(const StressParams p, double* __restrict__ G) {
  const int64_t r = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (r >= p.n_owned) return;
  const int n = p.owned_nodes[r];
  const int nn = p.nn, nq = p.nq;
  double acc = 0.;
  for (int64_t ke = p.n2e_ptr[n]; ke < p.n2e_ptr[n + 1]; ++ke) {
    const int64_t e = p.n2e[ke];
    const int32_t* en = p.conn + e * nn;
    int a = 0;
    for (int b = 0; b < nn; ++b)
      if (en[b] == n) a = b;
    double o = 0.;
    if (p.scratch_nodal) o = p.scratch[e * 8 + a];  // already summed over the points (sens_qp_hex8_kernel)
    else
      for (int q = 0; q < nq; ++q) o = fma(p.tab[(size_t)q * nn + a], p.scratch[e * nq + q], o);
    acc += o;
  }
  G[n] = acc;
}

}  // namespace

cudaError_t launch_density_sensitivities(const StressParams& p, const double* adj, double* G, cudaStream_t stream) {
  StressParams q = p;
  const bool fast8 = p.nn == 8 && !p.nonlinear;  // the HEX8 specialisation is the linear-strain kernel
  q.scratch_nodal = fast8 ? 1 : 0;
  if (fast8) sens_qp_hex8_kernel<<<(int)((p.n_elem * 8 + 255) / 256), 256, 0, stream>>>(q, adj);
  else sens_qp_kernel<<<stress_blocks(p.n_elem, p.nq), 128, 0, stream>>>(q, adj);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  sens_gather_kernel<<<(int)((p.n_owned + 127) / 128), 128, 0, stream>>>(q, G);
  return cudaGetLastError();
}

cudaError_t launch_stress_density_derivative(const StressParams& p, double* G, cudaStream_t stream) {
  StressParams q = p;
  q.scratch_nodal = 0;
  stress_drho_qp_kernel<<<stress_blocks(p.n_elem, p.nq), TB, 0, stream>>>(q);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  sens_gather_kernel<<<(int)((p.n_owned + 127) / 128), 128, 0, stream>>>(q, G);
  return cudaGetLastError();
}

}  // namespace mastb200
