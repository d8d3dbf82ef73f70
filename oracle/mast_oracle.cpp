// mast_oracle.cpp -- CPU ORACLE (test infrastructure only; see mast_oracle.h for the contract).
//
// A scalar, dynamic-size restatement of the reference's element-assembly + sensitivity path.
// Loop and summation orders follow the cited reference code so that the oracle is the closest
// thing to "MAST on one MPI rank" that can be built without libMesh/PETSc/Eigen:
//   * element dof order is variable-major (src/numerics/fem_operator_matrix.h:277-296)
//   * quadrature points ascending inside an element, elements in id order into each CSR entry
//     (src/base/nonlinear_implicit_assembly.cpp:111-186)
// Paths in comments are relative to /root/reference.
#include "mast_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <set>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// tiny dense helpers (stand-ins for Eigen RealMatrixX/RealVectorX, row-major, plain loops)
// ---------------------------------------------------------------------------------------------
struct Mat {
  int r = 0, c = 0;
  std::vector<double> a;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), a(size_t(r_) * c_, 0.0) {}
  double& operator()(int i, int j) { return a[size_t(i) * c + j]; }
  double operator()(int i, int j) const { return a[size_t(i) * c + j]; }
  void zero() { std::fill(a.begin(), a.end(), 0.0); }
  void resize(int r_, int c_) { r = r_; c = c_; a.assign(size_t(r_) * c_, 0.0); }
};
typedef std::vector<double> Vec;

static Mat matmul(const Mat& A, const Mat& B) {  // C = A*B
  Mat C(A.r, B.c);
  for (int i = 0; i < A.r; ++i)
    for (int j = 0; j < B.c; ++j) {
      double s = 0.;
      for (int k = 0; k < A.c; ++k) s += A(i, k) * B(k, j);
      C(i, j) = s;
    }
  return C;
}
static Mat matmul_tn(const Mat& A, const Mat& B) {  // C = A^T * B
  Mat C(A.c, B.c);
  for (int i = 0; i < A.c; ++i)
    for (int j = 0; j < B.c; ++j) {
      double s = 0.;
      for (int k = 0; k < A.r; ++k) s += A(k, i) * B(k, j);
      C(i, j) = s;
    }
  return C;
}
static Vec matvec(const Mat& A, const Vec& x) {
  Vec y(A.r, 0.);
  for (int i = 0; i < A.r; ++i) {
    double s = 0.;
    for (int k = 0; k < A.c; ++k) s += A(i, k) * x[k];
    y[i] = s;
  }
  return y;
}
static Vec matvec_t(const Mat& A, const Vec& x) {  // A^T x
  Vec y(A.c, 0.);
  for (int i = 0; i < A.c; ++i) {
    double s = 0.;
    for (int k = 0; k < A.r; ++k) s += A(k, i) * x[k];
    y[i] = s;
  }
  return y;
}

// ---------------------------------------------------------------------------------------------
// FEMOperatorMatrix  (src/numerics/fem_operator_matrix.h:33-467)
// Block-sparse operator: rows = interpolated variables, columns = (discrete var, dof) pairs,
// column index j*n_dofs_per_var+k.  Each product below keeps the reference's loop nest
// (interpolated var outer, discrete var next, then dof) because that fixes the summation order.
// ---------------------------------------------------------------------------------------------
class OperatorMatrix {
 public:
  void reinit(int n_interp, int n_disc, int n_dofs) {  // fem_operator_matrix.h:207-221
    ni_ = n_interp; nd_ = n_disc; nk_ = n_dofs;
    has_.assign(size_t(ni_) * nd_, 0);
    blk_.assign(size_t(ni_) * nd_ * nk_, 0.);
  }
  void reinit_diag(int n_vars, const Vec& shape) {  // fem_operator_matrix.h:253-272
    reinit(n_vars, n_vars, int(shape.size()));
    for (int i = 0; i < n_vars; ++i) set_shape_function(i, i, shape);
  }
  void set_shape_function(int interp, int disc, const Vec& shape) {  // :225-249
    has_[size_t(disc) * ni_ + interp] = 1;
    std::copy(shape.begin(), shape.end(), blk_.begin() + (size_t(disc) * ni_ + interp) * nk_);
  }
  int m() const { return ni_; }
  int n() const { return nd_ * nk_; }

  void vector_mult(Vec& res, const Vec& v) const {  // :276-296
    res.assign(ni_, 0.);
    for (int i = 0; i < ni_; ++i)
      for (int j = 0; j < nd_; ++j)
        if (has(i, j)) {
          const double* s = blk(i, j);
          for (int k = 0; k < nk_; ++k) res[i] += s[k] * v[j * nk_ + k];
        }
  }
  void vector_mult_transpose(Vec& res, const Vec& v) const {  // :299-319
    res.assign(n(), 0.);
    for (int i = 0; i < ni_; ++i)
      for (int j = 0; j < nd_; ++j)
        if (has(i, j)) {
          const double* s = blk(i, j);
          for (int k = 0; k < nk_; ++k) res[j * nk_ + k] += s[k] * v[i];
        }
  }
  void right_multiply(Mat& r, const Mat& mm) const {  // r = B * m   :323-346
    r.resize(ni_, mm.c);
    for (int i = 0; i < ni_; ++i)
      for (int j = 0; j < nd_; ++j)
        if (has(i, j)) {
          const double* s = blk(i, j);
          for (int l = 0; l < mm.c; ++l)
            for (int k = 0; k < nk_; ++k) r(i, l) += s[k] * mm(j * nk_ + k, l);
        }
  }
  void right_multiply_transpose(Mat& r, const Mat& mm) const {  // r = B^T * m   :351-374
    r.resize(n(), mm.c);
    for (int i = 0; i < ni_; ++i)
      for (int j = 0; j < nd_; ++j)
        if (has(i, j)) {
          const double* s = blk(i, j);
          for (int l = 0; l < mm.c; ++l)
            for (int k = 0; k < nk_; ++k) r(j * nk_ + k, l) += s[k] * mm(i, l);
        }
  }
  void right_multiply_transpose(Mat& r, const OperatorMatrix& o) const {  // r = B^T * O   :378-406
    r.resize(n(), o.n());
    for (int i = 0; i < nd_; ++i)
      for (int j = 0; j < o.nd_; ++j)
        for (int k = 0; k < ni_; ++k)
          if (has(k, i) && o.has(k, j)) {
            const double *n1 = blk(k, i), *n2 = o.blk(k, j);
            for (int a = 0; a < nk_; ++a)
              for (int b = 0; b < o.nk_; ++b) r(i * nk_ + a, j * o.nk_ + b) += n1[a] * n2[b];
          }
  }
  void left_multiply(Mat& r, const Mat& mm) const {  // r = m * B   :411-434
    r.resize(mm.r, n());
    for (int i = 0; i < ni_; ++i)
      for (int j = 0; j < nd_; ++j)
        if (has(i, j)) {
          const double* s = blk(i, j);
          for (int l = 0; l < mm.r; ++l)
            for (int k = 0; k < nk_; ++k) r(l, j * nk_ + k) += s[k] * mm(l, i);
        }
  }

 private:
  bool has(int interp, int disc) const { return has_[size_t(disc) * ni_ + interp] != 0; }
  const double* blk(int interp, int disc) const { return &blk_[(size_t(disc) * ni_ + interp) * nk_]; }
  int ni_ = 0, nd_ = 0, nk_ = 0;
  std::vector<char> has_;
  Vec blk_;
};

// ---------------------------------------------------------------------------------------------
// libMesh FE restatement (un-vendored dependency; libMesh 1.5.1 per .travis.yml).
// Lagrange shape functions on [-1,1]^d with libMesh node ordering, Gauss tensor rules with the
// default order 2p+1 (fe_base.cpp:103-105 -> FEType::default_quadrature_rule), i fastest.
// ---------------------------------------------------------------------------------------------
static const double GP2[2] = {-0.57735026918962576451, 0.57735026918962576451};
static const double GW2[2] = {1., 1.};
static const double GP3[3] = {-0.77459666924148337704, 0., 0.77459666924148337704};
static const double GW3[3] = {5. / 9., 8. / 9., 5. / 9.};

// 1D Lagrange: linear idx 0 -> node at -1, 1 -> node at +1; quadratic adds idx 2 -> node at 0
static double lag1(int order, int idx, double x) {
  if (order == 1) return idx == 0 ? 0.5 * (1. - x) : 0.5 * (1. + x);
  switch (idx) {
    case 0: return 0.5 * x * (x - 1.);
    case 1: return 0.5 * x * (x + 1.);
    default: return (1. - x * x);
  }
}
static double dlag1(int order, int idx, double x) {
  if (order == 1) return idx == 0 ? -0.5 : 0.5;
  switch (idx) {
    case 0: return x - 0.5;
    case 1: return x + 0.5;
    default: return -2. * x;
  }
}
// tensor index of each node along xi/eta/zeta
static const int H8_I[8] = {0, 1, 1, 0, 0, 1, 1, 0};
static const int H8_J[8] = {0, 0, 1, 1, 0, 0, 1, 1};
static const int H8_K[8] = {0, 0, 0, 0, 1, 1, 1, 1};
static const int H27_I[27] = {0, 1, 1, 0, 0, 1, 1, 0, 2, 1, 2, 0, 0, 1, 1, 0, 2, 1, 2, 0, 2, 2, 1, 2, 0, 2, 2};
static const int H27_J[27] = {0, 0, 1, 1, 0, 0, 1, 1, 0, 2, 1, 2, 0, 0, 1, 1, 0, 2, 1, 2, 2, 0, 2, 1, 2, 2, 2};
static const int H27_K[27] = {0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 2, 2, 2, 2, 1, 1, 1, 1, 0, 2, 2, 2, 2, 1, 2};
static const int Q4_I[4] = {0, 1, 1, 0};
static const int Q4_J[4] = {0, 0, 1, 1};
static const int Q9_I[9] = {0, 1, 1, 0, 2, 1, 2, 0, 2};
static const int Q9_J[9] = {0, 0, 1, 1, 0, 2, 1, 2, 2};
// libMesh side -> element node maps (outward-normal ordering)
static const int H8_SIDE[6][4] = {{0, 3, 2, 1}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};
static const int H27_SIDE[6][9] = {{0, 3, 2, 1, 11, 10, 9, 8, 20},  {0, 1, 5, 4, 8, 13, 16, 12, 21},
                                   {1, 2, 6, 5, 9, 14, 17, 13, 22}, {2, 3, 7, 6, 10, 15, 18, 14, 23},
                                   {3, 0, 4, 7, 11, 12, 19, 15, 24}, {4, 5, 6, 7, 16, 17, 18, 19, 25}};

struct ElemInfo {
  int nn, dim, order, n1d, nq;
};
static ElemInfo elem_info(int type) {
  switch (type) {
    case ORACLE_HEX8: return {8, 3, 1, 2, 8};
    case ORACLE_HEX27: return {27, 3, 2, 3, 27};
    case ORACLE_QUAD4: return {4, 2, 1, 2, 4};
  }
  return {0, 0, 0, 0, 0};
}

// what MAST::FEBase exposes after init (fe_base.cpp:64-131): phi, dphi (physical), JxW, xyz
struct FE {
  int nn = 0, nq = 0, dim = 0;
  std::vector<double> phi;   // [a][qp]
  std::vector<double> dphi;  // [a][qp][3]
  std::vector<double> JxW;   // [qp]
  std::vector<double> xyz;   // [qp][3]
  std::vector<double> normal;  // side FE only: [qp][3]
  double N(int a, int q) const { return phi[size_t(a) * nq + q]; }
  double dN(int a, int q, int d) const { return dphi[(size_t(a) * nq + q) * 3 + d]; }
};

static void fe_init(int type, const double* X /*[nn][3]*/, FE& fe) {
  const ElemInfo ei = elem_info(type);
  fe.nn = ei.nn; fe.nq = ei.nq; fe.dim = ei.dim;
  fe.phi.assign(size_t(ei.nn) * ei.nq, 0.);
  fe.dphi.assign(size_t(ei.nn) * ei.nq * 3, 0.);
  fe.JxW.assign(ei.nq, 0.);
  fe.xyz.assign(size_t(ei.nq) * 3, 0.);
  const double* gp = ei.n1d == 2 ? GP2 : GP3;
  const double* gw = ei.n1d == 2 ? GW2 : GW3;
  const int *I, *J, *K = nullptr;
  if (type == ORACLE_HEX8) { I = H8_I; J = H8_J; K = H8_K; }
  else if (type == ORACLE_HEX27) { I = H27_I; J = H27_J; K = H27_K; }
  else { I = Q4_I; J = Q4_J; }
  std::vector<double> dref(size_t(ei.nn) * 3);
  int q = 0;
  const int nk = ei.dim == 3 ? ei.n1d : 1;
  for (int kq = 0; kq < nk; ++kq)
    for (int jq = 0; jq < ei.n1d; ++jq)
      for (int iq = 0; iq < ei.n1d; ++iq, ++q) {
        const double xi = gp[iq], eta = gp[jq], zeta = ei.dim == 3 ? gp[kq] : 0.;
        const double w = gw[iq] * gw[jq] * (ei.dim == 3 ? gw[kq] : 1.);
        double Jm[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};  // Jm[r][c] = d x_c / d xi_r
        double pt[3] = {0, 0, 0};
        for (int a = 0; a < ei.nn; ++a) {
          const double li = lag1(ei.order, I[a], xi), lj = lag1(ei.order, J[a], eta);
          const double lk = ei.dim == 3 ? lag1(ei.order, K[a], zeta) : 1.;
          const double N = li * lj * lk;
          fe.phi[size_t(a) * ei.nq + q] = N;
          double* d = &dref[size_t(a) * 3];
          d[0] = dlag1(ei.order, I[a], xi) * lj * lk;
          d[1] = li * dlag1(ei.order, J[a], eta) * lk;
          d[2] = ei.dim == 3 ? li * lj * dlag1(ei.order, K[a], zeta) : 0.;
          for (int c = 0; c < 3; ++c) {
            pt[c] += N * X[a * 3 + c];
            for (int r = 0; r < ei.dim; ++r) Jm[r][c] += d[r] * X[a * 3 + c];
          }
        }
        for (int c = 0; c < 3; ++c) fe.xyz[size_t(q) * 3 + c] = pt[c];
        if (ei.dim == 3) {
          const double dx_dxi = Jm[0][0], dy_dxi = Jm[0][1], dz_dxi = Jm[0][2];
          const double dx_deta = Jm[1][0], dy_deta = Jm[1][1], dz_deta = Jm[1][2];
          const double dx_dzeta = Jm[2][0], dy_dzeta = Jm[2][1], dz_dzeta = Jm[2][2];
          const double det = dx_dxi * (dy_deta * dz_dzeta - dz_deta * dy_dzeta) +
                             dy_dxi * (dz_deta * dx_dzeta - dx_deta * dz_dzeta) +
                             dz_dxi * (dx_deta * dy_dzeta - dy_deta * dx_dzeta);
          const double inv = 1. / det;
          fe.JxW[q] = det * w;
          const double dxidx = (dy_deta * dz_dzeta - dz_deta * dy_dzeta) * inv;
          const double dxidy = (dz_deta * dx_dzeta - dx_deta * dz_dzeta) * inv;
          const double dxidz = (dx_deta * dy_dzeta - dy_deta * dx_dzeta) * inv;
          const double detadx = (dz_dxi * dy_dzeta - dy_dxi * dz_dzeta) * inv;
          const double detady = (dx_dxi * dz_dzeta - dz_dxi * dx_dzeta) * inv;
          const double detadz = (dy_dxi * dx_dzeta - dx_dxi * dy_dzeta) * inv;
          const double dzetadx = (dy_dxi * dz_deta - dz_dxi * dy_deta) * inv;
          const double dzetady = (dz_dxi * dx_deta - dx_dxi * dz_deta) * inv;
          const double dzetadz = (dx_dxi * dy_deta - dy_dxi * dx_deta) * inv;
          for (int a = 0; a < ei.nn; ++a) {
            const double* d = &dref[size_t(a) * 3];
            double* g = &fe.dphi[(size_t(a) * ei.nq + q) * 3];
            g[0] = d[0] * dxidx + d[1] * detadx + d[2] * dzetadx;
            g[1] = d[0] * dxidy + d[1] * detady + d[2] * dzetady;
            g[2] = d[0] * dxidz + d[1] * detadz + d[2] * dzetadz;
          }
        } else {  // planar element in the xy-plane (GeomElem uses the element itself, geom_elem.cpp:358-374)
          const double dx_dxi = Jm[0][0], dy_dxi = Jm[0][1], dx_deta = Jm[1][0], dy_deta = Jm[1][1];
          const double det = dx_dxi * dy_deta - dx_deta * dy_dxi;
          const double inv = 1. / det;
          fe.JxW[q] = det * w;
          const double dxidx = dy_deta * inv, dxidy = -dx_deta * inv;
          const double detadx = -dy_dxi * inv, detady = dx_dxi * inv;
          for (int a = 0; a < ei.nn; ++a) {
            const double* d = &dref[size_t(a) * 3];
            double* g = &fe.dphi[(size_t(a) * ei.nq + q) * 3];
            g[0] = d[0] * dxidx + d[1] * detadx;
            g[1] = d[0] * dxidy + d[1] * detady;
            g[2] = 0.;
          }
        }
      }
}

// MAST::FEBase::init_for_side (fe_base.cpp:137-204): element shape functions evaluated at the
// side's Gauss points (libMesh FE::reinit(elem, side)), side JxW and outward normals.
static void fe_init_side(int type, const double* X, int side, FE& fe) {
  const ElemInfo ei = elem_info(type);
  const int nfn = type == ORACLE_HEX8 ? 4 : 9;
  const int forder = type == ORACLE_HEX8 ? 1 : 2;
  const int n1d = type == ORACLE_HEX8 ? 2 : 3;
  const int* smap = type == ORACLE_HEX8 ? H8_SIDE[side] : H27_SIDE[side];
  const int *FI = type == ORACLE_HEX8 ? Q4_I : Q9_I, *FJ = type == ORACLE_HEX8 ? Q4_J : Q9_J;
  const double* gp = n1d == 2 ? GP2 : GP3;
  const double* gw = n1d == 2 ? GW2 : GW3;
  fe.nn = ei.nn; fe.nq = n1d * n1d; fe.dim = 3;
  fe.phi.assign(size_t(ei.nn) * fe.nq, 0.);
  fe.dphi.clear();
  fe.JxW.assign(fe.nq, 0.);
  fe.xyz.assign(size_t(fe.nq) * 3, 0.);
  fe.normal.assign(size_t(fe.nq) * 3, 0.);
  int q = 0;
  for (int jq = 0; jq < n1d; ++jq)
    for (int iq = 0; iq < n1d; ++iq, ++q) {
      double t1[3] = {0, 0, 0}, t2[3] = {0, 0, 0}, pt[3] = {0, 0, 0};
      for (int k = 0; k < nfn; ++k) {
        const double li = lag1(forder, FI[k], gp[iq]), lj = lag1(forder, FJ[k], gp[jq]);
        const double N = li * lj;
        const double dxi = dlag1(forder, FI[k], gp[iq]) * lj, deta = li * dlag1(forder, FJ[k], gp[jq]);
        fe.phi[size_t(smap[k]) * fe.nq + q] = N;  // element shape fn of a face node == face shape fn
        for (int c = 0; c < 3; ++c) {
          const double x = X[smap[k] * 3 + c];
          pt[c] += N * x; t1[c] += dxi * x; t2[c] += deta * x;
        }
      }
      double n[3] = {t1[1] * t2[2] - t1[2] * t2[1], t1[2] * t2[0] - t1[0] * t2[2], t1[0] * t2[1] - t1[1] * t2[0]};
      const double len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      fe.JxW[q] = len * gw[iq] * gw[jq];
      for (int c = 0; c < 3; ++c) {
        fe.normal[size_t(q) * 3 + c] = n[c] / len;
        fe.xyz[size_t(q) * 3 + c] = pt[c];
      }
    }
}

// ---------------------------------------------------------------------------------------------
// property cards
// ---------------------------------------------------------------------------------------------
// IsotropicMaterialProperty::StiffnessMatrix3D::operator()  (isotropic_material_property_card.cpp:512-528)
static void stiffness_3d(double E, double nu, Mat& m) {
  m.resize(6, 6);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      if (i == j) m(i, i) = E * (1. - nu) / (1. - nu - 2. * nu * nu);
      else m(i, j) = E * nu / (1. - nu - 2. * nu * nu);
    m(i + 3, i + 3) = E / 2. / (1. + nu);
  }
}
// ...::derivative  (isotropic_material_property_card.cpp:533-567)
static void stiffness_3d_derivative(double E, double nu, double dEdf, double dnudf, Mat& m) {
  m.resize(6, 6);
  Mat dm(6, 6);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      if (i == j) dm(i, i) = (1. - nu) / (1. - nu - 2. * nu * nu);
      else dm(i, j) = nu / (1. - nu - 2. * nu * nu);
    dm(i + 3, i + 3) = 1. / 2. / (1. + nu);
  }
  for (size_t k = 0; k < m.a.size(); ++k) m.a[k] += dEdf * dm.a[k];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      if (i == j)
        dm(i, i) = -E / (1. - nu - 2. * nu * nu) + E * (1. - nu) / std::pow(1. - nu - 2. * nu * nu, 2) * (1. + 4. * nu);
      else
        dm(i, j) = E / (1. - nu - 2. * nu * nu) + E * nu / std::pow(1. - nu - 2. * nu * nu, 2) * (1. + 4. * nu);
    dm(i + 3, i + 3) = -E / 2. / std::pow(1. + nu, 2);
  }
  for (size_t k = 0; k < m.a.size(); ++k) m.a[k] += dnudf * dm.a[k];
}
// IsotropicElementProperty3D::ThermalExpansionMatrix::operator() = D * {alpha,alpha,alpha,0,0,0}
// (isotropic_element_property_card_3D.cpp:275-282; material part isotropic_material_property_card.cpp:613-637)
static Vec thermal_expansion_A_3d(double E, double nu, double alpha) {
  Mat D;
  stiffness_3d(E, nu, D);
  Vec a(6, 0.);
  for (int i = 0; i < 3; ++i) a[i] = alpha;
  return matvec(D, a);
}

// SIMP modulus and its derivative at a quadrature point
// (examples/structural/example_6/example_6.cpp:100-137; density through MeshFieldFunction =
//  FE interpolation of the nodal field at the point, src/base/mesh_field_function.cpp:72-100)
struct Modulus {
  const OracleProblem* p;
  double rho_at(const FE& fe, const int32_t* en, int q) const {
    double r = 0.;
    for (int a = 0; a < fe.nn; ++a) r += fe.N(a, q) * p->rho[en[a]];
    return r;
  }
  double E(const FE& fe, const int32_t* en, int q) const {
    if (!p->rho) return p->E0;
    return p->E0 * (p->rho_min + (1. - p->rho_min) * std::pow(rho_at(fe, en, q), p->penalty));
  }
  // dE/dp for a DV whose nodal density sensitivity on this element is drho_nodal[a]
  double dE(const FE& fe, const int32_t* en, int q, const double* drho_nodal) const {
    if (!p->rho) return 0.;
    double dr = 0.;
    for (int a = 0; a < fe.nn; ++a) dr += fe.N(a, q) * drho_nodal[a];
    return p->E0 * (1. - p->rho_min) * p->penalty * std::pow(rho_at(fe, en, q), p->penalty - 1.) * dr;
  }
};

// ---------------------------------------------------------------------------------------------
// StructuralElement3D
// ---------------------------------------------------------------------------------------------
struct StrainOps {
  OperatorMatrix lin, nl_x, nl_y, nl_z, nl_u, nl_v, nl_w;
  Mat mat_x, mat_y, mat_z;
  Vec eps;
};

// StructuralElement3D::initialize_green_lagrange_strain_operator  (solid_element_3d.cpp:1308-1463)
static void init_strain_operator(const FE& fe, int qp, bool nonlinear, const Vec& disp, StrainOps& s) {
  const int nn = fe.nn;
  s.eps.assign(6, 0.);
  s.mat_x.resize(6, 3); s.mat_y.resize(6, 3); s.mat_z.resize(6, 3);
  Vec phi(nn);
  for (int a = 0; a < nn; ++a) phi[a] = fe.dN(a, qp, 0);
  s.lin.set_shape_function(0, 0, phi);
  s.lin.set_shape_function(3, 1, phi);
  s.lin.set_shape_function(5, 2, phi);
  if (nonlinear) {
    s.nl_x.set_shape_function(0, 0, phi); s.nl_x.set_shape_function(1, 1, phi); s.nl_x.set_shape_function(2, 2, phi);
    s.nl_u.set_shape_function(0, 0, phi); s.nl_v.set_shape_function(0, 1, phi); s.nl_w.set_shape_function(0, 2, phi);
  }
  for (int a = 0; a < nn; ++a) phi[a] = fe.dN(a, qp, 1);
  s.lin.set_shape_function(1, 1, phi);
  s.lin.set_shape_function(3, 0, phi);
  s.lin.set_shape_function(4, 2, phi);
  if (nonlinear) {
    s.nl_y.set_shape_function(0, 0, phi); s.nl_y.set_shape_function(1, 1, phi); s.nl_y.set_shape_function(2, 2, phi);
    s.nl_u.set_shape_function(1, 0, phi); s.nl_v.set_shape_function(1, 1, phi); s.nl_w.set_shape_function(1, 2, phi);
  }
  for (int a = 0; a < nn; ++a) phi[a] = fe.dN(a, qp, 2);
  s.lin.set_shape_function(2, 2, phi);
  s.lin.set_shape_function(4, 1, phi);
  s.lin.set_shape_function(5, 0, phi);
  if (nonlinear) {
    s.nl_z.set_shape_function(0, 0, phi); s.nl_z.set_shape_function(1, 1, phi); s.nl_z.set_shape_function(2, 2, phi);
    s.nl_u.set_shape_function(2, 0, phi); s.nl_v.set_shape_function(2, 1, phi); s.nl_w.set_shape_function(2, 2, phi);
    Vec ddx, ddy, ddz;
    s.nl_x.vector_mult(ddx, disp);  // {du/dx, dv/dx, dw/dx}
    s.nl_y.vector_mult(ddy, disp);
    s.nl_z.vector_mult(ddz, disp);
    double F[3][3], E[3][3];
    for (int i = 0; i < 3; ++i) { F[i][0] = ddx[i]; F[i][1] = ddy[i]; F[i][2] = ddz[i]; }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double ftf = 0.;
        for (int k = 0; k < 3; ++k) ftf += F[k][i] * F[k][j];
        E[i][j] = 0.5 * (F[i][j] + F[j][i] + ftf);
      }
    s.eps[0] = E[0][0]; s.eps[1] = E[1][1]; s.eps[2] = E[2][2];
    s.eps[3] = E[0][1] + E[1][0]; s.eps[4] = E[1][2] + E[2][1]; s.eps[5] = E[0][2] + E[2][0];
    for (int c = 0; c < 3; ++c) {
      s.mat_x(0, c) = ddx[c]; s.mat_x(3, c) = ddy[c]; s.mat_x(5, c) = ddz[c];
      s.mat_y(1, c) = ddy[c]; s.mat_y(3, c) = ddx[c]; s.mat_y(4, c) = ddz[c];
      s.mat_z(2, c) = ddz[c]; s.mat_z(4, c) = ddy[c]; s.mat_z(5, c) = ddx[c];
    }
  } else {
    s.lin.vector_mult(s.eps, disp);
  }
}

static void strain_ops_reinit(StrainOps& s, int nn) {
  s.lin.reinit(6, 3, nn);
  s.nl_x.reinit(3, 3, nn); s.nl_y.reinit(3, 3, nn); s.nl_z.reinit(3, 3, nn);
  s.nl_u.reinit(3, 3, nn); s.nl_v.reinit(3, 3, nn); s.nl_w.reinit(3, 3, nn);
}

static void add_scaled_top_left(Mat& jac, const Mat& m, double s) {  // jac.topLeftCorner(n2,n2) += s*m
  for (int i = 0; i < m.r; ++i)
    for (int j = 0; j < m.c; ++j) jac(i, j) += s * m(i, j);
}

// StructuralElement3D::internal_residual (solid_element_3d.cpp:135-423) and, with sens != nullptr,
// ::internal_residual_sensitivity (:566-788: identical except D <- dD/dp at :640).
// f has nd entries, jac nd x nd (nd = nv*nn); only the leading 3nn block is touched, plus the
// rotational diagonal when nv == 6 (:413-417).
static void solid3d_internal_residual(const OracleProblem* p, const FE& fe, const int32_t* en, const Vec& sol,
                                      bool request_jacobian, Vec& f, Mat& jac, const double* drho_nodal /*sens*/,
                                      int param_kind = 0 /* 1: d/dE0, 2: d/dnu of the property card (internal_residual_sensitivity) */) {
  const int nn = fe.nn, n2 = 3 * nn;
  const bool nl = p->nonlinear_strain != 0;
  Vec disp(sol.begin(), sol.begin() + n2), stress, vec2, vec3;
  Mat D, mat1, mat2, mat3, mat4(3, 3);
  StrainOps s;
  strain_ops_reinit(s, nn);
  Modulus mod{p};
  for (int qp = 0; qp < fe.nq; ++qp) {
    if (param_kind == 1) stiffness_3d_derivative(mod.E(fe, en, qp), p->nu, mod.E(fe, en, qp) / p->E0, 0., D);  // :640, dE(x)/dE0
    else if (param_kind == 2) stiffness_3d_derivative(mod.E(fe, en, qp), p->nu, 0., 1., D);              // :640, dnu/dnu = 1
    else if (!drho_nodal) stiffness_3d(mod.E(fe, en, qp), p->nu, D);                                  // :272
    else stiffness_3d_derivative(mod.E(fe, en, qp), p->nu, mod.dE(fe, en, qp, drho_nodal), 0., D);  // :640
    init_strain_operator(fe, qp, nl, disp, s);
    stress = matvec(D, s.eps);  // :289 (Gmat*alpha == 0: incompatible modes are inert)
    const double JxW = fe.JxW[qp];
    s.lin.vector_mult_transpose(vec2, stress);
    for (int i = 0; i < n2; ++i) f[i] += JxW * vec2[i];  // :296-297
    if (nl) {                                             // :299-316
      vec3 = matvec_t(s.mat_x, stress); s.nl_x.vector_mult_transpose(vec2, vec3);
      for (int i = 0; i < n2; ++i) f[i] += JxW * vec2[i];
      vec3 = matvec_t(s.mat_y, stress); s.nl_y.vector_mult_transpose(vec2, vec3);
      for (int i = 0; i < n2; ++i) f[i] += JxW * vec2[i];
      vec3 = matvec_t(s.mat_z, stress); s.nl_z.vector_mult_transpose(vec2, vec3);
      for (int i = 0; i < n2; ++i) f[i] += JxW * vec2[i];
    }
    if (request_jacobian) {
      s.lin.left_multiply(mat1, D);                    // :327  D*B_lin
      s.lin.right_multiply_transpose(mat2, mat1);      // :328  B_lin^T D B_lin
      add_scaled_top_left(jac, mat2, JxW);             // :329
      if (nl) {
        mat3 = matmul_tn(s.mat_x, mat1); s.nl_x.right_multiply_transpose(mat2, mat3); add_scaled_top_left(jac, mat2, JxW);  // :334-336
        mat3 = matmul_tn(s.mat_y, mat1); s.nl_y.right_multiply_transpose(mat2, mat3); add_scaled_top_left(jac, mat2, JxW);
        mat3 = matmul_tn(s.mat_z, mat1); s.nl_z.right_multiply_transpose(mat2, mat3); add_scaled_top_left(jac, mat2, JxW);
        for (int idim = 0; idim < 3; ++idim) {           // :349-384
          if (idim == 0) s.nl_x.left_multiply(mat1, s.mat_x);
          else if (idim == 1) s.nl_y.left_multiply(mat1, s.mat_y);
          else s.nl_z.left_multiply(mat1, s.mat_z);
          mat1 = matmul(D, mat1);
          s.lin.right_multiply_transpose(mat2, mat1); add_scaled_top_left(jac, mat2, JxW);
          mat3 = matmul_tn(s.mat_x, mat1); s.nl_x.right_multiply_transpose(mat2, mat3); add_scaled_top_left(jac, mat2, JxW);
          mat3 = matmul_tn(s.mat_y, mat1); s.nl_y.right_multiply_transpose(mat2, mat3); add_scaled_top_left(jac, mat2, JxW);
          mat3 = matmul_tn(s.mat_z, mat1); s.nl_z.right_multiply_transpose(mat2, mat3); add_scaled_top_left(jac, mat2, JxW);
        }
        mat4(0, 0) = stress[0]; mat4(1, 1) = stress[1]; mat4(2, 2) = stress[2];  // :388-393
        mat4(0, 1) = mat4(1, 0) = stress[3];
        mat4(1, 2) = mat4(2, 1) = stress[4];
        mat4(0, 2) = mat4(2, 0) = stress[5];
        s.nl_u.left_multiply(mat3, mat4); s.nl_u.right_multiply_transpose(mat2, mat3); add_scaled_top_left(jac, mat2, JxW);
        s.nl_v.left_multiply(mat3, mat4); s.nl_v.right_multiply_transpose(mat2, mat3); add_scaled_top_left(jac, mat2, JxW);
        s.nl_w.left_multiply(mat3, mat4); s.nl_w.right_multiply_transpose(mat2, mat3); add_scaled_top_left(jac, mat2, JxW);
      }
    }
  }
  if (request_jacobian && p->nv == 6) {  // :413-417  rotational dofs get 1e-20*max(diag)
    double mx = jac(0, 0);
    for (int i = 1; i < jac.r; ++i) mx = std::max(mx, jac(i, i));
    for (int i = 0; i < n2; ++i) jac(n2 + i, n2 + i) += 1.0e-20 * mx;
  }
}

// StructuralElement3D::surface_pressure_residual (side version, solid_element_3d.cpp:815-876)
static void solid3d_surface_pressure_residual(const OracleProblem* p, const double* X, int side, double press, Vec& f) {
  FE fe;
  fe_init_side(p->elem_type, X, side, fe);
  const int nn = fe.nn, n2 = 3 * nn;
  OperatorMatrix B;
  Vec phi(nn), force(3), vec, local_f(n2, 0.);
  for (int qp = 0; qp < fe.nq; ++qp) {
    for (int a = 0; a < nn; ++a) phi[a] = fe.N(a, qp);
    B.reinit_diag(3, phi);
    for (int d = 0; d < 3; ++d) force[d] = press * fe.normal[size_t(qp) * 3 + d];
    B.vector_mult_transpose(vec, force);
    for (int i = 0; i < n2; ++i) local_f[i] += fe.JxW[qp] * vec[i];
  }
  for (int i = 0; i < n2; ++i) f[i] -= local_f[i];
}

// StructuralElement3D::thermal_residual (solid_element_3d.cpp:950-1089)
static void solid3d_thermal_residual(const OracleProblem* p, const FE& fe, const int32_t* en, const Vec& sol,
                                     bool request_jacobian, Vec& f, Mat& jac) {
  const int nn = fe.nn, n2 = 3 * nn;
  const bool nl = p->nonlinear_strain != 0;
  Vec disp(sol.begin(), sol.begin() + n2), vec1, vec2, vec3;
  Mat mat2, mat3, mat4(3, 3);
  StrainOps s;
  strain_ops_reinit(s, nn);
  Modulus mod{p};
  for (int qp = 0; qp < fe.nq; ++qp) {
    Vec A = thermal_expansion_A_3d(mod.E(fe, en, qp), p->nu, p->alpha);
    double t = p->temperature_const;
    if (p->temperature) {  // nodal temperature field interpolated at the point (MeshFieldFunction)
      t = 0.;
      for (int a = 0; a < nn; ++a) t += fe.N(a, qp) * p->temperature[en[a]];
    }
    const double dT = t - p->ref_temperature;
    vec1.assign(6, 0.);
    for (int i = 0; i < 6; ++i) vec1[i] = A[i] * dT;  // :1016
    init_strain_operator(fe, qp, nl, disp, s);
    const double JxW = fe.JxW[qp];
    s.lin.vector_mult_transpose(vec3, vec1);
    for (int i = 0; i < n2; ++i) f[i] -= JxW * vec3[i];  // :1032-1033
    if (nl) {
      vec2 = matvec_t(s.mat_x, vec1); s.nl_x.vector_mult_transpose(vec3, vec2);
      for (int i = 0; i < n2; ++i) f[i] -= JxW * vec3[i];
      vec2 = matvec_t(s.mat_y, vec1); s.nl_y.vector_mult_transpose(vec3, vec2);
      for (int i = 0; i < n2; ++i) f[i] -= JxW * vec3[i];
      vec2 = matvec_t(s.mat_z, vec1); s.nl_z.vector_mult_transpose(vec3, vec2);
      for (int i = 0; i < n2; ++i) f[i] -= JxW * vec3[i];
      if (request_jacobian) {  // :1054-1082
        mat4(0, 0) = vec1[0]; mat4(1, 1) = vec1[1]; mat4(2, 2) = vec1[2];
        mat4(0, 1) = mat4(1, 0) = vec1[3];
        mat4(1, 2) = mat4(2, 1) = vec1[4];
        mat4(0, 2) = mat4(2, 0) = vec1[5];
        s.nl_u.left_multiply(mat3, mat4); s.nl_u.right_multiply_transpose(mat2, mat3); add_scaled_top_left(jac, mat2, -JxW);
        s.nl_v.left_multiply(mat3, mat4); s.nl_v.right_multiply_transpose(mat2, mat3); add_scaled_top_left(jac, mat2, -JxW);
        s.nl_w.left_multiply(mat3, mat4); s.nl_w.right_multiply_transpose(mat2, mat3); add_scaled_top_left(jac, mat2, -JxW);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// StructuralElement2D::internal_residual for a QUAD4 Mindlin plate in the xy-plane; linear strain or, with
// nonlinear_strain = 1, the Green-Lagrange membrane strain of ::initialize_green_lagrange_strain_operator (:141-250) plus the
// von Karman strain of ::initialize_von_karman_strain_operator (:57-100) with every tangent term of :1886-1998
// (structural_element_2d.cpp:1075-1227, ::_internal_residual_operation :1765-2019,
//  ::initialize_green_lagrange_strain_operator :141-250, MindlinBendingOperator
//  mindlin_bending_operator.cpp:47-80 (bending strain), :94-160 + :314-361 (transverse shear with the
//  quadrature order reduced by two: one Gauss point for QUAD4), section matrices A = C h,
//  B = C h off, D = C (h^3/12 + h off^2) of solid_2d_section_element_property_card.cpp:395-510,
//  transverse shear G kappa h (:122-134; isotropic_material_property_card.cpp:336-345), plane-stress /
//  plane-strain C of isotropic_material_property_card.cpp:395-426, drilling diagonal 1e-8 (:1217-1221)).
// Elements out of the xy-plane would need the T K T^T transform of structural_element_base.cpp:1565-1601;
// the element is in the plane here (geom_elem.cpp:358-374), so local == global.
// ---------------------------------------------------------------------------------------------
static void stiffness_2d(double E, double nu, bool plane_stress, Mat& m) {
  m.resize(3, 3);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) {
      if (plane_stress) m(i, j) = (i == j) ? E / (1. - nu * nu) : E * nu / (1. - nu * nu);
      else m(i, j) = (i == j) ? E * (1. - nu) / ((1. + nu) * (1. - 2. * nu)) : E * nu / ((1. + nu) * (1. - 2. * nu));
    }
  m(2, 2) = E / 2. / (1. + nu);
}

static Mat scaled(const Mat& m, double s) {
  Mat r = m;
  for (double& v : r.a) v *= s;
  return r;
}
static Mat transposed(const Mat& m) {
  Mat r(m.c, m.r);
  for (int i = 0; i < m.r; ++i)
    for (int j = 0; j < m.c; ++j) r(j, i) = m(i, j);
  return r;
}

// FE of the QUAD4 at the single point of the order-1 Gauss rule (xi = eta = 0, weight 4)
static void fe_init_quad4_reduced(const double* X, FE& fe) {
  fe.nn = 4; fe.nq = 1; fe.dim = 2;
  fe.phi.assign(4, 0.); fe.dphi.assign(12, 0.); fe.JxW.assign(1, 0.); fe.xyz.assign(3, 0.);
  double Jm[2][2] = {{0, 0}, {0, 0}}, dref[4][2];
  for (int a = 0; a < 4; ++a) {
    const double li = lag1(1, Q4_I[a], 0.), lj = lag1(1, Q4_J[a], 0.);
    fe.phi[a] = li * lj;
    dref[a][0] = dlag1(1, Q4_I[a], 0.) * lj;
    dref[a][1] = li * dlag1(1, Q4_J[a], 0.);
    for (int c = 0; c < 2; ++c) {
      Jm[0][c] += dref[a][0] * X[a * 3 + c];
      Jm[1][c] += dref[a][1] * X[a * 3 + c];
    }
    for (int c = 0; c < 3; ++c) fe.xyz[c] += fe.phi[a] * X[a * 3 + c];
  }
  const double det = Jm[0][0] * Jm[1][1] - Jm[1][0] * Jm[0][1], inv = 1. / det;
  fe.JxW[0] = det * 4.;
  const double dxidx = Jm[1][1] * inv, dxidy = -Jm[1][0] * inv, detadx = -Jm[0][1] * inv, detady = Jm[0][0] * inv;
  for (int a = 0; a < 4; ++a) {
    fe.dphi[a * 3 + 0] = dref[a][0] * dxidx + dref[a][1] * detadx;
    fe.dphi[a * 3 + 1] = dref[a][0] * dxidy + dref[a][1] * detady;
  }
}

static void plate_internal_residual(const OracleProblem* p, const FE& fe, const double* X, const Vec& sol,
                                    bool request_jacobian, Vec& f, Mat& jac) {
  const int nn = fe.nn, n2 = 6 * nn;
  Mat C, A, B, D, mat1, mat2, mat3;
  stiffness_2d(p->E0, p->nu, p->plane_strain == 0, C);
  const double h = p->thickness, off = p->offset;
  A = scaled(C, h);                                    // ExtensionStiffnessMatrix
  B = scaled(C, h * off);                              // ExtensionBendingStiffnessMatrix
  D = scaled(C, std::pow(h, 3) / 12. + h * std::pow(off, 2));  // BendingStiffnessMatrix
  OperatorMatrix Bl, Bb, Bnx, Bny, Bnu, Bnv, Bvk;
  Bl.reinit(3, 6, nn); Bb.reinit(3, 6, nn);
  const bool nl = p->nonlinear_strain != 0;  // strain_type() == NONLINEAR_STRAIN: if_vk (:1133)
  if (nl) { Bnx.reinit(2, 6, nn); Bny.reinit(2, 6, nn); Bnu.reinit(2, 6, nn); Bnv.reinit(2, 6, nn); Bvk.reinit(2, 6, nn); }
  Vec phi(nn), strain, vec1, vec2, vec3, vec4, vec6, local_f(n2, 0.);
  Mat local_jac(n2, n2), mat_x(3, 2), mat_y(3, 2), vk_dwdxi(3, 2), stress(2, 2), mat5;
  for (int qp = 0; qp < fe.nq; ++qp) {
    const double JxW = fe.JxW[qp];
    for (int a = 0; a < nn; ++a) phi[a] = fe.dN(a, qp, 0);
    Bl.set_shape_function(0, 0, phi); Bl.set_shape_function(2, 1, phi);     // :184-186
    if (nl) {                                                               // :188-197
      Bnx.set_shape_function(0, 0, phi); Bnx.set_shape_function(1, 1, phi);
      Bnu.set_shape_function(0, 0, phi); Bnv.set_shape_function(0, 1, phi);
      Bvk.set_shape_function(0, 2, phi);                                    // :84
    }
    Bb.set_shape_function(0, 4, phi);                                       // mindlin :69
    for (int a = 0; a < nn; ++a) phi[a] = -fe.dN(a, qp, 0);
    Bb.set_shape_function(2, 3, phi);                                       // :71
    for (int a = 0; a < nn; ++a) phi[a] = fe.dN(a, qp, 1);
    Bl.set_shape_function(1, 1, phi); Bl.set_shape_function(2, 0, phi);     // :203-204
    if (nl) {                                                               // :206-215
      Bny.set_shape_function(0, 0, phi); Bny.set_shape_function(1, 1, phi);
      Bnu.set_shape_function(1, 0, phi); Bnv.set_shape_function(1, 1, phi);
      Bvk.set_shape_function(1, 2, phi);                                    // :95
    }
    Bb.set_shape_function(2, 4, phi);                                       // :78
    for (int a = 0; a < nn; ++a) phi[a] = -fe.dN(a, qp, 1);
    Bb.set_shape_function(1, 3, phi);                                       // :81
    if (nl) {  // :217-246: E = 1/2 (F + F^T + F^T F) with F = [d disp/dx, d disp/dy]
      Vec ddx, ddy;
      Bnx.vector_mult(ddx, sol); Bny.vector_mult(ddy, sol);
      Mat F(2, 2), E(2, 2);
      F(0, 0) = ddx[0]; F(1, 0) = ddx[1]; F(0, 1) = ddy[0]; F(1, 1) = ddy[1];
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) E(i, j) = 0.5 * (F(i, j) + F(j, i) + (F(0, i) * F(0, j) + F(1, i) * F(1, j)));
      strain.assign(3, 0.);
      strain[0] = E(0, 0); strain[1] = E(1, 1); strain[2] = E(0, 1) + E(1, 0);
      mat_x.zero(); mat_y.zero();
      mat_x(0, 0) = ddx[0]; mat_x(0, 1) = ddx[1]; mat_x(2, 0) = ddy[0]; mat_x(2, 1) = ddy[1];
      mat_y(1, 0) = ddy[0]; mat_y(1, 1) = ddy[1]; mat_y(2, 0) = ddx[0]; mat_y(2, 1) = ddx[1];
    } else {
      Bl.vector_mult(strain, sol);
    }
    vec2 = matvec(A, strain);                       // membrane stress  :1815
    Bb.vector_mult(vec1, sol);
    { const Vec t = matvec(B, vec1); for (int i = 0; i < 3; ++i) vec2[i] += t[i]; }  // :1822
    if (nl) {  // von Karman strain (:57-100) joins the membrane strain and stress (:1824-1833)
      Vec dwv;
      Bvk.vector_mult(dwv, sol);
      const double dwx = dwv[0], dwy = dwv[1];
      vk_dwdxi.zero();
      vk_dwdxi(0, 0) = dwx; vk_dwdxi(2, 1) = dwx; vk_dwdxi(1, 1) = dwy; vk_dwdxi(2, 0) = dwy;
      const Vec evk = {0.5 * dwx * dwx, 0.5 * dwy * dwy, dwx * dwy};
      for (int i = 0; i < 3; ++i) strain[i] += evk[i];
      const Vec t = matvec(A, evk);
      for (int i = 0; i < 3; ++i) vec2[i] += t[i];
    }
    stress(0, 0) = vec2[0]; stress(0, 1) = vec2[2]; stress(1, 0) = vec2[2]; stress(1, 1) = vec2[1];  // :1838-1841
    Bl.vector_mult_transpose(vec3, vec2);
    for (int i = 0; i < n2; ++i) local_f[i] += JxW * vec3[i];               // :1846-1847
    if (nl) {
      vec4 = matvec_t(mat_x, vec2); Bnx.vector_mult_transpose(vec6, vec4);  // :1853-1855
      for (int i = 0; i < n2; ++i) local_f[i] += JxW * vec6[i];
      vec4 = matvec_t(mat_y, vec2); Bny.vector_mult_transpose(vec6, vec4);  // :1858-1860
      for (int i = 0; i < n2; ++i) local_f[i] += JxW * vec6[i];
      vec4 = matvec_t(vk_dwdxi, vec2); Bvk.vector_mult_transpose(vec3, vec4);  // :1867-1869
      for (int i = 0; i < n2; ++i) local_f[i] += JxW * vec3[i];
    }
    vec1 = matvec_t(B, strain);                                             // :1875
    Bb.vector_mult_transpose(vec3, vec1);
    for (int i = 0; i < n2; ++i) local_f[i] += JxW * vec3[i];
    Bb.vector_mult(vec2, sol);                                              // :1880-1883
    vec1 = matvec(D, vec2);
    Bb.vector_mult_transpose(vec3, vec1);
    for (int i = 0; i < n2; ++i) local_f[i] += JxW * vec3[i];
    if (request_jacobian) {
      Bl.left_multiply(mat1, A); Bl.right_multiply_transpose(mat2, mat1); add_scaled_top_left(local_jac, mat2, JxW);  // :1888-1890
      if (nl) {
        // the eight products of the Green-Lagrange membrane operator (:1895-1949) ...
        Bnx.left_multiply(mat1, mat_x); mat1 = matmul(A, mat1); Bl.right_multiply_transpose(mat2, mat1); add_scaled_top_left(local_jac, mat2, JxW);
        Bny.left_multiply(mat1, mat_y); mat1 = matmul(A, mat1); Bl.right_multiply_transpose(mat2, mat1); add_scaled_top_left(local_jac, mat2, JxW);
        Bl.left_multiply(mat1, A); mat5 = matmul_tn(mat_x, mat1); Bnx.right_multiply_transpose(mat2, mat5); add_scaled_top_left(local_jac, mat2, JxW);
        Bnx.left_multiply(mat1, mat_x); mat1 = matmul(A, mat1); mat5 = matmul_tn(mat_x, mat1); Bnx.right_multiply_transpose(mat2, mat5); add_scaled_top_left(local_jac, mat2, JxW);
        Bny.left_multiply(mat1, mat_y); mat1 = matmul(A, mat1); mat5 = matmul_tn(mat_x, mat1); Bnx.right_multiply_transpose(mat2, mat5); add_scaled_top_left(local_jac, mat2, JxW);
        Bl.left_multiply(mat1, A); mat5 = matmul_tn(mat_y, mat1); Bny.right_multiply_transpose(mat2, mat5); add_scaled_top_left(local_jac, mat2, JxW);
        Bnx.left_multiply(mat1, mat_x); mat1 = matmul(A, mat1); mat5 = matmul_tn(mat_y, mat1); Bny.right_multiply_transpose(mat2, mat5); add_scaled_top_left(local_jac, mat2, JxW);
        Bny.left_multiply(mat1, mat_y); mat1 = matmul(A, mat1); mat5 = matmul_tn(mat_y, mat1); Bny.right_multiply_transpose(mat2, mat5); add_scaled_top_left(local_jac, mat2, JxW);
        // ... and the initial-stress stiffness of u and v (:1954-1963)
        Bnu.left_multiply(mat5, stress); Bnu.right_multiply_transpose(mat2, mat5); add_scaled_top_left(local_jac, mat2, JxW);
        Bnv.left_multiply(mat5, stress); Bnv.right_multiply_transpose(mat2, mat5); add_scaled_top_left(local_jac, mat2, JxW);
        // von Karman terms (:1967-1998): membrane - vk, vk - membrane (with B_lin only, as the reference has them), vk - vk, bending - vk
        Bvk.left_multiply(mat3, vk_dwdxi); mat3 = matmul(A, mat3); Bl.right_multiply_transpose(mat2, mat3); add_scaled_top_left(local_jac, mat2, JxW);
        Bl.left_multiply(mat1, A); mat3 = matmul_tn(vk_dwdxi, mat1); Bvk.right_multiply_transpose(mat2, mat3); add_scaled_top_left(local_jac, mat2, JxW);
        Bvk.left_multiply(mat3, stress); Bvk.right_multiply_transpose(mat2, mat3); add_scaled_top_left(local_jac, mat2, JxW);
        Bvk.left_multiply(mat3, vk_dwdxi); mat3 = matmul_tn(vk_dwdxi, matmul(A, mat3)); Bvk.right_multiply_transpose(mat2, mat3); add_scaled_top_left(local_jac, mat2, JxW);
        Bvk.left_multiply(mat3, vk_dwdxi); mat3 = matmul_tn(B, mat3); Bb.right_multiply_transpose(mat2, mat3); add_scaled_top_left(local_jac, mat2, JxW);
        Bb.left_multiply(mat1, B); mat3 = matmul_tn(vk_dwdxi, mat1); Bvk.right_multiply_transpose(mat2, mat3); add_scaled_top_left(local_jac, mat2, JxW);
      }
      mat3 = transposed(B);
      Bl.left_multiply(mat1, mat3); Bb.right_multiply_transpose(mat2, mat1); add_scaled_top_left(local_jac, mat2, JxW);  // :2001-2004
      Bb.left_multiply(mat1, B); Bl.right_multiply_transpose(mat2, mat1); add_scaled_top_left(local_jac, mat2, JxW);     // :2007-2009
      Bb.left_multiply(mat1, D); Bb.right_multiply_transpose(mat2, mat1); add_scaled_top_left(local_jac, mat2, JxW);     // :2012-2014
    }
  }
  {  // MindlinBendingOperator::calculate_transverse_shear_residual  (mindlin_bending_operator.cpp:94-160, 314-361)
    FE fs;
    fe_init_quad4_reduced(X, fs);
    Mat G(2, 2);
    G(0, 0) = G(1, 1) = p->E0 / 2. / (1. + p->nu);
    G = scaled(scaled(G, h), p->kappa);
    OperatorMatrix Bt;
    Bt.reinit(2, 6, nn);
    for (int a = 0; a < nn; ++a) phi[a] = fs.dN(a, 0, 0);
    Bt.set_shape_function(0, 2, phi);
    for (int a = 0; a < nn; ++a) phi[a] = fs.dN(a, 0, 1);
    Bt.set_shape_function(1, 2, phi);
    for (int a = 0; a < nn; ++a) phi[a] = fs.N(a, 0);
    Bt.set_shape_function(0, 4, phi);
    for (int a = 0; a < nn; ++a) phi[a] = -fs.N(a, 0);
    Bt.set_shape_function(1, 3, phi);
    Bt.vector_mult(vec1, sol);
    vec1 = matvec(G, vec1);
    Bt.vector_mult_transpose(vec3, vec1);
    for (int i = 0; i < n2; ++i) local_f[i] += fs.JxW[0] * vec3[i];
    if (request_jacobian) {
      Bt.left_multiply(mat1, G); Bt.right_multiply_transpose(mat2, mat1); add_scaled_top_left(local_jac, mat2, fs.JxW[0]);
    }
  }
  for (int i = 0; i < n2; ++i) f[i] += local_f[i];  // transform_vector_to_global_system: identity in the xy-plane
  if (request_jacobian) {
    for (int i = 0; i < nn; ++i) local_jac(5 * nn + i, 5 * nn + i) = 1.0e-8;  // :1217-1221
    add_scaled_top_left(jac, local_jac, 1.);
  }
}

// A QUAD4 plate anywhere in space: GeomElem builds a local element in the plane of the first three nodes when the
// nodes do not share one z (geom_elem.cpp:358-374, ::_init_local_elem_2d :461-511), the element is evaluated there with
// the solution rotated into the local frame (structural_element_base.cpp:1606-1640) and the residual / Jacobian are
// rotated back: f = T_g f_l, K = T_g K_l T_g^T (:1565-1601, 1642-1674), T_g repeating the 3x3 T for the translations
// and the rotations of every node.
static void plate_elem_calculations(const OracleProblem* p, const double* X, const Vec& sol, bool request_jacobian,
                                    Vec& f, Mat& jac) {
  const int nn = 4, n2 = 24;
  bool use_local = false;
  for (int i = 1; i < nn; ++i)
    if (std::fabs(X[2] - X[i * 3 + 2]) != 0.) use_local = true;
  if (!use_local) {
    FE fe;
    fe_init(ORACLE_QUAD4, X, fe);
    plate_internal_residual(p, fe, X, sol, request_jacobian, f, jac);
    return;
  }
  auto sub = [&](int a, int b, double* o) { for (int c = 0; c < 3; ++c) o[c] = X[a * 3 + c] - X[b * 3 + c]; };
  auto normalise = [](double* v) { const double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); for (int c = 0; c < 3; ++c) v[c] /= n; };
  auto cross = [](const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
  };
  double v1[3], v2[3], v3[3];
  sub(1, 0, v1); normalise(v1);
  sub(2, 0, v2); normalise(v2);
  cross(v1, v2, v3); normalise(v3);
  cross(v3, v1, v2); normalise(v2);
  double T[3][3];
  for (int i = 0; i < 3; ++i) { T[i][0] = v1[i]; T[i][1] = v2[i]; T[i][2] = v3[i]; }
  std::vector<double> Xl(nn * 3, 0.);
  for (int i = 0; i < nn; ++i)
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) Xl[i * 3 + j] += T[k][j] * (X[i * 3 + k] - X[k]);
  Mat Tg(n2, n2);
  for (int i = 0; i < nn; ++i)
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) { Tg(j * nn + i, k * nn + i) = T[j][k]; Tg((j + 3) * nn + i, (k + 3) * nn + i) = T[j][k]; }
  const Vec sol_l = matvec_t(Tg, sol);
  FE fe;
  fe_init(ORACLE_QUAD4, Xl.data(), fe);  // planar FE in the local x-y coordinates
  Vec fl(n2, 0.);
  Mat Kl(n2, n2);
  plate_internal_residual(p, fe, Xl.data(), sol_l, request_jacobian, fl, Kl);
  const Vec fg = matvec(Tg, fl);
  for (int i = 0; i < n2; ++i) f[i] += fg[i];
  if (request_jacobian) {
    Mat TK = matmul(Tg, Kl), Tt(n2, n2);
    for (int i = 0; i < n2; ++i)
      for (int j = 0; j < n2; ++j) Tt(i, j) = Tg(j, i);
    const Mat Kg = matmul(TK, Tt);
    add_scaled_top_left(jac, Kg, 1.);
  }
}

// HeatConductionElementBase::surface_flux_residual (heat_conduction_elem_base.cpp:876-914) and
// ::surface_convection_residual (:1100-1155) on side `side` of a QUAD4 (an EDGE2 with the 2-point Gauss rule of
// MAST::FEBase::init_for_side, fe_base.cpp:137-204): f += JxW N th flux ; f += JxW N th h_c (T - T_amb),
// jac += JxW N N^T th h_c.  th = section thickness (solid_2d_section_element_property_card: "h").
static void heat_side_residual(const OracleProblem* p, const double* X, const Vec& sol, int side, double flux, double hc,
                               double Tamb, double emiss, bool request_jacobian, Vec& f, Mat& jac) {
  const int n1 = side, n2 = (side + 1) % 4;
  const double dx = X[n2 * 3] - X[n1 * 3], dy = X[n2 * 3 + 1] - X[n1 * 3 + 1], dz = X[n2 * 3 + 2] - X[n1 * 3 + 2];
  const double L = std::sqrt(dx * dx + dy * dy + dz * dz);
  for (int q = 0; q < 2; ++q) {
    const double xi = GP2[q], JxW = 0.5 * L * GW2[q];
    const double N[2] = {0.5 * (1. - xi), 0.5 * (1. + xi)};
    const int nd[2] = {n1, n2};
    const double temp = N[0] * sol[n1] + N[1] * sol[n2];
    for (int a = 0; a < 2; ++a) {
      f[nd[a]] += JxW * N[a] * flux * p->thickness;
      if (hc != 0.) f[nd[a]] += JxW * N[a] * p->thickness * hc * (temp - Tamb);
      if (request_jacobian && hc != 0.)
        for (int b = 0; b < 2; ++b) jac(nd[a], nd[b]) += JxW * (N[a] * N[b]) * p->thickness * hc;
      if (emiss != 0.) {
        // ::surface_radiation_residual (heat_conduction_elem_base.cpp:1404-1458):
        // f += JxW phi sbc emiss th ((T + zero_ref)^4 - (T_amb + zero_ref)^4), jac += JxW phi phi^T sbc emiss th 4 (T + zero_ref)^3
        f[nd[a]] += JxW * N[a] * p->rad_sb_const * emiss * p->thickness *
                    (std::pow(temp + p->rad_T_ref_zero, 4.) - std::pow(Tamb + p->rad_T_ref_zero, 4.));
        if (request_jacobian)
          for (int b = 0; b < 2; ++b)
            jac(nd[a], nd[b]) += JxW * (N[a] * N[b]) * p->rad_sb_const * emiss * p->thickness * 4. * std::pow(temp + p->rad_T_ref_zero, 3.);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// HeatConductionElementBase::internal_residual (heat_conduction_elem_base.cpp:53-149), QUAD4 with
// Solid2DSectionProperty::ThermalConductanceMatrix = k*h (solid_2d_section_element_property_card.cpp:940-953)
// and volume_heat_source_residual (:1709-1745)
// ---------------------------------------------------------------------------------------------
static void heat_elem_calculations(const OracleProblem* p, const FE& fe, const Vec& sol, bool request_jacobian,
                                   Vec& f, Mat& jac) {
  const int nn = fe.nn, dim = fe.dim;
  OperatorMatrix Bmat;
  std::vector<OperatorMatrix> dB(dim);
  Vec phi(nn), vec2;
  Mat vec1, mat_n2n2, solm(nn, 1);
  for (int a = 0; a < nn; ++a) solm(a, 0) = sol[a];
  for (int qp = 0; qp < fe.nq; ++qp) {
    for (int a = 0; a < nn; ++a) phi[a] = fe.N(a, qp);
    Bmat.reinit_diag(1, phi);
    Bmat.right_multiply(vec1, solm);  // T at the quadrature point  (:84)
    const double Tqp = vec1(0, 0);
    const double k = (p->k_of_T ? p->k0 * (1. + p->k_beta * Tqp) : p->k0) * p->thickness;
    const double dk = p->k_of_T ? p->k0 * p->k_beta * p->thickness : 0.;
    for (int d = 0; d < dim; ++d) {
      for (int a = 0; a < nn; ++a) phi[a] = fe.dN(a, qp, d);
      dB[d].reinit_diag(1, phi);
    }
    Vec flux(dim, 0.);
    for (int j = 0; j < dim; ++j) {
      dB[j].right_multiply(vec1, solm);  // dT/dx_j
      for (int i = 0; i < dim; ++i) flux[i] += vec1(0, 0) * (i == j ? k : 0.);
    }
    for (int i = 0; i < dim; ++i) {
      Vec v1(1, flux[i]);
      dB[i].vector_mult_transpose(vec2, v1);
      for (int a = 0; a < nn; ++a) f[a] += fe.JxW[qp] * vec2[a];
    }
    if (request_jacobian) {
      for (int i = 0; i < dim; ++i)
        for (int j = 0; j < dim; ++j) {
          dB[i].right_multiply_transpose(mat_n2n2, dB[j]);
          const double kij = (i == j ? k : 0.);
          for (int a = 0; a < nn; ++a)
            for (int b = 0; b < nn; ++b) jac(a, b) += fe.JxW[qp] * kij * mat_n2n2(a, b);
        }
      if (p->k_of_T) {  // :125-141
        for (int j = 0; j < dim; ++j) {
          dB[j].right_multiply(vec1, solm);
          for (int i = 0; i < dim; ++i) {
            const double dkij = (i == j ? dk : 0.);
            if (dkij != 0.) {
              dB[i].right_multiply_transpose(mat_n2n2, Bmat);
              for (int a = 0; a < nn; ++a)
                for (int b = 0; b < nn; ++b) jac(a, b) += fe.JxW[qp] * vec1(0, 0) * dkij * mat_n2n2(a, b);
            }
          }
        }
      }
    }
  }
  if (p->heat_source != 0.) {  // volume_heat_source_residual
    for (int qp = 0; qp < fe.nq; ++qp)
      for (int a = 0; a < nn; ++a) f[a] -= fe.JxW[qp] * fe.N(a, qp) * p->heat_source * p->thickness;
  }
}

// HeatConductionNonlinearAssemblyElemOperations::elem_sensitivity_calculations (heat_conduction_nonlinear_assembly.cpp:
// 100-116) for a parameter of the property cards: kind 1 = the conductivity k0 of the material card, kind 2 = the section
// thickness h.  ::internal_residual_sensitivity (heat_conduction_elem_base.cpp:552-617) with the derivative of the
// conductance matrix k(T) h (solid_2d_section_element_property_card.cpp:955-975: dk h + k dh),
// ::volume_heat_source_residual_sensitivity (:1750-1794: f -= JxW N (dq th + q dth)) and the side-load sensitivities
// (:957-1002 flux, :1217-1275 convection, :1524-1583 radiation: every term is (value dth + dvalue th) with dvalue = 0 here).
static void heat_elem_sensitivity(const OracleProblem* p, const FE& fe, const Vec& sol, int kind, Vec& f) {
  const int nn = fe.nn, dim = fe.dim;
  const double dth = kind == 2 ? 1. : 0.;
  for (int qp = 0; qp < fe.nq; ++qp) {
    double Tqp = 0.;
    for (int a = 0; a < nn; ++a) Tqp += fe.N(a, qp) * sol[a];
    const double kT = p->k_of_T ? p->k0 * (1. + p->k_beta * Tqp) : p->k0;
    const double dkT = kind == 1 ? (p->k_of_T ? (1. + p->k_beta * Tqp) : 1.) : 0.;
    const double dcond = dkT * p->thickness + kT * dth;  // d(k h)/dp
    for (int i = 0; i < dim; ++i) {
      double dTi = 0.;
      for (int a = 0; a < nn; ++a) dTi += fe.dN(a, qp, i) * sol[a];
      for (int a = 0; a < nn; ++a) f[a] += fe.JxW[qp] * fe.dN(a, qp, i) * (dcond * dTi);
    }
    if (p->heat_source != 0.)
      for (int a = 0; a < nn; ++a) f[a] -= fe.JxW[qp] * fe.N(a, qp) * (p->heat_source * dth);
  }
}
static void heat_side_sensitivity(const OracleProblem* p, const double* X, const Vec& sol, int side, double flux, double hc,
                                  double Tamb, double emiss, int kind, Vec& f) {
  const double dth = kind == 2 ? 1. : 0.;
  const int n1 = side, n2 = (side + 1) % 4;
  const double dx = X[n2 * 3] - X[n1 * 3], dy = X[n2 * 3 + 1] - X[n1 * 3 + 1], dz = X[n2 * 3 + 2] - X[n1 * 3 + 2];
  const double L = std::sqrt(dx * dx + dy * dy + dz * dz);
  for (int q = 0; q < 2; ++q) {
    const double xi = GP2[q], JxW = 0.5 * L * GW2[q];
    const double N[2] = {0.5 * (1. - xi), 0.5 * (1. + xi)};
    const int nd[2] = {n1, n2};
    const double temp = N[0] * sol[n1] + N[1] * sol[n2];
    for (int a = 0; a < 2; ++a) {
      f[nd[a]] += JxW * N[a] * (flux * dth);
      if (hc != 0.) f[nd[a]] += JxW * N[a] * (dth * hc * (temp - Tamb));
      if (emiss != 0.)
        f[nd[a]] += JxW * N[a] * p->rad_sb_const * (emiss * dth) *
                    (std::pow(temp + p->rad_T_ref_zero, 4.) - std::pow(Tamb + p->rad_T_ref_zero, 4.));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// elem-ops layer
// ---------------------------------------------------------------------------------------------
struct SideLoadIndex {  // (element -> its loaded sides), GeomElem::external_side_loads_for_quadrature_elem
  std::vector<int64_t> ptr;
  std::vector<int32_t> side;
  std::vector<double> press;
  std::vector<int64_t> idx;  // original load index (heat side loads carry more than one value)
  void build(const OracleProblem* p) {
    ptr.assign(p->n_elem + 1, 0);
    for (int64_t l = 0; l < p->n_side_loads; ++l) ptr[p->side_elem[l] + 1]++;
    for (int64_t e = 0; e < p->n_elem; ++e) ptr[e + 1] += ptr[e];
    side.resize(p->n_side_loads); press.resize(p->n_side_loads); idx.resize(p->n_side_loads);
    std::vector<int64_t> pos(ptr.begin(), ptr.end() - 1);
    // std::map<unsigned, ...> iteration is by ascending side id; stable within a side
    std::vector<int64_t> order(p->n_side_loads);
    for (int64_t l = 0; l < p->n_side_loads; ++l) order[l] = l;
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
      if (p->side_elem[a] != p->side_elem[b]) return p->side_elem[a] < p->side_elem[b];
      return p->side_id[a] < p->side_id[b];
    });
    for (int64_t l : order) {
      const int64_t k = pos[p->side_elem[l]]++;
      side[k] = p->side_id[l]; press[k] = p->side_pressure[l]; idx[k] = l;
    }
  }
};

static void gather_coords(const OracleProblem* p, int64_t e, int nn, std::vector<double>& X) {
  X.resize(size_t(nn) * 3);
  for (int a = 0; a < nn; ++a)
    for (int c = 0; c < 3; ++c) X[a * 3 + c] = p->xyz[size_t(p->conn[e * nn + a]) * 3 + c];
}
// dof_map.dof_indices(elem): variable-major list, dof(node,var) = nv*node+var
static void dof_indices(const OracleProblem* p, int64_t e, int nn, std::vector<int64_t>& di) {
  di.resize(size_t(p->nv) * nn);
  for (int v = 0; v < p->nv; ++v)
    for (int a = 0; a < nn; ++a) di[v * nn + a] = int64_t(p->nv) * p->conn[e * nn + a] + v;
}

// StructuralNonlinearAssemblyElemOperations::elem_calculations (structural_nonlinear_assembly.cpp:112-139)
// HeatConductionNonlinearAssemblyElemOperations::elem_calculations (heat_conduction_nonlinear_assembly.cpp:83-101)
static void elem_calculations(const OracleProblem* p, const SideLoadIndex& sl, int64_t e, const Vec& sol, bool if_jac,
                              Vec& vec, Mat& mat) {
  const ElemInfo ei = elem_info(p->elem_type);
  std::vector<double> X;
  gather_coords(p, e, ei.nn, X);
  FE fe;
  fe_init(p->elem_type, X.data(), fe);
  const int nd = p->nv * ei.nn;
  vec.assign(nd, 0.);
  mat.resize(nd, nd);
  if (p->elem_type == ORACLE_QUAD4) {
    if (p->nv == 6) {
      plate_elem_calculations(p, X.data(), sol, if_jac, vec, mat);  // StructuralElement2D
    } else {
      heat_elem_calculations(p, fe, sol, if_jac, vec, mat);
      // side_external_residual of heat conduction (heat_conduction_elem_base.cpp:428-483): HEAT_FLUX, CONVECTION_HEAT_FLUX
      for (int64_t k = sl.ptr[e]; k < sl.ptr[e + 1]; ++k) {
        const int64_t l = sl.idx[k];
        heat_side_residual(p, X.data(), sol, sl.side[k], sl.press[k], p->side_conv_coeff ? p->side_conv_coeff[l] : 0.,
                           p->side_T_amb ? p->side_T_amb[l] : 0., p->side_emissivity ? p->side_emissivity[l] : 0., if_jac, vec, mat);
      }
    }
    return;
  }
  const int32_t* en = &p->conn[e * ei.nn];
  solid3d_internal_residual(p, fe, en, sol, if_jac, vec, mat, nullptr);
  for (int64_t k = sl.ptr[e]; k < sl.ptr[e + 1]; ++k)  // side_external_residual (structural_element_base.cpp:628-701)
    solid3d_surface_pressure_residual(p, X.data(), sl.side[k], sl.press[k], vec);
  if (p->thermal)  // volume_external_residual (structural_element_base.cpp:803-855)
    solid3d_thermal_residual(p, fe, en, sol, if_jac, vec, mat);
}

// external (side + volume) residual only, as ComplianceOutput uses it (compliance_output.cpp:71-102)
static void elem_external_residual(const OracleProblem* p, const SideLoadIndex& sl, int64_t e, const Vec& sol, Vec& vec) {
  const ElemInfo ei = elem_info(p->elem_type);
  std::vector<double> X;
  gather_coords(p, e, ei.nn, X);
  const int nd = p->nv * ei.nn;
  vec.assign(nd, 0.);
  for (int64_t k = sl.ptr[e]; k < sl.ptr[e + 1]; ++k)
    solid3d_surface_pressure_residual(p, X.data(), sl.side[k], sl.press[k], vec);
  if (p->thermal) {
    FE fe;
    fe_init(p->elem_type, X.data(), fe);
    Mat dummy(nd, nd);
    solid3d_thermal_residual(p, fe, &p->conn[e * ei.nn], sol, false, vec, dummy);
  }
}

// libMesh DofMap::constrain_element_matrix_and_vector for homogeneous Dirichlet rows without
// dependencies: K <- C^T K C (row & column zeroed), K(i,i) = 1, f(i) = 0.
static void constrain(const OracleProblem* p, const std::vector<int64_t>& di, Mat* m, Vec* v) {
  if (!p->constrained) return;
  const int nd = int(di.size());
  for (int i = 0; i < nd; ++i)
    if (p->constrained[di[i]]) {
      if (m) {
        for (int j = 0; j < nd; ++j) { (*m)(i, j) = 0.; (*m)(j, i) = 0.; }
        (*m)(i, i) = 1.;
      }
      if (v) (*v)[i] = 0.;
    }
}

static inline int64_t csr_find(const int64_t* row_ptr, const int32_t* col_idx, int64_t row, int64_t col) {
  const int32_t* b = col_idx + row_ptr[row];
  const int32_t* e = col_idx + row_ptr[row + 1];
  const int32_t* it = std::lower_bound(b, e, int32_t(col));  // PETSc MatSetValues row search
  return (it != e && *it == col) ? (it - col_idx) : -1;
}

// inertial / capacitance element matrix (jac_xddot resp. jac_xdot) and vector f_m = M rate:
//   3D solids: StructuralElement3D::inertial_residual (solid_element_3d.cpp:48-130): consistent sum_q JxW N^T rho N on the
//     three translations, or with if_diagonal_mass_matrix vol / n_shape rho on the diagonal;
//   heat conduction: HeatConductionElementBase::velocity_residual (heat_conduction_elem_base.cpp:155-243): sum_q JxW
//     (rho cp) N^T N (times the section thickness for the 2D card), row-sum lumped with if_diagonal_mass_matrix.
static void elem_mass(const OracleProblem* p, const FE& fe, double rho_m, bool lumped, const Vec& rate, Vec& f, Mat& m) {
  const int nn = fe.nn, nd = p->nv * nn;
  m.resize(nd, nd);
  f.assign(nd, 0.);
  if (p->elem_type == ORACLE_QUAD4 && p->nv == 1) {
    const double c = rho_m * p->thickness;
    for (int q = 0; q < fe.nq; ++q)
      for (int a = 0; a < nn; ++a)
        for (int b = 0; b < nn; ++b) m(a, b) += fe.JxW[q] * c * fe.N(a, q) * fe.N(b, q);   // :213
    if (lumped)
      for (int a = 0; a < nn; ++a) {                                                        // :232-236
        double s = 0.;
        for (int b = 0; b < nn; ++b) { s += m(a, b); m(a, b) = 0.; }
        m(a, a) = s;
      }
  } else {
    const int ndim = 3;
    if (lumped) {                                                                            // :84-99
      double vol = 0.;
      for (int q = 0; q < fe.nq; ++q) vol += fe.JxW[q];
      vol /= (1. * nn);
      for (int v = 0; v < ndim; ++v)
        for (int a = 0; a < nn; ++a) m(v * nn + a, v * nn + a) = vol * rho_m;
    } else {                                                                                 // :101-126
      for (int q = 0; q < fe.nq; ++q)
        for (int v = 0; v < ndim; ++v)
          for (int a = 0; a < nn; ++a)
            for (int b = 0; b < nn; ++b) m(v * nn + a, v * nn + b) += fe.JxW[q] * rho_m * fe.N(a, q) * fe.N(b, q);
    }
  }
  for (int i = 0; i < nd; ++i) {
    double s = 0.;
    for (int j = 0; j < nd; ++j) s += m(i, j) * rate[j];
    f[i] = s;
  }
}

// the element loop of NonlinearImplicitAssembly::residual_and_jacobian (nonlinear_implicit_assembly.cpp:111-186)
// restricted to elements [e0,e1) and rows [r0,r1) (r0=0,r1=n_dof for the whole system)
static int assemble_range(const OracleProblem* p, const SideLoadIndex& sl, const double* Xg, double* R, double* Jv,
                          const int64_t* row_ptr, const int32_t* col_idx, int64_t e0, int64_t e1) {
  const ElemInfo ei = elem_info(p->elem_type);
  std::vector<int64_t> di;
  Vec sol, vec;
  Mat mat;
  for (int64_t e = e0; e < e1; ++e) {
    dof_indices(p, e, ei.nn, di);
    const int nd = int(di.size());
    if (p->diag_elem && p->diag_elem[e]) {  // :117-130
      if (Jv) {
        mat.resize(nd, nd);
        for (int i = 0; i < nd; ++i) mat(i, i) = 1.e-24;
        constrain(p, di, &mat, nullptr);
        for (int i = 0; i < nd; ++i)
          for (int j = 0; j < nd; ++j) {
            const int64_t k = csr_find(row_ptr, col_idx, di[i], di[j]);
            if (k < 0) return 2;
            Jv[k] += mat(i, j);
          }
      }
      continue;
    }
    sol.resize(nd);
    for (int i = 0; i < nd; ++i) sol[i] = Xg[di[i]];  // :145-146
    elem_calculations(p, sl, e, sol, Jv != nullptr, vec, mat);  // :157
    constrain(p, di, Jv ? &mat : nullptr, R ? &vec : nullptr);  // :174-179
    if (R)
      for (int i = 0; i < nd; ++i) R[di[i]] += vec[i];  // :182
    if (Jv)
      for (int i = 0; i < nd; ++i)
        for (int j = 0; j < nd; ++j) {  // :183 (MatSetValues ADD_VALUES, explicit zeros kept)
          const int64_t k = csr_find(row_ptr, col_idx, di[i], di[j]);
          if (k < 0) return 2;
          Jv[k] += mat(i, j);
        }
  }
  return 0;
}

}  // namespace

// =============================================================================================
// C API
// =============================================================================================
extern "C" {

int64_t oracle_build_pattern(const OracleProblem* p, int64_t* row_ptr, int32_t* col_idx) {
  // libMesh SparsityPattern::Build: a row couples to every dof of every element touching it (all
  // variables couple); PETSc AIJ keeps columns sorted.  Done node-wise: dofs of a node share columns.
  const ElemInfo ei = elem_info(p->elem_type);
  std::vector<std::vector<int32_t>> nbr(p->n_nodes);
  for (int64_t e = 0; e < p->n_elem; ++e)
    for (int a = 0; a < ei.nn; ++a) {
      std::vector<int32_t>& l = nbr[p->conn[e * ei.nn + a]];
      for (int b = 0; b < ei.nn; ++b) l.push_back(p->conn[e * ei.nn + b]);
    }
  const int nv = p->nv;
  int64_t nnz = 0;
  row_ptr[0] = 0;
  for (int64_t n = 0; n < p->n_nodes; ++n) {
    std::vector<int32_t>& l = nbr[n];
    std::sort(l.begin(), l.end());
    l.erase(std::unique(l.begin(), l.end()), l.end());
    for (int v = 0; v < nv; ++v) {
      if (col_idx) {
        int64_t k = nnz;
        for (int32_t m : l)
          for (int w = 0; w < nv; ++w) col_idx[k++] = m * nv + w;
      }
      nnz += int64_t(l.size()) * nv;
      row_ptr[n * nv + v + 1] = nnz;
    }
  }
  return nnz;
}

int oracle_residual_and_jacobian(const OracleProblem* p, const double* X, double* R, double* J_vals,
                                 const int64_t* row_ptr, const int32_t* col_idx) {
  const int64_t n_dof = p->n_nodes * p->nv;
  if (R) std::fill(R, R + n_dof, 0.);                       // nonlinear_implicit_assembly.cpp:81
  if (J_vals) std::fill(J_vals, J_vals + row_ptr[n_dof], 0.);  // :82
  SideLoadIndex sl;
  sl.build(p);
  return assemble_range(p, sl, X, R, J_vals, row_ptr, col_idx, 0, p->n_elem);
}

int oracle_elem_calculations(const OracleProblem* p, int64_t e, const double* X, int want_jac, double* vec, double* mat) {
  const ElemInfo ei = elem_info(p->elem_type);
  SideLoadIndex sl;
  sl.build(p);
  std::vector<int64_t> di;
  dof_indices(p, e, ei.nn, di);
  Vec sol(di.size()), v;
  for (size_t i = 0; i < di.size(); ++i) sol[i] = X[di[i]];
  Mat m;
  elem_calculations(p, sl, e, sol, want_jac != 0, v, m);
  std::copy(v.begin(), v.end(), vec);
  if (want_jac && mat) std::copy(m.a.begin(), m.a.end(), mat);
  return 0;
}

// NonlinearImplicitAssembly::sensitivity_assemble (nonlinear_implicit_assembly.cpp:515-686) for one DV whose
// nodal density sensitivity is the sparse vector (dv_node, dv_weight)[dv_begin:dv_end]
int oracle_sensitivity_assemble(const OracleProblem* p, const double* X, int64_t dv_begin, int64_t dv_end,
                                const int32_t* dv_node, const double* dv_weight, double* dres_dp) {
  const ElemInfo ei = elem_info(p->elem_type);
  if (p->elem_type == ORACLE_QUAD4) return 3;
  const int64_t n_dof = p->n_nodes * p->nv;
  std::fill(dres_dp, dres_dp + n_dof, 0.);  // :529
  // elements in the DV's support (ElemParameterDependence, :570-573): here the elements touching a node
  // with non-zero density sensitivity; everywhere else dD/dp == 0 identically.
  std::vector<double> drho(p->n_nodes, 0.);
  std::vector<char> touched(p->n_nodes, 0);
  for (int64_t k = dv_begin; k < dv_end; ++k) { drho[dv_node[k]] += dv_weight[k]; touched[dv_node[k]] = 1; }
  std::vector<int64_t> di;
  std::vector<double> Xe, dre(ei.nn);
  Vec sol, vec;
  Mat dummy;
  for (int64_t e = 0; e < p->n_elem; ++e) {
    if (p->diag_elem && p->diag_elem[e]) continue;  // :566-567
    bool dep = false;
    for (int a = 0; a < ei.nn; ++a) dep = dep || touched[p->conn[e * ei.nn + a]];
    if (!dep) continue;
    dof_indices(p, e, ei.nn, di);
    const int nd = int(di.size());
    sol.resize(nd);
    for (int i = 0; i < nd; ++i) sol[i] = X[di[i]];
    gather_coords(p, e, ei.nn, Xe);
    FE fe;
    fe_init(p->elem_type, Xe.data(), fe);
    for (int a = 0; a < ei.nn; ++a) dre[a] = drho[p->conn[e * ei.nn + a]];
    // StructuralNonlinearAssemblyElemOperations::elem_sensitivity_calculations (structural_nonlinear_assembly.cpp:172-197):
    // internal_residual_sensitivity; pressure loads are parameter independent (surface_pressure_residual_sensitivity
    // differentiates the pressure field only); 3D thermal_residual_sensitivity is libmesh_error() (solid_element_3d.cpp:1093-1104)
    if (p->thermal) return 4;
    vec.assign(nd, 0.);
    dummy.resize(nd, nd);
    solid3d_internal_residual(p, fe, &p->conn[e * ei.nn], sol, false, vec, dummy, dre.data());
    constrain(p, di, nullptr, &vec);                          // :613
    for (int i = 0; i < nd; ++i) dres_dp[di[i]] += vec[i];  // :616
  }
  return 0;
}

// NonlinearImplicitAssembly::sensitivity_assemble (nonlinear_implicit_assembly.cpp:515-686) for a Parameter of the
// isotropic property card: kind 1 = E0, 2 = nu; every element depends on it (no ElemParameterDependence skip)
int oracle_sensitivity_assemble_param(const OracleProblem* p, const double* X, int kind, double* dres_dp) {
  const ElemInfo ei = elem_info(p->elem_type);
  if (p->elem_type == ORACLE_QUAD4) {  // heat conduction: kind 1 = k0, kind 2 = section thickness
    if (p->nv != 1) return 3;
    const int64_t n_dof = p->n_nodes;
    std::fill(dres_dp, dres_dp + n_dof, 0.);
    SideLoadIndex sl;
    sl.build(p);
    std::vector<int64_t> di;
    std::vector<double> Xe;
    Vec sol, vec;
    for (int64_t e = 0; e < p->n_elem; ++e) {
      if (p->diag_elem && p->diag_elem[e]) continue;
      dof_indices(p, e, ei.nn, di);
      const int nd = int(di.size());
      sol.resize(nd);
      for (int i = 0; i < nd; ++i) sol[i] = X[di[i]];
      gather_coords(p, e, ei.nn, Xe);
      FE fe;
      fe_init(p->elem_type, Xe.data(), fe);
      vec.assign(nd, 0.);
      heat_elem_sensitivity(p, fe, sol, kind, vec);
      for (int64_t k = sl.ptr[e]; k < sl.ptr[e + 1]; ++k) {
        const int64_t l = sl.idx[k];
        heat_side_sensitivity(p, Xe.data(), sol, sl.side[k], sl.press[k], p->side_conv_coeff ? p->side_conv_coeff[l] : 0.,
                              p->side_T_amb ? p->side_T_amb[l] : 0., p->side_emissivity ? p->side_emissivity[l] : 0., kind, vec);
      }
      constrain(p, di, nullptr, &vec);
      for (int i = 0; i < nd; ++i) dres_dp[di[i]] += vec[i];
    }
    return 0;
  }
  if (p->thermal) return 4;  // 3D thermal_residual_sensitivity is libmesh_error() (solid_element_3d.cpp:1093-1104)
  const int64_t n_dof = p->n_nodes * p->nv;
  std::fill(dres_dp, dres_dp + n_dof, 0.);
  std::vector<int64_t> di;
  std::vector<double> Xe;
  Vec sol, vec;
  Mat dummy;
  for (int64_t e = 0; e < p->n_elem; ++e) {
    if (p->diag_elem && p->diag_elem[e]) continue;
    dof_indices(p, e, ei.nn, di);
    const int nd = int(di.size());
    sol.resize(nd);
    for (int i = 0; i < nd; ++i) sol[i] = X[di[i]];
    gather_coords(p, e, ei.nn, Xe);
    FE fe;
    fe_init(p->elem_type, Xe.data(), fe);
    vec.assign(nd, 0.);
    dummy.resize(nd, nd);
    solid3d_internal_residual(p, fe, &p->conn[e * ei.nn], sol, false, vec, dummy, nullptr, kind);
    constrain(p, di, nullptr, &vec);
    for (int i = 0; i < nd; ++i) dres_dp[di[i]] += vec[i];
  }
  return 0;
}

// NonlinearImplicitAssembly::linearized_jacobian_solution_product (nonlinear_implicit_assembly.cpp:270-383) with
// StructuralNonlinearAssemblyElemOperations::elem_linearized_jacobian_solution_product
// (structural_nonlinear_assembly.cpp:142-166): vec = K_e(sol) dsol (StructuralElementBase::linearized_internal_residual,
// structural_element_base.cpp:224-252; the non-follower pressure and the linear-strain thermal load have no tangent),
// constrain_element_vector, add
int oracle_linearized_jacobian_solution_product(const OracleProblem* p, const double* X, const double* dX, double* out) {
  const ElemInfo ei = elem_info(p->elem_type);
  if (p->elem_type == ORACLE_QUAD4) return 3;
  const int64_t n_dof = p->n_nodes * p->nv;
  std::fill(out, out + n_dof, 0.);
  std::vector<int64_t> di;
  std::vector<double> Xe;
  Vec sol, dsol, v, vec;
  Mat m;
  for (int64_t e = 0; e < p->n_elem; ++e) {
    if (p->diag_elem && p->diag_elem[e]) continue;
    dof_indices(p, e, ei.nn, di);
    const int nd = int(di.size());
    sol.resize(nd); dsol.resize(nd);
    for (int i = 0; i < nd; ++i) { sol[i] = X[di[i]]; dsol[i] = dX[di[i]]; }
    gather_coords(p, e, ei.nn, Xe);
    FE fe;
    fe_init(p->elem_type, Xe.data(), fe);
    v.assign(nd, 0.);
    m.resize(nd, nd);
    solid3d_internal_residual(p, fe, &p->conn[e * ei.nn], sol, true, v, m, nullptr);
    vec = matvec(m, dsol);
    constrain(p, di, nullptr, &vec);
    for (int i = 0; i < nd; ++i) out[di[i]] += vec[i];
  }
  return 0;
}

// the point loads of NonlinearImplicitAssembly::residual_and_jacobian (nonlinear_implicit_assembly.cpp:190-247): for every
// (node, load[nv]) entry vec = -load, constrain_element_vector (constrained dofs zeroed), R.add_vector
int oracle_add_point_loads(const OracleProblem* p, int64_t n, const int32_t* node, const double* load, double* R) {
  for (int64_t i = 0; i < n; ++i)
    for (int v = 0; v < p->nv; ++v) {
      const int64_t d = int64_t(node[i]) * p->nv + v;
      if (p->constrained && p->constrained[d]) continue;
      R[d] += -load[i * p->nv + v];
    }
  return 0;
}

// Mass (capacitance) matrix in the system's pattern: the B matrix of EigenproblemAssembly::eigenproblem_assemble
// (src/base/eigenproblem_assembly.cpp:90-180: constrain_element_matrix on A and B separately, :167-168)
int oracle_mass_matrix(const OracleProblem* p, double rho_m, int lumped, const int64_t* row_ptr, const int32_t* col_idx,
                       double* M_vals) {
  const ElemInfo ei = elem_info(p->elem_type);
  const int64_t n_dof = p->n_nodes * p->nv;
  std::fill(M_vals, M_vals + row_ptr[n_dof], 0.);
  std::vector<int64_t> di;
  std::vector<double> Xe;
  Vec rate, f;
  Mat m;
  for (int64_t e = 0; e < p->n_elem; ++e) {
    if (p->diag_elem && p->diag_elem[e]) continue;
    dof_indices(p, e, ei.nn, di);
    const int nd = int(di.size());
    gather_coords(p, e, ei.nn, Xe);
    FE fe;
    fe_init(p->elem_type, Xe.data(), fe);
    rate.assign(nd, 0.);
    elem_mass(p, fe, rho_m, lumped != 0, rate, f, m);
    constrain(p, di, &m, nullptr);
    for (int i = 0; i < nd; ++i)
      for (int j = 0; j < nd; ++j) {
        const int64_t k = csr_find(row_ptr, col_idx, di[i], di[j]);
        if (k < 0) return 2;
        M_vals[k] += m(i, j);
      }
  }
  return 0;
}

// TransientAssembly::residual_and_jacobian (src/base/transient_assembly.cpp:63-180) with the element combination of the
// transient solvers: vec = f_m + f_x, mat = jac_coeff * df_m/d(rate) + df_x/dx -- SecondOrderNewmarkTransientSolver::
// elem_calculations (src/solver/second_order_newmark_transient_solver.cpp:277-380: jac_coeff = 1 / beta / dt^2, rate = the
// acceleration it derives from x) and FirstOrderNewmarkTransientSolver (jac_coeff = 1 / beta / dt, rate = the velocity);
// the combined element matrix and vector are constrained together (:150-156)
int oracle_transient_residual_and_jacobian(const OracleProblem* p, const double* X, const double* rate, double rho_m, int lumped,
                                           double jac_coeff, double* R, double* J_vals, const int64_t* row_ptr,
                                           const int32_t* col_idx) {
  const ElemInfo ei = elem_info(p->elem_type);
  const int64_t n_dof = p->n_nodes * p->nv;
  if (R) std::fill(R, R + n_dof, 0.);
  if (J_vals) std::fill(J_vals, J_vals + row_ptr[n_dof], 0.);
  SideLoadIndex sl;
  sl.build(p);
  std::vector<int64_t> di;
  std::vector<double> Xe;
  Vec sol, rt, vec, fm;
  Mat mat, mm;
  for (int64_t e = 0; e < p->n_elem; ++e) {
    dof_indices(p, e, ei.nn, di);
    const int nd = int(di.size());
    sol.resize(nd); rt.resize(nd);
    for (int i = 0; i < nd; ++i) { sol[i] = X[di[i]]; rt[i] = rate[di[i]]; }
    elem_calculations(p, sl, e, sol, J_vals != nullptr, vec, mat);
    gather_coords(p, e, ei.nn, Xe);
    FE fe;
    fe_init(p->elem_type, Xe.data(), fe);
    elem_mass(p, fe, rho_m, lumped != 0, rt, fm, mm);
    for (int i = 0; i < nd; ++i) {
      vec[i] += fm[i];
      if (J_vals)
        for (int j = 0; j < nd; ++j) mat(i, j) += jac_coeff * mm(i, j);
    }
    constrain(p, di, J_vals ? &mat : nullptr, R ? &vec : nullptr);
    if (R)
      for (int i = 0; i < nd; ++i) R[di[i]] += vec[i];
    if (J_vals)
      for (int i = 0; i < nd; ++i)
        for (int j = 0; j < nd; ++j) {
          const int64_t k = csr_find(row_ptr, col_idx, di[i], di[j]);
          if (k < 0) return 2;
          J_vals[k] += mat(i, j);
        }
  }
  return 0;
}

// AssemblyBase::calculate_output_adjoint_sensitivity_multiple_parameters_no_direct (assembly_base.cpp:594-639)
int oracle_adjoint_sensitivity(const OracleProblem* p, const double* X, const double* adj, int64_t n_dv,
                               const int64_t* dv_ptr, const int32_t* dv_node, const double* dv_weight, double* sens) {
  const int64_t n_dof = p->n_nodes * p->nv;
  std::vector<double> dres(n_dof);
  for (int64_t i = 0; i < n_dv; ++i) {
    const int rc = oracle_sensitivity_assemble(p, X, dv_ptr[i], dv_ptr[i + 1], dv_node, dv_weight, dres.data());
    if (rc) return rc;
    double s = 0.;
    for (int64_t k = 0; k < n_dof; ++k) s += adj[k] * dres[k];  // adj_sol.dot(dres_dp)  :636-637
    sens[i] = s;
  }
  return 0;
}

// AssemblyBase::calculate_output (assembly_base.cpp:253-333) with ComplianceOutput::evaluate (compliance_output.cpp:71-102)
double oracle_compliance(const OracleProblem* p, const double* X) {
  const ElemInfo ei = elem_info(p->elem_type);
  SideLoadIndex sl;
  sl.build(p);
  std::vector<int64_t> di;
  Vec sol, vec;
  double q = 0.;
  for (int64_t e = 0; e < p->n_elem; ++e) {
    if (p->diag_elem && p->diag_elem[e]) continue;
    dof_indices(p, e, ei.nn, di);
    sol.resize(di.size());
    for (size_t i = 0; i < di.size(); ++i) sol[i] = X[di[i]];
    elem_external_residual(p, sl, e, sol, vec);
    double d = 0.;
    for (size_t i = 0; i < di.size(); ++i) d += vec[i] * sol[i];
    q -= d;  // _compliance -= vec.dot(e.sol())
  }
  return q;
}

// AssemblyBase::calculate_output_derivative (assembly_base.cpp:338-430) with
// ComplianceOutput::output_derivative_for_elem (compliance_output.cpp:265-296)
int oracle_compliance_derivative(const OracleProblem* p, const double* X, double* dq_dX) {
  const ElemInfo ei = elem_info(p->elem_type);
  const int64_t n_dof = p->n_nodes * p->nv;
  std::fill(dq_dX, dq_dX + n_dof, 0.);
  SideLoadIndex sl;
  sl.build(p);
  std::vector<int64_t> di;
  Vec sol, vec;
  for (int64_t e = 0; e < p->n_elem; ++e) {
    if (p->diag_elem && p->diag_elem[e]) continue;
    dof_indices(p, e, ei.nn, di);
    sol.resize(di.size());
    for (size_t i = 0; i < di.size(); ++i) sol[i] = X[di[i]];
    elem_external_residual(p, sl, e, sol, vec);
    for (double& v : vec) v *= -1.;
    constrain(p, di, nullptr, &vec);
    for (size_t i = 0; i < di.size(); ++i) dq_dX[di[i]] += vec[i];
  }
  return 0;
}

// StructuralElement3D::calculate_stress (solid_element_3d.cpp:1146-1265) + Data::von_Mises_stress
// (stress_output_base.cpp:182-195)
static double von_mises(const double* s) {
  return std::pow(0.5 * (std::pow(s[0] - s[1], 2) + std::pow(s[1] - s[2], 2) + std::pow(s[2] - s[0], 2)) +
                      3.0 * (std::pow(s[3], 2) + std::pow(s[4], 2) + std::pow(s[5], 2)),
                  0.5);
}

int oracle_stress(const OracleProblem* p, const double* X, double* stress, double* strain, double* vm, double* jxw) {
  const ElemInfo ei = elem_info(p->elem_type);
  if (ei.dim != 3) return 3;
  std::vector<int64_t> di;
  std::vector<double> Xe;
  Vec sol;
  Mat D;
  StrainOps s;
  Modulus mod{p};
  for (int64_t e = 0; e < p->n_elem; ++e) {
    dof_indices(p, e, ei.nn, di);
    sol.resize(di.size());
    for (size_t i = 0; i < di.size(); ++i) sol[i] = X[di[i]];
    gather_coords(p, e, ei.nn, Xe);
    FE fe;
    fe_init(p->elem_type, Xe.data(), fe);
    Vec disp(sol.begin(), sol.begin() + 3 * ei.nn);
    strain_ops_reinit(s, ei.nn);
    for (int qp = 0; qp < fe.nq; ++qp) {
      stiffness_3d(mod.E(fe, &p->conn[e * ei.nn], qp), p->nu, D);
      init_strain_operator(fe, qp, p->nonlinear_strain != 0, disp, s);
      Vec sig = matvec(D, s.eps);
      const size_t o = size_t(e) * ei.nq + qp;
      for (int i = 0; i < 6; ++i) { stress[o * 6 + i] = sig[i]; strain[o * 6 + i] = s.eps[i]; }
      vm[o] = von_mises(sig.data());
      jxw[o] = fe.JxW[qp];
    }
  }
  return 0;
}

// StressStrainOutputBase::functional_for_all_elems (stress_output_base.cpp:851-909) and, if dq_dX,
// functional_state_derivartive_for_elem (:1100-1172) with dstress/dX = D*B_lin (the 3D derivative is
// libmesh_error() in the reference, solid_element_3d.cpp:1258-1261; pattern of structural_element_2d.cpp:555-611),
// driven by AssemblyBase::calculate_output_derivative (assembly_base.cpp:338-430).
double oracle_stress_functional(const OracleProblem* p, const double* X, double p_norm, double rho_exp, double sigma0,
                                double* dq_dX) {
  const ElemInfo ei = elem_info(p->elem_type);
  const double p_norm_weight = 1., exp_arg_lim = 1.e2;  // stress_output_base.cpp:278-281
  const size_t nqp = size_t(p->n_elem) * ei.nq;
  std::vector<double> st(nqp * 6), en_(nqp * 6), vm(nqp), jw(nqp);
  if (oracle_stress(p, X, st.data(), en_.data(), vm.data(), jw.data())) return std::nan("");
  double sigma_vm_int = 0., JxW_val = 0.;
  for (size_t k = 0; k < nqp; ++k) {
    const double e_val = vm[k];
    const double sp = std::pow((e_val - sigma0) / sigma0, p_norm_weight);
    const double exp_sp = (rho_exp * sp > exp_arg_lim) ? std::exp(exp_arg_lim) : std::exp(rho_exp * sp);
    sigma_vm_int += std::pow(e_val / sigma0, p_norm) * exp_sp * jw[k];
    JxW_val += exp_sp * jw[k];
  }
  const double q = sigma0 * std::pow(sigma_vm_int / JxW_val, 1. / p_norm);
  if (!dq_dX) return q;
  const int64_t n_dof = p->n_nodes * p->nv;
  std::fill(dq_dX, dq_dX + n_dof, 0.);
  std::vector<int64_t> di;
  std::vector<double> Xe;
  Mat D, DB;
  StrainOps s;
  Modulus mod{p};
  for (int64_t e = 0; e < p->n_elem; ++e) {
    dof_indices(p, e, ei.nn, di);
    const int nd = int(di.size()), n2 = 3 * ei.nn;
    gather_coords(p, e, ei.nn, Xe);
    FE fe;
    fe_init(p->elem_type, Xe.data(), fe);
    Vec disp(n2);
    for (int i = 0; i < n2; ++i) disp[i] = X[di[i]];
    strain_ops_reinit(s, ei.nn);
    Vec denom(nd, 0.), num(nd, 0.), de(nd, 0.), out(nd, 0.);
    for (int qp = 0; qp < fe.nq; ++qp) {
      const size_t o = size_t(e) * ei.nq + qp;
      stiffness_3d(mod.E(fe, &p->conn[e * ei.nn], qp), p->nu, D);
      init_strain_operator(fe, qp, false, disp, s);
      s.lin.left_multiply(DB, D);  // dstress_dX (6 x 3nn)
      const double* sg = &st[o * 6];
      // Data::dvon_Mises_stress_dX (stress_output_base.cpp:199-230)
      const double pp = 0.5 * (std::pow(sg[0] - sg[1], 2) + std::pow(sg[1] - sg[2], 2) + std::pow(sg[2] - sg[0], 2)) +
                        3.0 * (std::pow(sg[3], 2) + std::pow(sg[4], 2) + std::pow(sg[5], 2));
      std::fill(de.begin(), de.end(), 0.);
      if (std::fabs(pp) > 0.)
        for (int j = 0; j < n2; ++j)
          de[j] = (((DB(0, j) - DB(1, j)) * (sg[0] - sg[1]) + (DB(1, j) - DB(2, j)) * (sg[1] - sg[2]) +
                    (DB(2, j) - DB(0, j)) * (sg[2] - sg[0])) +
                   6.0 * (DB(3, j) * sg[3] + DB(4, j) * sg[4] + DB(5, j) * sg[5])) *
                  0.5 * std::pow(pp, -0.5);
      const double e_val = vm[o], JxW = jw[o];
      const double sp_stress = std::pow(e_val / sigma0, p_norm), sp1_stress = std::pow(e_val / sigma0, p_norm - 1.);
      const double sp_weight = std::pow((e_val - sigma0) / sigma0, p_norm_weight);
      const double sp1_weight = std::pow((e_val - sigma0) / sigma0, p_norm_weight - 1.);
      if (rho_exp * sp_weight > exp_arg_lim) {
        const double exp_sp = std::exp(exp_arg_lim);
        for (int j = 0; j < nd; ++j) num[j] += 1. * p_norm * sp1_stress * exp_sp / sigma0 * JxW * de[j];
      } else {
        const double exp_sp = std::exp(rho_exp * sp_weight);
        for (int j = 0; j < nd; ++j) {
          denom[j] += exp_sp * rho_exp * p_norm_weight * sp1_weight / sigma0 * JxW * de[j];
          num[j] += (de[j] / sigma0 * JxW * exp_sp * (p_norm * sp1_stress + sp_stress * rho_exp * p_norm_weight * sp1_weight));
        }
      }
    }
    for (int j = 0; j < nd; ++j)
      out[j] = sigma0 / p_norm * std::pow(sigma_vm_int / JxW_val, 1. / p_norm - 1.) *
               (num[j] / JxW_val - sigma_vm_int / std::pow(JxW_val, 2.) * denom[j]);
    constrain(p, di, nullptr, &out);
    for (int j = 0; j < nd; ++j) dq_dX[di[j]] += out[j];
  }
  return q;
}

// StressStrainOutputBase::functional_sensitivity_for_elem / _for_all_elems (stress_output_base.cpp:911-1043) for
// nodal-density design variables: the partial derivative of the aggregate at fixed state.  The stress sensitivity of
// the 3D element is absent from the reference (solid_element_3d.cpp:1146-1265 has no parameter branch, so
// Data::dvon_Mises_stress_dp returns 0 there); as for dstress/dX it follows the 2D element's pattern
// (structural_element_2d.cpp:664-677): dstress/dp = dD/dp * strain, with dD/dp of
// isotropic_material_property_card.cpp:533-567 and the SIMP law; dvon_Mises_stress_dp of stress_output_base.cpp:237-271.
int oracle_stress_functional_density_sensitivity(const OracleProblem* p, const double* X, double p_norm, double rho_exp,
                                                 double sigma0, int64_t n_dv, const int64_t* dv_ptr, const int32_t* dv_node,
                                                 const double* dv_weight, double* sens) {
  const ElemInfo ei = elem_info(p->elem_type);
  if (ei.dim != 3 || !p->rho) return 3;
  const double p_norm_weight = 1., exp_arg_lim = 1.e2;
  const size_t nqp = size_t(p->n_elem) * ei.nq;
  std::vector<double> st(nqp * 6), en_(nqp * 6), vm(nqp), jw(nqp);
  if (oracle_stress(p, X, st.data(), en_.data(), vm.data(), jw.data())) return 3;
  double sigma_vm_int = 0., JxW_val = 0.;
  for (size_t k = 0; k < nqp; ++k) {
    const double sp = std::pow((vm[k] - sigma0) / sigma0, p_norm_weight);
    const double exp_sp = (rho_exp * sp > exp_arg_lim) ? std::exp(exp_arg_lim) : std::exp(rho_exp * sp);
    sigma_vm_int += std::pow(vm[k] / sigma0, p_norm) * exp_sp * jw[k];
    JxW_val += exp_sp * jw[k];
  }
  std::vector<double> drho(size_t(p->n_nodes));
  std::vector<double> Xe, dn(ei.nn);
  Modulus mod{p};
  Mat dD;
  for (int64_t i = 0; i < n_dv; ++i) {
    std::fill(drho.begin(), drho.end(), 0.);
    for (int64_t k = dv_ptr[i]; k < dv_ptr[i + 1]; ++k) drho[dv_node[k]] += dv_weight[k];
    double total = 0.;
    for (int64_t e = 0; e < p->n_elem; ++e) {
      const int32_t* en = &p->conn[e * ei.nn];
      bool dep = false;
      for (int a = 0; a < ei.nn; ++a) { dn[a] = drho[en[a]]; dep = dep || dn[a] != 0.; }
      if (!dep || (p->diag_elem && p->diag_elem[e])) continue;
      gather_coords(p, e, ei.nn, Xe);
      FE fe;
      fe_init(p->elem_type, Xe.data(), fe);
      double num_sens = 0., denom_sens = 0.;
      for (int qp = 0; qp < fe.nq; ++qp) {
        const size_t o = size_t(e) * ei.nq + qp;
        stiffness_3d_derivative(mod.E(fe, en, qp), p->nu, mod.dE(fe, en, qp, dn.data()), 0., dD);
        const Vec eps(en_.begin() + o * 6, en_.begin() + o * 6 + 6);
        const Vec ds = matvec(dD, eps);  // dstress/dp
        const double* sg = &st[o * 6];
        const double pp = 0.5 * (std::pow(sg[0] - sg[1], 2) + std::pow(sg[1] - sg[2], 2) + std::pow(sg[2] - sg[0], 2)) +
                          3.0 * (std::pow(sg[3], 2) + std::pow(sg[4], 2) + std::pow(sg[5], 2));
        double de_val = 0.;
        if (std::fabs(pp) > 0.)
          de_val = (((ds[0] - ds[1]) * (sg[0] - sg[1]) + (ds[1] - ds[2]) * (sg[1] - sg[2]) + (ds[2] - ds[0]) * (sg[2] - sg[0])) +
                    6.0 * (ds[3] * sg[3] + ds[4] * sg[4] + ds[5] * sg[5])) * 0.5 * std::pow(pp, -0.5);
        const double e_val = vm[o], JxW = jw[o];
        const double sp_stress = std::pow(e_val / sigma0, p_norm), sp1_stress = std::pow(e_val / sigma0, p_norm - 1.);
        const double sp_weight = std::pow((e_val - sigma0) / sigma0, p_norm_weight);
        const double sp1_weight = std::pow((e_val - sigma0) / sigma0, p_norm_weight - 1.);
        if (rho_exp * sp_weight > exp_arg_lim) {
          num_sens += p_norm * sp1_stress * std::exp(exp_arg_lim) * de_val / sigma0 * JxW;
        } else {
          const double exp_sp = std::exp(rho_exp * sp_weight);
          denom_sens += exp_sp * rho_exp * p_norm_weight * sp1_weight * de_val / sigma0 * JxW;
          num_sens += (de_val / sigma0 * JxW * exp_sp * (p_norm * sp1_stress + sp_stress * rho_exp * p_norm_weight * sp1_weight));
        }
      }
      total += sigma0 / p_norm * std::pow(sigma_vm_int / JxW_val, 1. / p_norm - 1.) *
               (num_sens / JxW_val - sigma_vm_int / std::pow(JxW_val, 2.) * denom_sens);
    }
    sens[i] = total;
  }
  return 0;
}

// MAST::FilterBase::_init (src/level_set/filter_base.cpp:378-447, the linear search; _init2 :275-372 finds the same set
// with a KD-tree): for every node 1 all nodes 2 with |x_1 - x_2| <= radius, weight (radius - d_12) / sum, node 2 ascending.
// row_ptr[n_nodes+1]; call with col == NULL to size.  Returns nnz.
int64_t oracle_filter_weights(int64_t n_nodes, const double* xyz, double radius, int64_t* row_ptr, int32_t* col, double* weight) {
  int64_t k = 0;
  row_ptr[0] = 0;
  for (int64_t a = 0; a < n_nodes; ++a) {
    double sum = 0.;
    const int64_t k0 = k;
    for (int64_t b = 0; b < n_nodes; ++b) {
      const double dx = xyz[a * 3] - xyz[b * 3], dy = xyz[a * 3 + 1] - xyz[b * 3 + 1], dz = xyz[a * 3 + 2] - xyz[b * 3 + 2];
      const double d12 = std::sqrt(dx * dx + dy * dy + dz * dz);
      if (d12 <= radius) {                                                  // :419
        sum += radius - d12;                                                // :421
        if (col) { col[k] = (int32_t)b; weight[k] = radius - d12; }        // :424
        ++k;
      }
    }
    if (col)
      for (int64_t i = k0; i < k; ++i) weight[i] /= sum;                   // :434-438
    row_ptr[a + 1] = k;
  }
  return k;
}

// FilterBase::compute_filtered_values (filter_base.cpp:59-97): filtered value for the design-variable nodes (dv == NULL:
// all), the input itself for the others
void oracle_filtered_values(int64_t n_nodes, const int64_t* row_ptr, const int32_t* col, const double* weight,
                            const uint8_t* dv, const double* input, double* output) {
  for (int64_t n = 0; n < n_nodes; ++n) {
    output[n] = 0.;
    for (int64_t k = row_ptr[n]; k < row_ptr[n + 1]; ++k) {
      if (!dv || dv[n]) output[n] += input[col[k]] * weight[k];              // :86
      else output[n] = input[n];                                           // :88
    }
  }
}

// material volume of examples/structural/example_6/example_6.cpp:714-745: sum over the elements of
// volume(e) * mean nodal density; dV_drho (may be NULL): derivative w.r.t. every nodal density (:748-788 for one DV)
double oracle_volume(const OracleProblem* p, const double* rho, double* dV_drho) {
  const ElemInfo ei = elem_info(p->elem_type);
  if (dV_drho) std::fill(dV_drho, dV_drho + p->n_nodes, 0.);
  std::vector<double> Xe;
  double V = 0.;
  for (int64_t e = 0; e < p->n_elem; ++e) {
    if (p->diag_elem && p->diag_elem[e]) continue;
    gather_coords(p, e, ei.nn, Xe);
    FE fe;
    fe_init(p->elem_type, Xe.data(), fe);
    double vol = 0.;
    for (int q = 0; q < fe.nq; ++q) vol += fe.JxW[q];
    double r = 0.;
    for (int a = 0; a < ei.nn; ++a) r += rho[p->conn[e * ei.nn + a]];
    r /= (1. * ei.nn);
    V += vol * r;
    if (dV_drho)
      for (int a = 0; a < ei.nn; ++a) dV_drho[p->conn[e * ei.nn + a]] += vol / (1. * ei.nn);
  }
  return V;
}


void oracle_plate_section_matrices(double E, double nu, double h, double off, double kappa, int plane_strain,
                                   double* A9, double* B9, double* D9, double* S4) {
  Mat C;
  stiffness_2d(E, nu, plane_strain == 0, C);
  const Mat A = scaled(C, h), B = scaled(C, h * off), D = scaled(C, std::pow(h, 3) / 12. + h * std::pow(off, 2));
  std::copy(A.a.begin(), A.a.end(), A9);
  std::copy(B.a.begin(), B.a.end(), B9);
  std::copy(D.a.begin(), D.a.end(), D9);
  Mat G(2, 2);
  G(0, 0) = G(1, 1) = E / 2. / (1. + nu);
  G = scaled(scaled(G, h), kappa);
  std::copy(G.a.begin(), G.a.end(), S4);
}

void oracle_stiffness_3d(double E, double nu, double* D36) {
  Mat m;
  stiffness_3d(E, nu, m);
  std::copy(m.a.begin(), m.a.end(), D36);
}
void oracle_stiffness_3d_derivative(double E, double nu, double dE, double dnu, double* D36) {
  Mat m;
  stiffness_3d_derivative(E, nu, dE, dnu, m);
  std::copy(m.a.begin(), m.a.end(), D36);
}
void oracle_thermal_expansion_A_3d(double E, double nu, double alpha, double* v6) {
  Vec v = thermal_expansion_A_3d(E, nu, alpha);
  std::copy(v.begin(), v.end(), v6);
}

int oracle_fe_tables(int32_t elem_type, const double* coords, double* phi, double* dphi, double* jxw, double* xyz_qp) {
  FE fe;
  fe_init(elem_type, coords, fe);
  std::copy(fe.phi.begin(), fe.phi.end(), phi);
  std::copy(fe.dphi.begin(), fe.dphi.end(), dphi);
  std::copy(fe.JxW.begin(), fe.JxW.end(), jxw);
  std::copy(fe.xyz.begin(), fe.xyz.end(), xyz_qp);
  return 0;
}

// Timed CPU baseline: `threads` workers stand in for MPI ranks.  Each owns a contiguous element block and adds into
// a private copy of the rows it touches outside its own row block; here, because rows are shared only at block
// interfaces, workers run the element loop on disjoint element ranges into the SAME arrays in two passes
// (even blocks, then odd blocks) so that no two concurrent workers touch the same row.  The arithmetic and the
// per-row binary search are exactly those of the serial path.
// threads_used (may be NULL): the number of threads the sweep really ran with -- the block scheme below needs blocks of
// equal parity to touch disjoint node ranges and halves the thread count until they do
double oracle_time_residual_and_jacobian_ex(const OracleProblem* p, const double* X, double* R, double* J_vals,
                                            const int64_t* row_ptr, const int32_t* col_idx, int threads, int* threads_used) {
  const int64_t n_dof = p->n_nodes * p->nv;
  const ElemInfo ei = elem_info(p->elem_type);
  SideLoadIndex sl;
  sl.build(p);
  // 2*threads contiguous element blocks, even blocks first, then odd ones: one std::thread per block of a parity, like
  // MPI ranks owning contiguous element ranges.  Two blocks of equal parity must not share a node (they ADD into the same
  // rows otherwise): checked on the actual node ranges, not assumed.
  auto disjoint = [&](int nb) {
    std::vector<int64_t> lo(nb), hi(nb);
    for (int b = 0; b < nb; ++b) {
      const int64_t e0 = p->n_elem * b / nb, e1 = p->n_elem * (b + 1) / nb;
      int64_t mn = INT64_MAX, mx = -1;
      for (int64_t k = e0 * ei.nn; k < e1 * ei.nn; ++k) { mn = std::min<int64_t>(mn, p->conn[k]); mx = std::max<int64_t>(mx, p->conn[k]); }
      lo[b] = mn; hi[b] = mx;
    }
    for (int b = 0; b + 2 < nb; ++b)
      for (int c = b + 2; c < nb; c += 2)
        if (hi[b] >= lo[c] && hi[c] >= lo[b]) return false;
    return true;
  };
  while (threads > 1 && !disjoint(2 * threads)) threads /= 2;
  if (threads_used) *threads_used = std::max(threads, 1);
  const auto t0 = std::chrono::steady_clock::now();
  if (R) std::fill(R, R + n_dof, 0.);
  if (J_vals) std::fill(J_vals, J_vals + row_ptr[n_dof], 0.);
  if (threads <= 1) {
    assemble_range(p, sl, X, R, J_vals, row_ptr, col_idx, 0, p->n_elem);
  } else {
    const int nb = 2 * threads;
    for (int parity = 0; parity < 2; ++parity) {
      std::vector<std::thread> pool;
      for (int b = parity; b < nb; b += 2) {
        const int64_t e0 = p->n_elem * b / nb, e1 = p->n_elem * (b + 1) / nb;
        pool.emplace_back([=, &sl] { assemble_range(p, sl, X, R, J_vals, row_ptr, col_idx, e0, e1); });
      }
      for (auto& t : pool) t.join();
    }
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

double oracle_time_residual_and_jacobian(const OracleProblem* p, const double* X, double* R, double* J_vals,
                                         const int64_t* row_ptr, const int32_t* col_idx, int threads) {
  return oracle_time_residual_and_jacobian_ex(p, X, R, J_vals, row_ptr, col_idx, threads, nullptr);
}

}  // extern "C"
