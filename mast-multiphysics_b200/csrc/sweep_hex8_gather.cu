// sweep_hex8_gather.cu -- HEX8 residual + Jacobian with Green-Lagrange (or linear) strain as a row-owner gather
// (sm_100a, fp64).
//
// Device counterpart of the element loop of MAST::NonlinearImplicitAssembly::residual_and_jacobian
// (/root/reference/src/base/nonlinear_implicit_assembly.cpp:111-186) with StructuralElement3D::internal_residual for
// strain_type() == NONLINEAR_STRAIN (src/elasticity/solid_element_3d.cpp:135-423: material tangent :327-384, geometric
// stiffness :386-409, rotational diagonal :413-417), the strain operator of :1308-1463 and ::thermal_residual
// (:950-1089, including the geometric term of :1054-1082).
//
// The reference expands B^T D B with B = B_lin + mat_x nl_x + mat_y nl_y + mat_z nl_z into sixteen operator
// products; here B_a (6x3 per node) is formed directly from the deformation gradient F = I + grad u:
//   B_a[(ij)][k] = F_ki d_jN_a + F_kj d_iN_a (shear rows, engineering), F_ki d_iN_a (normal rows),
// so K_ab = sum_q JxW (B_a^T D B_b + (grad N_a . S grad N_b) I) and f_a = sum_q JxW B_a^T S, S = D E_GL - D alpha dT.
// With linear strain F = I and the geometric term is dropped, which reproduces the linear element.
//
// One warp per owned node (= nv CSR rows).  The warp walks the elements adjacent to its node in ascending element id
// (the order of the reference's element loop).  Per element, lanes 0-7 evaluate one quadrature point each (J^-1, F,
// stress, weights -> shared memory) and the node's own force row; then lane c < nnb, which owns the 3x3 block of the
// row's c-th neighbour, adds the element's contribution if that neighbour is a node of the element.  No atomics, no
// colours: every CSR value is accumulated in one register in a fixed order and stored once.  Redundant geometry (an
// element is evaluated by each of its eight nodes) is the price; the linear-strain production path is the streaming
// sweep of sweep_hex8.cu, this kernel serves the nonlinear strain measure.
#include <cuda_runtime.h>

#include "kernels.h"

namespace mastb200 {
namespace {

constexpr int GW = 4;            // warps (nodes) per CTA
constexpr int QD = 27;           // doubles staged per quadrature point: inv[9], F[9], S[6], wE, w, sth
constexpr double GP = 0.57735026918962576451;

__device__ __forceinline__ void ref_shape(int b, int q, double& N, double d[3]) {
  const double sx = (b == 1 || b == 2 || b == 5 || b == 6) ? 1. : -1.;
  const double sy = (b == 2 || b == 3 || b == 6 || b == 7) ? 1. : -1.;
  const double sz = (b >= 4) ? 1. : -1.;
  const double fx = 1. + sx * ((q & 1) ? GP : -GP), fy = 1. + sy * ((q & 2) ? GP : -GP), fz = 1. + sz * ((q & 4) ? GP : -GP);
  N = 0.125 * fx * fy * fz;
  d[0] = 0.125 * sx * fy * fz; d[1] = 0.125 * sy * fx * fz; d[2] = 0.125 * sz * fx * fy;
}

// B_a (6x3, Voigt rows xx,yy,zz,xy,yz,zx with engineering shear) from F and g = grad N_a
__device__ __forceinline__ void make_B(const double F[3][3], const double g[3], double B[6][3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    B[0][k] = F[k][0] * g[0];
    B[1][k] = F[k][1] * g[1];
    B[2][k] = F[k][2] * g[2];
    B[3][k] = F[k][0] * g[1] + F[k][1] * g[0];
    B[4][k] = F[k][1] * g[2] + F[k][2] * g[1];
    B[5][k] = F[k][0] * g[2] + F[k][2] * g[0];
  }
}

__global__ void __launch_bounds__(GW * 32) hex8_gather_kernel(const SolidGatherParams p) {
  __shared__ double stage_all[GW][8 * QD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * GW + warp;
  if (r >= p.n_owned) return;
  double* stage = stage_all[warp];
  const int n = p.owned_nodes[r];
  const int64_t nb0 = p.nbr_ptr[r];
  const int nnb = (int)(p.nbr_ptr[r + 1] - nb0);
  const int my_nbr = lane < nnb ? p.nbr[nb0 + lane] : -1;
  const unsigned cmn = p.cmask ? p.cmask[n] : 0u;
  const unsigned cmm = (p.cmask && my_nbr >= 0) ? p.cmask[my_nbr] : 0u;
  const bool nl = p.nonlinear != 0;

  double K[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};
  double f[3] = {0., 0., 0.};
  double rot = 0.;
  int nadj = 0, nadjd = 0;

  for (int64_t ke = p.n2e_ptr[n]; ke < p.n2e_ptr[n + 1]; ++ke) {
    const int e = p.n2e[ke];
    if (p.diag_elem && p.diag_elem[e]) { ++nadjd; continue; }  // nonlinear_implicit_assembly.cpp:117-130
    ++nadj;
    int m[8];
    int a = 0, bme = -1;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      m[b] = p.conn[(int64_t)e * 8 + b];
      if (m[b] == n) a = b;
      if (m[b] == my_nbr) bme = b;
    }
    __syncwarp();
    double fq[3] = {0., 0., 0.};
    if (lane < 8) {
      const int q = lane;
      double Jm[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};
      double Nn[8], dref[8][3];
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        ref_shape(b, q, Nn[b], dref[b]);
        const double* X = p.xyz + (int64_t)m[b] * 3;
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
#pragma unroll
          for (int c = 0; c < 3; ++c) Jm[rr][c] = fma(dref[b][rr], X[c], Jm[rr][c]);
      }
      const double c00 = Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1];
      const double c01 = Jm[1][2] * Jm[2][0] - Jm[1][0] * Jm[2][2];
      const double c02 = Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0];
      const double det = Jm[0][0] * c00 + Jm[0][1] * c01 + Jm[0][2] * c02;
      const double idet = 1. / det;
      double inv[3][3];  // inv[c][r] = d xi_r / d x_c
      inv[0][0] = c00 * idet; inv[1][0] = c01 * idet; inv[2][0] = c02 * idet;
      inv[0][1] = (Jm[0][2] * Jm[2][1] - Jm[0][1] * Jm[2][2]) * idet;
      inv[1][1] = (Jm[0][0] * Jm[2][2] - Jm[0][2] * Jm[2][0]) * idet;
      inv[2][1] = (Jm[0][1] * Jm[2][0] - Jm[0][0] * Jm[2][1]) * idet;
      inv[0][2] = (Jm[0][1] * Jm[1][2] - Jm[0][2] * Jm[1][1]) * idet;
      inv[1][2] = (Jm[0][2] * Jm[1][0] - Jm[0][0] * Jm[1][2]) * idet;
      inv[2][2] = (Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0]) * idet;
      double H[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};  // H[i][j] = d u_i / d x_j
      double rho = 0., T = 0., ga[3] = {0., 0., 0.};
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        double g[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) g[c] = dref[b][0] * inv[c][0] + dref[b][1] * inv[c][1] + dref[b][2] * inv[c][2];
        const double* U = p.X + (int64_t)m[b] * p.nv;
        const double u0 = U[0], u1 = U[1], u2 = U[2];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          H[0][j] = fma(u0, g[j], H[0][j]); H[1][j] = fma(u1, g[j], H[1][j]); H[2][j] = fma(u2, g[j], H[2][j]);
        }
        if (p.rho) rho = fma(Nn[b], p.rho[m[b]], rho);
        if (p.temperature) T = fma(Nn[b], p.temperature[m[b]], T);
        if (b == a) { ga[0] = g[0]; ga[1] = g[1]; ga[2] = g[2]; }
      }
      double E = p.E0;
      if (p.rho) E = p.E0 * (p.rho_min + (1. - p.rho_min) * simp_pow(rho, p.penalty));  // example_6.cpp:100-137
      double eps[6];
      double F[3][3];
      if (nl) {  // E = 1/2 (H + H^T + H^T H), solid_element_3d.cpp:1404-1440
        double Eg[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j)
            Eg[i][j] = 0.5 * (H[i][j] + H[j][i] + (H[0][i] * H[0][j] + H[1][i] * H[1][j] + H[2][i] * H[2][j]));
        eps[0] = Eg[0][0]; eps[1] = Eg[1][1]; eps[2] = Eg[2][2];
        eps[3] = Eg[0][1] + Eg[1][0]; eps[4] = Eg[1][2] + Eg[2][1]; eps[5] = Eg[0][2] + Eg[2][0];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) F[i][j] = H[i][j] + (i == j ? 1. : 0.);
      } else {
        eps[0] = H[0][0]; eps[1] = H[1][1]; eps[2] = H[2][2];
        eps[3] = H[0][1] + H[1][0]; eps[4] = H[1][2] + H[2][1]; eps[5] = H[0][2] + H[2][0];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) F[i][j] = (i == j ? 1. : 0.);
      }
      const double lam = E * p.lam1, mu = E * p.mu1, tr = eps[0] + eps[1] + eps[2];
      double S[6];
      S[0] = lam * tr + 2. * mu * eps[0]; S[1] = lam * tr + 2. * mu * eps[1]; S[2] = lam * tr + 2. * mu * eps[2];
      S[3] = mu * eps[3]; S[4] = mu * eps[4]; S[5] = mu * eps[5];
      double sth = 0.;
      if (p.thermal) {
        if (!p.temperature) T = p.T_const;
        sth = E * p.alpha_k * (T - p.T_ref);  // D alpha dT (isotropic: on the normal components)
      }
      double* o = stage + q * QD;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) { o[c * 3 + rr] = inv[c][rr]; o[9 + c * 3 + rr] = F[c][rr]; }
#pragma unroll
      for (int i = 0; i < 6; ++i) o[18 + i] = S[i];
      o[24] = det * E; o[25] = det; o[26] = sth;
      if (p.R) {  // f_a += JxW B_a^T (S - S_th)
        double B[6][3];
        make_B(F, ga, B);
        const double Se[6] = {S[0] - sth, S[1] - sth, S[2] - sth, S[3], S[4], S[5]};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double s = 0.;
#pragma unroll
          for (int rr = 0; rr < 6; ++rr) s = fma(B[rr][k], Se[rr], s);
          fq[k] = det * s;
        }
      }
    }
    if (p.R) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double v = fq[k];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        f[k] += __shfl_sync(0xffffffffu, v, 0);
      }
    }
    __syncwarp();
    if (p.J) {
      // lanes 0-7 also form the diagonal of K_e for the rotational term (nv = 6); lane c its own neighbour block
      double dmax = 0.;
      const bool do_blk = bme >= 0;
      const bool do_diag = p.nv == 6 && lane < 8;
      if (do_blk || do_diag) {
        double Kd[3] = {0., 0., 0.};
        for (int q = 0; q < 8; ++q) {
          const double* o = stage + q * QD;
          double F[3][3];
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) F[c][rr] = o[9 + c * 3 + rr];
          const double wE = o[24], w = o[25], sth = o[26];
          auto grad = [&](int b, double g[3]) {
            double Nn, d[3];
            ref_shape(b, q, Nn, d);
#pragma unroll
            for (int c = 0; c < 3; ++c) g[c] = d[0] * o[c * 3] + d[1] * o[c * 3 + 1] + d[2] * o[c * 3 + 2];
          };
          if (do_blk) {
            double g1[3], g2[3], Ba[6][3], Bb[6][3];
            grad(a, g1); grad(bme, g2);
            make_B(F, g1, Ba); make_B(F, g2, Bb);
            double geo = 0.;
            if (nl) {
              const double s0 = o[18] - sth, s1 = o[19] - sth, s2 = o[20] - sth, s3 = o[21], s4 = o[22], s5 = o[23];
              geo = w * (g1[0] * (s0 * g2[0] + s3 * g2[1] + s5 * g2[2]) + g1[1] * (s3 * g2[0] + s1 * g2[1] + s4 * g2[2]) +
                         g1[2] * (s5 * g2[0] + s4 * g2[1] + s2 * g2[2]));
            }
#pragma unroll
            for (int l = 0; l < 3; ++l) {
              const double t = Bb[0][l] + Bb[1][l] + Bb[2][l];
              const double d0 = p.lam1 * t + 2. * p.mu1 * Bb[0][l], d1 = p.lam1 * t + 2. * p.mu1 * Bb[1][l],
                           d2 = p.lam1 * t + 2. * p.mu1 * Bb[2][l];
              const double d3 = p.mu1 * Bb[3][l], d4 = p.mu1 * Bb[4][l], d5 = p.mu1 * Bb[5][l];
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                const double s = Ba[0][k] * d0 + Ba[1][k] * d1 + Ba[2][k] * d2 + Ba[3][k] * d3 + Ba[4][k] * d4 + Ba[5][k] * d5;
                K[k][l] = fma(wE, s, K[k][l]);
              }
            }
            K[0][0] += geo; K[1][1] += geo; K[2][2] += geo;
          }
          if (do_diag) {  // diagonal entries of K_e (internal_residual only: the thermal term is added afterwards)
            double g[3], B[6][3];
            grad(lane, g);
            make_B(F, g, B);
            double geo = 0.;
            if (nl) {
              const double s0 = o[18], s1 = o[19], s2 = o[20], s3 = o[21], s4 = o[22], s5 = o[23];
              geo = w * (g[0] * (s0 * g[0] + s3 * g[1] + s5 * g[2]) + g[1] * (s3 * g[0] + s1 * g[1] + s4 * g[2]) +
                         g[2] * (s5 * g[0] + s4 * g[1] + s2 * g[2]));
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const double t = B[0][k] + B[1][k] + B[2][k];
              const double s = B[0][k] * (p.lam1 * t + 2. * p.mu1 * B[0][k]) + B[1][k] * (p.lam1 * t + 2. * p.mu1 * B[1][k]) +
                               B[2][k] * (p.lam1 * t + 2. * p.mu1 * B[2][k]) +
                               p.mu1 * (B[3][k] * B[3][k] + B[4][k] * B[4][k] + B[5][k] * B[5][k]);
              Kd[k] = fma(wE, s, Kd[k]);
              Kd[k] += geo;
            }
          }
        }
        if (do_diag) dmax = fmax(Kd[0], fmax(Kd[1], Kd[2]));
      }
      if (p.nv == 6) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
        rot += 1.0e-20 * dmax;  // solid_element_3d.cpp:413-417
      }
    }
  }

  // ---- write the node's rows: libMesh constrain_element_matrix_and_vector semantics per adjacent element ----
  const double nadj_all = double(nadj + nadjd);
  if (p.J && lane < nnb) {
    const bool self = my_nbr == n;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double k = K[i][j];
        if (((cmn >> i) | (cmm >> j)) & 1u) k = 0.;
        if (self && i == j) {
          if ((cmn >> i) & 1u) k = nadj_all;
          else k += 1.0e-24 * nadjd;
        }
        K[i][j] = k;
      }
    if (p.nv == 3) {
      double* dst = p.J + 9 * nb0 + 3 * lane;
      const int rs = 3 * nnb;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) dst[i * rs + j] = K[i][j];
    } else {
      double* dst = p.J + 36 * nb0 + 6 * lane;
      const int rl = 6 * nnb;
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          double val = 0.;
          if (i < 3 && j < 3) val = K[i][j];
          else if (i == j && self) val = ((cmn >> i) & 1u) ? nadj_all : rot + 1.0e-24 * nadjd;
          dst[i * rl + j] = val;
        }
    }
  }
  if (p.R && lane < p.nv) {
    double v = 0.;
    if (lane < 3) v = ((cmn >> lane) & 1u) ? 0. : (lane == 0 ? f[0] : lane == 1 ? f[1] : f[2]);
    p.R[(int64_t)n * p.nv + lane] = v;
  }
}

}  // namespace

cudaError_t launch_hex8_gather(const SolidGatherParams& p, cudaStream_t stream) {
  const int blocks = (int)((p.n_owned + GW - 1) / GW);
  hex8_gather_kernel<<<blocks, GW * 32, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace mastb200
