// sweep_hex27_gather.cu -- HEX27 residual + Jacobian with Green-Lagrange strain as a row-owner gather (sm_100a, fp64).
//
// Device counterpart of the element loop of MAST::NonlinearImplicitAssembly::residual_and_jacobian
// (/root/reference/src/base/nonlinear_implicit_assembly.cpp:111-186) with StructuralElement3D::internal_residual for
// strain_type() == NONLINEAR_STRAIN on the tri-quadratic element (src/elasticity/solid_element_3d.cpp:135-423: material
// tangent :327-384, geometric stiffness :386-409, rotational diagonal :413-417; 3x3x3 Gauss rule, src/mesh/fe_base.cpp:
// 103-105), the strain operator of :1308-1463 and ::thermal_residual (:950-1089 with the geometric term of :1054-1082).
//
// Same scheme as sweep_hex8_gather.cu, sized for 27 nodes / 27 quadrature points: one warp per owned node (= nv CSR rows)
// walks the adjacent elements in ascending id (the order of the reference's element loop); per element lanes 0-26
// evaluate one quadrature point each (J^-1, F = I + grad u, second Piola-Kirchhoff stress, weights -> shared memory) and
// the node's own force row; then every lane adds the element's contribution to the 3x3 blocks of the row's neighbours
// lane, lane + 32, lane + 64, lane + 96 (a vertex node of a HEX27 mesh has 125) that are nodes of the element:
//   K_ab += sum_q JxW E (B_a^T D_1 B_b) + JxW (grad N_a . S grad N_b) I,   B_a from F as in sweep_hex8_gather.cu.
// No atomics, no colours: every CSR value is accumulated in one register in a fixed order and stored once.  The linear
// production path of HEX27 is the DMMA sweep of sweep_hex27.cu; this kernel serves the nonlinear strain measure.
#include <cuda_runtime.h>

#include "kernels.h"

namespace mastb200 {
namespace {

constexpr int GW = 4;            // warps (nodes) per CTA
constexpr int NN = 27;
constexpr int QD = 27;           // doubles staged per quadrature point: inv[9], F[9], S[6], JxW E, JxW, sth
constexpr int KB = 4;            // neighbour blocks per lane (<= 128 neighbours per row)

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

__global__ void __launch_bounds__(GW * 32) hex27_gather_kernel(const SolidGatherParams p) {
  __shared__ double stage_all[GW][NN * QD];
  __shared__ int32_t conn_all[GW][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * GW + warp;
  if (r >= p.n_owned) return;
  double* stage = stage_all[warp];
  int32_t* m = conn_all[warp];
  const double* __restrict__ tabN = p.tab;               // [q][b]
  const double* __restrict__ tabD = p.tab + NN * NN;     // [q][b][r]
  const double* __restrict__ tabW = p.tab + NN * NN * 4; // [q]
  const int n = p.owned_nodes[r];
  const int64_t nb0 = p.nbr_ptr[r];
  const int nnb = (int)(p.nbr_ptr[r + 1] - nb0);
  int my_nbr[KB];
  unsigned cmm[KB];
#pragma unroll
  for (int k = 0; k < KB; ++k) {
    const int c = lane + 32 * k;
    my_nbr[k] = c < nnb ? p.nbr[nb0 + c] : -1;
    cmm[k] = (p.cmask && my_nbr[k] >= 0) ? p.cmask[my_nbr[k]] : 0u;
  }
  const unsigned cmn = p.cmask ? p.cmask[n] : 0u;
  const bool nl = p.nonlinear != 0;

  double K[KB][3][3];
#pragma unroll
  for (int k = 0; k < KB; ++k)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) K[k][i][j] = 0.;
  double f[3] = {0., 0., 0.};
  double rot = 0.;
  int nadj = 0, nadjd = 0;

  for (int64_t ke = p.n2e_ptr[n]; ke < p.n2e_ptr[n + 1]; ++ke) {
    const int e = p.n2e[ke];
    if (p.diag_elem && p.diag_elem[e]) { ++nadjd; continue; }  // nonlinear_implicit_assembly.cpp:117-130
    ++nadj;
    __syncwarp();
    if (lane < NN) m[lane] = p.conn[(int64_t)e * NN + lane];
    __syncwarp();
    int a = 0, bme[KB];
#pragma unroll
    for (int k = 0; k < KB; ++k) bme[k] = -1;
    for (int b = 0; b < NN; ++b) {
      const int mb = m[b];
      if (mb == n) a = b;
#pragma unroll
      for (int k = 0; k < KB; ++k)
        if (mb == my_nbr[k]) bme[k] = b;
    }
    double fq[3] = {0., 0., 0.};
    if (lane < NN) {
      const int q = lane;
      double Jm[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};
      for (int b = 0; b < NN; ++b) {
        const double* d = tabD + (q * NN + b) * 3;
        const double* X = p.xyz + (int64_t)m[b] * 3;
        const double x0 = X[0], x1 = X[1], x2 = X[2];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
          Jm[rr][0] = fma(d[rr], x0, Jm[rr][0]); Jm[rr][1] = fma(d[rr], x1, Jm[rr][1]); Jm[rr][2] = fma(d[rr], x2, Jm[rr][2]);
        }
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
      for (int b = 0; b < NN; ++b) {
        const double* d = tabD + (q * NN + b) * 3;
        double g[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) g[c] = d[0] * inv[c][0] + d[1] * inv[c][1] + d[2] * inv[c][2];
        const double* U = p.X + (int64_t)m[b] * p.nv;
        const double u0 = U[0], u1 = U[1], u2 = U[2];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          H[0][j] = fma(u0, g[j], H[0][j]); H[1][j] = fma(u1, g[j], H[1][j]); H[2][j] = fma(u2, g[j], H[2][j]);
        }
        const double Nb = tabN[q * NN + b];
        if (p.rho) rho = fma(Nb, p.rho[m[b]], rho);
        if (p.temperature) T = fma(Nb, p.temperature[m[b]], T);
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
      const double JxW = det * tabW[q];
      double* o = stage + q * QD;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) { o[c * 3 + rr] = inv[c][rr]; o[9 + c * 3 + rr] = F[c][rr]; }
#pragma unroll
      for (int i = 0; i < 6; ++i) o[18 + i] = S[i];
      o[24] = JxW * E; o[25] = JxW; o[26] = sth;
      if (p.R) {  // f_a += JxW B_a^T (S - S_th)
        double B[6][3];
        make_B(F, ga, B);
        const double Se[6] = {S[0] - sth, S[1] - sth, S[2] - sth, S[3], S[4], S[5]};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double s = 0.;
#pragma unroll
          for (int rr = 0; rr < 6; ++rr) s = fma(B[rr][k], Se[rr], s);
          fq[k] = JxW * s;
        }
      }
    }
    if (p.R) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double v = fq[k];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        f[k] += v;
      }
    }
    __syncwarp();
    if (p.J) {
      // lanes 0-26 also form the diagonal of K_e for the rotational term (nv = 6); every lane its own neighbour blocks
      double dmax = 0.;
      const bool do_diag = p.nv == 6 && lane < NN;
      bool any_blk = false;
#pragma unroll
      for (int k = 0; k < KB; ++k) any_blk = any_blk || bme[k] >= 0;
      if (any_blk || do_diag) {
        double Kd[3] = {0., 0., 0.};
        for (int q = 0; q < NN; ++q) {
          const double* o = stage + q * QD;
          double F[3][3];
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) F[c][rr] = o[9 + c * 3 + rr];
          const double wE = o[24], w = o[25], sth = o[26];
          auto grad = [&](int b, double g[3]) {
            const double* d = tabD + (q * NN + b) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) g[c] = d[0] * o[c * 3] + d[1] * o[c * 3 + 1] + d[2] * o[c * 3 + 2];
          };
          if (any_blk) {
            double g1[3], Ba[6][3];
            grad(a, g1);
            make_B(F, g1, Ba);
#pragma unroll
            for (int k = 0; k < KB; ++k) {
              if (bme[k] < 0) continue;
              double g2[3], Bb[6][3];
              grad(bme[k], g2);
              make_B(F, g2, Bb);
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
                for (int kk = 0; kk < 3; ++kk) {
                  const double s = Ba[0][kk] * d0 + Ba[1][kk] * d1 + Ba[2][kk] * d2 + Ba[3][kk] * d3 + Ba[4][kk] * d4 + Ba[5][kk] * d5;
                  K[k][kk][l] = fma(wE, s, K[k][kk][l]);
                }
              }
              K[k][0][0] += geo; K[k][1][1] += geo; K[k][2][2] += geo;
            }
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
  if (p.J) {
#pragma unroll
    for (int k = 0; k < KB; ++k) {
      const int c = lane + 32 * k;
      if (c >= nnb) continue;
      const bool self = my_nbr[k] == n;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          double v = K[k][i][j];
          if (((cmn >> i) | (cmm[k] >> j)) & 1u) v = 0.;
          if (self && i == j) {
            if ((cmn >> i) & 1u) v = nadj_all;
            else v += 1.0e-24 * nadjd;
          }
          K[k][i][j] = v;
        }
      if (p.nv == 3) {
        double* dst = p.J + 9 * nb0 + 3 * c;
        const int rs = 3 * nnb;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) dst[i * rs + j] = K[k][i][j];
      } else {
        double* dst = p.J + 36 * nb0 + 6 * c;
        const int rl = 6 * nnb;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            double val = 0.;
            if (i < 3 && j < 3) val = K[k][i][j];
            else if (i == j && self) val = ((cmn >> i) & 1u) ? nadj_all : rot + 1.0e-24 * nadjd;
            dst[i * rl + j] = val;
          }
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

cudaError_t launch_hex27_gather(const SolidGatherParams& p, cudaStream_t stream) {
  const int blocks = (int)((p.n_owned + GW - 1) / GW);
  hex27_gather_kernel<<<blocks, GW * 32, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace mastb200
