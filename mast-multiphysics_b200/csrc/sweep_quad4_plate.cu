// sweep_quad4_plate.cu -- QUAD4 Mindlin plate residual + Jacobian, linear or von Karman strain, elements anywhere in space
// (sm_100a, fp64).
//
// Device counterpart of the element loop of MAST::NonlinearImplicitAssembly::residual_and_jacobian
// (/root/reference/src/base/nonlinear_implicit_assembly.cpp:111-186) with StructuralElement2D::internal_residual
// (src/elasticity/structural_element_2d.cpp:1075-1227) and ::_internal_residual_operation (:1765-2019): membrane
// B_lin^T A B_lin, coupling B_lin^T B B_bend + B_bend^T B^T B_lin, bending B_bend^T D B_bend with the Mindlin bending
// strain operator (src/elasticity/mindlin_bending_operator.cpp:47-80), the transverse shear energy integrated with
// the rule two orders lower -- one Gauss point for QUAD4 (:94-160, 314-361), the section matrices A = C h, B = C h off,
// D = C (h^3/12 + h off^2), G kappa h (src/property_cards/solid_2d_section_element_property_card.cpp:395-510, 122-134;
// C of isotropic_material_property_card.cpp:395-426) and the 1e-8 drilling diagonal (:1217-1221).
// Elements whose nodes do not share one z are evaluated in the local frame of their first three nodes (GeomElem,
// src/mesh/geom_elem.cpp:358-374, 461-511): solution rotated in (structural_element_base.cpp:1606-1640), f = T f_l and
// K = T K_l T^T rotated out (:1565-1601, 1642-1674), per element, before the block joins the row (LOCAL variant; the
// host picks it when the mesh has such an element).
// NONLINEAR_STRAIN (NL variant): Green-Lagrange membrane strain (::initialize_green_lagrange_strain_operator, :141-250:
// B_m = B_lin + mat_x B_x + mat_y B_y, i.e. the columns [(1+u,x) N,x; u,y N,y; (1+u,x) N,y + u,y N,x] and
// [v,x N,x; (1+v,y) N,y; v,x N,y + (1+v,y) N,x]) plus the von Karman strain 1/2 w,x^2, 1/2 w,y^2, w,x w,y
// (::initialize_von_karman_strain_operator, :57-100), with the tangent terms exactly as the reference has them
// (:1886-1998): B_m^T A B_m and the initial-stress stiffness of u, v and w from N = A eps + B kappa, but B_lin -- not
// B_m -- in the membrane / von Karman / bending couplings.
//
// Row-owner gather: one thread per (owned node, neighbour) pair, i.e. per 6x6 block of the node-level CSR.  The thread
// walks the elements adjacent to the row node in ascending id (the order of the reference's element loop) and adds
// K_ab of every element that also holds its neighbour; the diagonal pair also forms the node's force.  No atomics, no
// colours; every CSR value is accumulated in registers in a fixed order and stored once.
// Algorithmic bytes per element on a structured mesh: 9 blocks x 36 values x 8 B = 2592 B of J + vectors: HBM-bound.
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"

namespace mastb200 {
namespace {

constexpr int TB = 128;
constexpr double GP = 0.57735026918962576451;

struct Q4 {
  double N[4], gx[4], gy[4], JxW;
};

__device__ __forceinline__ void q4_eval(const double* x, const double* y, double xi, double eta, double w, Q4& o) {
  const double sx[4] = {-1., 1., 1., -1.}, sy[4] = {-1., -1., 1., 1.};
  double d0[4], d1[4];
  double j00 = 0., j01 = 0., j10 = 0., j11 = 0.;
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const double fx = 0.5 * (1. + sx[b] * xi), fy = 0.5 * (1. + sy[b] * eta);
    o.N[b] = fx * fy;
    d0[b] = 0.5 * sx[b] * fy;
    d1[b] = fx * 0.5 * sy[b];
    j00 += d0[b] * x[b]; j01 += d0[b] * y[b];
    j10 += d1[b] * x[b]; j11 += d1[b] * y[b];
  }
  const double det = j00 * j11 - j10 * j01;
  const double inv = 1. / det;
  o.JxW = det * w;
  const double dxidx = j11 * inv, dxidy = -j10 * inv, detadx = -j01 * inv, detady = j00 * inv;
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    o.gx[b] = d0[b] * dxidx + d1[b] * detadx;
    o.gy[b] = d0[b] * dxidy + d1[b] * detady;
  }
}

#ifndef MASTB200_PLATE_CTAS
#define MASTB200_PLATE_CTAS 3  // 168 registers: 12 warps per SM (2: 191 registers, 320 M elem/s; 3: 380 M; 4: spills, 309 M at 1024^2)
#endif
template <bool LOCAL, bool NL>
__global__ void __launch_bounds__(TB, MASTB200_PLATE_CTAS) quad4_plate_kernel(const PlateParams p) {
  const int64_t t = (int64_t)blockIdx.x * TB + threadIdx.x;
  if (t >= p.n_entries) return;
  const int64_t r = p.ent_row[t];  // row of this neighbour entry (precomputed with the plan)
  const int n = p.owned_nodes[r];
  const int64_t nb0 = p.nbr_ptr[r];
  const int nnb = (int)(p.nbr_ptr[r + 1] - nb0);
  const int c = (int)(t - nb0);
  const int m = p.nbr[t];
  const bool self = m == n;
  const unsigned cmn = p.cmask ? p.cmask[n] : 0u, cmm = p.cmask ? p.cmask[m] : 0u;

  const double c11 = p.c11, c12 = p.c12, c33 = p.c33;
  const double h = p.thickness, hoff = p.thickness * p.offset;
  const double hD = p.thickness * p.thickness * p.thickness / 12. + p.thickness * p.offset * p.offset;
  const double Gs = p.shear_G * p.thickness * p.kappa;

  double K[6][6];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) K[i][j] = 0.;
  double f[6] = {0., 0., 0., 0., 0., 0.};
  int nadj = 0, nadjd = 0;

  for (int64_t ke = p.n2e_ptr[n]; ke < p.n2e_ptr[n + 1]; ++ke) {
    const int e = p.n2e[ke];
    if (p.diag_elem && p.diag_elem[e]) { ++nadjd; continue; }
    ++nadj;
    int en[4], a = 0, b = -1;
    double x[4], y[4], z[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      en[k] = p.conn[(int64_t)e * 4 + k];
      x[k] = p.xyz[(int64_t)en[k] * 3];
      y[k] = p.xyz[(int64_t)en[k] * 3 + 1];
      if (LOCAL) z[k] = p.xyz[(int64_t)en[k] * 3 + 2];
      if (en[k] == n) a = k;
      if (en[k] == m) b = k;
    }
    if (b < 0) continue;
    // local frame of an element out of the xy-plane: T columns = v1 (node 0 -> 1), v2 = v3 x v1, v3 = v1 x (node 0 -> 2)
    bool use_local = false;
    double T[3][3] = {{1., 0., 0.}, {0., 1., 0.}, {0., 0., 1.}};
    if (LOCAL) {
      use_local = z[1] != z[0] || z[2] != z[0] || z[3] != z[0];
      if (use_local) {
        double v1[3] = {x[1] - x[0], y[1] - y[0], z[1] - z[0]}, v2[3] = {x[2] - x[0], y[2] - y[0], z[2] - z[0]}, v3[3];
        double nr = sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]);
        v1[0] /= nr; v1[1] /= nr; v1[2] /= nr;
        nr = sqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
        v2[0] /= nr; v2[1] /= nr; v2[2] /= nr;
        v3[0] = v1[1] * v2[2] - v1[2] * v2[1]; v3[1] = v1[2] * v2[0] - v1[0] * v2[2]; v3[2] = v1[0] * v2[1] - v1[1] * v2[0];
        nr = sqrt(v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2]);
        v3[0] /= nr; v3[1] /= nr; v3[2] /= nr;
        v2[0] = v3[1] * v1[2] - v3[2] * v1[1]; v2[1] = v3[2] * v1[0] - v3[0] * v1[2]; v2[2] = v3[0] * v1[1] - v3[1] * v1[0];
        nr = sqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
        v2[0] /= nr; v2[1] /= nr; v2[2] /= nr;
#pragma unroll
        for (int i = 0; i < 3; ++i) { T[i][0] = v1[i]; T[i][1] = v2[i]; T[i][2] = v3[i]; }
        const double x0 = x[0], y0 = y[0], z0 = z[0];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double dx = x[k] - x0, dy = y[k] - y0, dz = z[k] - z0;
          x[k] = T[0][0] * dx + T[1][0] * dy + T[2][0] * dz;
          y[k] = T[0][1] * dx + T[1][1] * dy + T[2][1] * dz;
        }
      }
    }
    // LOCAL: the element's block and force are formed in Ke / fe and rotated into K / f afterwards
    double Ke[6][6], fe[6];
    double (&KK)[6][6] = LOCAL ? Ke : K;
    double (&ff)[6] = LOCAL ? fe : f;
    if (LOCAL) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        fe[i] = 0.;
#pragma unroll
        for (int j = 0; j < 6; ++j) Ke[i][j] = 0.;
      }
    }
    const bool want_f = self && p.R;
    const bool want_u = want_f || NL;  // the nonlinear tangent depends on the state
    double U[4][6];
    if (want_u)
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int v = 0; v < 6; ++v) U[k][v] = p.X[(int64_t)en[k] * 6 + v];
    if (LOCAL && want_u && use_local)
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int h3 = 0; h3 < 6; h3 += 3) {
          const double g0 = U[k][h3], g1 = U[k][h3 + 1], g2 = U[k][h3 + 2];
#pragma unroll
          for (int j = 0; j < 3; ++j) U[k][h3 + j] = T[0][j] * g0 + T[1][j] * g1 + T[2][j] * g2;
        }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      Q4 g;
      q4_eval(x, y, (q & 1) ? GP : -GP, (q & 2) ? GP : -GP, 1., g);
      const double gxa = g.gx[a], gya = g.gy[a], gxb = g.gx[b], gyb = g.gy[b];
      // NL: displacement gradients, total membrane strain (Green-Lagrange + von Karman), curvatures, stress resultants
      double ux = 0., uy = 0., vx = 0., vy = 0., wx = 0., wy = 0., N0 = 0., N1 = 0., N2 = 0., Q0 = 0., Q1 = 0., Q2 = 0.;
      if (NL) {
        double k0 = 0., k1 = 0., k2 = 0.;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          ux += g.gx[k] * U[k][0]; uy += g.gy[k] * U[k][0];
          vx += g.gx[k] * U[k][1]; vy += g.gy[k] * U[k][1];
          wx += g.gx[k] * U[k][2]; wy += g.gy[k] * U[k][2];
          k0 += g.gx[k] * U[k][4]; k1 += -g.gy[k] * U[k][3]; k2 += -g.gx[k] * U[k][3] + g.gy[k] * U[k][4];
        }
        const double e0 = ux + 0.5 * (ux * ux + vx * vx) + 0.5 * wx * wx, e1 = vy + 0.5 * (uy * uy + vy * vy) + 0.5 * wy * wy;
        const double e2 = (uy + vx) + (ux * uy + vx * vy) + wx * wy;
        const double Ce0 = c11 * e0 + c12 * e1, Ce1 = c12 * e0 + c11 * e1, Ce2 = c33 * e2;
        const double Ck0 = c11 * k0 + c12 * k1, Ck1 = c12 * k0 + c11 * k1, Ck2 = c33 * k2;
        N0 = h * Ce0 + hoff * Ck0; N1 = h * Ce1 + hoff * Ck1; N2 = h * Ce2 + hoff * Ck2;
        Q0 = hoff * Ce0 + hD * Ck0; Q1 = hoff * Ce1 + hD * Ck1; Q2 = hoff * Ce2 + hD * Ck2;
      }
      auto cdot = [&](double x0, double x1, double x2, double y0, double y1, double y2) {
        return x0 * (c11 * y0 + c12 * y1) + x1 * (c12 * y0 + c11 * y1) + x2 * (c33 * y2);
      };
      if (p.J) {
        // M[p][q] = s_p(a)^T C s_q(b),  s_1 = [gx, 0, gy], s_2 = [0, gy, gx]
        const double M00 = gxa * c11 * gxb + gya * c33 * gyb, M01 = gxa * c12 * gyb + gya * c33 * gxb;
        const double M10 = gya * c12 * gxb + gxa * c33 * gyb, M11 = gya * c11 * gyb + gxa * c33 * gxb;
        const double wA = g.JxW * h, wB = g.JxW * hoff, wD = g.JxW * hD;
        if (!NL) {
          KK[0][0] += wA * M00; KK[0][1] += wA * M01; KK[1][0] += wA * M10; KK[1][1] += wA * M11;      // membrane
        } else {
          // columns of B_m for u and v at the two nodes, the von Karman column of w
          const double au0 = (1. + ux) * gxa, au1 = uy * gya, au2 = (1. + ux) * gya + uy * gxa;
          const double av0 = vx * gxa, av1 = (1. + vy) * gya, av2 = vx * gya + (1. + vy) * gxa;
          const double bu0 = (1. + ux) * gxb, bu1 = uy * gyb, bu2 = (1. + ux) * gyb + uy * gxb;
          const double bv0 = vx * gxb, bv1 = (1. + vy) * gyb, bv2 = vx * gyb + (1. + vy) * gxb;
          const double aw0 = wx * gxa, aw1 = wy * gya, aw2 = wy * gxa + wx * gya;
          const double bw0 = wx * gxb, bw1 = wy * gyb, bw2 = wy * gxb + wx * gyb;
          const double geo = g.JxW * (gxa * (N0 * gxb + N2 * gyb) + gya * (N2 * gxb + N1 * gyb));  // initial-stress stiffness
          KK[0][0] += wA * cdot(au0, au1, au2, bu0, bu1, bu2) + geo; KK[0][1] += wA * cdot(au0, au1, au2, bv0, bv1, bv2);
          KK[1][0] += wA * cdot(av0, av1, av2, bu0, bu1, bu2); KK[1][1] += wA * cdot(av0, av1, av2, bv0, bv1, bv2) + geo;
          // membrane - vk and vk - membrane with B_lin (s_1, s_2), vk - vk, bending - vk and vk - bending
          KK[0][2] += wA * cdot(gxa, 0., gya, bw0, bw1, bw2); KK[1][2] += wA * cdot(0., gya, gxa, bw0, bw1, bw2);
          KK[2][0] += wA * cdot(aw0, aw1, aw2, gxb, 0., gyb); KK[2][1] += wA * cdot(aw0, aw1, aw2, 0., gyb, gxb);
          KK[2][2] += geo + wA * cdot(aw0, aw1, aw2, bw0, bw1, bw2);
          KK[3][2] += wB * cdot(0., -gya, -gxa, bw0, bw1, bw2); KK[4][2] += wB * cdot(gxa, 0., gya, bw0, bw1, bw2);
          KK[2][3] += wB * cdot(aw0, aw1, aw2, 0., -gyb, -gxb); KK[2][4] += wB * cdot(aw0, aw1, aw2, gxb, 0., gyb);
        }
        KK[0][3] += wB * (-M01); KK[0][4] += wB * M00; KK[1][3] += wB * (-M11); KK[1][4] += wB * M10;  // membrane - bending
        KK[3][0] += wB * (-M10); KK[3][1] += wB * (-M11); KK[4][0] += wB * M00; KK[4][1] += wB * M01;  // bending - membrane
        KK[3][3] += wD * M11; KK[3][4] += wD * (-M10); KK[4][3] += wD * (-M01); KK[4][4] += wD * M00;  // bending
      }
      if (want_f && NL) {  // B_m^T N, the von Karman column of w, B_bend^T (B^T eps + D kappa) with the total membrane strain
        ff[0] += g.JxW * ((1. + ux) * gxa * N0 + uy * gya * N1 + ((1. + ux) * gya + uy * gxa) * N2);
        ff[1] += g.JxW * (vx * gxa * N0 + (1. + vy) * gya * N1 + (vx * gya + (1. + vy) * gxa) * N2);
        ff[2] += g.JxW * (wx * gxa * N0 + wy * gya * N1 + (wy * gxa + wx * gya) * N2);
        ff[3] += g.JxW * (-(gya * Q1 + gxa * Q2));
        ff[4] += g.JxW * (gxa * Q0 + gya * Q2);
      }
      if (want_f && !NL) {
        double e0 = 0., e1 = 0., e2 = 0., k0 = 0., k1 = 0., k2 = 0.;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          e0 += g.gx[k] * U[k][0]; e1 += g.gy[k] * U[k][1]; e2 += g.gy[k] * U[k][0] + g.gx[k] * U[k][1];
          k0 += g.gx[k] * U[k][4]; k1 += -g.gy[k] * U[k][3]; k2 += -g.gx[k] * U[k][3] + g.gy[k] * U[k][4];
        }
        const double Ce0 = c11 * e0 + c12 * e1, Ce1 = c12 * e0 + c11 * e1, Ce2 = c33 * e2;
        const double Ck0 = c11 * k0 + c12 * k1, Ck1 = c12 * k0 + c11 * k1, Ck2 = c33 * k2;
        const double N0 = h * Ce0 + hoff * Ck0, N1 = h * Ce1 + hoff * Ck1, N2 = h * Ce2 + hoff * Ck2;
        const double Q0 = hoff * Ce0 + hD * Ck0, Q1 = hoff * Ce1 + hD * Ck1, Q2 = hoff * Ce2 + hD * Ck2;
        ff[0] += g.JxW * (gxa * N0 + gya * N2);
        ff[1] += g.JxW * (gya * N1 + gxa * N2);
        ff[3] += g.JxW * (-(gya * Q1 + gxa * Q2));
        ff[4] += g.JxW * (gxa * Q0 + gya * Q2);
      }
    }
    {  // transverse shear at the single point of the reduced rule (weight 4)
      Q4 g;
      q4_eval(x, y, 0., 0., 4., g);
      const double w = g.JxW * Gs;
      const double gxa = g.gx[a], gya = g.gy[a], Na = g.N[a], gxb = g.gx[b], gyb = g.gy[b], Nb = g.N[b];
      if (p.J) {
        KK[2][2] += w * (gxa * gxb + gya * gyb); KK[2][3] += w * (-(gya * Nb)); KK[2][4] += w * (gxa * Nb);
        KK[3][2] += w * (-(Na * gyb)); KK[3][3] += w * (Na * Nb);
        KK[4][2] += w * (Na * gxb); KK[4][4] += w * (Na * Nb);
      }
      if (want_f) {
        double g0 = 0., g1 = 0.;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          g0 += g.gx[k] * U[k][2] + g.N[k] * U[k][4];
          g1 += g.gy[k] * U[k][2] - g.N[k] * U[k][3];
        }
        ff[2] += w * (gxa * g0 + gya * g1);
        ff[3] += w * (-(Na * g1));
        ff[4] += w * (Na * g0);
      }
    }
    if (self) KK[5][5] += 1.0e-8;  // drilling dof (structural_element_2d.cpp:1217-1221)
    if (LOCAL) {
      if (use_local) {  // K += T6 Ke T6^T, f += T6 fe with T6 = diag(T, T)
#pragma unroll
        for (int bi = 0; bi < 6; bi += 3)
#pragma unroll
          for (int bj = 0; bj < 6; bj += 3) {
            double W[3][3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j)
                W[i][j] = T[i][0] * Ke[bi][bj + j] + T[i][1] * Ke[bi + 1][bj + j] + T[i][2] * Ke[bi + 2][bj + j];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j) K[bi + i][bj + j] += W[i][0] * T[j][0] + W[i][1] * T[j][1] + W[i][2] * T[j][2];
          }
#pragma unroll
        for (int bi = 0; bi < 6; bi += 3)
#pragma unroll
          for (int i = 0; i < 3; ++i) f[bi + i] += T[i][0] * fe[bi] + T[i][1] * fe[bi + 1] + T[i][2] * fe[bi + 2];
      } else {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          f[i] += fe[i];
#pragma unroll
          for (int j = 0; j < 6; ++j) K[i][j] += Ke[i][j];
        }
      }
    }
  }

  if (p.J) {
    const double nadj_all = double(nadj + nadjd);
    double* dst = p.J + 36 * nb0 + 6 * c;
    const int rl = 6 * nnb;
    // a six-value row piece starts at a multiple of 48 bytes (36 nb0 + 6 c + i * 6 nnb doubles): 16-byte stores when the
    // caller's J is 16-byte aligned
    const bool al16 = (reinterpret_cast<uintptr_t>(p.J) & 15) == 0;
    const bool al32 = (reinterpret_cast<uintptr_t>(p.J) & 31) == 0;  // every other piece starts on a 32-byte boundary
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double v[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        double k = K[i][j];
        if (((cmn >> i) | (cmm >> j)) & 1u) k = 0.;
        if (self && i == j) {
          if ((cmn >> i) & 1u) k = nadj_all;
          else k += 1.0e-24 * nadjd;
        }
        v[j] = k;
      }
      if (al32) {  // one 32-byte and one 16-byte store, chosen without a branch
        double* d = dst + i * rl;
        const bool odd = ((c + i * nnb) & 1) != 0;
        double* p16 = odd ? d : d + 4;
        double* p32 = odd ? d + 2 : d;
        asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p16), "d"(odd ? v[0] : v[4]), "d"(odd ? v[1] : v[5]) : "memory");
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p32), "d"(odd ? v[2] : v[0]), "d"(odd ? v[3] : v[1]),
                     "d"(odd ? v[4] : v[2]), "d"(odd ? v[5] : v[3]) : "memory");
      } else if (al16) {
        double2* d2 = reinterpret_cast<double2*>(dst + i * rl);
        d2[0] = make_double2(v[0], v[1]); d2[1] = make_double2(v[2], v[3]); d2[2] = make_double2(v[4], v[5]);
      } else {
#pragma unroll
        for (int j = 0; j < 6; ++j) dst[i * rl + j] = v[j];
      }
    }
  }
  if (p.R && self)
#pragma unroll
    for (int i = 0; i < 6; ++i) p.R[(int64_t)n * 6 + i] = ((cmn >> i) & 1u) ? 0. : f[i];
}

}  // namespace

cudaError_t launch_quad4_plate(const PlateParams& p, cudaStream_t stream) {
  const int blocks = (int)((p.n_entries + TB - 1) / TB);
  if (p.nonlinear) {
    if (p.local_frames) quad4_plate_kernel<true, true><<<blocks, TB, 0, stream>>>(p);
    else quad4_plate_kernel<false, true><<<blocks, TB, 0, stream>>>(p);
  } else {
    if (p.local_frames) quad4_plate_kernel<true, false><<<blocks, TB, 0, stream>>>(p);
    else quad4_plate_kernel<false, false><<<blocks, TB, 0, stream>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace mastb200
