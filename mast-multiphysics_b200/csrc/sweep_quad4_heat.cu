// sweep_quad4_heat.cu -- QUAD4 heat-conduction residual + Jacobian (sm_100a, fp64).
//
// Device counterpart of the element loop of MAST::NonlinearImplicitAssembly::residual_and_jacobian
// (/root/reference/src/base/nonlinear_implicit_assembly.cpp:111-186) with
// HeatConductionNonlinearAssemblyElemOperations::elem_calculations (src/heat_conduction/heat_conduction_nonlinear_assembly.cpp:83-101),
// HeatConductionElementBase::internal_residual (src/heat_conduction/heat_conduction_elem_base.cpp:53-149, including the
// dk/dT tangent of :125-141 when a solution function is attached), the section conductance k*h
// (src/property_cards/solid_2d_section_element_property_card.cpp:940-953) and volume_heat_source_residual (:1709-1745).
//
// Row-owner gather: one thread per owned node (= CSR row, one temperature dof per node).  The thread walks the
// elements adjacent to its node in ascending element id -- the order in which the reference's element loop adds into
// that row -- recomputes each element's 2x2 Gauss quadrature (points in libMesh order, xi fastest) and keeps only its
// own row: f_a and K_ab, b = 0..3.  Redundant flops (every element is evaluated by its four nodes) buy a kernel without
// atomics or colours whose every CSR value is summed in the reference order and stored exactly once.  The row
// accumulators live in shared memory (acc[k][thread], conflict-free) because the column position is a run-time index.
//
// Algorithmic bytes per element on a structured mesh: 9 values of J (72 B) + R (8) + T (8) + xy (16) + connectivity
// (16) = 120 B: HBM-bound; C2 (1024^2) is 126 MB in total, i.e. it lives in L2 and the sweep takes tens of microseconds.
#include <cuda_runtime.h>

#include "kernels.h"

namespace mastb200 {
namespace {

constexpr int TB = 128;
constexpr double GP = 0.57735026918962576451;

__global__ void __launch_bounds__(TB) quad4_heat_kernel(const GatherParams p) {
  __shared__ double acc[kGatherMaxNbr * TB];
  const int64_t r = (int64_t)blockIdx.x * TB + threadIdx.x;
  if (r >= p.n_owned) return;
  const int n = p.owned_nodes[r];
  const int64_t nb = p.nbr_ptr[r];
  const int nnb = (int)(p.nbr_ptr[r + 1] - nb);
  const int32_t* nbr = p.nbr + nb;
  double* my = acc + threadIdx.x;
  for (int k = 0; k < nnb; ++k) my[k * TB] = 0.;
  const bool con_row = p.cmask && (p.cmask[n] & 1u);
  int self = 0;
  while (self < nnb && nbr[self] != n) ++self;
  double res = 0.;

  for (int64_t ke = p.n2e_ptr[n]; ke < p.n2e_ptr[n + 1]; ++ke) {
    const int e = p.n2e[ke];
    if (p.diag_elem && p.diag_elem[e]) {  // diagonal_elem_subdomain_id (nonlinear_implicit_assembly.cpp:117-130)
      if (p.J) my[self * TB] += con_row ? 1. : 1.0e-24;
      continue;
    }
    int m[4];
    double x[4], y[4], T[4];
    int a = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      m[b] = p.conn[(int64_t)e * 4 + b];
      x[b] = p.xyz[(int64_t)m[b] * 3];
      y[b] = p.xyz[(int64_t)m[b] * 3 + 1];
      T[b] = p.X[m[b]];
      if (m[b] == n) a = b;
    }
    double f = 0., K[4] = {0., 0., 0., 0.};
    double jxw[4], Na[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double xi = (q & 1) ? GP : -GP, eta = (q & 2) ? GP : -GP;
      // libMesh QUAD4 node order (-,-), (+,-), (+,+), (-,+)
      const double sx[4] = {-1., 1., 1., -1.}, sy[4] = {-1., -1., 1., 1.};
      double N[4], d0[4], d1[4];
      double j00 = 0., j01 = 0., j10 = 0., j11 = 0.;  // j[r][c] = d x_c / d xi_r
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const double fx = 0.5 * (1. + sx[b] * xi), fy = 0.5 * (1. + sy[b] * eta);
        N[b] = fx * fy;
        d0[b] = 0.5 * sx[b] * fy;
        d1[b] = fx * 0.5 * sy[b];
        j00 += d0[b] * x[b]; j01 += d0[b] * y[b];
        j10 += d1[b] * x[b]; j11 += d1[b] * y[b];
      }
      const double det = j00 * j11 - j10 * j01;
      const double inv = 1. / det;
      const double JxW = det;  // Gauss weights are 1
      const double dxidx = j11 * inv, dxidy = -j10 * inv, detadx = -j01 * inv, detady = j00 * inv;
      double gx[4], gy[4];
      double Tq = 0., dTx = 0., dTy = 0.;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        gx[b] = d0[b] * dxidx + d1[b] * detadx;
        gy[b] = d0[b] * dxidy + d1[b] * detady;
        Tq += N[b] * T[b];
        dTx += gx[b] * T[b];
        dTy += gy[b] * T[b];
      }
      const double k = (p.k_of_T ? p.k0 * (1. + p.k_beta * Tq) : p.k0) * p.thickness;
      const double dk = p.k_of_T ? p.k0 * p.k_beta * p.thickness : 0.;
      if (p.R) {
        f += JxW * (gx[a] * (dTx * k));
        f += JxW * (gy[a] * (dTy * k));
      }
      if (p.J) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          K[b] += JxW * k * (gx[a] * gx[b]);
          K[b] += JxW * k * (gy[a] * gy[b]);
        }
        if (p.k_of_T) {
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            K[b] += JxW * dTx * dk * (gx[a] * N[b]);
            K[b] += JxW * dTy * dk * (gy[a] * N[b]);
          }
        }
      }
      jxw[q] = JxW; Na[q] = N[a];
    }
    if (p.sl_ptr) {
      // HeatConductionElementBase::surface_flux_residual / surface_convection_residual / surface_radiation_residual
      // (/root/reference/src/heat_conduction/heat_conduction_elem_base.cpp:876-914, 1100-1155, 1404-1458) on the loaded sides of
      // this element that hold the row node: EDGE2 with the 2-point Gauss rule, f += JxW N_a th (flux + h_c (T - T_amb)),
      // K_ab += JxW N_a N_b th h_c
      for (int k = p.sl_ptr[e]; k < p.sl_ptr[e + 1]; ++k) {
        const int sd = p.sl_side[k], n1 = sd, n2 = (sd + 1) & 3;
        if (a != n1 && a != n2) continue;
        const double flux = p.sl_val[k * 4], hc = p.sl_val[k * 4 + 1], Tamb = p.sl_val[k * 4 + 2], eps = p.sl_val[k * 4 + 3];
        const double ex = x[n2] - x[n1], ey = y[n2] - y[n1];
        const double JxWs = 0.5 * sqrt(ex * ex + ey * ey);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const double xi = q ? GP : -GP;
          const double N1 = 0.5 * (1. - xi), N2 = 0.5 * (1. + xi);
          const double Nrow = a == n1 ? N1 : N2;
          const double temp = N1 * T[n1] + N2 * T[n2];
          if (p.R) {
            f += JxWs * Nrow * flux * p.thickness;
            if (hc != 0.) f += JxWs * Nrow * p.thickness * hc * (temp - Tamb);
          }
          if (p.J && hc != 0.) {
            K[n1] += JxWs * (Nrow * N1) * p.thickness * hc;
            K[n2] += JxWs * (Nrow * N2) * p.thickness * hc;
          }
          if (eps != 0.) {  // ::surface_radiation_residual (:1404-1458)
            const double tz = temp + p.rad_T0, az = Tamb + p.rad_T0;
            if (p.R) f += JxWs * Nrow * p.rad_sbc * eps * p.thickness * (tz * tz * tz * tz - az * az * az * az);
            if (p.J) {
              const double d = p.rad_sbc * eps * p.thickness * 4. * (tz * tz * tz);
              K[n1] += JxWs * (Nrow * N1) * d;
              K[n2] += JxWs * (Nrow * N2) * d;
            }
          }
        }
      }
    }
    if (p.R && p.heat_source != 0.) {
#pragma unroll
      for (int q = 0; q < 4; ++q) f -= jxw[q] * Na[q] * p.heat_source * p.thickness;
    }
    // libMesh constrain_element_matrix_and_vector: constrained row/column zeroed, diagonal 1, rhs row 0 -- per element
    if (p.R) res += con_row ? 0. : f;
    if (p.J) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        double v = K[b];
        const bool con_col = p.cmask && (p.cmask[m[b]] & 1u);
        if (con_row || con_col) v = (b == a && con_row) ? 1. : 0.;
        int pos = 0;
        while (pos < nnb - 1 && nbr[pos] != m[b]) ++pos;
        my[pos * TB] += v;
      }
    }
  }
  if (p.J)
    for (int k = 0; k < nnb; ++k) p.J[nb + k] = my[k * TB];
  if (p.R) p.R[n] = res;
}

}  // namespace

cudaError_t launch_quad4_heat(const GatherParams& p, cudaStream_t stream) {
  const int blocks = (int)((p.n_owned + TB - 1) / TB);
  quad4_heat_kernel<<<blocks, TB, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace mastb200
