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

// ---- build_square-numbered meshes: implicit connectivity, one evaluation per element and CTA tile -----------------------
// CTA tile = 32 x 8 elements and the 31 x 7 nodes interior to them (plus the mesh-boundary nodes of edge tiles).
// Element phase: thread = element -> the full element vector / matrix (4 + 16 values: conduction, k(T) tangent, heat
// source, side loads -- the reference's elem_calculations) into shared memory, ke[value][element] (conflict-free).
// Node phase: thread = owned node = CSR row; its <= 4 adjacent elements in ascending id (the order of the reference's
// element loop) with libMesh's per-element Dirichlet semantics; the 9 column positions of a row are compile-time
// constants, so the row lives in registers and is stored as one contiguous run.  Nothing but nodal state is read:
// CSR row pointers have a closed form for a box.
constexpr int SEX = 32, SEY = 8, SNX = SEX - 1, SNY = SEY - 1, STB = SEX * SEY;
__device__ __forceinline__ double sel4(const double (&v)[4], int i) { return i == 0 ? v[0] : (i == 1 ? v[1] : (i == 2 ? v[2] : v[3])); }

__global__ void __launch_bounds__(STB) quad4_heat_struct_kernel(const GatherParams p, const int nx, const int ny, const int tiles_x) {
  __shared__ double ke[20 * STB];
  const int tid = threadIdx.x;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int o0x = tx * SNX, o0y = ty * SNY;  // first owned node; elements start one column / row before
  const int npx = nx + 1;
  {
    const int lex = tid & (SEX - 1), ley = tid / SEX;
    const int ex = o0x - 1 + lex, ey = o0y - 1 + ley;
    if (ex >= 0 && ex < nx && ey >= 0 && ey < ny) {
      const int e = ex + ey * nx;
      const int m[4] = {ex + ey * npx, ex + 1 + ey * npx, ex + 1 + (ey + 1) * npx, ex + (ey + 1) * npx};
      double x[4], y[4], T[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        x[b] = p.xyz[(int64_t)m[b] * 3];
        y[b] = p.xyz[(int64_t)m[b] * 3 + 1];
        T[b] = p.X[m[b]];
      }
      double f[4] = {0., 0., 0., 0.}, K[4][4], Kt[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { K[a][b] = 0.; Kt[a][b] = 0.; }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double xi = (q & 1) ? GP : -GP, eta = (q & 2) ? GP : -GP;
        const double sx[4] = {-1., 1., 1., -1.}, sy[4] = {-1., -1., 1., 1.};
        double N[4], d0[4], d1[4];
        double j00 = 0., j01 = 0., j10 = 0., j11 = 0.;
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
        const double JxW = det;
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
        // the element is fp64-pipe-bound: weights folded into the row gradients, so that a matrix entry costs two FMAs for
        // the conduction part (upper triangle only, k grad N_a . grad N_b is symmetric) and one for the k(T) tangent
        const double w = JxW * k, ws = JxW * p.heat_source * p.thickness;
        const double wt = JxW * dk;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const double wgx = w * gx[a], wgy = w * gy[a];
          f[a] = fma(wgx, dTx, f[a]);
          f[a] = fma(wgy, dTy, f[a]);
          if (p.heat_source != 0.) f[a] = fma(-ws, N[a], f[a]);
#pragma unroll
          for (int b = a; b < 4; ++b) K[a][b] = fma(wgy, gy[b], fma(wgx, gx[b], K[a][b]));
          if (p.k_of_T) {
            const double tg = wt * (dTx * gx[a] + dTy * gy[a]);
#pragma unroll
            for (int b = 0; b < 4; ++b) Kt[a][b] = fma(tg, N[b], Kt[a][b]);
          }
        }
      }
      // full matrix: symmetric conduction part + tangent
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const double sym = a <= b ? K[a][b] : K[b][a];
          Kt[a][b] += sym;
        }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) K[a][b] = Kt[a][b];
      if (p.sl_ptr) {  // HEAT_FLUX / CONVECTION_HEAT_FLUX / SURFACE_RADIATION_HEAT_FLUX sides of this element (see above)
        for (int k = p.sl_ptr[e]; k < p.sl_ptr[e + 1]; ++k) {
          const int sd = p.sl_side[k], n1 = sd, n2 = (sd + 1) & 3;
          const double flux = p.sl_val[k * 4], hc = p.sl_val[k * 4 + 1], Tamb = p.sl_val[k * 4 + 2], eps = p.sl_val[k * 4 + 3];
          const double ex_ = sel4(x, n2) - sel4(x, n1), ey_ = sel4(y, n2) - sel4(y, n1), T1 = sel4(T, n1), T2 = sel4(T, n2);
          const double JxWs = 0.5 * sqrt(ex_ * ex_ + ey_ * ey_);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const double xi = q ? GP : -GP;
            const double Ns[2] = {0.5 * (1. - xi), 0.5 * (1. + xi)};
            const double temp = Ns[0] * T1 + Ns[1] * T2;
            const double tz = temp + p.rad_T0, az = Tamb + p.rad_T0;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
              const int a = s ? n2 : n1;
              double df = JxWs * Ns[s] * flux * p.thickness;
              if (hc != 0.) df += JxWs * Ns[s] * p.thickness * hc * (temp - Tamb);
              if (eps != 0.) df += JxWs * Ns[s] * p.rad_sbc * eps * p.thickness * (tz * tz * tz * tz - az * az * az * az);
              double dK1 = 0., dK2 = 0.;
              if (hc != 0.) { dK1 += JxWs * (Ns[s] * Ns[0]) * p.thickness * hc; dK2 += JxWs * (Ns[s] * Ns[1]) * p.thickness * hc; }
              if (eps != 0.) {
                const double d = p.rad_sbc * eps * p.thickness * 4. * (tz * tz * tz);
                dK1 += JxWs * (Ns[s] * Ns[0]) * d; dK2 += JxWs * (Ns[s] * Ns[1]) * d;
              }
#pragma unroll
              for (int aa = 0; aa < 4; ++aa)
                if (aa == a) {
                  f[aa] += df;
#pragma unroll
                  for (int bb = 0; bb < 4; ++bb) {
                    if (bb == n1) K[aa][bb] += dK1;
                    if (bb == n2) K[aa][bb] += dK2;
                  }
                }
            }
          }
        }
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        ke[a * STB + tid] = f[a];
#pragma unroll
        for (int b = 0; b < 4; ++b) ke[(4 + a * 4 + b) * STB + tid] = K[a][b];
      }
    }
  }
  __syncthreads();
  const int li = tid & (SEX - 1), lj = tid / SEX;
  const int i = o0x + li, j = o0y + lj;
  if (li >= SNX || lj >= SNY || i > nx || j > ny) return;
  const int n = i + j * npx;
  const bool con_row = p.cmask && (p.cmask[n] & 1u);
  double acc[9] = {0., 0., 0., 0., 0., 0., 0., 0., 0.};
  double res = 0.;
  // adjacent elements in ascending id: (i-1,j-1) holds the node as its local node 2, (i,j-1) as 3, (i-1,j) as 1, (i,j) as 0
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int dx = (c & 1) - 1, dy = (c >> 1) - 1;  // element (i + dx, j + dy)
    const int a = c == 0 ? 2 : (c == 1 ? 3 : (c == 2 ? 1 : 0));
    if (i + dx < 0 || i + dx >= nx || j + dy < 0 || j + dy >= ny) continue;
    const double* k = ke + (lj + 1 + dy) * SEX + (li + 1 + dx);
    if (p.R) res += con_row ? 0. : k[a * STB];
    if (p.J) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int bx = dx + ((b == 1 || b == 2) ? 1 : 0), by = dy + (b >= 2 ? 1 : 0);  // neighbour offset of local node b
        double v = k[(4 + a * 4 + b) * STB];
        const bool con_col = p.cmask && (p.cmask[n + bx + by * npx] & 1u);
        if (con_row || con_col) v = (b == a && con_row) ? 1. : 0.;
        acc[(by + 1) * 3 + bx + 1] += v;
      }
    }
  }
  if (p.J) {
    // node-level CSR row pointer of a box: rows in node order with xcnt*ycnt neighbours each
    const int ycnt = 3 - (j == 0 ? 1 : 0) - (j == ny ? 1 : 0);
    const int64_t SX = 3 * (int64_t)nx + 1, py = 3 * (int64_t)j - (j > 0 ? 1 : 0), px = 3 * (int64_t)i - (i > 0 ? 1 : 0);
    double* dst = p.J + SX * py + ycnt * px;
#pragma unroll
    for (int by = 0; by < 3; ++by)
#pragma unroll
      for (int bx = 0; bx < 3; ++bx) {
        const bool on = i + bx - 1 >= 0 && i + bx - 1 <= nx && j + by - 1 >= 0 && j + by - 1 <= ny;
        if (on) *dst++ = acc[by * 3 + bx];
      }
  }
  if (p.R) p.R[n] = res;
}

}  // namespace

cudaError_t launch_quad4_heat_struct(const GatherParams& p, int32_t nx, int32_t ny, cudaStream_t stream) {
  const int tiles_x = (nx + 1 + SNX - 1) / SNX, tiles_y = (ny + 1 + SNY - 1) / SNY;
  quad4_heat_struct_kernel<<<tiles_x * tiles_y, STB, 0, stream>>>(p, nx, ny, tiles_x);
  return cudaGetLastError();
}

cudaError_t launch_quad4_heat(const GatherParams& p, cudaStream_t stream) {
  const int blocks = (int)((p.n_owned + TB - 1) / TB);
  quad4_heat_kernel<<<blocks, TB, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace mastb200
