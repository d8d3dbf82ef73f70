// sweep_hex8_sens_struct.cu -- density sensitivities alone (SURVEY 8a row S without R / J, the second half of the BASELINE
// metric) for ijk-numbered (build_cube) HEX8 meshes with implicit connectivity (sm_100a, fp64).
//
// Device counterpart of NonlinearImplicitAssembly::sensitivity_assemble (/root/reference/src/base/
// nonlinear_implicit_assembly.cpp:515-686) + StructuralElement3D::internal_residual_sensitivity
// (src/elasticity/solid_element_3d.cpp:566-788, D <- dD/dp at :640) + the dot products of
// AssemblyBase::calculate_output_adjoint_sensitivity_multiple_parameters_no_direct (src/base/assembly_base.cpp:594-639)
// for nodal-density design variables under the SIMP law (examples/structural/example_6/example_6.cpp:100-137), linear
// strain:  G_n = adj . dR/d(rho_n) = sum_e sum_q N_n(q) JxW E'(rho_q) eps(adj)^T D_1 eps(u), adj zeroed on constrained dofs.
//
// The element-parallel kernels of stress_output.cu (sens_qp_hex8_kernel + sens_gather_kernel) read the eight nodes of
// every (element, qp) from global memory (80 loads per thread) and pass 8 doubles per element through HBM.  Here the work
// unit is the structured sweep's: a column of <= 15x11 owned nodes over a z-range with the <= 16x12 elements that touch it (its own unit list);
// node planes (coordinates, state, adjoint, density) are staged once per plane with cp.async, lane = (element, qp) forms
// s_q, the eight lanes of an element turn the point values into nodal sums by a three-stage shuffle butterfly and leave them in a
// per-element slot; one thread per owned node then adds its <= 8 corner values in ascending element id (the reference's
// ADD order) and writes G -- while the other warps already compute the next element plane (three slot buffers, one CTA
// barrier per plane).  Nothing but G leaves the SM; bitwise reproducible; ~33 KB of shared memory, two CTAs per SM.
#include <cuda_runtime.h>

#include "kernels.h"

namespace mastb200 {
namespace {

constexpr int SW = 8;                  // warps per CTA
constexpr int ST = SW * 32;
// Tile of EX x EY elements = (EX + 1) x (EY + 1) node columns, of which <= kSensTileX x kSensTileY are owned.  The kernel keeps
// only node planes and one value per (element, node) in shared memory, so it affords a larger tile than the R / J sweeps:
// 16 x 12 elements visit 1.16 elements per element where 8 x 8 visit 1.31 -- the kernel is fp64-bound, the halo is pure loss.
constexpr int EX = kSensTileX + 1, EY = kSensTileY + 1, NE = EX * EY;
constexpr int TCX = EX + 1, TCY = EY + 1;
constexpr int NS = TCX * TCY;
static_assert(EX == 16 && NE % 32 == 0 && NS <= ST && kSensTileX * kSensTileY <= ST, "tile shape vs. thread mapping");
constexpr int NSB = (NS + 3) & ~3;
constexpr double GP = 0.57735026918962576451;

struct SmemS {
  double xyz[3][3][NS];    // [plane % 3][component][node column]
  double u[3][3][NS];
  double adj[3][3][NS];
  double rho[3][NS];
  double slot[3][NE][8];   // [element plane % 3][element column][node in ijk order]: sum_q N_a(q) s_q
  uint8_t cm[3][NSB];
  int2 node[kSensTileX * kSensTileY];  // owned nodes of the unit: in-plane node id, lx | ly << 8
  int misc[4];
};

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// lane = (element, qp): s_q = JxW E'(rho_q) eps(adj)^T D_1 eps(u); returns the nodal sum of the lane's node a = q
__device__ __forceinline__ double element_point(const StructParams& p, const SmemS& sm, int s, int base, int q, bool live, bool same,
                                               int lane) {
  const double xi = (q & 1) ? GP : -GP, eta = (q & 2) ? GP : -GP, zeta = (q & 4) ? GP : -GP;
  const double xm = 1. - xi, xp = 1. + xi, ym = 1. - eta, yp = 1. + eta, zm = 1. - zeta, zp = 1. + zeta;
  double sq = 0.;
  if (live) {
    double Jm[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};  // Jm[r][c] = d x_c / d xi_r
    double Hr[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}}, Ar[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};
    double rho_q = 0.;
    unsigned cmany = 0u;
    if (p.cmask) {
#pragma unroll
      for (int m = 0; m < 8; ++m) cmany |= sm.cm[((m & 4) ? s + 1 : s) % 3][base + ((m >> 1) & 1) * TCX + (m & 1)];
    }
    const bool need_adj = !same || cmany != 0u;  // the adjoint's gradient is the state's unless a dof of the element is constrained
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const double fx = (m & 1) ? xp : xm, fy = (m & 2) ? yp : ym, fz = (m & 4) ? zp : zm;
      const double d0 = ((m & 1) ? 0.125 : -0.125) * fy * fz, d1 = ((m & 2) ? 0.125 : -0.125) * fx * fz,
                   d2 = ((m & 4) ? 0.125 : -0.125) * fx * fy;
      const int z = ((m & 4) ? s + 1 : s) % 3, col = base + ((m >> 1) & 1) * TCX + (m & 1);
      const double x0 = sm.xyz[z][0][col], x1 = sm.xyz[z][1][col], x2 = sm.xyz[z][2][col];
      Jm[0][0] = fma(d0, x0, Jm[0][0]); Jm[0][1] = fma(d0, x1, Jm[0][1]); Jm[0][2] = fma(d0, x2, Jm[0][2]);
      Jm[1][0] = fma(d1, x0, Jm[1][0]); Jm[1][1] = fma(d1, x1, Jm[1][1]); Jm[1][2] = fma(d1, x2, Jm[1][2]);
      Jm[2][0] = fma(d2, x0, Jm[2][0]); Jm[2][1] = fma(d2, x1, Jm[2][1]); Jm[2][2] = fma(d2, x2, Jm[2][2]);
      const double u0 = sm.u[z][0][col], u1 = sm.u[z][1][col], u2 = sm.u[z][2][col];
      Hr[0][0] = fma(u0, d0, Hr[0][0]); Hr[0][1] = fma(u0, d1, Hr[0][1]); Hr[0][2] = fma(u0, d2, Hr[0][2]);
      Hr[1][0] = fma(u1, d0, Hr[1][0]); Hr[1][1] = fma(u1, d1, Hr[1][1]); Hr[1][2] = fma(u1, d2, Hr[1][2]);
      Hr[2][0] = fma(u2, d0, Hr[2][0]); Hr[2][1] = fma(u2, d1, Hr[2][1]); Hr[2][2] = fma(u2, d2, Hr[2][2]);
      if (need_adj) {
        const unsigned cm = p.cmask ? sm.cm[z][col] : 0u;
        const double* L0 = same ? &sm.u[z][0][col] : &sm.adj[z][0][col];
        const double l0 = (cm & 1u) ? 0. : L0[0], l1 = (cm & 2u) ? 0. : L0[NS], l2 = (cm & 4u) ? 0. : L0[2 * NS];
        Ar[0][0] = fma(l0, d0, Ar[0][0]); Ar[0][1] = fma(l0, d1, Ar[0][1]); Ar[0][2] = fma(l0, d2, Ar[0][2]);
        Ar[1][0] = fma(l1, d0, Ar[1][0]); Ar[1][1] = fma(l1, d1, Ar[1][1]); Ar[1][2] = fma(l1, d2, Ar[1][2]);
        Ar[2][0] = fma(l2, d0, Ar[2][0]); Ar[2][1] = fma(l2, d1, Ar[2][1]); Ar[2][2] = fma(l2, d2, Ar[2][2]);
      }
      rho_q = fma(0.125 * fx * fy * fz, sm.rho[z][col], rho_q);
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
    double H[3][3], A[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        H[i][j] = Hr[i][0] * inv[j][0] + Hr[i][1] * inv[j][1] + Hr[i][2] * inv[j][2];
        A[i][j] = need_adj ? Ar[i][0] * inv[j][0] + Ar[i][1] * inv[j][1] + Ar[i][2] * inv[j][2] : H[i][j];
      }
    const double e0 = H[0][0], e1 = H[1][1], e2 = H[2][2];
    const double e3 = H[0][1] + H[1][0], e4 = H[1][2] + H[2][1], e5 = H[0][2] + H[2][0];
    const double a0 = A[0][0], a1 = A[1][1], a2 = A[2][2];
    const double a3 = A[0][1] + A[1][0], a4 = A[1][2] + A[2][1], a5 = A[0][2] + A[2][0];
    const double dens = p.lam1 * (a0 + a1 + a2) * (e0 + e1 + e2) +
                        p.mu1 * (2. * (a0 * e0 + a1 * e1 + a2 * e2) + a3 * e3 + a4 * e4 + a5 * e5);
    const double pm1 = p.penalty - 1.;
    const double rpow = pm1 == 3. ? rho_q * rho_q * rho_q : (pm1 == 2. ? rho_q * rho_q : (pm1 == 1. ? rho_q : (pm1 == 0. ? 1. : pow(rho_q, pm1))));
    sq = det * p.E0 * (1. - p.rho_min) * p.penalty * rpow * dens;
  }
  // lane a of the element's eight lanes ends with sum_q N_a(q) s_q (a in ijk order: dx = a & 1, dy = (a >> 1) & 1, dz = a >> 2).
  // N_a(q) = prod_d (1 +- GP) / 2 is a tensor product, so the sum is three butterfly stages (x, y, z): v <- w_same v +
  // w_other v(partner), partner = the lane whose bit d differs -- 6 shuffles and 6 fp64 operations instead of 16 and 24.
  // The summation order is fixed (x, then y, then z), so runs stay bitwise reproducible.
  double acc = sq;
  {
    const double ws = 0.5 * (1. + GP), wo = 0.5 * (1. - GP);
#pragma unroll
    for (int bit = 1; bit < 8; bit <<= 1) {
      const double other = __shfl_xor_sync(0xffffffffu, acc, bit);
      acc = ws * acc + wo * other;
    }
  }
  return acc;
}

__global__ void __launch_bounds__(ST, 2) hex8_sens_struct_kernel(const StructParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemS& sm = *reinterpret_cast<SmemS*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool same = p.adj == p.X;

  for (;;) {
    __syncthreads();
    if (tid == 0) sm.misc[0] = atomicAdd(p.counter, 1);
    __syncthreads();
    const int unit = sm.misc[0];
    if (unit >= p.n_units) break;
    const StructUnit u = p.units[unit];
    const int s_begin = u.z0 > 0 ? u.z0 - 1 : 0;
    const int64_t npx = p.nx + 1, plane_nodes = npx * (p.ny + 1);
    const int ncx = u.nex + 1;
    const int n_own = u.nox * u.noy;
    // the node column this thread stages every plane (tile shape 9 x 9: constant divisors)
    const int col = tid, cly = tid / TCX, clx = tid - cly * TCX;
    const bool col_on = tid < NS && clx < ncx && cly <= u.ney;
    const int64_t col_src = (u.e0x + clx) + (u.e0y + cly) * npx;
    auto load_plane = [&](int z) {
      if (!col_on) return;
      const int zz = z % 3;
      const int64_t n = (int64_t)z * plane_nodes + col_src;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        cp_async8(&sm.xyz[zz][c][col], p.xyz + n * 3 + c);
        cp_async8(&sm.u[zz][c][col], p.X + n * 3 + c);
        if (!same) cp_async8(&sm.adj[zz][c][col], p.adj + n * 3 + c);
      }
      cp_async8(&sm.rho[zz][col], p.rho + n);
      if (p.cmask) sm.cm[zz][col] = p.cmask[n];
    };
    if (tid < n_own) {
      const int oyi = tid / u.nox, oxi = tid - oyi * u.nox;
      const int ax = u.o0x + oxi, ay = u.o0y + oyi;
      sm.node[tid] = make_int2((int)(ax + ay * npx), (ax - u.e0x) | ((ay - u.e0y) << 8));
    }
    load_plane(s_begin);
    if (s_begin + 1 <= p.nz) load_plane(s_begin + 1);
    cp_async_wait_all();
    __syncthreads();

    for (int s = s_begin; s < u.z1; ++s) {
      if (s + 2 <= p.nz) load_plane(s + 2);  // fetched while this plane is processed (its slot held plane s - 1)
      if (s < p.nz) {
        // element plane s: four elements per warp and pass (lane = (element, qp)), 32 per pass and CTA
#pragma unroll 1
        for (int pass = 0; pass < NE / (SW * 4); ++pass) {
          const int el = pass * (SW * 4) + warp * 4 + (lane >> 3), q = lane & 7;
          const int lex = el & (EX - 1), ley = el / EX;
          if ((pass * (SW * 4)) / EX >= u.ney) break;  // the remaining rows of the tile hold no element of this unit
          const bool live = lex < u.nex && ley < u.ney;
          const double g = element_point(p, sm, s, ley * TCX + lex, q, live, same, lane);
          if (live) sm.slot[s % 3][el][q] = g;
        }
      }
      cp_async_wait_all();
      __syncthreads();
      // node plane s is complete: corner elements of plane s - 1 (the node is one of their top nodes) and of plane s.  ONE
      // barrier per plane: three slot buffers, so the next element plane is computed while these few threads add corners
      if (s >= u.z0 && tid < n_own) {
        const int2 ni = sm.node[tid];
        const int lx = ni.y & 255, ly = ni.y >> 8;
        double acc = 0.;
#pragma unroll
        for (int cz = 0; cz < 2; ++cz) {
          const int layer = s - 1 + cz;
          if (layer < s_begin || layer < 0 || layer >= p.nz) continue;
#pragma unroll
          for (int cy = 0; cy < 2; ++cy) {
            const int ley = ly - 1 + cy;
            if ((unsigned)ley >= (unsigned)u.ney) continue;
#pragma unroll
            for (int cx = 0; cx < 2; ++cx) {
              const int lex = lx - 1 + cx;
              if ((unsigned)lex >= (unsigned)u.nex) continue;
              acc += sm.slot[layer % 3][ley * EX + lex][(1 - cx) + 2 * (1 - cy) + 4 * (1 - cz)];
            }
          }
        }
        p.G[(int64_t)s * plane_nodes + ni.x] = acc;
      }
    }
  }
}

}  // namespace

size_t hex8_sens_struct_smem_bytes() { return sizeof(SmemS); }

cudaError_t launch_hex8_sens_struct(const StructParams& p, int sm_count, cudaStream_t stream) {
  if (p.nv != 3 || !p.G || !p.rho || !p.adj) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(p.counter, 0, sizeof(int32_t), stream);
  if (e != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(hex8_sens_struct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemS))) != cudaSuccess) return e;
  const int grid = p.n_units < 2 * sm_count ? p.n_units : 2 * sm_count;
  hex8_sens_struct_kernel<<<grid > 0 ? grid : 1, ST, sizeof(SmemS), stream>>>(p);
  return cudaGetLastError();
}

}  // namespace mastb200
