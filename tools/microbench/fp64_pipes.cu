// fp64_pipes.cu -- which fp64 pipe should the element kernels use on B200 (sm_100a)?
//
// Measures, with CUDA events, the sustained throughput of
//   (1) scalar DFMA                      (the "FMA path")
//   (2) mma.sync m8n8k4 f64  (DMMA)      (the "tensor path" named by BASELINE.json north_star)
//   (3) mma.sync m16n8k8 / m16n8k16 f64 (wider DMMA shapes, sm_90+)
//   (4) DFMA and DMMA interleaved in one instruction stream (do the two pipes overlap?)
// and prints one JSON line.  Not part of the product path; results are recorded in
// profiles/r01_fp64_pipes.md and decide the inner loop of the HEX8/HEX27 kernels.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void __launch_bounds__(256) k_dmma884(double* out, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dmma1688(double* out, double a, double b) {
  double c[4][4]; double af[4] = {a, a, b, b}; double bf[2] = {b, a};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-9 + i + j;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dmma1688(c[i], af, bf);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dmma16816(double* out, double a, double b) {
  double c[4][4]; double af[8] = {a, a, b, b, a, b, a, b}; double bf[4] = {b, a, a, b};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-9 + i + j;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dmma16816(c[i], af, bf);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 8 DMMA(8x8x4) + NF DFMA per iteration, independent chains: if the pipes overlap the time is
// max(t_dmma, t_dfma), otherwise the sum.
template <int NF>
__global__ void __launch_bounds__(256) k_mixed(double* out, double a, double b) {
  double c[8][2]; double acc[NF];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
#pragma unroll
  for (int i = 0; i < NF; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
      for (int j = 0; j < NF / 8; ++j) acc[i * (NF / 8) + j] = fma(acc[i * (NF / 8) + j], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < NF; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int blocks = p.multiProcessorCount * 8, threads = 256;
  double* out; CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
  const double a = 0.999999, b = 1e-7;
  const double nthr = double(blocks) * threads, nwarp = nthr / 32;

  double t_fma = time_ms([&] { k_dfma<<<blocks, threads>>>(out, a, b); });
  double t_884 = time_ms([&] { k_dmma884<<<blocks, threads>>>(out, a, b); });
  double t_1688 = time_ms([&] { k_dmma1688<<<blocks, threads>>>(out, a, b); });
  double t_16816 = time_ms([&] { k_dmma16816<<<blocks, threads>>>(out, a, b); });
  double t_mix16 = time_ms([&] { k_mixed<16><<<blocks, threads>>>(out, a, b); });
  double t_mix64 = time_ms([&] { k_mixed<64><<<blocks, threads>>>(out, a, b); });
  CK(cudaGetLastError());

  const double tf_fma = nthr * ITERS * 16 * 2 / (t_fma * 1e-3) / 1e12;
  const double tf_884 = nwarp * ITERS * 8 * (8. * 8 * 4 * 2) / (t_884 * 1e-3) / 1e12;
  const double tf_1688 = nwarp * ITERS * 4 * (16. * 8 * 8 * 2) / (t_1688 * 1e-3) / 1e12;
  const double tf_16816 = nwarp * ITERS * 4 * (16. * 8 * 16 * 2) / (t_16816 * 1e-3) / 1e12;
  // mixed: flops of each part over the common time
  const double mix16_dmma = nwarp * ITERS * 8 * 512. / (t_mix16 * 1e-3) / 1e12;
  const double mix16_dfma = nthr * ITERS * 16 * 2 / (t_mix16 * 1e-3) / 1e12;
  const double mix64_dmma = nwarp * ITERS * 8 * 512. / (t_mix64 * 1e-3) / 1e12;
  const double mix64_dfma = nthr * ITERS * 64 * 2 / (t_mix64 * 1e-3) / 1e12;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_m8n8k4_tflops\": %.2f, "
         "\"dmma_m16n8k8_tflops\": %.2f, \"dmma_m16n8k16_tflops\": %.2f, "
         "\"mixed_8dmma_16dfma\": {\"ms\": %.3f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f}, "
         "\"mixed_8dmma_64dfma\": {\"ms\": %.3f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f}, "
         "\"ms\": {\"dfma\": %.3f, \"m8n8k4\": %.3f, \"m16n8k8\": %.3f, \"m16n8k16\": %.3f}}\n",
         p.name, p.multiProcessorCount, tf_fma, tf_884, tf_1688, tf_16816,
         t_mix16, mix16_dmma, mix16_dfma, t_mix64, mix64_dmma, mix64_dfma,
         t_fma, t_884, t_1688, t_16816);
  return 0;
}
