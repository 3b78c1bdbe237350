// Probe 2: (a) do DFMA and DMMA share one pipe on B200?  (b) DMMA dependent-issue latency.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// NM independent DMMA chains + NF independent DFMA chains per thread per iteration
template <int NM, int NF>
__global__ void k_mix(double* out, int iters, double a0, double b0) {
  double c[NM > 0 ? NM : 1][2], f[NF > 0 ? NF : 1];
#pragma unroll
  for (int i = 0; i < NM; ++i) c[i][0] = c[i][1] = 0.0;
#pragma unroll
  for (int i = 0; i < NF; ++i) f[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < (NM > NF ? NM : NF); ++i) {
      if (i < NM) dmma884(c[i][0], c[i][1], a, b);
      if (i < NF) f[i] = fma(a, f[i], b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NM; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < NF; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// one warp per SM sub-partition issue slot, ONE dependent DMMA chain: latency = cycles / iters
__global__ void k_lat(double* out, long long* cyc, int iters) {
  double c0 = 0, c1 = 0, a = 1.0 + threadIdx.x, b = 0.5;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) dmma884(c0, c1, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = c0 + c1;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
// chain where the accumulator result feeds the NEXT DMMA's A operand (register chaining as in the BP kernel)
__global__ void k_lat_chainA(double* out, long long* cyc, int iters) {
  double c0 = 1e-3 * threadIdx.x, c1 = 0, b = 0.5;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) { double d0 = 0, d1 = 0; dmma884(d0, d1, c0, b); c0 = d0; c1 = d1; }
  long long t1 = clock64();
  out[threadIdx.x] = c0 + c1;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <typename F> float time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
  return best;
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount, iters = 20000, grid = sms * 2, block = 512;
  double* out; CK(cudaMalloc(&out, sizeof(double) * grid * block));
  long long* cyc; CK(cudaMalloc(&cyc, 8));
  auto rep = [&](const char* name, float ms, int nm, int nf) {
    double fm = 2.0 * 256 * nm * (double)iters * (block / 32) * grid, ff = 2.0 * 32 * nf * (double)iters * (block / 32) * grid;
    printf("%-28s %8.3f ms  DMMA %6.2f + DFMA %6.2f = %6.2f TFLOP/s\n", name, ms, fm / ms * 1e-9, ff / ms * 1e-9, (fm + ff) / ms * 1e-9);
  };
  rep("DMMA x8 only", time_ms([&] { k_mix<8, 0><<<grid, block>>>(out, iters, 1.0000001, 1e-9); }), 8, 0);
  rep("DFMA x8 only", time_ms([&] { k_mix<0, 8><<<grid, block>>>(out, iters, 1.0000001, 1e-9); }), 0, 8);
  rep("DMMA x8 + DFMA x8", time_ms([&] { k_mix<8, 8><<<grid, block>>>(out, iters, 1.0000001, 1e-9); }), 8, 8);
  rep("DMMA x8 + DFMA x64", time_ms([&] { k_mix<8, 64><<<grid, block>>>(out, iters / 4, 1.0000001, 1e-9); }) * 4, 8, 64);
  rep("DMMA x4 + DFMA x32", time_ms([&] { k_mix<4, 32><<<grid, block>>>(out, iters, 1.0000001, 1e-9); }), 4, 32);
  long long h;
  k_lat<<<1, 32>>>(out, cyc, 4096); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("DMMA accumulate-chain latency: %.1f clk\n", (double)h / 4096);
  k_lat_chainA<<<1, 32>>>(out, cyc, 4096); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("DMMA D->A operand chain latency: %.1f clk\n", (double)h / 4096);
  return 0;
}
