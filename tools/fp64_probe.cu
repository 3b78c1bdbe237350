// FP64 pipe probe for B200: DFMA vs DMMA (mma.sync m8n8k4 f64) issue rate, plus LDS.64 bandwidth.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_probe tools/fp64_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double* out, int iters, double a0, double b0) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = a0 + threadIdx.x, b = b0 - threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a0, double b0) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(a, c[i], b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA fed from shared memory: per 8 DMMAs, 4 LDS.64 for A fragments (B fragments in registers) + 4 STS.64
__global__ void k_dmma_smem(double* out, int iters, int stride) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1e-3 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  double bfrag[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) bfrag[i] = 1.0 + i + lane;
  double acc = 0;
  for (int it = 0; it < iters; ++it) {
    const int base = ((it * 8 + warp) * 128) & 8191;
    double c[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double a = sm[(base + g * stride + k * 4 + t) & 8191];
      dmma884(c[0][0], c[0][1], a, bfrag[2 * k]);
      dmma884(c[1][0], c[1][1], a, bfrag[2 * k + 1]);
    }
    acc += c[0][0] + c[0][1] + c[1][0] + c[1][1];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void k_lds(double* out, int iters, int stride) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double acc = 0;
  int idx = (threadIdx.x * stride) & 8191;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += sm[(idx + u * 1024 + it) & 8191];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
float time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("device %s, %d SMs, clock attr %d kHz\n", p.name, p.multiProcessorCount, clk);
  const int sms = p.multiProcessorCount;
  double* out;
  CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 1024));
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    for (int bps : {1, 2}) {
      const int grid = sms * bps, block = warps * 32;
      float ms = time_ms([&] { k_dmma<8><<<grid, block>>>(out, iters, 1.0, 2.0); });
      double flops = 2.0 * 256 * 8 * (double)iters * warps * grid;
      printf("DMMA m8n8k4  x8 acc: warps/CTA %2d CTAs/SM %d: %8.3f ms  %7.2f TFLOP/s\n", warps, bps, ms, flops / ms * 1e-9);
    }
  }
  for (int warps : {8, 16, 32}) {
    const int grid = sms * 2, block = warps * 32;
    float ms = time_ms([&] { k_dfma<16><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
    double flops = 2.0 * 32 * 16 * (double)iters * warps * grid;
    printf("DFMA x16 chains:     warps/CTA %2d CTAs/SM 2: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flops / ms * 1e-9);
  }
  {
    const int grid = sms * 2, block = 16 * 32;
    float ms = time_ms([&] { k_dmma<2><<<grid, block>>>(out, iters, 1.0, 2.0); });
    double flops = 2.0 * 256 * 2 * (double)iters * 16 * grid;
    printf("DMMA m8n8k4  x2 acc (latency bound): %8.3f ms  %7.2f TFLOP/s\n", ms, flops / ms * 1e-9);
  }
  CK(cudaFuncSetAttribute(k_dmma_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  for (int stride : {4, 16, 20, 128}) {
    const int grid = sms * 2, block = 8 * 32;
    float ms = time_ms([&] { k_dmma_smem<<<grid, block, 65536>>>(out, iters, stride); });
    double flops = 2.0 * 256 * 8 * (double)iters * 8 * grid;
    printf("DMMA from smem, A row stride %3d doubles: %8.3f ms  %7.2f TFLOP/s\n", stride, ms, flops / ms * 1e-9);
  }
  for (int stride : {1, 2, 16}) {
    const int grid = sms * 2, block = 512;
    float ms = time_ms([&] { k_lds<<<grid, block, 65536>>>(out, iters / 4, stride); });
    double bytes = 8.0 * 8 * (double)(iters / 4) * block * grid;
    printf("LDS.64 lane stride %2d: %8.3f ms  %8.1f GB/s  (%.1f B/clk/SM at %d MHz attr)\n", stride, ms, bytes / ms * 1e-6,
           bytes / ms * 1e-6 / sms / (clk * 1e-6), clk / 1000);
  }
  return 0;
}
