// Micro-benchmark (GPU box): issue rate of FFMA (3-register), FFMA with a shared broadcast operand, and FFMA2
// (fma.rn.f32x2) on this GPU. Prints lane-FMAs per clock per SM. Build: nvcc -O3 -arch=sm_100a -o ffma_bench ffma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float w, float2 f, float2 c) {
  unsigned long long rw, rf, rc, rd;
  float2 w2 = make_float2(w, w);
  rw = *reinterpret_cast<unsigned long long*>(&w2); rf = *reinterpret_cast<unsigned long long*>(&f); rc = *reinterpret_cast<unsigned long long*>(&c);
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(rw), "l"(rf), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float acc[16];
  for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 0.001f + i;
  float2 acc2[8];
  for (int i = 0; i < 8; i++) acc2[i] = make_float2(acc[2 * i], acc[2 * i + 1]);
  float w = a + threadIdx.x * 1e-6f, f0 = b, f1 = b * 1.5f;
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; i++) acc[i] = fmaf(w, (i & 1) ? f1 : f0, acc[i]);  // 16 FFMA
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) acc2[i] = ffma2(w, make_float2(f0, f1), acc2[i]);  // 8 FFMA2 = 16 lane-FMA
    }
  }
  float s = 0;
  for (int i = 0; i < 16; i++) s += acc[i];
  for (int i = 0; i < 8; i++) s += acc2[i].x + acc2[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * 4);
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int iters = 20000;
  for (int mode = 0; mode < 2; mode++) {
    for (int warps = 4; warps <= 32; warps *= 2) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      auto run = [&]() { if (mode == 0) k<0><<<p.multiProcessorCount, warps * 32>>>(out, iters, 1.0001f, 0.5f); else k<1><<<p.multiProcessorCount, warps * 32>>>(out, iters, 1.0001f, 0.5f); };
      run(); cudaDeviceSynchronize();
      cudaEventRecord(e0); run(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = (double)p.multiProcessorCount * warps * 32 * 16.0 * iters;
      int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
      printf("%s warps/SM %2d: %.3f ms  %.1f TFLOP/s  %.1f lane-FMA/clk/SM (at %d MHz nominal)\n", mode ? "FFMA2" : "FFMA ", warps, ms,
             2 * fma / ms / 1e9, fma / (ms * 1e-3) / p.multiProcessorCount / (khz * 1e3), khz / 1000);
    }
  }
  return 0;
}
