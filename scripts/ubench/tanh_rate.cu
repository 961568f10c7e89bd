// Micro-benchmark: throughput of the tanh variants the MLP epilogue can use (elements / clk / SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/tanh_rate scripts/ubench/tanh_rate.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float v[8];
  unsigned u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = seed + 0.01f * (threadIdx.x + i); u[i] = __float_as_uint(v[i]); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { asm volatile("tanh.approx.f32 %0, %1;" : "=f"(v[i]) : "f"(v[i])); }
      if (MODE == 1) { asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(u[i]) : "r"(u[i])); }
      if (MODE == 2) { asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(u[i]) : "r"(u[i])); }
      if (MODE == 3) { asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(v[i]) : "f"(v[i])); }
      if (MODE == 4) {  // FMA-pipe odd polynomial x*P(x^2), degree 9, clamped
        float x = fminf(fmaxf(v[i], -4.f), 4.f), x2 = x * x;
        float p = fmaf(x2, -2.7607684e-6f, 1.0e-4f);
        p = fmaf(x2, p, -2.0e-3f); p = fmaf(x2, p, 2.1e-2f); p = fmaf(x2, p, -1.3e-1f); p = fmaf(x2, p, 1.0f);
        v[i] = x * p;
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int per_instr) {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 4096, grid = 148 * 2, block = 1024;
  k<MODE><<<grid, block>>>(out, 16, 0.1f);
  cudaEventRecord(a);
  k<MODE><<<grid, block>>>(out, iters, 0.1f);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double elems = (double)grid * block * iters * 8 * per_instr;
  printf("%-22s %.3f ms  %.2f elements/clk/SM (at %d MHz)\n", name, ms, elems / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
  cudaFree(out);
}

int main() {
  run<0>("tanh.approx.f32", 1);
  run<1>("tanh.approx.bf16x2", 2);
  run<2>("tanh.approx.f16x2", 2);
  run<3>("ex2.approx.f32", 1);
  run<4>("poly9 (FMA pipe)", 1);
  return 0;
}
