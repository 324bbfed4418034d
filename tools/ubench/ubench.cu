// Micro-benchmarks of per-SM instruction throughput on sm_100a (results/clk/SM), to size the FMHA softmax:
//   ex2.approx.ftz.f32, ex2.approx.ftz.f16x2, fma.rn.f32x2, fma.rn.f32, fma.rn.f16x2, cvt.rn.f16x2.f32
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ubench tools/ubench/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int UNROLL = 8;

template <int KIND>
__global__ void __launch_bounds__(512) bench(float* out, long long* cycles) {
  float r[UNROLL];
  uint32_t h[UNROLL];
  unsigned long long d[UNROLL];
#pragma unroll
  for (int i = 0; i < UNROLL; ++i) {
    r[i] = -0.001f * (threadIdx.x + i + 1);
    h[i] = 0xb800b400u + threadIdx.x + i;  // two negative halves
    d[i] = (static_cast<unsigned long long>(__float_as_uint(r[i])) << 32) | __float_as_uint(0.5f * r[i]);
  }
  const unsigned long long c2 = (static_cast<unsigned long long>(__float_as_uint(0.999f)) << 32) | __float_as_uint(0.998f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      if (KIND == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[i]));
      if (KIND == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (KIND == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(c2));
      if (KIND == 3) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(r[i]) : "f"(0.999f));
      if (KIND == 4) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(h[i]) : "r"(0x3bff3bfeu));
      if (KIND == 5) asm volatile("{.reg .f32 a; mov.b32 a, %0; cvt.rn.f16x2.f32 %0, a, a;}" : "+r"(h[i]));
      if (KIND == 6) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(c2));
      if (KIND == 7) asm volatile("ex2.approx.f16 %0, %0;" : "+h"(*reinterpret_cast<unsigned short*>(&h[i])));
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < UNROLL; ++i) acc += r[i] + __uint_as_float(h[i]) + __uint_as_float(static_cast<uint32_t>(d[i]));
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char* name, int results_per_op) {
  const int blocks = 148, threads = 512;
  float* out;
  long long* cyc;
  cudaMalloc(&out, blocks * threads * sizeof(float));
  cudaMalloc(&cyc, blocks * sizeof(long long));
  bench<KIND><<<blocks, threads>>>(out, cyc);
  bench<KIND><<<blocks, threads>>>(out, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks; ++i) avg += h[i];
  avg /= blocks;
  const double ops = static_cast<double>(threads) * ITERS * UNROLL;
  printf("%-28s %8.1f instr-lanes/clk/SM  %8.1f results/clk/SM  (err=%s)\n", name, ops / avg, ops * results_per_op / avg,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.ftz.f16x2", 2);
  run<7>("ex2.approx.f16", 1);
  run<2>("fma.rn.f32x2", 2);
  run<3>("fma.rn.f32", 1);
  run<4>("fma.rn.f16x2", 2);
  run<5>("cvt.rn.f16x2.f32", 2);
  run<6>("add.rn.f32x2", 2);
  return 0;
}
