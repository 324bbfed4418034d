// Which pipe limits the FMHA softmax inner loop on sm_100a?  Instruction MIXES, clocks per "pair of scores" per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/mix tools/ubench/mix.cu && tools/ubench/mix
// Each kind runs 16 warps per SM (4 per scheduler), 8 independent chains per thread.
//   0  2 x ex2.f32                                   (MUFU alone: 16 clk per pair and scheduler expected)
//   1  cvt.rn.f16x2.f32                              (F2FP alone)
//   2  2 x ex2 + cvt                                 (does F2FP share the MUFU pipe?)
//   3  2 x ex2 + fma.f32x2 + add.f32x2 + max3 + cvt  (the softmax loop)
//   4  2 x ex2 + fma.f32x2 + add.f32x2 + max3 + 2 shf + prmt   (integer pack instead of F2FP)
//   5  2 x ex2 + fma.f32x2 + add.f32x2 + max3        (no pack at all)
//   6  max3 alone, 7 shf+shf+prmt alone
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;
constexpr int U = 8;

template <int KIND>
__global__ void __launch_bounds__(512) bench(float* out, long long* cycles) {
  float a[U], b[U], mx[U];
  uint32_t h[U];
  unsigned long long d[U], s[U];
#pragma unroll
  for (int i = 0; i < U; ++i) {
    a[i] = -0.001f * (threadIdx.x + i + 1);
    b[i] = -0.002f * (threadIdx.x + i + 1);
    mx[i] = a[i];
    h[i] = threadIdx.x + i;
    d[i] = (static_cast<unsigned long long>(__float_as_uint(a[i])) << 32) | __float_as_uint(b[i]);
    s[i] = 0ull;
  }
  const unsigned long long c2 = (static_cast<unsigned long long>(__float_as_uint(0.999f)) << 32) | __float_as_uint(0.998f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < U; ++i) {
      if (KIND == 0 || KIND == 2 || KIND == 3 || KIND == 4 || KIND == 5) {
        if (KIND >= 3) {
          asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(c2));
          asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(b[i]) : "l"(d[i]));
        }
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b[i]));
        if (KIND >= 3) {
          unsigned long long pr;
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pr) : "f"(a[i]), "f"(b[i]));
          asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(s[i]) : "l"(pr));
          asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(mx[i]) : "f"(a[i]), "f"(b[i]));
        }
      }
      if (KIND == 1 || KIND == 2 || KIND == 3) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(b[i]));
      if (KIND == 4 || KIND == 7) {
        uint32_t x, y;
        asm volatile("shl.b32 %0, %1, 3;" : "=r"(x) : "r"(__float_as_uint(a[i])));
        asm volatile("shl.b32 %0, %1, 3;" : "=r"(y) : "r"(__float_as_uint(b[i])));
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(h[i]) : "r"(x), "r"(y));
        if (KIND == 7) a[i] = __uint_as_float(h[i]);
      }
      if (KIND == 6) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(mx[i]) : "f"(a[i]), "f"(b[i]));
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < U; ++i)
    acc += a[i] + b[i] + mx[i] + __uint_as_float(h[i]) + __uint_as_float(static_cast<uint32_t>(d[i])) +
           __uint_as_float(static_cast<uint32_t>(s[i]));
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char* name, int threads = 512) {
  const int blocks = 148;
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
  // groups per scheduler = 4 warps x ITERS x U; clocks per warp-group on one scheduler
  printf("%-62s %7.2f clk per warp-group per scheduler (err=%s)\n", name, avg / ((threads / 128.0) * ITERS * U),
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("2 ex2");
  run<1>("cvt.rn.f16x2.f32");
  run<2>("2 ex2 + cvt");
  run<3>("2 ex2 + ffma2 + fadd2 + max3 + cvt (softmax loop)");
  run<4>("2 ex2 + ffma2 + fadd2 + max3 + shl,shl,prmt");
  run<5>("2 ex2 + ffma2 + fadd2 + max3");
  run<6>("max3");
  run<7>("shl,shl,prmt");
  // how many warps does a scheduler need to keep the MUFU pipe full?
  run<0>("2 ex2, ONE warp per scheduler", 128);
  run<0>("2 ex2, TWO warps per scheduler", 256);
  run<3>("softmax loop, ONE warp per scheduler", 128);
  run<3>("softmax loop, TWO warps per scheduler", 256);
  run<3>("softmax loop, THREE warps per scheduler", 384);
  return 0;
}
