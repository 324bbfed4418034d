// tcgen05.mma issue/throughput per instruction shape on sm_100a (clocks per MMA, one issuing thread per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I video-diffusion-pipeline-parallel_b200/csrc -o tools/ubench/mma tools/ubench/mma.cu
// Shapes (kind::f16, M = 128, K = 16 per instruction):
//   SS N=64 / 128 / 256  (A and B from shared memory, K-major, 128B swizzle)
//   TS N=64              (A from tensor memory, B MN-major from shared memory: the FMHA's P V product)
// and two interleaved streams from TWO issuing warps (as the two-tile FMHA does).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace svdpp;

__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}

// mode 0: SS, N given; mode 1: TS N = 64; issuers = 1 or 2 warps, each with its own accumulator columns
__global__ void __launch_bounds__(128, 1) bench(int mode, int n, int issuers, int accs, int iters, long long* cycles, int style) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (warp < issuers && style == 1) {
    // the WHOLE warp runs the loop (addresses / descriptors stay in uniform registers); one elected lane issues
    const uint32_t a_addr = smem_u32(smem + warp * 16384);
    const uint32_t b_addr = smem_u32(smem + 32768);
    const uint32_t d0 = tm + warp * (issuers > 1 ? 256 : 0);
    const uint32_t idesc = make_idesc_f16(n, mode == 1);
    const bool leader = elect_one();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t d = d0 + (accs > 1 ? ((it & 1) * n) : 0);
        const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 1024, 0);
        const uint64_t db = mode == 0 ? make_smem_desc_sw128(b_addr + k * 32, 1024, 0)
                                      : make_smem_desc_sw128(b_addr + k * 2048, 1024, 8192);
        if (leader) {
          if (mode == 0)
            umma_f16(d, da, db, idesc, 1u);
          else
            mma_ts(d, tm + 448 + k * 8, db, idesc, 1u);
        }
      }
    }
    if (leader) {
      umma_commit(&bar[warp]);
      mbar_wait(&bar[warp], 0, 1);
      const long long t1 = clock64();
      cycles[blockIdx.x * 2 + warp] = t1 - t0;
    }
    __syncwarp();
  } else if (warp < issuers && lane == 0) {
    const uint32_t a_addr = smem_u32(smem + warp * 16384);
    const uint32_t b_addr = smem_u32(smem + 32768);
    const uint32_t d0 = tm + warp * (issuers > 1 ? 256 : 0);
    const uint32_t idesc = make_idesc_f16(n, mode == 1);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t d = d0 + (accs > 1 ? ((it & 1) * n) : 0);  // alternate between two independent accumulators
        if (mode == 0)
          umma_f16(d, make_smem_desc_sw128(a_addr + k * 32, 1024, 0), make_smem_desc_sw128(b_addr + k * 32, 1024, 0), idesc, 1u);
        else
          mma_ts(d, tm + 448 + k * 8, make_smem_desc_sw128(b_addr + k * 2048, 1024, 8192), idesc, 1u);
      }
    }
    umma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0, 1);
    const long long t1 = clock64();
    cycles[blockIdx.x * 2 + warp] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tm, 512);
  }
}

static void run(const char* name, int mode, int n, int issuers, int accs, int style) {
  const int blocks = 148, iters = 2048;
  long long* cyc;
  cudaMalloc(&cyc, blocks * 2 * sizeof(long long));
  cudaMemset(cyc, 0, blocks * 2 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  bench<<<blocks, 128, 66 * 1024>>>(mode, n, issuers, accs, iters, cyc, style);
  bench<<<blocks, 128, 66 * 1024>>>(mode, n, issuers, accs, iters, cyc, style);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[296];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks; ++i) avg += h[2 * i] > h[2 * i + 1] ? h[2 * i] : h[2 * i + 1];
  avg /= blocks;
  const double per = avg / (4.0 * iters * issuers);
  const double flop = 2.0 * 128 * n * 16;
  printf("%-34s %7.1f clk per MMA (all issuers)  %7.0f FLOP/clk/SM  (%s)\n", name, per, flop / per, cudaGetErrorString(e));
  cudaFree(cyc);
}


// lean issue loop: compile-time shape, descriptors computed once, 8 MMAs per iteration, converged warp, elected lane
template <int N, int ACCS>
__global__ void __launch_bounds__(128, 1) bench_lean(int issuers, int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (warp < issuers) {
    const uint32_t a_addr = smem_u32(smem + warp * 16384);
    const uint32_t b_addr = smem_u32(smem + 32768);
    const uint32_t d0 = tm + warp * 256;
    constexpr uint32_t idesc = make_idesc_f16(N, false);
    uint64_t da[4], db[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      da[k] = make_smem_desc_sw128(a_addr + k * 32, 1024, 0);
      db[k] = make_smem_desc_sw128(b_addr + k * 32, 1024, 0);
    }
    const bool leader = elect_one();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it += 2) {
      if (leader) {
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d0 + (ACCS > 1 ? u * N : 0), da[k], db[k], idesc, 1u);
      }
      __syncwarp();
    }
    if (leader) {
      umma_commit(&bar[warp]);
      mbar_wait(&bar[warp], 0, 1);
      cycles[blockIdx.x * 2 + warp] = clock64() - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tm, 512);
  }
}
template <int N, int ACCS>
static void run_lean(const char* name, int issuers) {
  const int blocks = 148, iters = 2048;
  long long* cyc;
  cudaMalloc(&cyc, blocks * 2 * sizeof(long long));
  cudaMemset(cyc, 0, blocks * 2 * sizeof(long long));
  cudaFuncSetAttribute(bench_lean<N, ACCS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  bench_lean<N, ACCS><<<blocks, 128, 66 * 1024>>>(issuers, iters, cyc);
  bench_lean<N, ACCS><<<blocks, 128, 66 * 1024>>>(issuers, iters, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[296];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks; ++i) avg += h[2 * i] > h[2 * i + 1] ? h[2 * i] : h[2 * i + 1];
  avg /= blocks;
  const double per = avg / (4.0 * iters * issuers);
  printf("%-34s %7.1f clk per MMA (all issuers)  %7.0f FLOP/clk/SM  (%s)\n", name, per, 2.0 * 128 * N * 16 / per, cudaGetErrorString(e));
  cudaFree(cyc);
}
int main() {
  for (int style = 0; style < 2; ++style) {
    printf("--- issue style %d (%s)\n", style, style ? "converged warp, elected lane issues" : "if (lane == 0) around the loop");
    run("SS N=64, 1 issuer", 0, 64, 1, 1, style);
    run("SS N=128, 1 issuer", 0, 128, 1, 1, style);
    run("SS N=256, 1 issuer", 0, 256, 1, 1, style);
    run("TS N=64, 1 issuer", 1, 64, 1, 1, style);
    run("SS N=64, 1 issuer, 2 accumulators", 0, 64, 1, 2, style);
    run("SS N=128, 1 issuer, 2 accumulators", 0, 128, 1, 2, style);
    run("SS N=256, 1 issuer, 2 accumulators", 0, 256, 1, 2, style);
    run("TS N=64, 1 issuer, 2 accumulators", 1, 64, 1, 2, style);
    run("SS N=64, 2 issuers", 0, 64, 2, 1, style);
    run("SS N=128, 2 issuers", 0, 128, 2, 1, style);
    run("TS N=64, 2 issuers", 1, 64, 2, 1, style);
    run("SS N=64, 2 issuers x 2 accumulators", 0, 64, 2, 2, style);
    run("SS N=128, 2 issuers x 2 accumulators", 0, 128, 2, 2, style);
  }
  printf("--- lean loop (descriptors precomputed, converged warp)\n");
  run_lean<64, 1>("lean SS N=64, 1 issuer", 1);
  run_lean<128, 1>("lean SS N=128, 1 issuer", 1);
  run_lean<160, 1>("lean SS N=160, 1 issuer", 1);
  run_lean<256, 1>("lean SS N=256, 1 issuer", 1);
  run_lean<64, 2>("lean SS N=64, 1 issuer, 2 acc", 1);
  run_lean<128, 2>("lean SS N=128, 1 issuer, 2 acc", 1);
  run_lean<64, 1>("lean SS N=64, 2 issuers", 2);
  run_lean<128, 1>("lean SS N=128, 2 issuers", 2);
  return 0;
}
