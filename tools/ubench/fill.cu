// Does the fill of a k-block (TMA boxes of 128-byte rows, L2-resident source) compete with the tcgen05.mma operand reads
// for the shared-memory port, and how fast can one SM / the whole chip fill?  ("What bounds a narrow tile".)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I video-diffusion-pipeline-parallel_b200/csrc -o tools/ubench/fill tools/ubench/fill.cu
// One CTA per SM, three roles in separate warps:
//   warp 1  MMA issuer: 4 x tcgen05.mma (M = 128, N = 128, K = 16, A and B from two fixed 16 KB smem tiles) per iteration
//           = one k-block of the 128x128 tile, 256 clocks at the tensor-pipe rate
//   warp 2  TMA producer: per iteration one A box (128 rows x 128 B) and one B box (rows_b rows x 128 B) into a ring of
//           2 / 4 / 6 stages, each box row a separate 128-byte chunk of a 256-byte-pitch matrix (a C = 128 channels-last
//           activation); the source (64 MB) stays in L2
// mode 1: MMAs only; 2: fills only; 3: both.  Prints clocks per iteration of each role, for 1 CTA and for 148.
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace svdpp;

constexpr int STAGE_BYTES = 32768;
constexpr int MAX_STAGES = 6;
constexpr int SMEM_BYTES = 32768 + MAX_STAGES * STAGE_BYTES + 1024;

template <int STAGES>
__global__ void __launch_bounds__(128, 1) fill_bench(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                                                     int mode, int rows_b, int a_split, int iters, int row_blocks, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[STAGES];
  __shared__ uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    mbar_init(&mma_bar, 1);
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
  if (warp == 1 && (mode & 1)) {
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 16384);
    constexpr uint32_t idesc = make_idesc_f16(128, false);
    uint64_t da[4], db[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      da[k] = make_smem_desc_sw128(a_addr + k * 32, 1024, 0);
      db[k] = make_smem_desc_sw128(b_addr + k * 32, 1024, 0);
    }
    const bool leader = elect_one();
    // with fills running beside it the issuer keeps going for 4 x iters so that every fill is overlapped by MMAs; the
    // reported time is that of the first `iters` iterations (the issue queue is shallow, so issue time = execution time)
    const int total = (mode & 2) ? 4 * iters : iters;
    const long long t0 = clock64();
    long long t_first = 0;
    for (int it = 0; it < total; ++it) {
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tm + (it & 1) * 128, da[k], db[k], idesc, 1u);
        if (it == iters - 1) t_first = clock64() - t0;
      }
      __syncwarp();
    }
    if (leader) {
      umma_commit(&mma_bar);
      mbar_wait(&mma_bar, 0, 1);
      cycles[blockIdx.x * 2 + 0] = (mode & 2) ? t_first : clock64() - t0;
    }
    __syncwarp();
  }
  // producers: warp 2 alone (n_prod = 1), or warps 2 and 3 taking alternate iterations (n_prod = 2: disjoint stages).
  // lean = 1: coordinates advance by an add and a compare (no division in the single issuing thread's dependency chain)
  const int n_prod = (mode >> 2) & 1 ? 2 : 1, lean = (mode >> 3) & 1;
  if (warp >= 2 && warp < 2 + n_prod && (mode & 2)) {
    if (elect_one()) {
      const int me = warp - 2;
      const uint32_t tx = 16384u + static_cast<uint32_t>(rows_b) * 128u;
      const int a_rows = 128 / a_split;
      int r = static_cast<int>((static_cast<unsigned>(blockIdx.x) * 977u + static_cast<unsigned>(me) * 131u) % static_cast<unsigned>(row_blocks));
      const int step = 131 * n_prod;
      const long long t0 = clock64();
      for (int it = me; it < iters + STAGES; it += n_prod) {
        const int s = it % STAGES;
        if (it >= STAGES) mbar_wait(&full[s], ((it / STAGES) - 1) & 1, 2);      // the previous fill of this stage has landed
        if (it >= iters) continue;
        uint8_t* dst = smem + 32768 + s * STAGE_BYTES;
        int r0, rb0;
        if (lean) {
          r += step;
          if (r >= row_blocks) r -= row_blocks;
          r0 = r * 128;
          rb0 = (r0 + 1031 * 128) & (row_blocks * 128 - 1);                      // row_blocks is a power of two
        } else {
          r0 = static_cast<int>((static_cast<unsigned>(blockIdx.x) * 977u + static_cast<unsigned>(it) * 131u) %
                                static_cast<unsigned>(row_blocks)) * 128;
          rb0 = static_cast<int>((static_cast<unsigned>(r0 / 128) + 1031u) % static_cast<unsigned>(row_blocks)) * 128;
        }
        mbar_expect_tx(&full[s], tx);
        for (int j2 = 0; j2 < a_split; ++j2) tma_load_2d(dst + j2 * a_rows * 128, &tm_a, &full[s], (it & 1) * 64, r0 + j2 * a_rows);
        if (rows_b > 0) tma_load_2d(dst + 16384, &tm_b, &full[s], (it & 1) * 64, rb0);
      }
      cycles[blockIdx.x * 2 + 1] = clock64() - t0;       // with two producers: whichever writes last (they finish together)
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

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(EncodeFn fn, CUtensorMap* m, void* base, uint64_t rows, uint32_t box_rows) {
  cuuint64_t gdim[2] = {128, rows};
  cuuint64_t gstr[1] = {256};
  cuuint32_t bdim[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? 0
             : 1;
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  EncodeFn fn = reinterpret_cast<EncodeFn>(p);
  const uint64_t rows = 262144;                       // x 256 B = 64 MB: stays in the 126 MB L2
  void* src;
  cudaMalloc(&src, rows * 256);
  cudaMemset(src, 0, rows * 256);
  long long* cyc;
  cudaMalloc(&cyc, 148 * 2 * sizeof(long long));
  cudaFuncSetAttribute(fill_bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  cudaFuncSetAttribute(fill_bench<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  cudaFuncSetAttribute(fill_bench<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  const int iters = 4096;
  printf("%-66s %10s %10s\n", "case (clocks per iteration = per k-block)", "MMA role", "fill role");
  struct Case { const char* name; int mode, rows_b, a_split; };
  // mode bits: 1 MMAs, 2 fills, 4 two producer warps, 8 lean coordinate arithmetic
  const Case cases[] = {
      {"MMAs only", 1, 0, 1},
      {"fills, 1 producer, division in the loop, A + B 16 KB", 2, 128, 1},
      {"fills, 1 producer, lean loop,            A + B 16 KB", 2 | 8, 128, 1},
      {"fills, 1 producer, lean loop,            A + B  8 KB", 2 | 8, 64, 1},
      {"fills, 1 producer, lean loop,            A only", 2 | 8, 0, 1},
      {"fills, 1 producer, lean loop,            A as 4 boxes only", 2 | 8, 0, 4},
      {"fills, 2 producers, lean loop,           A + B 16 KB", 2 | 4 | 8, 128, 1},
      {"fills, 2 producers, lean loop,           A + B  8 KB", 2 | 4 | 8, 64, 1},
      {"fills, 2 producers, lean loop,           A only", 2 | 4 | 8, 0, 1},
      {"both,  1 producer, lean loop,            A + B 16 KB", 3 | 8, 128, 1},
      {"both,  2 producers, lean loop,           A + B 16 KB", 3 | 4 | 8, 128, 1},
  };
  for (int blocks = 1; blocks <= 148; blocks += 147) {
    for (const Case& c : cases) {
      CUtensorMap ma, mb;
      if (make_map(fn, &ma, src, rows, 128 / c.a_split) || make_map(fn, &mb, src, rows, c.rows_b > 0 ? c.rows_b : 64)) {
        printf("tensor map failed\n");
        return 1;
      }
      cudaMemset(cyc, 0, 148 * 2 * sizeof(long long));
      for (int rep = 0; rep < 2; ++rep)
        fill_bench<4><<<blocks, 128, SMEM_BYTES>>>(ma, mb, c.mode, c.rows_b, c.a_split, iters, static_cast<int>(rows / 128), cyc);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[296];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double mma = 0, fill = 0;
      for (int i = 0; i < blocks; ++i) {
        mma += h[2 * i];
        fill += h[2 * i + 1];
      }
      printf("%3d CTAs, 4 stages, %-56s %10.1f %10.1f  (%s)\n", blocks, c.name, mma / blocks / iters, fill / blocks / iters,
             cudaGetErrorString(e));
      if (e != cudaSuccess) return 1;
    }
  }
  return 0;
}
