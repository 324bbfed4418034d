// HBM-bound kernels of the SVD step: GroupNorm(+SiLU), LayerNorm, embedding MLPs, layout packers,
// nearest upsample, im2col gather, CFG + Euler update, and the DummyUNet step.
// All vectorised to 16-byte accesses where the layout allows, coalesced along the contiguous axis,
// fp32 statistics, deterministic (no floating-point atomics).
#include <cuda_fp16.h>
#include <math_constants.h>
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace svdpp {

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// SiLU with ONE MUFU op: x*sigmoid(x) = h*tanh(h) + h, h = x/2 (tanh.approx.f32, relative error 2^-11, below the
// fp16 rounding of the result).  The exp + reciprocal form costs two MUFU per element, and at 16 MUFU/clk/SM that
// - not HBM - bound the GroupNorm+SiLU apply pass (ncu: 63.6 us for 295 MB at level 0).
__device__ __forceinline__ float silu_fast(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&o)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    o[2 * i] = f.x;
    o[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 u;
  __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
  __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  u.z = *reinterpret_cast<uint32_t*>(&h2);
  u.w = *reinterpret_cast<uint32_t*>(&h3);
  return u;
}

// ------------------------------------------------------------------------------------ GroupNorm
// Thread mapping shared by the statistics and the apply kernel: blockDim = (C/8, rows); thread (tx, ty)
// owns the 8 channels [8*tx, 8*tx+8) and walks the pixels ty, ty+rows, ... of a 64-pixel chunk, so
// consecutive threads touch consecutive 16-byte vectors (also across pixel boundaries) and per-channel
// quantities (running sums, or scale/shift) live in registers.  grid = (chunks, images).
constexpr int GN_MAX_PIX = 64;         // pixels per block at high resolution; halved (down to 8) until the grid
constexpr int GN_MIN_PIX = 8;          // has >= 4 blocks per SM (low-resolution levels: 144 pixels x 1280 channels)
constexpr int GN_GROUPS = 32;
// The workspace starts with a fixed block of arrival counters (per image, then per statistics group), so calls
// with different (n_img, HW) that share one workspace never alias a counter with another call's partial sums.
constexpr int GN_MAX_STATS = 4096;
constexpr size_t GN_COUNTER_BYTES = 2 * GN_MAX_STATS * sizeof(unsigned);

__device__ __forceinline__ const __half* gn_src(const __half* x1, int C1, const __half* x2, int C2, long long pix,
                                                int c0) {
  return c0 < C1 ? x1 + pix * C1 + c0 : x2 + pix * C2 + (c0 - C1);
}

// Statistics pass.  Work item = (image n, chunk of `pix` pixels), items numbered image-major.  Persistent blocks
// (two per SM); block c owns the contiguous item range [c*ipc, (c+1)*ipc) and streams it through a two-slot
// shared-memory ring: a chunk is one contiguous range of each source, so thread 0 fetches the NEXT item with (at
// most two) bulk async copies on an mbarrier while the block adds up the current one.  Per-thread channel sums stay
// in registers across all chunks of an image; only when the image changes (once or twice per block) are they
// reduced to partial[n][slot][g] = (sum, sumsq), slot = block - first block of image n.  The last block to finish
// an image reduces that image's slots (fixed order, fp64); for the temporal GroupNorm (statistics over
// frames_per_stat images) the last image to finish reduces the per-image sums -> stats[st][g] = (mean, rstd).
// No finalize launch, no floating-point atomics (deterministic); counters are left at zero for the next call.
// History (ncu, level 0, 147 MB read): register-staged loads, one chunk per block, per-chunk reduce + fence +
// atomic: 55.9 us; one bulk copy per block: 50.1 us (a wave's blocks all fetch, then all compute).
__device__ __forceinline__ void gn_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__global__ void __launch_bounds__(512)
gn_partial_kernel(const __half* __restrict__ x1, int C1, const __half* __restrict__ x2, int C2, int HW, int pix,
                  int n_chunks, int n_items, int ipc, int n_slots, float2* partial, double2* imgsum, unsigned* counters,
                  int frames_per_stat, float count, float eps, float2* __restrict__ stats, int rev) {
  pdl_launch_dependents();  // programmatic dependent launch: see ptx.cuh
  pdl_wait();
  extern __shared__ __align__(128) uint8_t gn_smem[];
  __shared__ int s_last;
  __shared__ __align__(8) uint64_t s_bar[8];
  // rev: the blocks that are scheduled first take the LAST items (the end of the tensor, which the producing kernel
  // wrote last and which is therefore still in L2); the item -> partial-slot mapping is unchanged
  const int vblk = rev ? static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x) : static_cast<int>(blockIdx.x);
  const int begin = vblk * ipc;
  const int end = min(begin + ipc, n_items);
  if (begin >= end) return;
  const int C = C1 + C2;
  const int rows = blockDim.y;
  const int c0 = threadIdx.x * 8;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nthreads = blockDim.x * blockDim.y;
  const int max_slots = n_chunks + 1;
  const uint32_t slot_bytes = static_cast<uint32_t>(pix) * C * 2;
  float2* red = reinterpret_cast<float2*>(gn_smem + n_slots * slot_bytes);  // [rows][C]; later the fp64 finalize scratch
  const uint32_t smem0 = static_cast<uint32_t>(__cvta_generic_to_shared(gn_smem));
  const uint32_t bar0 = static_cast<uint32_t>(__cvta_generic_to_shared(&s_bar[0]));

  auto issue = [&](int item, int slot) {  // thread 0 only
    const int n = item / n_chunks, ch = item - n * n_chunks;
    const int p0 = ch * pix;
    const int npx = min(pix, HW - p0);
    const uint32_t b1 = static_cast<uint32_t>(npx) * C1 * 2, b2 = static_cast<uint32_t>(npx) * C2 * 2;
    const uint32_t bar = bar0 + slot * 8, dst = smem0 + slot * slot_bytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b1 + b2) : "memory");
    gn_bulk_load(dst, x1 + (static_cast<long long>(n) * HW + p0) * C1, b1, bar);
    if (C2 > 0) gn_bulk_load(dst + pix * C1 * 2, x2 + (static_cast<long long>(n) * HW + p0) * C2, b2, bar);
  };

  if (tid == 0) {
    for (int i = 0; i < n_slots; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + i * 8) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < n_slots - 1 && begin + i < end; ++i) issue(begin + i, i);  // ring prologue
  }
  __syncthreads();  // barriers initialised before anyone polls them

  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;

  // block-level reduction of the register sums of image n, partial write, arrival count, finalize by the last block
  auto flush = [&](int n) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.y * C + c0 + j] = make_float2(s[j], ss[j]);
    __syncthreads();
    const int first_blk = (n * n_chunks) / ipc;
    const int n_blk = ((n + 1) * n_chunks - 1) / ipc - first_blk + 1;  // blocks that hold a piece of image n
    {
      // group sums: LPG lanes per group walk the group's rows x channels entries in a fixed interleave and
      // combine with shuffles
      const int lpg = nthreads >= 128 ? 4 : (nthreads >= 64 ? 2 : 1);
      if (tid < GN_GROUPS * lpg) {
        const int g = tid / lpg, j = tid - g * lpg;
        const int cpg = C / GN_GROUPS;
        const int items = rows * cpg;
        float a = 0.f, b = 0.f;
        for (int it = j; it < items; it += lpg) {
          const int r = it / cpg, i = it - r * cpg;
          const float2 v = red[r * C + g * cpg + i];
          a += v.x;
          b += v.y;
        }
        for (int o = 1; o < lpg; o <<= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (j == 0) {
          const int slot = vblk - first_blk;
          partial[(static_cast<long long>(n) * max_slots + slot) * GN_GROUPS + g] = make_float2(a, b);
          __threadfence();
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      const unsigned old = atomicAdd(&counters[n], 1u);
      s_last = old == static_cast<unsigned>(n_blk - 1);
      if (s_last) counters[n] = 0u;
    }
    __syncthreads();
    if (s_last) {  // block-uniform: this block completed image n
      __threadfence();
      double a = 0.0, b = 0.0;
      if (tid < GN_GROUPS) {
        const float2* base = partial + static_cast<long long>(n) * max_slots * GN_GROUPS + tid;
        for (int i = 0; i < n_blk; ++i) {
          const float2 v = __ldcg(base + static_cast<long long>(i) * GN_GROUPS);
          a += v.x;
          b += v.y;
        }
      }
      const int st = n / frames_per_stat;
      bool emit = true;
      if (frames_per_stat > 1) {
        if (tid < GN_GROUPS) {
          imgsum[n * GN_GROUPS + tid] = make_double2(a, b);
          __threadfence();
        }
        __syncthreads();
        if (tid == 0) {
          const unsigned old = atomicAdd(&counters[GN_MAX_STATS + st], 1u);
          s_last = old == static_cast<unsigned>(frames_per_stat - 1);
          if (s_last) counters[GN_MAX_STATS + st] = 0u;
        }
        __syncthreads();
        emit = s_last != 0;
        if (emit) {
          __threadfence();
          if (tid < GN_GROUPS) {
            a = 0.0;
            b = 0.0;
            for (int f = 0; f < frames_per_stat; ++f) {
              const double2 v = __ldcg(imgsum + (static_cast<long long>(st) * frames_per_stat + f) * GN_GROUPS + tid);
              a += v.x;
              b += v.y;
            }
          }
        }
      }
      if (emit && tid < GN_GROUPS) {
        const double mean = a / count;
        double var = b / count - mean * mean;
        if (var < 0.0) var = 0.0;
        stats[st * GN_GROUPS + tid] = make_float2(static_cast<float>(mean),
                                                  static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps))));
      }
    }
    __syncthreads();  // scratch and s_last are free again
  };

  int slot = 0;
  uint32_t phases = 0u;  // bit i = parity to wait for on slot i
  int cur = begin / n_chunks;
  for (int item = begin; item < end; ++item) {
    const int n = item / n_chunks, ch = item - n * n_chunks;
    if (n != cur) {
      flush(cur);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
      cur = n;
    }
    const int npx = min(pix, HW - ch * pix);
    // keep n_slots - 1 items in flight; the slot refilled here was consumed before the barrier that ended the
    // previous iteration
    if (tid == 0 && item + n_slots - 1 < end) {
      int ns = slot + n_slots - 1;
      if (ns >= n_slots) ns -= n_slots;
      issue(item + n_slots - 1, ns);
    }
    {
      const uint32_t bar = bar0 + slot * 8;
      const uint32_t ph = (phases >> slot) & 1u;
      uint32_t ok = 0;
      while (!ok)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 1000000;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(bar), "r"(ph)
            : "memory");
      phases ^= 1u << slot;
    }
    const __half* s1 = reinterpret_cast<const __half*>(gn_smem + slot * slot_bytes);  // [npx][C1]
    const __half* s2 = s1 + static_cast<size_t>(pix) * C1;                             // [npx][C2]
    const __half* src = c0 < C1 ? s1 + c0 : s2 + (c0 - C1);
    const int pitch = c0 < C1 ? C1 : C2;
#pragma unroll 4
    for (int p = threadIdx.y; p < npx; p += rows) {
      float v[8];
      unpack8(*reinterpret_cast<const uint4*>(src + static_cast<size_t>(p) * pitch), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += v[j];
        ss[j] += v[j] * v[j];
      }
    }
    __syncthreads();  // slot consumed
    if (++slot == n_slots) slot = 0;
  }
  flush(cur);
}

__global__ void __launch_bounds__(512)
gn_apply_kernel(const __half* __restrict__ x1, int C1, const __half* __restrict__ x2, int C2,
                const __half* __restrict__ gamma, const __half* __restrict__ beta, const float2* __restrict__ stats,
                __half* __restrict__ out, int HW, int pix, int frames_per_stat, int silu, int rev) {
  pdl_launch_dependents();  // programmatic dependent launch: see ptx.cuh
  pdl_wait();
  const int C = C1 + C2;
  const int rows = blockDim.y;
  const int c0 = threadIdx.x * 8;
  const int n = rev ? static_cast<int>(gridDim.y) - 1 - static_cast<int>(blockIdx.y) : static_cast<int>(blockIdx.y);
  const int p0 = (rev ? static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x) : static_cast<int>(blockIdx.x)) * pix;
  const int p1 = min(p0 + pix, HW);
  const int cpg = C / GN_GROUPS;
  const float2* st = stats + (n / frames_per_stat) * GN_GROUPS;
  float sc[8], sh[8];
  {
    float g[8], b[8];
    unpack8(*reinterpret_cast<const uint4*>(gamma + c0), g);
    unpack8(*reinterpret_cast<const uint4*>(beta + c0), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 mr = st[(c0 + j) / cpg];
      sc[j] = mr.y * g[j];
      sh[j] = b[j] - mr.x * sc[j];
    }
  }
  for (int pb = p0 + threadIdx.y; pb < p1; pb += 4 * rows) {
    uint4 u[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {  // four independent 16-byte loads in flight
      const int p = pb + t * rows;
      if (p < p1) u[t] = *reinterpret_cast<const uint4*>(gn_src(x1, C1, x2, C2, static_cast<long long>(n) * HW + p, c0));
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int p = pb + t * rows;
      if (p < p1) {
        float v[8];
        unpack8(u[t], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float y = fmaf(v[j], sc[j], sh[j]);
          v[j] = silu ? silu_fast(y) : y;
        }
        *reinterpret_cast<uint4*>(out + (static_cast<long long>(n) * HW + p) * C + c0) = pack8(v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ LayerNorm
constexpr int LN_MAX_VEC = 5;  // per lane: 5 * 8 * 32 = 1280 channels max

__global__ void __launch_bounds__(256)
layernorm_kernel(const __half* __restrict__ x, long long ldx, const __half* __restrict__ addvec, int add_hw,
                 int add_mod, const __half* __restrict__ gamma, const __half* __restrict__ beta,
                 __half* __restrict__ out, long long ldo, int M, int C, float eps, int rev) {
  pdl_launch_dependents();  // programmatic dependent launch: see ptx.cuh
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m = static_cast<long long>(rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * 8 + warp;
  if (m >= M) return;
  const int nvec = C >> 3;
  const __half* xr = x + m * ldx;
  const __half* ar = addvec ? addvec + static_cast<long long>((m / add_hw) % add_mod) * C : nullptr;
  float v[LN_MAX_VEC][8];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < LN_MAX_VEC; ++k) {
    const int vi = lane + k * 32;
    if (vi < nvec) {
      unpack8(*reinterpret_cast<const uint4*>(xr + vi * 8), v[k]);
      if (ar) {
        float a[8];
        unpack8(*reinterpret_cast<const uint4*>(ar + vi * 8), a);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[k][j] = __half2float(__float2half_rn(v[k][j] + a[j]));  // fp16 add, as torch
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[k][j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / C;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < LN_MAX_VEC; ++k) {
    const int vi = lane + k * 32;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[k][j] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = 1.0f / sqrtf(sq / C + eps);
#pragma unroll
  for (int k = 0; k < LN_MAX_VEC; ++k) {
    const int vi = lane + k * 32;
    if (vi < nvec) {
      float g[8], b[8], y[8];
      unpack8(*reinterpret_cast<const uint4*>(gamma + vi * 8), g);
      unpack8(*reinterpret_cast<const uint4*>(beta + vi * 8), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = (v[k][j] - mean) * rstd * g[j] + b[j];
      *reinterpret_cast<uint4*>(out + m * ldo + vi * 8) = pack8(y);
    }
  }
}

// C = 40 * LPR channels (320 / 640 / 1280): LPR lanes per row, 5 vectors per lane, 32/LPR rows per warp, so
// every lane has 5 independent 16-byte loads in flight (the generic kernel leaves 3/4 of the lanes idle on
// the second vector at C = 320).
template <int LPR>
__global__ void __launch_bounds__(256)
layernorm_rows_kernel(const __half* __restrict__ x, long long ldx, const __half* __restrict__ addvec, int add_hw,
                      int add_mod, const __half* __restrict__ gamma, const __half* __restrict__ beta,
                      __half* __restrict__ out, long long ldo, int M, float eps, int rev) {
  pdl_launch_dependents();  // programmatic dependent launch: see ptx.cuh
  pdl_wait();
  constexpr int RPW = 32 / LPR;
  constexpr int C = 40 * LPR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPR, li = lane % LPR;
  const long long m = (static_cast<long long>(rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * 8 + warp) * RPW + sub;
  const bool active = m < M;
  const __half* xr = x + (active ? m : 0) * ldx;
  const __half* ar = addvec ? addvec + static_cast<long long>(((active ? m : 0) / add_hw) % add_mod) * C : nullptr;
  float v[5][8];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) unpack8(*reinterpret_cast<const uint4*>(xr + (li + k * LPR) * 8), v[k]);
  if (ar) {
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      float a[8];
      unpack8(*reinterpret_cast<const uint4*>(ar + (li + k * LPR) * 8), a);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[k][j] = __half2float(__float2half_rn(v[k][j] + a[j]));  // fp16 add, as torch
    }
  }
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += v[k][j];
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / C;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = v[k][j] - mean;
      sq += d * d;
    }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = 1.0f / sqrtf(sq / C + eps);
  if (active) {
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int c0 = (li + k * LPR) * 8;
      float g[8], b[8], y[8];
      unpack8(*reinterpret_cast<const uint4*>(gamma + c0), g);
      unpack8(*reinterpret_cast<const uint4*>(beta + c0), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = (v[k][j] - mean) * rstd * g[j] + b[j];
      *reinterpret_cast<uint4*>(out + m * ldo + c0) = pack8(y);
    }
  }
}

// ------------------------------------------------------------------------------------ small linear
// One warp per output feature, 16-byte loads along K, up to 8 rows at a time.
__global__ void __launch_bounds__(128)
linear_small_kernel(const __half* __restrict__ x, const __half* __restrict__ x_add, long long ldx,
                    const __half* __restrict__ W, long long ldw,
                    const __half* __restrict__ bias, __half* __restrict__ y, long long ldy, int R, int N, int K,
                    int act_in, int act_out) {
  pdl_launch_dependents();  // programmatic dependent launch: see ptx.cuh
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 4 + warp;
  if (n >= N) return;
  const __half* wr = W + static_cast<long long>(n) * ldw;
  for (int r0 = 0; r0 < R; r0 += 8) {
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
    for (int k8 = lane; k8 < (K >> 3); k8 += 32) {
      float w[8];
      unpack8(*reinterpret_cast<const uint4*>(wr + 8 * k8), w);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (r0 + r < R) {
          float a[8];
          unpack8(*reinterpret_cast<const uint4*>(x + static_cast<long long>(r0 + r) * ldx + 8 * k8), a);
          if (x_add != nullptr) {
            float a2[8];
            unpack8(*reinterpret_cast<const uint4*>(x_add + static_cast<long long>(r0 + r) * ldx + 8 * k8), a2);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = __half2float(__float2half_rn(a[j] + a2[j]));  // fp16 add, as torch
          }
          if (act_in == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = __half2float(__float2half_rn(silu_f(a[j])));
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[r] = fmaf(a[j], w[j], acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
    }
    if (lane == 0) {
      const float b = bias ? __half2float(bias[n]) : 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (r0 + r < R) {
          float v = acc[r] + b;
          if (act_out == 1) v = silu_f(__half2float(__float2half_rn(v)));
          y[static_cast<long long>(r0 + r) * ldy + n] = __float2half_rn(v);
        }
      }
    }
  }
}

// Many independent small linears in one launch (the 1-token cross-attention output projections of all
// 32 transformer blocks): group = blockIdx.y, y[r, y_off + n] = x[r, x_off : x_off + K] . W[n, :] + bias[n].
struct SmallGroup {
  const __half* W;     // [N, K], K contiguous
  const __half* bias;  // [N] or null
  int x_off, y_off, N, K;
};
static_assert(sizeof(SmallGroup) == sizeof(svdpp_small_group), "ABI struct mismatch");

__global__ void __launch_bounds__(128)
linear_small_grouped_kernel(const __half* __restrict__ x, long long ldx, const SmallGroup* __restrict__ groups,
                            __half* __restrict__ y, long long ldy, int R) {
  pdl_launch_dependents();  // programmatic dependent launch: see ptx.cuh
  pdl_wait();
  const SmallGroup gr = groups[blockIdx.y];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 4 + warp;
  if (n >= gr.N) return;
  const __half* wr = gr.W + static_cast<long long>(n) * gr.K;
  const __half* xg = x + gr.x_off;
  for (int r0 = 0; r0 < R; r0 += 8) {
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
    for (int k8 = lane; k8 < (gr.K >> 3); k8 += 32) {
      float w[8];
      unpack8(*reinterpret_cast<const uint4*>(wr + 8 * k8), w);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (r0 + r < R) {
          float a[8];
          unpack8(*reinterpret_cast<const uint4*>(xg + static_cast<long long>(r0 + r) * ldx + 8 * k8), a);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[r] = fmaf(a[j], w[j], acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
    }
    if (lane == 0) {
      const float b = gr.bias ? __half2float(gr.bias[n]) : 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r0 + r < R) y[static_cast<long long>(r0 + r) * ldy + gr.y_off + n] = __float2half_rn(acc[r] + b);
    }
  }
}

// ------------------------------------------------------------------------------------ sinusoid
__global__ void sinusoid_kernel(const void* __restrict__ src, int src_kind, int src_mod, int n_vals, int dim,
                                __half* __restrict__ out) {
  pdl_launch_dependents();  // programmatic dependent launch: see ptx.cuh
  pdl_wait();
  const int half_dim = dim >> 1;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_vals * half_dim) return;
  const int i = idx / half_dim, k = idx % half_dim;
  float t;
  if (src_kind == 0)
    t = static_cast<const float*>(src)[i];
  else if (src_kind == 1)
    t = __half2float(static_cast<const __half*>(src)[i]);
  else
    t = static_cast<float>(i % src_mod);
  // diffusers get_timestep_embedding: exp(-ln(10000) * k / half_dim), fp32
  const float expo = (-9.210340371976184f * static_cast<float>(k)) / static_cast<float>(half_dim);
  const float arg = t * expf(expo);
  out[static_cast<long long>(i) * dim + k] = __float2half_rn(cosf(arg));             // flip_sin_to_cos: cos first
  out[static_cast<long long>(i) * dim + half_dim + k] = __float2half_rn(sinf(arg));
}

// ------------------------------------------------------------------------------------ upsample / im2col / packers
__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, long long n_vec_out, int H,
                                  int W, int vpc) {
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < n_vec_out;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(idx % vpc);
    long long pix = idx / vpc;
    const int wo = static_cast<int>(pix % (2 * W));
    pix /= (2 * W);
    const int ho = static_cast<int>(pix % (2 * H));
    const long long n = pix / (2 * H);
    out[idx] = x[((n * H + (ho >> 1)) * W + (wo >> 1)) * vpc + v];
  }
}

struct Taps {
  int8_t t[SVDPP_MAX_TAPS][4];
};

__global__ void im2col_kernel(const __half* __restrict__ x, __half* __restrict__ out, long long ldo, int F, int H,
                              int W, int C, int Ho, int Wo, int stride, int ntaps, Taps taps, long long n_vec) {
  pdl_launch_dependents();  // programmatic dependent launch: see ptx.cuh
  pdl_wait();
  const int vpr = static_cast<int>(ldo >> 3);
  const int K = ntaps * C;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < n_vec;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = idx / vpr;
    const int k0 = static_cast<int>(idx - m * vpr) << 3;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (k0 < K) {
      const int tap = k0 / C, c = k0 - tap * C;
      const int wo = static_cast<int>(m % Wo);
      long long r = m / Wo;
      const int ho = static_cast<int>(r % Ho);
      r /= Ho;
      const int f = static_cast<int>(r % F);
      const long long b = r / F;
      const int w = wo * stride + taps.t[tap][0], h = ho * stride + taps.t[tap][1], ff = f + taps.t[tap][2];
      if (w >= 0 && w < W && h >= 0 && h < H && ff >= 0 && ff < F)
        val = *reinterpret_cast<const uint4*>(x + ((((b * F + ff) * H + h) * W) + w) * C + c);
    }
    *reinterpret_cast<uint4*>(out + m * ldo + k0) = val;
  }
}

__global__ void pack_unet_input_kernel(const __half* __restrict__ s0, long long s0b, long long s0f, long long s0c,
                                       int C0, float in_div, const __half* __restrict__ s1, long long s1b,
                                       long long s1f, long long s1c, int C1, __half* __restrict__ out,
                                       int out_bfchw, int B, int F, int HW) {
  pdl_launch_dependents();  // programmatic dependent launch: see ptx.cuh
  pdl_wait();
  const long long total = static_cast<long long>(B) * F * HW;
  const int C = C0 + C1;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(idx % HW);
    const long long bf = idx / HW;
    const int f = static_cast<int>(bf % F);
    const long long b = bf / F;
    __half* o = out_bfchw ? out + bf * C * HW + p : out + idx * C;
    const long long oc = out_bfchw ? HW : 1;
    for (int c = 0; c < C0; ++c) {
      const float v = __half2float(s0[b * s0b + f * s0f + c * s0c + p]);
      o[c * oc] = __float2half_rn(__fdiv_rn(v, in_div));
    }
    for (int c = 0; c < C1; ++c) o[(C0 + c) * oc] = s1[b * s1b + f * s1f + c * s1c + p];
  }
}

__global__ void nhwc_to_bfchw_kernel(const __half* __restrict__ x, __half* __restrict__ out, long long total, int C,
                                     int HW) {
  pdl_launch_dependents();  // programmatic dependent launch: see ptx.cuh
  pdl_wait();
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(idx % HW);
    const long long bf = idx / HW;
    for (int c = 0; c < C; ++c) out[(bf * C + c) * HW + p] = x[idx * C + c];
  }
}

// ------------------------------------------------------------------------------------ CFG + Euler
__global__ void euler_vpred_kernel(const __half* __restrict__ latent, const __half* __restrict__ va,
                                   const __half* __restrict__ vc, const __half* __restrict__ gs, int v_nhwc,
                                   float c_v, float c_x, float sigma, float dt, __half* __restrict__ out, int B,
                                   int C, int F, int HW, unsigned* done_counter, unsigned* ready_flag,
                                   unsigned flag_value) {
  pdl_launch_dependents();  // programmatic dependent launch: see ptx.cuh
  pdl_wait();
  const long long total = static_cast<long long>(B) * F * HW;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(idx % HW);
    const long long bf = idx / HW;
    const int f = static_cast<int>(bf % F);
    const long long b = bf / F;
    for (int c = 0; c < C; ++c) {
      const long long vi = v_nhwc ? idx * C + c : (bf * C + c) * HW + p;
      __half v = va[vi];
      if (vc != nullptr) {
        // fp16 arithmetic, one rounding per op, as torch does at svd_unet.py:411
        const __half d = __hsub_rn(vc[vi], v);   // _rn: no contraction of mul+add into an fma
        const __half e = __hmul_rn(gs[f], d);
        v = __hadd_rn(v, e);
      }
      const long long li = ((b * C + c) * F + f) * HW + p;
      const float x = __half2float(latent[li]);
      // svd_unet.py:428-437, every op rounded to fp32 separately (no FMA contraction)
      const float x0 = __fadd_rn(__fmul_rn(__half2float(v), c_v), __fdiv_rn(x, c_x));
      const float d = __fdiv_rn(__fsub_rn(x, x0), sigma);
      out[li] = __float2half_rn(__fadd_rn(x, __fmul_rn(d, dt)));
    }
  }
  // Stage-to-stage handoff fused into the producer: `out` may be a peer-mapped buffer on the next stage's GPU (the
  // stores above then travel over NVLink).  Once every block's stores are visible system-wide, the last block to finish
  // raises the consumer's flag with a release store; the consumer's svdpp_flag_wait kernel acquires it.
  if (ready_flag != nullptr) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned old = atomicAdd(done_counter, 1u);
      if (old == gridDim.x - 1) {
        *done_counter = 0u;   // re-armed for the next launch on this stream
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ready_flag), "r"(flag_value) : "memory");
      }
    }
  }
}

// one thread: spin until *flag == value (acquire, system scope: the flag is written by another GPU), then optionally
// re-arm it.  Bounded: after `timeout_ns` it reports and traps instead of hanging the stream for ever.
__global__ void flag_wait_kernel(unsigned* flag, unsigned value, int reset, unsigned reset_to, unsigned long long timeout_ns) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  unsigned v;
  unsigned spins = 0;
  while (true) {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v == value) break;
    if ((++spins & 1023u) == 0u) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) {
        printf("svdpp: flag_wait timeout: flag=%p value=%u wanted=%u\n", static_cast<void*>(flag), v, value);
        __trap();
      }
    }
    __nanosleep(200);
  }
  if (reset) *flag = reset_to;
}

__global__ void flag_set_kernel(unsigned* flag, unsigned value) {
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

// ------------------------------------------------------------------------------------ DummyUNet
// conv3d 3x3x3 pad 1, fp32, [B, Cin, F, H, W] -> [B, Cout, F, H, W]; thread per (voxel, cout)
__global__ void dummy_conv3d_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                    const float* __restrict__ bias, float* __restrict__ y, int B, int Cin, int Cout,
                                    int F, int H, int W, int silu) {
  const long long vox = static_cast<long long>(F) * H * W;
  const long long total = static_cast<long long>(B) * Cout * vox;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int wq = static_cast<int>(idx % W);
  long long r = idx / W;
  const int h = static_cast<int>(r % H);
  r /= H;
  const int f = static_cast<int>(r % F);
  r /= F;
  const int co = static_cast<int>(r % Cout);
  const long long b = r / Cout;
  float acc = bias[co];
  for (int ci = 0; ci < Cin; ++ci) {
    const float* xp = x + (b * Cin + ci) * vox;
    const float* wp = w + (static_cast<long long>(co) * Cin + ci) * 27;
    for (int kf = 0; kf < 3; ++kf) {
      const int ff = f + kf - 1;
      if (ff < 0 || ff >= F) continue;
      for (int kh = 0; kh < 3; ++kh) {
        const int hh = h + kh - 1;
        if (hh < 0 || hh >= H) continue;
        for (int kw = 0; kw < 3; ++kw) {
          const int ww = wq + kw - 1;
          if (ww < 0 || ww >= W) continue;
          acc += xp[(static_cast<long long>(ff) * H + hh) * W + ww] * wp[(kf * 3 + kh) * 3 + kw];
        }
      }
    }
  }
  y[idx] = silu ? acc / (1.0f + expf(-acc)) : acc;
}

// out = x + scale * conv_out + LN_C(x); thread per voxel
__global__ void dummy_combine_kernel(const float* __restrict__ x, const float* conv,
                                     const float* __restrict__ g, const float* __restrict__ bta, float eps,
                                     float scale, float* out, int B, int C, long long vox) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(B) * vox) return;
  const long long b = idx / vox, v = idx % vox;
  float mean = 0.f;
  for (int c = 0; c < C; ++c) mean += x[(b * C + c) * vox + v];
  mean /= C;
  float var = 0.f;
  for (int c = 0; c < C; ++c) {
    const float d = x[(b * C + c) * vox + v] - mean;
    var += d * d;
  }
  var /= C;
  const float rstd = g ? 1.0f / sqrtf(var + eps) : 0.f;
  for (int c = 0; c < C; ++c) {
    const long long i = (b * C + c) * vox + v;
    float o = x[i] + scale * conv[i];
    if (g) o += (x[i] - mean) * rstd * g[c] + bta[c];
    out[i] = o;
  }
}

static inline unsigned grid_for(long long n, int threads, int max_blocks = 148 * 16) {
  long long b = (n + threads - 1) / threads;
  if (b > max_blocks) b = max_blocks;
  if (b < 1) b = 1;
  return static_cast<unsigned>(b);
}

}  // namespace svdpp

using namespace svdpp;

static int gn_pixels_per_block(int n_img, int HW, int C) {
  static const int stats_max = [] {
    const char* e = getenv("SVDPP_GN_STATS_PIX");
    return e ? atoi(e) : GN_MAX_PIX;
  }();
  int pix = stats_max;
  while (pix > GN_MIN_PIX && static_cast<long long>(pix) * C * 2 > 40 * 1024) pix >>= 1;  // ~40 KB of pixels per block
  while (pix > GN_MIN_PIX && static_cast<long long>(n_img) * ((HW + pix - 1) / pix) < 4LL * num_sms()) pix >>= 1;
  return pix;
}

extern "C" size_t svdpp_groupnorm_workspace_bytes(int32_t n_img, int32_t HW) {
  const size_t n_chunks = (static_cast<size_t>(HW) + GN_MIN_PIX - 1) / GN_MIN_PIX;  // worst case (smallest blocks)
  return GN_COUNTER_BYTES + static_cast<size_t>(n_img) * GN_GROUPS * sizeof(double2) +
         (static_cast<size_t>(n_img) * (n_chunks + 1) * GN_GROUPS + static_cast<size_t>(n_img) * GN_GROUPS) * sizeof(float2);
}

extern "C" int svdpp_groupnorm_silu(const void* x1, int32_t C1, const void* x2, int32_t C2, const void* gamma,
                                    const void* beta, void* out, int32_t n_img, int32_t HW, int32_t frames_per_stat,
                                    float eps, int32_t apply_silu, void* workspace, size_t workspace_bytes,
                                    svdpp_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (x2 == nullptr) C2 = 0;
  const int C = C1 + C2;
  SVDPP_CHECK_ARG(x1 && gamma && beta && out && workspace, "groupnorm: null pointer");
  SVDPP_CHECK_ARG(C % GN_GROUPS == 0 && C1 % 8 == 0 && C2 % 8 == 0, "groupnorm: C1=%d C2=%d unsupported", C1, C2);
  SVDPP_CHECK_ARG(C / 8 <= 512, "groupnorm: C=%d too large", C);
  SVDPP_CHECK_ARG(frames_per_stat >= 1 && n_img % frames_per_stat == 0, "groupnorm: frames_per_stat=%d", frames_per_stat);
  SVDPP_CHECK_ARG(workspace_bytes >= svdpp_groupnorm_workspace_bytes(n_img, HW), "groupnorm: workspace too small");
  SVDPP_CHECK_ARG(n_img <= GN_MAX_STATS, "groupnorm: more than %d images", GN_MAX_STATS);
  const int pix = gn_pixels_per_block(n_img, HW, C);
  const int n_chunks = (HW + pix - 1) / pix;
  // traversal direction (svdpp_set_tuning("reverse")): the statistics pass starts at the end of the tensor the
  // producer has just written when rev = 1, and the apply pass then runs the other way - it starts where the
  // statistics pass stopped, i.e. on the part of the input that is hot in L2 now
  // "reverse": 0 = both passes forward; 1 = statistics from the end, apply forward; 2 = statistics forward, apply from the end
  const int rev_mode = tuning().reverse;
  const int rev = rev_mode == 1 ? 1 : 0;
  int rev_apply = rev_mode == 2 ? 1 : 0;
  if (rev_mode != 0 && tuning().reverse_gn_apply_same) rev_apply ^= 1;
  unsigned* counters = static_cast<unsigned*>(workspace);
  double2* imgsum = reinterpret_cast<double2*>(static_cast<uint8_t*>(workspace) + GN_COUNTER_BYTES);
  float2* partial = reinterpret_cast<float2*>(imgsum + static_cast<size_t>(n_img) * GN_GROUPS);
  float2* stats = partial + static_cast<size_t>(n_img) * (n_chunks + 1) * GN_GROUPS;
  const int nvc = C / 8;
  int rows = 256 / nvc;
  if (rows < 1) rows = 1;
  if (rows > pix) rows = pix;
  while (nvc * rows < 32) ++rows;  // the in-kernel reductions need at least one full warp
  dim3 block(nvc, rows);
  dim3 grid(n_chunks, n_img);
  const size_t red_bytes = static_cast<size_t>(rows) * C * sizeof(float2);   // reduction scratch / fp64 finalize
  static const int slots_env = [] {
    const char* e = getenv("SVDPP_GN_SLOTS");
    return e ? atoi(e) : 0;
  }();
  // ring depth: as many pixel slots as fit in ~80 KB (two blocks per SM), at least 2, at most 8
  int n_slots = slots_env ? slots_env : static_cast<int>((80 * 1024) / (static_cast<size_t>(pix) * C * 2));
  if (n_slots < 2) n_slots = 2;
  if (n_slots > 8) n_slots = 8;
  const size_t smem_bytes = static_cast<size_t>(n_slots) * pix * C * 2 + red_bytes;
  static bool configured = false;
  if (!configured) {
    SVDPP_CUDA(cudaFuncSetAttribute(gn_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    configured = true;
  }
  SVDPP_CHECK_ARG(smem_bytes <= 112 * 1024, "groupnorm: shared-memory tile too large");
  const float count = static_cast<float>(frames_per_stat) * HW * (C / GN_GROUPS);
  const int n_items = n_chunks * n_img;
  int stat_blocks = 2 * num_sms();  // persistent: two blocks per SM, each prefetching its next chunk
  if (stat_blocks > n_items) stat_blocks = n_items;
  const int ipc = (n_items + stat_blocks - 1) / stat_blocks;  // contiguous items per block
  launch_kernel(gn_partial_kernel, dim3(stat_blocks), dim3(block), smem_bytes, stream, 1, static_cast<const __half*>(x1), C1,
                                                                static_cast<const __half*>(x2), C2, HW, pix, n_chunks,
                                                                n_items, ipc, n_slots, partial, imgsum, counters,
                                                                frames_per_stat, count, eps, stats, rev);
  if (int e = check_launch("gn_partial_kernel")) return e;
  // the apply pass may use larger blocks than the statistics pass (its per-block prologue - gamma, beta, statistics -
  // amortises better); SVDPP_GN_APPLY_PIX overrides for experiments
  static const int apply_max = [] {
    const char* e = getenv("SVDPP_GN_APPLY_PIX");
    return e ? atoi(e) : 64;
  }();
  int apix = pix;
  while (apix < apply_max && static_cast<long long>(n_img) * ((HW + 2 * apix - 1) / (2 * apix)) >= 4LL * num_sms()) apix <<= 1;
  dim3 grid_apply((HW + apix - 1) / apix, n_img);
  launch_kernel(gn_apply_kernel, dim3(grid_apply), dim3(block), 0, stream, 1, static_cast<const __half*>(x1), C1, static_cast<const __half*>(x2), C2,
                                              static_cast<const __half*>(gamma), static_cast<const __half*>(beta),
                                              stats, static_cast<__half*>(out), HW, apix, frames_per_stat, apply_silu,
                                              rev_apply);
  return check_launch("gn_apply_kernel");
}

extern "C" int svdpp_layernorm(const void* x, int64_t ldx, const void* addvec, int32_t add_hw, int32_t add_mod,
                               const void* gamma, const void* beta, void* out, int64_t ldo, int32_t M, int32_t C,
                               float eps, svdpp_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SVDPP_CHECK_ARG(x && gamma && beta && out, "layernorm: null pointer");
  SVDPP_CHECK_ARG(C % 8 == 0 && C <= LN_MAX_VEC * 256, "layernorm: C=%d unsupported", C);
  SVDPP_CHECK_ARG(ldx % 8 == 0 && ldo % 8 == 0, "layernorm: pitches must be multiples of 8");
  if (addvec) SVDPP_CHECK_ARG(add_hw > 0 && add_mod > 0, "layernorm: bad addvec indexing");
  const __half* xh = static_cast<const __half*>(x);
  const __half* ah = static_cast<const __half*>(addvec);
  const __half* gh = static_cast<const __half*>(gamma);
  const __half* bh = static_cast<const __half*>(beta);
  __half* oh = static_cast<__half*>(out);
  const int hw = add_hw > 0 ? add_hw : 1, md = add_mod > 0 ? add_mod : 1;
  const int rev = tuning().reverse == 1 ? 1 : 0;
  if (C == 320)
    launch_kernel(layernorm_rows_kernel<8>, dim3((M + 31) / 32), dim3(256), 0, stream, 1, xh, ldx, ah, hw, md, gh, bh, oh, ldo, M, eps, rev);
  else if (C == 640)
    launch_kernel(layernorm_rows_kernel<16>, dim3((M + 15) / 16), dim3(256), 0, stream, 1, xh, ldx, ah, hw, md, gh, bh, oh, ldo, M, eps, rev);
  else if (C == 1280)
    launch_kernel(layernorm_rows_kernel<32>, dim3((M + 7) / 8), dim3(256), 0, stream, 1, xh, ldx, ah, hw, md, gh, bh, oh, ldo, M, eps, rev);
  else
    launch_kernel(layernorm_kernel, dim3((M + 7) / 8), dim3(256), 0, stream, 1, xh, ldx, ah, hw, md, gh, bh, oh, ldo, M, C, eps, rev);
  return check_launch("layernorm_kernel");
}

extern "C" int svdpp_linear_small(const void* x, const void* x_add, int64_t ldx, const void* W, int64_t ldw, const void* bias, void* y,
                                  int64_t ldy, int32_t R, int32_t N, int32_t K, int32_t act_in, int32_t act_out,
                                  svdpp_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SVDPP_CHECK_ARG(x && W && y, "linear_small: null pointer");
  SVDPP_CHECK_ARG(K % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0, "linear_small: K and pitches must be multiples of 8");
  SVDPP_CHECK_ARG(R >= 1 && N >= 1, "linear_small: bad shape");
  launch_kernel(linear_small_kernel, dim3((N + 3) / 4), dim3(128), 0, stream, 1, static_cast<const __half*>(x),
                                                       static_cast<const __half*>(x_add), ldx,
                                                       static_cast<const __half*>(W), ldw,
                                                       static_cast<const __half*>(bias), static_cast<__half*>(y), ldy,
                                                       R, N, K, act_in, act_out);
  return check_launch("linear_small_kernel");
}

extern "C" int svdpp_linear_small_grouped(const void* x, int64_t ldx, const svdpp_small_group* groups_dev,
                                          int32_t n_groups, int32_t max_n, void* y, int64_t ldy, int32_t R,
                                          svdpp_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SVDPP_CHECK_ARG(x && groups_dev && y, "linear_small_grouped: null pointer");
  SVDPP_CHECK_ARG(n_groups >= 1 && n_groups <= 65535 && max_n >= 1 && R >= 1, "linear_small_grouped: bad shape");
  SVDPP_CHECK_ARG(ldx % 8 == 0, "linear_small_grouped: ldx must be a multiple of 8");
  dim3 grid((max_n + 3) / 4, n_groups);
  launch_kernel(linear_small_grouped_kernel, dim3(grid), dim3(128), 0, stream, 1, static_cast<const __half*>(x), ldx,
                                                        reinterpret_cast<const SmallGroup*>(groups_dev),
                                                        static_cast<__half*>(y), ldy, R);
  return check_launch("linear_small_grouped_kernel");
}

extern "C" int svdpp_sinusoid_embed(const void* src, int32_t src_kind, int32_t src_mod, int32_t n_vals, int32_t dim,
                                    void* out, svdpp_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SVDPP_CHECK_ARG(out && dim % 2 == 0 && n_vals > 0, "sinusoid: bad arguments");
  SVDPP_CHECK_ARG(src_kind == 2 ? src_mod > 0 : src != nullptr, "sinusoid: bad source");
  const int total = n_vals * (dim / 2);
  launch_kernel(sinusoid_kernel, dim3((total + 127) / 128), dim3(128), 0, stream, 1, src, src_kind, src_mod, n_vals, dim,
                                                           static_cast<__half*>(out));
  return check_launch("sinusoid_kernel");
}

extern "C" int svdpp_upsample2x_nhwc(const void* x, void* out, int32_t n_img, int32_t H, int32_t W, int32_t C,
                                     svdpp_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SVDPP_CHECK_ARG(x && out && C % 8 == 0, "upsample: bad arguments");
  const long long n_vec = static_cast<long long>(n_img) * 4 * H * W * (C / 8);
  upsample2x_kernel<<<grid_for(n_vec, 256), 256, 0, stream>>>(static_cast<const uint4*>(x), static_cast<uint4*>(out),
                                                              n_vec, H, W, C / 8);
  return check_launch("upsample2x_kernel");
}

extern "C" int svdpp_im2col_nhwc(const void* x, void* out, int64_t ldo, int32_t B, int32_t F, int32_t H, int32_t W,
                                 int32_t C, int32_t Ho, int32_t Wo, int32_t stride, int32_t ntaps, const int8_t* taps,
                                 svdpp_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SVDPP_CHECK_ARG(x && out && taps, "im2col: null pointer");
  SVDPP_CHECK_ARG(C % 8 == 0 && ldo % 8 == 0 && ldo >= static_cast<int64_t>(ntaps) * C, "im2col: bad C/ldo");
  SVDPP_CHECK_ARG(ntaps >= 1 && ntaps <= SVDPP_MAX_TAPS, "im2col: ntaps=%d", ntaps);
  Taps t{};
  for (int i = 0; i < ntaps; ++i)
    for (int j = 0; j < 4; ++j) t.t[i][j] = taps[i * 4 + j];
  const long long n_vec = static_cast<long long>(B) * F * Ho * Wo * (ldo / 8);
  launch_kernel(im2col_kernel, dim3(grid_for(n_vec, 256)), dim3(256), 0, stream, 1, static_cast<const __half*>(x), static_cast<__half*>(out),
                                                          ldo, F, H, W, C, Ho, Wo, stride, ntaps, t, n_vec);
  return check_launch("im2col_kernel");
}

extern "C" int svdpp_pack_unet_input(const void* src0, int64_t s0_b, int64_t s0_f, int64_t s0_c, int32_t C0,
                                     float in_div, const void* src1, int64_t s1_b, int64_t s1_f, int64_t s1_c,
                                     int32_t C1, void* out, int32_t out_bfchw, int32_t B, int32_t F, int32_t H,
                                     int32_t W, svdpp_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (src1 == nullptr) C1 = 0;
  SVDPP_CHECK_ARG(src0 && out && C0 > 0, "pack_unet_input: bad arguments");
  const long long total = static_cast<long long>(B) * F * H * W;
  launch_kernel(pack_unet_input_kernel, dim3(grid_for(total, 256)), dim3(256), 0, stream, 1, 
      static_cast<const __half*>(src0), s0_b, s0_f, s0_c, C0, in_div, static_cast<const __half*>(src1), s1_b, s1_f,
      s1_c, C1, static_cast<__half*>(out), out_bfchw, B, F, H * W);
  return check_launch("pack_unet_input_kernel");
}

extern "C" int svdpp_nhwc_to_bfchw(const void* x, void* out, int32_t B, int32_t F, int32_t C, int32_t H, int32_t W,
                                   svdpp_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SVDPP_CHECK_ARG(x && out, "nhwc_to_bfchw: null pointer");
  const long long total = static_cast<long long>(B) * F * H * W;
  launch_kernel(nhwc_to_bfchw_kernel, dim3(grid_for(total, 256)), dim3(256), 0, stream, 1, static_cast<const __half*>(x),
                                                                 static_cast<__half*>(out), total, C, H * W);
  return check_launch("nhwc_to_bfchw_kernel");
}

extern "C" int svdpp_euler_vpred_step_signal(const void* latent, const void* v_a, const void* v_cond, const void* gs,
                                             int32_t v_nhwc, float c_v, float c_x, float sigma, float dt, void* out,
                                             int32_t B, int32_t C, int32_t F, int32_t H, int32_t W,
                                             const svdpp_handoff* ho, svdpp_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SVDPP_CHECK_ARG(latent && v_a && out, "euler: null pointer");
  SVDPP_CHECK_ARG(v_cond == nullptr || gs != nullptr, "euler: guidance needs gs");
  SVDPP_CHECK_ARG(ho == nullptr || ho->ready_flag == nullptr || ho->done_counter != nullptr,
                  "euler: a handoff needs a (zeroed, local) completion counter");
  const long long total = static_cast<long long>(B) * F * H * W;
  launch_kernel(euler_vpred_kernel, dim3(grid_for(total, 256)), dim3(256), 0, stream, 1, 
      static_cast<const __half*>(latent), static_cast<const __half*>(v_a), static_cast<const __half*>(v_cond),
      static_cast<const __half*>(gs), v_nhwc, c_v, c_x, sigma, dt, static_cast<__half*>(out), B, C, F, H * W,
      ho ? static_cast<unsigned*>(ho->done_counter) : nullptr, ho ? static_cast<unsigned*>(ho->ready_flag) : nullptr,
      ho ? ho->flag_value : 0u);
  return check_launch("euler_vpred_kernel");
}

extern "C" int svdpp_euler_vpred_step(const void* latent, const void* v_a, const void* v_cond, const void* gs,
                                      int32_t v_nhwc, float c_v, float c_x, float sigma, float dt, void* out,
                                      int32_t B, int32_t C, int32_t F, int32_t H, int32_t W, svdpp_stream stream_) {
  return svdpp_euler_vpred_step_signal(latent, v_a, v_cond, gs, v_nhwc, c_v, c_x, sigma, dt, out, B, C, F, H, W, nullptr,
                                       stream_);
}

extern "C" int svdpp_flag_wait(void* flag, uint32_t value, int32_t reset, uint32_t reset_to, int32_t timeout_s,
                               svdpp_stream stream_) {
  SVDPP_CHECK_ARG(flag != nullptr, "flag_wait: null flag");
  const unsigned long long ns = static_cast<unsigned long long>(timeout_s > 0 ? timeout_s : 600) * 1000000000ull;
  flag_wait_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream_)>>>(static_cast<unsigned*>(flag), value, reset, reset_to, ns);
  return check_launch("flag_wait_kernel");
}

extern "C" int svdpp_flag_set(void* flag, uint32_t value, svdpp_stream stream_) {
  SVDPP_CHECK_ARG(flag != nullptr, "flag_set: null flag");
  flag_set_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream_)>>>(static_cast<unsigned*>(flag), value);
  return check_launch("flag_set_kernel");
}

extern "C" int svdpp_dummy_unet_step(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                                     const float* ln_g, const float* ln_b, float ln_eps, float tanh_scale,
                                     float* hidden_ws, float* out, int32_t B, int32_t C, int32_t Ch, int32_t F,
                                     int32_t H, int32_t W, svdpp_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SVDPP_CHECK_ARG(x && w1 && b1 && w2 && b2 && hidden_ws && out, "dummy_unet: null pointer");
  const long long vox = static_cast<long long>(F) * H * W;
  const long long n1 = static_cast<long long>(B) * Ch * vox;
  dummy_conv3d_kernel<<<static_cast<unsigned>((n1 + 255) / 256), 256, 0, stream>>>(x, w1, b1, hidden_ws, B, C, Ch, F, H,
                                                                                   W, 1);
  if (int e = check_launch("dummy_conv3d_kernel")) return e;
  // second conv writes into `out`, then the combine kernel rewrites `out` in place
  const long long n2 = static_cast<long long>(B) * C * vox;
  dummy_conv3d_kernel<<<static_cast<unsigned>((n2 + 255) / 256), 256, 0, stream>>>(hidden_ws, w2, b2, out, B, Ch, C, F,
                                                                                   H, W, 0);
  if (int e = check_launch("dummy_conv3d_kernel")) return e;
  dummy_combine_kernel<<<static_cast<unsigned>((B * vox + 255) / 256), 256, 0, stream>>>(x, out, ln_g, ln_b, ln_eps,
                                                                                         tanh_scale, out, B, C, vox);
  return check_launch("dummy_combine_kernel");
}
