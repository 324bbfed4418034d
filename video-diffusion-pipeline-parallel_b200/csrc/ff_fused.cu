// svdpp_ff_geglu_f16: the transformer feed-forward  y = epilogue( GEGLU(x W1^T + b1) W2^T + b2 )  as ONE kernel, so that the
// [M, 4C] gated intermediate never reaches HBM (level 0: 590 MB written and read again per call, 15 calls per forward).
// Replaces the pair ff.net.0 (GEGLU) -> ff.net.2 of BasicTransformerBlock / TemporalBasicTransformerBlock inside
// UNetSpatioTemporalConditionModel (called at reference src/models/svd_unet.py:389-395) for C <= 320.
//
// One CTA per 128-row tile of x (persistent over tiles), the x tile [128, C] resident in shared memory.  The inner
// dimension 4C is walked in chunks of 64 gated columns:
//   GEMM1  S[128, 128] = x_tile W1_j^T          W1 rows interleaved per chunk [64 value | 64 gate]; A, B from smem
//   GEGLU  H_j[128, 64] = (S_v + b1_v) * gelu(S_g + b1_g), rounded to fp16 exactly like the unfused path, written to
//          TENSOR MEMORY (packed fp16, 32 columns)
//   GEMM2  Y[128, C] += H_j W2_j^T              A = H_j from tensor memory (tcgen05.mma, A operand in TMEM), B from smem
// and after the last chunk Y goes through the usual epilogue (bias, per-frame row vector, alpha, two scaled residuals).
// TMEM (512 columns): Y [0, 320) as two N = 160 halves | S [320, 448) | H0 [448, 480) | H1 [480, 512).
// S is single-buffered but released as soon as the GEGLU warps hold it in registers, so GEMM1 of chunk j+1 runs under the
// GEGLU arithmetic of chunk j and GEMM2 of chunk j follows it: the tensor pipe sees G1(j+1), G2(j), G1(j+2), ...
//   warp 0      TMA producer: x tile; W1 chunk k-blocks ([128 rows, 64 k] = 16 KB) through a 3-slot ring; W2 chunks
//               ([320 rows, 64 k] = 40 KB as two 160-row boxes) through a 2-slot ring
//   warp 1      MMA issuer (converged warp, elected lane)
//   warp 2      TMEM allocation;  warp 3: L2 prefetch of the residual rows the final epilogue will read
//   warps 4..11 GEGLU + final epilogue (thread = row = TMEM lane; the two warps of a lane quarter split the columns)
#include <cuda_fp16.h>

#include "common.h"
#include "ptx.cuh"

namespace svdpp {

struct FfParams {
  int M, C, n_chunks, kb;  // kb = C / 64
  int m_tiles;
  const __half* b1;  // [8C] interleaved like W1
  const __half* b2;  // [C]
  const __half* rowvec;
  long long rv_ld;
  int rv_hw, rv_div, rv_mod;
  const __half* R1;
  long long ldr1;
  float beta1;
  const __half* R2;
  long long ldr2;
  float beta2;
  float alpha;
  __half* D;
  long long ldd;
  int dbg;  // timing experiments only (tuning "ff_dbg"): 1 = no GELU arithmetic, 2 = no GEMM2, 4 = no final epilogue stores
};

constexpr int FF_BM = 128;
constexpr int FF_CMAX = 320;
constexpr int FF_XBLK = FF_BM * 64 * 2;           // one k-block of the x tile: 16 KB
constexpr int FF_W1_STAGE = 128 * 64 * 2;         // 16 KB
constexpr int FF_W1_STAGES = 3;                   // (measured: 4 or 5 stages with W2 in three 20 KB half-stages changes nothing)
constexpr int FF_W2_HALF = 160 * 64 * 2;          // 20 KB
constexpr int FF_W2_STAGE = 2 * FF_W2_HALF;       // 40 KB
constexpr int FF_W2_STAGES = 2;
constexpr int FF_B1_BYTES = 8 * FF_CMAX * 2;      // 5 KB: the whole interleaved b1
constexpr int FF_SMEM_BYTES = (FF_CMAX / 64) * FF_XBLK + FF_W1_STAGES * FF_W1_STAGE + FF_W2_STAGES * FF_W2_STAGE +
                              FF_B1_BYTES + 1024 /*align*/ + 512 /*barriers*/;
constexpr int FF_THREADS = 12 * 32;
constexpr int FF_EPI_WARPS = 8;

// exact-erf GELU, one MUFU (see gemm_tc.cu: degree-6 fit of log2 Phi(-a), max abs error 3.8e-6 over every fp16 input)
__device__ __forceinline__ float ff_gelu_erf(float x) {
  const float a = fminf(fabsf(x), 5.5f);
  float q = fmaf(a, 2.6153752e-05f, -6.6098123e-04f);
  q = fmaf(q, a, 7.4883099e-03f);
  q = fmaf(q, a, -5.1970404e-02f);
  q = fmaf(q, a, -4.6032947e-01f);
  q = fmaf(q, a, -1.1505840e+00f);
  q = fmaf(q, a, -1.0000361e+00f);
  return fmaf(-a, fast_exp2(q), fmaxf(x, 0.f));
}
// torch fp16 semantics of GEGLU, two columns at once: projection rounded to fp16, gelu(gate) rounded, product rounded
__device__ __forceinline__ uint32_t ff_geglu_fp16x2(float v0, float v1, float g0, float g1) {
  const __half2 v16 = __floats2half2_rn(v0, v1);
  const float2 gr = __half22float2(__floats2half2_rn(g0, g1));
  const __half2 ge = __floats2half2_rn(ff_gelu_erf(gr.x), ff_gelu_erf(gr.y));
  const __half2 r = __hmul2(v16, ge);
  return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ void ff_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void ff_umma2_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// TMA tensor load that lands at the same smem offset in every CTA of `mask` and signals the mbarrier at the same offset there
__device__ __forceinline__ void ff_tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// completion of all MMAs issued so far by this thread -> the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void ff_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

// PAIR: clusters of two CTAs work on two adjacent row tiles and SHARE the weight stream: each CTA fetches half of every W1
// k-block / W2 chunk and multicasts it into both CTAs' rings (the kernel re-reads all of W1 and W2 for every row tile -
// 2.4 MB per tile, 4.3 GB per level-0 call, which is what the L2 -> SM path can deliver in ~0.36 ms; the pair halves it).
// A ring slot is rewritten only when BOTH CTAs' MMAs have retired from it: every "empty" commit is multicast to the pair.
template <int MODE>
__global__ void __launch_bounds__(FF_THREADS, 1)
ff_geglu_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                const __grid_constant__ CUtensorMap tmW1h, const __grid_constant__ CUtensorMap tmW2,
                const __grid_constant__ CUtensorMap tmW2q, const FfParams p) {
  constexpr bool PAIR = MODE == 1;   // two CTAs share the weight stream through TMA multicast, independent MMAs
  constexpr bool TWO = MODE == 2;    // cta_group::2: the leader CTA issues M = 256 MMAs for both CTAs' row tiles
  constexpr bool CLUSTER = MODE != 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;                                         // [kb][128][64]
  uint8_t* sW1 = sX + (FF_CMAX / 64) * FF_XBLK;               // [3][128][64]
  uint8_t* sW2 = sW1 + FF_W1_STAGES * FF_W1_STAGE;            // [2][2][160][64]
  __half* sB1 = reinterpret_cast<__half*>(sW2 + FF_W2_STAGES * FF_W2_STAGE);   // [8C]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sB1) + FF_B1_BYTES);
  uint64_t* x_full = bars;                     // [1]
  uint64_t* x_empty = x_full + 1;              // [1]  every GEMM1 of the tile has retired
  uint64_t* w1_full = x_empty + 1;             // [FF_W1_STAGES]
  uint64_t* w1_empty = w1_full + FF_W1_STAGES; // [FF_W1_STAGES]
  uint64_t* w2_full = w1_empty + FF_W1_STAGES; // [FF_W2_STAGES]
  uint64_t* w2_empty = w2_full + FF_W2_STAGES; // [FF_W2_STAGES]
  uint64_t* s_full = w2_empty + FF_W2_STAGES;  // [1]  GEMM1 of a chunk retired
  uint64_t* s_free = s_full + 1;               // [1]  S is in the GEGLU warps' registers (8 warp arrivals)
  uint64_t* h_ready = s_free + 1;              // [2]  H buffer written (8 warp arrivals)
  uint64_t* h_free = h_ready + 2;              // [2]  GEMM2 that read the H buffer retired
  uint64_t* y_full = h_free + 2;               // [1]  every GEMM2 of the tile retired
  uint64_t* y_free = y_full + 1;               // [1]  Y has been read by the final epilogue (8 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(y_free + 1);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // work item w = one row tile (PAIR: two adjacent row tiles, one per CTA of the cluster; both CTAs run the same number of
  // items - an odd tile count gives the last CTA an all-zero tile whose rows are never stored)
  const int cta_rank = CLUSTER ? static_cast<int>(cluster_ctarank()) : 0;
  const int w_first = CLUSTER ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int w_stride = CLUSTER ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int w_total = CLUSTER ? (p.m_tiles + 1) >> 1 : p.m_tiles;
  const int n_my_tiles = w_first < w_total ? (w_total - w_first + w_stride - 1) / w_stride : 0;
  auto tile_of = [&](int it) { return CLUSTER ? 2 * (w_first + it * w_stride) + cta_rank : w_first + it * w_stride; };
  // TWO: every barrier the MMA issuer waits on lives in the LEADER CTA (rank 0); the follower's threads arrive remotely
  auto arrive_leader = [&](uint64_t* bar) {
    if constexpr (TWO)
      mbar_arrive_cluster(bar, 0);
    else
      mbar_arrive(bar);
  };
  constexpr int W1_BYTES = TWO ? FF_W1_STAGE / 2 : FF_W1_STAGE;   // TWO: each CTA holds half of the B rows of an MMA
  constexpr int W2_HALF_BYTES = TWO ? FF_W2_HALF / 2 : FF_W2_HALF;
  constexpr int EPI_ARRIVALS = TWO ? 2 * FF_EPI_WARPS : FF_EPI_WARPS;
  constexpr uint16_t MC_MASK = 3;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW1h);
    tma_prefetch_desc(&tmW2);
    mbar_init(x_full, TWO ? 2 : 1);  // TWO: both CTAs' producers arrive on the leader's barrier
    mbar_init(x_empty, 1);
    for (int s = 0; s < FF_W1_STAGES; ++s) {
      mbar_init(&w1_full[s], TWO ? 2 : 1);
      mbar_init(&w1_empty[s], PAIR ? 2 : 1);  // pairs: both CTAs' MMAs must have retired from the slot
    }
    for (int s = 0; s < FF_W2_STAGES; ++s) {
      mbar_init(&w2_full[s], TWO ? 2 : 1);
      mbar_init(&w2_empty[s], PAIR ? 2 : 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, EPI_ARRIVALS);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&h_ready[s], EPI_ARRIVALS);
      mbar_init(&h_free[s], 1);
    }
    mbar_init(y_full, 1);
    mbar_init(y_free, EPI_ARRIVALS);
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (TWO) {
      tmem_alloc2(tmem_slot, 512);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  pdl_launch_dependents();
  pdl_wait();
  // the interleaved b1 -> smem (read by every GEGLU thread, the same values for all rows)
  for (int i = threadIdx.x; i < p.n_chunks * 128 / 8; i += FF_THREADS)
    reinterpret_cast<uint4*>(sB1)[i] = reinterpret_cast<const uint4*>(p.b1)[i];
  tc_fence_before();
  if constexpr (CLUSTER)
    cluster_sync_all();  // the peer's barriers are initialised before anything is multicast at them
  else
    __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_Y = tmem_base;
  const uint32_t tmem_S = tmem_base + 320;
  const uint32_t tmem_H = tmem_base + 448;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int g1 = 0, g2 = 0;  // running W1 k-block / W2 chunk counters (ring positions)
      for (int it = 0; it < n_my_tiles; ++it) {
        const int m0 = tile_of(it) * FF_BM;
        mbar_wait(x_empty, (it & 1) ^ 1, 81);
        if constexpr (TWO) {
          if (cta_rank == 0)
            mbar_expect_tx(x_full, 2 * p.kb * FF_XBLK);  // both CTAs' bytes land on the leader's barrier
          else
            mbar_arrive_cluster(x_full, 0);
          for (int kb = 0; kb < p.kb; ++kb) tma2_load_2d(sX + kb * FF_XBLK, &tmX, x_full, kb * 64, m0);
        } else {
          mbar_expect_tx(x_full, p.kb * FF_XBLK);
          for (int kb = 0; kb < p.kb; ++kb) tma_load_2d(sX + kb * FF_XBLK, &tmX, x_full, kb * 64, m0);
        }
        auto load_w2 = [&](int j) {
          const int s = g2 % FF_W2_STAGES;
          mbar_wait(&w2_empty[s], ((g2 / FF_W2_STAGES) & 1) ^ 1, 83);
          if constexpr (TWO) {  // my 80 rows of each N = 160 half; the MMA reads the other 80 from the peer's smem
            if (cta_rank == 0)
              mbar_expect_tx(&w2_full[s], 2 * 2 * W2_HALF_BYTES);
            else
              mbar_arrive_cluster(&w2_full[s], 0);
            tma2_load_2d(sW2 + s * FF_W2_STAGE, &tmW2q, &w2_full[s], j * 64, cta_rank * 80);
            tma2_load_2d(sW2 + s * FF_W2_STAGE + W2_HALF_BYTES, &tmW2q, &w2_full[s], j * 64, 160 + cta_rank * 80);
            ++g2;
            return;
          }
          mbar_expect_tx(&w2_full[s], FF_W2_STAGE);
          if constexpr (PAIR) {  // my 160-row half, into both CTAs
            ff_tma_load_2d_mc(sW2 + s * FF_W2_STAGE + cta_rank * FF_W2_HALF, &tmW2, &w2_full[s], j * 64, cta_rank * 160, MC_MASK);
          } else {
            tma_load_2d(sW2 + s * FF_W2_STAGE, &tmW2, &w2_full[s], j * 64, 0);
            tma_load_2d(sW2 + s * FF_W2_STAGE + FF_W2_HALF, &tmW2, &w2_full[s], j * 64, 160);
          }
          ++g2;
        };
        for (int j = 0; j < p.n_chunks; ++j) {
          for (int kb = 0; kb < p.kb; ++kb, ++g1) {
            const int s = g1 % FF_W1_STAGES;
            mbar_wait(&w1_empty[s], ((g1 / FF_W1_STAGES) & 1) ^ 1, 82);
            if constexpr (TWO) {  // my 64 of the chunk's 128 rows
              if (cta_rank == 0)
                mbar_expect_tx(&w1_full[s], 2 * W1_BYTES);
              else
                mbar_arrive_cluster(&w1_full[s], 0);
              tma2_load_2d(sW1 + s * FF_W1_STAGE, &tmW1h, &w1_full[s], kb * 64, j * 128 + cta_rank * 64);
              continue;
            }
            mbar_expect_tx(&w1_full[s], FF_W1_STAGE);
            if constexpr (PAIR)  // my half of the 128 rows, into both CTAs
              ff_tma_load_2d_mc(sW1 + s * FF_W1_STAGE + cta_rank * (FF_W1_STAGE / 2), &tmW1h, &w1_full[s], kb * 64,
                                j * 128 + cta_rank * 64, MC_MASK);
            else
              tma_load_2d(sW1 + s * FF_W1_STAGE, &tmW1, &w1_full[s], kb * 64, j * 128);
          }
          load_w2(j);
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ residual prefetch: the final epilogue of a tile reads
    // R1 / R2 [128 rows, C] a whole tile time from now; pull the lines into L2 so that its loads do not wait for HBM
    for (int it = 0; it < n_my_tiles; ++it) {
      const long long m0 = static_cast<long long>(tile_of(it)) * FF_BM;
      const int lpr = (p.C * 2 + 127) / 128;  // 128-byte lines per row
      for (int i = lane; i < FF_BM * lpr; i += 32) {
        const int r = i / lpr, l = i - r * lpr;
        if (m0 + r < p.M) {
          if (p.R1 != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.R1 + (m0 + r) * p.ldr1 + l * 64));
          if (p.R2 != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.R2 + (m0 + r) * p.ldr2 + l * 64));
        }
      }
      if (it + 1 < n_my_tiles) {  // pace: one tile ahead of the epilogue is enough
        if constexpr (TWO)
          mbar_wait(y_full, it & 1, 93);  // (x_full only completes in the leader CTA)
        else
          mbar_wait(x_full, it & 1, 93);
      }
    }
  } else if (warp == 1 && (!TWO || cta_rank == 0)) {
    // ------------------------------------------------------------------ MMA issuer (converged warp, elected lane; TWO: leader CTA only)
    const bool leader = elect_one();
    constexpr uint32_t idesc1 = make_idesc_f16(128, false, TWO ? 256 : 128);  // S: N = 128 (64 value + 64 gate rows of W1)
    constexpr uint32_t idesc2 = make_idesc_f16(160, false, TWO ? 256 : 128);  // Y half: N = 160 rows of W2, K-major
    auto mma_ss = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
      if constexpr (TWO)
        umma2_f16(d, da, db, idesc, acc);
      else
        umma_f16(d, da, db, idesc, acc);
    };
    auto mma_ts = [&](uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
      if constexpr (TWO)
        ff_umma2_ts(d, a, db, idesc, acc);
      else
        ff_umma_ts(d, a, db, idesc, acc);
    };
    auto commit = [&](uint64_t* bar) {  // TWO: the completion is signalled at this offset in BOTH CTAs
      if constexpr (TWO)
        umma2_commit(bar);
      else
        umma_commit(bar);
    };
    const uint32_t x_addr = smem_u32(sX);
    const uint64_t a_desc0 = make_smem_desc_sw128(x_addr, 1024, 0);
    const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(sW1), 1024, 0);
    const uint64_t w2_desc0 = make_smem_desc_sw128(smem_u32(sW2), 1024, 0);
    int g1 = 0, g2 = 0, gs = 0;  // W1 k-blocks, W2 chunks, S uses (all running over the tiles of this CTA)
    bool w1_ready = false;  // w1_full of the next ring position already seen complete (probed one k-block ahead)
    auto gemm1 = [&](int /*j*/) {
      if (gs > 0) {  // S of the previous chunk must be in the GEGLU warps' registers
        mbar_wait(s_free, (gs - 1) & 1, 84);
        tc_fence_after();
      }
#pragma unroll
      for (int kb = 0; kb < FF_CMAX / 64; ++kb) {
        if (kb < p.kb) {
          const int s = g1 % FF_W1_STAGES;
          if (!w1_ready) mbar_wait(&w1_full[s], (g1 / FF_W1_STAGES) & 1, 85);
          tc_fence_after();
          // probe the next position before this one's MMAs go out (a try_wait costs ~90 clocks even on a completed phase)
          w1_ready = mbar_test_wait(&w1_full[(g1 + 1) % FF_W1_STAGES], ((g1 + 1) / FF_W1_STAGES) & 1);
          // lean issue: descriptors are a constant plus (address >> 4) in the low word - one add per operand and MMA
          const uint64_t da = a_desc0 + static_cast<uint64_t>(kb * (FF_XBLK >> 4));
          const uint64_t db = b_desc0 + static_cast<uint64_t>(s * (FF_W1_STAGE >> 4));
          if (leader) {
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_ss(tmem_S, da + 2 * k, db + 2 * k, idesc1, (kb | k) != 0 ? 1u : 0u);
            if constexpr (PAIR)
              ff_commit_mc(&w1_empty[s], MC_MASK);
            else
              commit(&w1_empty[s]);
          }
          __syncwarp();
          ++g1;
        }
      }
      if (leader) commit(s_full);
      __syncwarp();
      ++gs;
    };
    auto gemm2 = [&](int j, int it) {
      const int hb = g2 & 1;
      const int s = g2 % FF_W2_STAGES;
      mbar_wait(&w2_full[s], (g2 / FF_W2_STAGES) & 1, 86);
      mbar_wait(&h_ready[hb], (g2 >> 1) & 1, 87);
      if (j == 0 && it > 0) mbar_wait(y_free, (it - 1) & 1, 88);  // the previous tile's Y has been read
      tc_fence_after();
      const uint64_t db = w2_desc0 + static_cast<uint64_t>(s * (FF_W2_STAGE >> 4));
      const uint32_t a_h = tmem_H + hb * 32;
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (p.dbg & 2) break;
          mma_ts(tmem_Y, a_h + k * 8, db + 2 * k, idesc2, (j | k) != 0 ? 1u : 0u);
          mma_ts(tmem_Y + 160, a_h + k * 8, db + (W2_HALF_BYTES >> 4) + 2 * k, idesc2, (j | k) != 0 ? 1u : 0u);
        }
        commit(&h_free[hb]);
        if constexpr (PAIR)
          ff_commit_mc(&w2_empty[s], MC_MASK);
        else
          commit(&w2_empty[s]);
      }
      __syncwarp();
      ++g2;
    };
    for (int it = 0; it < n_my_tiles; ++it) {
      mbar_wait(x_full, it & 1, 89);
      tc_fence_after();
      // Issue order G1(j+1), G2(j).  Measured alternatives (level 0, burst clocks, this order: 0.568 ms): G2 two chunks
      // behind with the producer order matched 0.584; a second issuer warp for the GEMM2s 0.626 (the GEGLU arithmetic then
      // sits in the chain); W1 ring 4 / 5 deep with W2 in three half-stages 0.618.
      gemm1(0);
      for (int j = 0; j < p.n_chunks; ++j) {
        if (j + 1 < p.n_chunks) {
          gemm1(j + 1);
        } else {
          if (leader) commit(x_empty);  // every GEMM1 of this tile issued: x tile reusable once they retire
          __syncwarp();
        }
        gemm2(j, it);
      }
      if (leader) commit(y_full);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ GEGLU + final epilogue (thread = row)
    const int we = warp & 3;                 // TMEM lane quarter
    const int half = (warp - 4) >> 2;        // which half of the columns this warp takes
    const int row = we * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(we * 32) << 16;
    int gs = 0;  // chunks processed by this warp (over all tiles)
    for (int it = 0; it < n_my_tiles; ++it) {
      const int m_tile = tile_of(it);
      for (int j = 0; j < p.n_chunks; ++j, ++gs) {
        const int hb = gs & 1;
        mbar_wait(s_full, gs & 1, 90);
        tc_fence_after();
        uint32_t v[32], g[32];
        tmem_ld_x32(tmem_S + lane_sel + half * 32, v);        // value columns [half*32, +32)
        tmem_ld_x32(tmem_S + lane_sel + 64 + half * 32, g);   // gate columns
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(s_free);   // GEMM1 of the next chunk may overwrite S
        const __half* bv = sB1 + j * 128 + half * 32;
        const __half* bg = bv + 64;
        uint32_t hpk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 b_v = __half22float2(*reinterpret_cast<const __half2*>(bv + i));
          const float2 b_g = __half22float2(*reinterpret_cast<const __half2*>(bg + i));
          if (p.dbg & 1) {
            const __half2 t = __floats2half2_rn(__uint_as_float(v[i]) + b_v.x, __uint_as_float(g[i + 1]) + b_g.y);
            hpk[i / 2] = *reinterpret_cast<const uint32_t*>(&t);
          } else
          hpk[i / 2] = ff_geglu_fp16x2(__uint_as_float(v[i]) + b_v.x, __uint_as_float(v[i + 1]) + b_v.y,
                                       __uint_as_float(g[i]) + b_g.x, __uint_as_float(g[i + 1]) + b_g.y);
        }
        if (gs >= 2) {  // the GEMM2 that read this H buffer two chunks ago has retired
          mbar_wait(&h_free[hb], ((gs - 2) >> 1) & 1, 91);
          tc_fence_after();
        }
        tmem_st_x16(tmem_H + lane_sel + hb * 32 + half * 16, hpk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(&h_ready[hb]);
      }
      // ---- final epilogue of the tile: this thread's row, columns [half*160, +160)
      mbar_wait(y_full, it & 1, 92);
      tc_fence_after();
      const long long m = static_cast<long long>(m_tile) * FF_BM + row;
      const bool m_ok = m < p.M;
      const __half* rv_row = nullptr;
      if (p.rowvec != nullptr && m_ok) rv_row = p.rowvec + static_cast<long long>(((m / p.rv_hw) / p.rv_div) % p.rv_mod) * p.rv_ld;
#pragma unroll 1
      for (int c = 0; c < 5; ++c) {  // 32 columns per round; the residual loads are in flight during the TMEM round trip
        const int col = half * 160 + c * 32;
        const bool on = m_ok && col < p.C && !(p.dbg & 4);
        uint4 r1v[4], r2v[4];
        if (on && p.R1 != nullptr) {
          const uint4* src = reinterpret_cast<const uint4*>(p.R1 + m * p.ldr1 + col);
#pragma unroll
          for (int q = 0; q < 4; ++q) r1v[q] = src[q];
        }
        if (on && p.R2 != nullptr) {
          const uint4* src = reinterpret_cast<const uint4*>(p.R2 + m * p.ldr2 + col);
#pragma unroll
          for (int q = 0; q < 4; ++q) r2v[q] = src[q];
        }
        uint32_t y[32];
        tmem_ld_x32(tmem_Y + lane_sel + col, y);
        tmem_ld_wait();
        if (c == 4) {  // Y is in registers: the next tile's GEMM2 may start
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_leader(y_free);
        }
        if (on) {
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 b2 = __half22float2(*reinterpret_cast<const __half2*>(p.b2 + col + i));
            f[i] = __uint_as_float(y[i]) + b2.x;
            f[i + 1] = __uint_as_float(y[i + 1]) + b2.y;
          }
          if (rv_row != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float2 r = __half22float2(*reinterpret_cast<const __half2*>(rv_row + col + i));
              f[i] += r.x;
              f[i + 1] += r.y;
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] *= p.alpha;
          if (p.R1 != nullptr) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const __half2* h2 = reinterpret_cast<const __half2*>(&r1v[q]);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 r = __half22float2(h2[i]);
                f[q * 8 + 2 * i] += p.beta1 * r.x;
                f[q * 8 + 2 * i + 1] += p.beta1 * r.y;
              }
            }
          }
          if (p.R2 != nullptr) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const __half2* h2 = reinterpret_cast<const __half2*>(&r2v[q]);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 r = __half22float2(h2[i]);
                f[q * 8 + 2 * i] += p.beta2 * r.x;
                f[q * 8 + 2 * i + 1] += p.beta2 * r.y;
              }
            }
          }
          __align__(16) __half2 o[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
          uint4* dst = reinterpret_cast<uint4*>(p.D + m * p.ldd + col);
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[q] = reinterpret_cast<const uint4*>(o)[q];
        }
      }
    }
  }

  tc_fence_before();
  if constexpr (CLUSTER)
    cluster_sync_all();  // no CTA leaves while its peer may still multicast into it or arrive on its barriers
  else
    __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (TWO)
      tmem_dealloc2(tmem_base, 512);
    else
      tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace svdpp

using namespace svdpp;

extern "C" int svdpp_ff_geglu_f16(const svdpp_ff_desc* d, svdpp_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SVDPP_CHECK_ARG(d && d->X && d->W1 && d->b1 && d->W2 && d->b2 && d->D, "ff: null pointers");
  SVDPP_CHECK_ARG(d->M > 0 && d->C > 0 && d->C % 64 == 0 && d->C <= FF_CMAX, "ff: C must be a multiple of 64, at most %d (got %d)",
                  FF_CMAX, d->C);
  SVDPP_CHECK_ARG(d->ldx % 8 == 0 && d->ldd % 8 == 0 && d->ldw2 % 8 == 0, "ff: pitches must be multiples of 8");
  SVDPP_CHECK_ARG((d->R1 == nullptr || d->ldr1 % 8 == 0) && (d->R2 == nullptr || d->ldr2 % 8 == 0), "ff: residual pitches must be multiples of 8");
  SVDPP_CHECK_ARG(d->w2_rows >= d->C, "ff: W2 needs at least C rows (got %d)", d->w2_rows);
  FfParams p{};
  p.M = d->M;
  p.C = d->C;
  p.kb = d->C / 64;
  p.n_chunks = 4 * d->C / 64;
  p.m_tiles = (d->M + FF_BM - 1) / FF_BM;
  p.b1 = static_cast<const __half*>(d->b1);
  p.b2 = static_cast<const __half*>(d->b2);
  p.rowvec = static_cast<const __half*>(d->rowvec);
  p.rv_ld = d->rv_ld;
  p.rv_hw = d->rv_hw > 0 ? d->rv_hw : 1;
  p.rv_div = d->rv_div > 0 ? d->rv_div : 1;
  p.rv_mod = d->rv_mod > 0 ? d->rv_mod : 0x7fffffff;
  p.R1 = static_cast<const __half*>(d->R1);
  p.ldr1 = d->ldr1;
  p.beta1 = d->beta1;
  p.R2 = static_cast<const __half*>(d->R2);
  p.ldr2 = d->ldr2;
  p.beta2 = d->beta2;
  p.alpha = d->alpha;
  p.D = static_cast<__half*>(d->D);
  p.ldd = d->ldd;
  p.dbg = tuning().ff_dbg;
  CUtensorMap tmX, tmW1, tmW1h, tmW2, tmW2q;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(d->C), static_cast<uint64_t>(d->M)};
    uint64_t str[1] = {static_cast<uint64_t>(d->ldx) * 2};
    uint32_t box[2] = {64, FF_BM};
    if (encode_tmap_f16(&tmX, d->X, 2, dims, str, box)) return -5;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(d->C), static_cast<uint64_t>(8 * d->C)};
    uint64_t str[1] = {static_cast<uint64_t>(d->C) * 2};
    uint32_t box[2] = {64, 128};
    if (encode_tmap_f16(&tmW1, d->W1, 2, dims, str, box)) return -5;
    uint32_t boxh[2] = {64, 64};
    if (encode_tmap_f16(&tmW1h, d->W1, 2, dims, str, boxh)) return -5;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(4 * d->C), static_cast<uint64_t>(d->w2_rows)};
    uint64_t str[1] = {static_cast<uint64_t>(d->ldw2) * 2};
    uint32_t box[2] = {64, 160};
    if (encode_tmap_f16(&tmW2, d->W2, 2, dims, str, box)) return -5;
    uint32_t boxq[2] = {64, 80};
    if (encode_tmap_f16(&tmW2q, d->W2, 2, dims, str, boxq)) return -5;
  }
  static bool configured = false;
  if (!configured) {
    SVDPP_CUDA(cudaFuncSetAttribute(ff_geglu_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM_BYTES));
    SVDPP_CUDA(cudaFuncSetAttribute(ff_geglu_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM_BYTES));
    SVDPP_CUDA(cudaFuncSetAttribute(ff_geglu_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM_BYTES));
    configured = true;
  }
  const int mode = p.m_tiles >= 2 ? tuning().ff_pair : 0;   // 0 single CTAs, 1 multicast pairs, 2 cta_group::2 pairs
  if (mode != 0) {
    const int pairs = (p.m_tiles + 1) / 2;
    const int max_pairs = num_sms() / 2;
    const int grid = 2 * (pairs < max_pairs ? pairs : max_pairs);
    if (mode == 2)
      SVDPP_CUDA(launch_kernel(ff_geglu_kernel<2>, dim3(grid), dim3(FF_THREADS), FF_SMEM_BYTES, stream, 2, tmX, tmW1, tmW1h, tmW2, tmW2q, p));
    else
      SVDPP_CUDA(launch_kernel(ff_geglu_kernel<1>, dim3(grid), dim3(FF_THREADS), FF_SMEM_BYTES, stream, 2, tmX, tmW1, tmW1h, tmW2, tmW2q, p));
  } else {
    const int grid = p.m_tiles < num_sms() ? p.m_tiles : num_sms();
    SVDPP_CUDA(launch_kernel(ff_geglu_kernel<0>, dim3(grid), dim3(FF_THREADS), FF_SMEM_BYTES, stream, 1, tmX, tmW1, tmW1h, tmW2, tmW2q, p));
  }
  return check_launch("ff_geglu_kernel");
}
