// svdpp_gemm_f16: fp16 GEMM / implicit-GEMM convolution with a fused epilogue, sm_100a.
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0      TMA producer   global -> 128B-swizzled smem ring (A tile 128x64, B tile BNx64)
//   warp 1      MMA issuer     one thread issues tcgen05.mma (M=128, N=BN, K=16) into TMEM
//   warp 2      TMEM allocator (512 columns = two BN-wide fp32 accumulator stages)
//   warp 3      epilogue DMA lane (p.epi_dma): issues the output tensor stores and residual tensor loads; otherwise
//               (p.two_prod, long main loops) a SECOND TMA producer taking every other k-block
//   warps 4..11 epilogue       tcgen05.ld accumulator rows -> bias/rowvec/residual/GEGLU -> fp16 into a staging tile in
//               smem; two warps per TMEM lane quarter, each taking every other column chunk.  The staging tile is
//               kept in the 64-byte swizzle of a tensor map over the output and leaves through TMA tensor stores
//               (one thread per epilogue group, or the DMA lane); the residual tile R1 arrives in it through TMA
//               tensor loads.  A padded staging tile with per-thread coalesced copies remains for outputs the
//               tensor map cannot describe (sub-pixel scatter, rows that are not 16-byte multiples).
// The two accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
// Convolutions never materialise im2col: for tap (dw,dh,df) the producer loads the activation
// window shifted by the tap through a 5-D tensor map [C, W, H, F, B]; TMA zero-fills the halo.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace svdpp {

struct GemmParams {
  int M, N;
  int num_kb, kb_split;
  int conv, cF, cH, cW, cpk, nrows, bw, cstride;
  int8_t taps[SVDPP_MAX_TAPS][4];
  const __half* bias;
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
  int n_store;
  int m_tiles, n_tiles;
  int rev_m;   // walk the M tiles from the last row block to the first (the end of A is hot in L2 after a forward producer)
  int prefetch_r1;  // producer warp pulls the residual tile into L2 one main loop ahead of the epilogue
  // sub-pixel output map (conv mode, up > 1): GEMM row m = pixel (img, h, w) of the cH x cW grid is stored at pixel
  // (up*h + up_y, up*w + up_x) of the (up*cH) x (up*cW) output image
  int up, up_y, up_x;
  // 1: the output tile leaves through TMA tensor stores (tmD) from a 64-byte-swizzled staging tile instead of the
  // per-thread smem -> global copy loop (ncu: that loop held 48 % of the epilogue warps' samples on the K = 320 layers)
  int tma_store;
  // 1 (TMA-store mode only): the residual tile R1 arrives in the staging tile through TMA tensor loads (tmR, same
  // boxes and swizzle as tmD) instead of per-thread global loads + st.shared
  int tma_r1;
  // 1 (TMA-store mode, one epilogue group, short K): a dedicated DMA lane (warp 3) issues the tensor stores and the
  // residual loads, and two staging tiles alternate - the second one lives in the last pipeline stages, which a
  // short main loop does not need.  The epilogue warps never wait for a store to drain or a residual to arrive.
  int epi_dma;
  // 1: warp 3 is a second TMA producer taking every other k-block (never together with epi_dma, whose lane it is).
  // tools/ubench/fill.cu: one issuing thread sustains one k-block per ~500-590 clocks whatever its size, two producers
  // 300-440 - more than the 256 / 320 / 512 clocks of MMAs a 128 / 160 / 256-wide tile spends on a k-block.
  int two_prod;
  // Split-K tail (256x320 pair tiles only).  The tiles of the last, partial wave (sk_r of them, after sk_full tiles
  // in full waves) are each computed by `splitk` CTA pairs over disjoint K ranges; the fp32 partial accumulators meet
  // in the workspace sk_ws, and once a tile's partials are all there (counter in sk_cnt) every participating CTA
  // finishes a share of its 16-column chunks, summing the partials in slice order (deterministic).
  int splitk, sk_r, sk_full;
  float* sk_ws;
  unsigned* sk_cnt;
  long long sk_ws_bytes;  // host side only: size of the caller's scratch (counters + partials)
};

// Exact-erf GELU (torch F.gelu, approximate='none') in 10 instructions and ONE MUFU:  gelu(x) = max(x,0) - a*Phi(-a),
// a = |x|, with Phi(-a) = 2^Q(a), Q a degree-6 minimax fit of log2 Phi(-a) on [0, 5.5] (relative error of Phi
// 2.6e-5, 20x below the fp16 rounding that follows; max abs error 3.8e-6 over every fp16 input; fitted and
// checked by tools/fit_gelu.py).  For a > 5.5 the term is below 1e-7 and a is clamped.  The Abramowitz-Stegun
// erfc form used before cost 16 instructions and two MUFU, which made the GEGLU epilogue of the K = 320 layers
// (128 x 128 outputs, 2 x 16384 MUFU per tile = 2048 cycles against a 2560-cycle main loop) the bound of the
// largest GEMMs of the network.
__device__ __forceinline__ float gelu_erf(float x) {
  const float a = fminf(fabsf(x), 5.5f);
  float q = fmaf(a, 2.6153752e-05f, -6.6098123e-04f);
  q = fmaf(q, a, 7.4883099e-03f);
  q = fmaf(q, a, -5.1970404e-02f);
  q = fmaf(q, a, -4.6032947e-01f);
  q = fmaf(q, a, -1.1505840e+00f);
  q = fmaf(q, a, -1.0000361e+00f);
  return fmaf(-a, fast_exp2(q), fmaxf(x, 0.f));
}

// torch fp16 semantics of GEGLU: proj output rounded to fp16, gelu(gate) rounded, product rounded
__device__ __forceinline__ __half geglu_fp16(float val, float gate) {
  __half v16 = __float2half_rn(val);
  __half g16 = __float2half_rn(gate);
  __half ge = __float2half_rn(gelu_erf(__half2float(g16)));
  return __hmul(v16, ge);
}
// Two columns at once: packed f32->f16x2 conversions and one HMUL2 (scalar F2F conversions are
// quarter-rate and were the epilogue's bottleneck).
__device__ __forceinline__ __half2 geglu_fp16x2(float v0, float v1, float g0, float g1) {
  const __half2 v16 = __floats2half2_rn(v0, v1);
  const float2 gr = __half22float2(__floats2half2_rn(g0, g1));
  const __half2 ge = __floats2half2_rn(gelu_erf(gr.x), gelu_erf(gr.y));
  return __hmul2(v16, ge);
}

__device__ __forceinline__ void load8(const __half* p, float (&o)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    o[2 * i] = f.x;
    o[2 * i + 1] = f.y;
  }
}

// Output row of GEMM row m (identity unless the sub-pixel map is on)
__device__ __forceinline__ long long out_row(const GemmParams& p, int m) {
  if (p.up <= 1) return m;
  const int w = m % p.cW;
  const int r = m / p.cW;
  const int h = r % p.cH;
  const long long img = r / p.cH;
  return (img * (p.up * p.cH) + p.up * h + p.up_y) * (p.up * p.cW) + p.up * w + p.up_x;
}

// TWO: CTA pairs (cta_group::2).  The pair computes a 256 x BN tile: each CTA loads its own 128 rows of A
// and HALF of the B tile, the leader issues M = 256 MMAs that read B from both CTAs' smem, and each CTA
// keeps the accumulator of its own 128 rows in its own TMEM.  Per SM and k-block this moves 26 KB instead of
// 36 KB from L2 (read in round 1 as an L2 -> SM cap of the one-CTA kernel near 1 PFLOP/s; tools/ubench/fill.cu later showed
// the cap was the single producer thread's k-block rate - see two_prod).
template <int BN, bool GEGLU, bool TWO, int EW_ = 0>
struct GemmCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_ROWS = TWO ? BN / 2 : BN;  // B rows this CTA loads
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // GROUPS = 2 (the 16-warp GEGLU epilogue): two independent groups of 8 epilogue warps, each with its own TMEM
  // stage, staging tile, bias tile and named barrier, take alternate tiles - the barrier and TMEM-load stalls of one
  // group are covered by the other group's arithmetic instead of idling all 16 warps (ncu: 26 % of the K = 320
  // GEGLU's samples sat on the epilogue's named barrier).  Costs one pipeline stage of smem.
  static constexpr int GROUPS = (GEGLU && BN == 256 && (EW_ ? EW_ : 16) == 16) ? 2 : 1;
  static constexpr int STAGES = (TWO && BN <= 160) ? 6 : (GROUPS == 2 ? 4 : 5);
  static constexpr int ACC_STRIDE = 256;  // TMEM columns between the two accumulator stages
  // BN = 320 (CTA pairs only): the whole N = 320 of the level-0 layers in one 256 x 320 pair tile, issued as two
  // N = 160 MMAs per k-step.  Per CTA and k-block 36 KB come from L2 for 5.2 MFLOP (145 FLOP/B against 71 for the
  // one-CTA 128x160 tile): the narrow tiles sat near 0.9 PFLOP/s (producer-issue-bound, see two_prod; the L2->SM path delivers 10.8 KB/clk).  320 fp32
  // columns leave room for ONE accumulator stage in the 512 TMEM columns, so TMEM reads of the epilogue are not
  // overlapped with the next main loop (stores still are); worth it for K >= 1280.
  static constexpr int NACC = BN > 256 ? 1 : 2;
  // BN = 320: the two 160-column halves of a tile rotate through THREE TMEM buffers (480 of 512 columns): tile i
  // uses buffers (2i) % 3 and (2i+1) % 3, so the next tile's MMAs need only the FIRST half of this tile to have
  // been read by the epilogue, not the whole tile - most of a second accumulator stage without its columns.
  static constexpr bool ROT3 = BN > 256;
  static constexpr int MMA_N = BN > 256 ? BN / 2 : BN;   // N of one tcgen05.mma
  static constexpr int MMAS = BN / MMA_N;                 // MMAs per k-step
  static constexpr int B_BOX_ROWS = B_ROWS / MMAS;        // rows of one B TMA box
  static constexpr int NOUT = GEGLU ? BN / 2 : BN;
  // epilogue staging tile [128][SW + 8] fp16, SW output columns per round (a 256-wide tile is staged in two
  // rounds of 128): the 16-byte pad makes a quarter-warp's 16-byte row accesses hit 32 distinct banks
  static constexpr int ROUNDS = NOUT > 160 ? 2 : 1;
  static constexpr int SW = NOUT / ROUNDS;
  static constexpr int C_PITCH = SW + 8;
  static constexpr int C_BYTES = BM * C_PITCH * 2;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + GROUPS * C_BYTES + 1024 /*align*/ + 2048 /*barriers + bias tiles*/;
  // Epilogue warps: the GEGLU epilogue (two accumulators, exact GELU, fp16 roundings per output) needs about as
  // many issue slots per tile as the K = 320 main loop has cycles; with 8 warps (2 per scheduler) dependency
  // stalls made it the bound.  It uses few registers, so the 256-wide variant runs 16 epilogue warps.
  static constexpr int EPI_WARPS = EW_ ? EW_ : ((GEGLU && BN == 256) ? 16 : 8);
  static constexpr int EPI_THREADS = EPI_WARPS * 32;
  // DMA-lane epilogue: bytes of one dense staging tile (SW columns as 32-column boxes) and the pipeline stages the
  // second tile takes over
  static constexpr int STG_TILE = (SW / 32) * 8192;
  static constexpr int STEAL = (STG_TILE + STAGE_BYTES - 1) / STAGE_BYTES;
  static constexpr int THREADS = 128 + EPI_THREADS;
};

template <int BN, bool GEGLU, bool TWO, int EW_ = 0>
__global__ void __launch_bounds__(GemmCfg<BN, GEGLU, TWO, EW_>::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD,
               const __grid_constant__ CUtensorMap tmR, const GemmParams p) {
  using Cfg = GemmCfg<BN, GEGLU, TWO, EW_>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __half* sC = reinterpret_cast<__half*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::GROUPS * Cfg::C_BYTES);
  uint64_t* empty = full + Cfg::STAGES;
  uint64_t* tfull = empty + Cfg::STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* r1_full = tempty + 3;  // [2] residual tile landed in the staging tile (one per epilogue group)
  uint64_t* stg_full = r1_full + 2;    // [2] DMA-lane epilogue: results of a round are in staging tile b
  uint64_t* stg_ready = stg_full + 2;  // [2] ... staging tile b is free again (and holds the residual tile, if any)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stg_ready + 2);
  static_assert((2 * Cfg::STAGES + 2 + 3 + 2 + 2 + 2) * 8 + 4 <= 192, "barrier area");
  const int nstages = p.epi_dma ? Cfg::STAGES - Cfg::STEAL : Cfg::STAGES;  // smem ring depth of this launch
  __half* sBias = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(full) + 192);  // [BN] this tile's bias

  // warp index through a shuffle: provably warp-uniform, so that what the MMA issuer derives from it stays in uniform registers
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // work item = one tile (or, for CTA pairs, two vertically adjacent tiles sharing the N tile)
  const int cta_rank = TWO ? static_cast<int>(cluster_ctarank()) : 0;
  const int work_first = TWO ? (blockIdx.x >> 1) : blockIdx.x;
  const int work_stride = TWO ? (gridDim.x >> 1) : gridDim.x;
  const int work_total_all = TWO ? ((p.m_tiles + 1) >> 1) * p.n_tiles : p.m_tiles * p.n_tiles;
  // with a split-K tail the regular loops cover the full waves only; this CTA pair's share of the tail is one extra
  // item (tile tail_w, k-blocks [tail_kb0, tail_kb1)) appended to the producer's and the MMA issuer's loops
  const int work_total = (Cfg::ROT3 && p.splitk > 1) ? p.sk_full : work_total_all;
  const int n_full_items = work_first < work_total ? (work_total - work_first + work_stride - 1) / work_stride : 0;
  int tail_w = -1, tail_kb0 = 0, tail_kb1 = 0, tail_j = 0, tail_s = 0;
  if constexpr (Cfg::ROT3) {
    if (p.splitk > 1) {
      tail_j = work_first / p.splitk;
      tail_s = work_first - tail_j * p.splitk;
      if (tail_j < p.sk_r) {
        tail_w = p.sk_full + tail_j;
        tail_kb0 = (tail_s * p.num_kb) / p.splitk;
        tail_kb1 = ((tail_s + 1) * p.num_kb) / p.splitk;
      }
    }
  }
  const int n_items = n_full_items + (tail_w >= 0 ? 1 : 0);
  const int m_rows_total = TWO ? (p.m_tiles + 1) >> 1 : p.m_tiles;
  auto m_tile_of = [&](int w) {
    int r = w / p.n_tiles;
    if (p.rev_m) r = m_rows_total - 1 - r;
    return TWO ? 2 * r + cta_rank : r;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) tma_prefetch_desc(&tmD);
    if (p.tma_r1) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full[s], TWO ? 2 : 1);  // pairs: both producers arrive on the leader's barrier
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) mbar_init(&tfull[s], 1);
    for (int s = 0; s < 2; ++s) mbar_init(&r1_full[s], 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&stg_full[s], Cfg::EPI_THREADS / Cfg::GROUPS);
      mbar_init(&stg_ready[s], 1);
    }
    for (int s = 0; s < 3; ++s)
      mbar_init(&tempty[s], (TWO ? 2 : 1) * Cfg::EPI_THREADS / Cfg::GROUPS);  // pairs: both CTAs' epilogue threads, on the leader
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
  tc_fence_before();
  if constexpr (TWO)
    cluster_sync_all();
  else
    __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, tensor-map prefetch) may run
  // while the previous kernel drains; nothing below starts before that kernel has completed.
  pdl_launch_dependents();
  pdl_wait();

  // ------------------------------------------------------------------ TMA producer (warp 0; with p.two_prod also warp 3)
  auto produce = [&](const int me, const int nprod) {
    // Lane 0 owns the ring (waits, expect_tx, the B tile); in conv mode lane r < nrows issues the
    // window box of image row r, whose coordinates are computed once per tile.  Keep this loop lean: ONE issuing
    // thread gets through a k-block (barrier wait, expect_tx, the TMA instructions) in 500-590 clocks however small the
    // boxes are (tools/ubench/fill.cu), which is more than the MMAs of a 128 / 160 / 256-wide tile take - hence the
    // second producer warp for long main loops.  (Tried and removed: an L2 prefetch stream 16 k-blocks
    // ahead via cp.async.bulk.prefetch.tensor lowered throughput by ~40 % on B200.)
    int stage = 0;
    uint32_t phase = 0;
    int kcount = 0;  // k-blocks of this CTA so far: with two producers, warp `me` fills those with kcount % 2 == me
    for (int item = 0; item < n_items; ++item) {
      const bool is_tail = item >= n_full_items;
      const int w = is_tail ? tail_w : work_first + item * work_stride;
      const int kb_begin = is_tail ? tail_kb0 : 0, kb_end = is_tail ? tail_kb1 : p.num_kb;
      const int m0 = m_tile_of(w) * Cfg::BM;
      const int n0 = (w % p.n_tiles) * BN + cta_rank * Cfg::B_ROWS;
      int cw = 0, ch = 0, cf = 0, cb = 0;
      if (p.conv && lane < p.nrows) {
        const int px = m0 + lane * p.bw;
        cw = px % p.cW;
        const int row = px / p.cW;
        ch = row % p.cH;
        const int img = row / p.cH;
        cf = img % p.cF;
        cb = img / p.cF;
      }
      if (!GEGLU && p.prefetch_r1 && me == 0) {
        // the epilogue of this tile will read R1[m0 .. m0+127][n0 .. n0+BN) a whole main loop from now: start the
        // HBM -> L2 fetch here so that its loads find the lines in L2 (K = 320 layers are epilogue-latency bound)
        const int n_first = (w % p.n_tiles) * BN;
        constexpr int LPR = (BN * 2 + 127) / 128;  // 128-byte lines per tile row
        for (int i = lane; i < Cfg::BM * LPR; i += 32) {
          const int r = i / LPR, l = i - r * LPR;
          if (m0 + r < p.M && n_first + l * 64 < p.n_store)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.R1 + static_cast<long long>(m0 + r) * p.ldr1 + n_first + l * 64));
        }
      }
      int tap = 0, kc = 0;  // k-block = (tap, kc)
      if (p.conv && kb_begin > 0) {
        tap = kb_begin / p.cpk;
        kc = kb_begin - tap * p.cpk;
      }
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + Cfg::A_BYTES;
        const bool mine = nprod == 1 || (kcount & 1) == me;
        ++kcount;
        if (lane == 0 && mine) {
          mbar_wait(&empty[stage], phase ^ 1, 1);
          if constexpr (TWO) {
            if (cta_rank == 0)
              mbar_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);  // both CTAs' bytes land on this barrier
            else
              mbar_arrive_cluster(&full[stage], 0);
            if constexpr (Cfg::MMAS == 2) {
              // this CTA's half of each of the two N = 160 MMAs: rows [80 r, 80 r + 80) and [160 + 80 r, ...)
              const int nt = (w % p.n_tiles) * BN + cta_rank * Cfg::B_BOX_ROWS;
              tma2_load_2d(sb, &tmB, &full[stage], kb * 64, nt);
              tma2_load_2d(sb + Cfg::B_BOX_ROWS * 128, &tmB, &full[stage], kb * 64, nt + Cfg::MMA_N);
            } else {
              tma2_load_2d(sb, &tmB, &full[stage], kb * 64, n0);
            }
            if (!p.conv) {
              if (kb < p.kb_split)
                tma2_load_2d(sa, &tmA, &full[stage], kb * 64, m0);
              else
                tma2_load_2d(sa, &tmA2, &full[stage], (kb - p.kb_split) * 64, m0);
            }
          } else {
            mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
            tma_load_2d(sb, &tmB, &full[stage], kb * 64, n0);
            if (!p.conv) {
              if (kb < p.kb_split)
                tma_load_2d(sa, &tmA, &full[stage], kb * 64, m0);
              else
                tma_load_2d(sa, &tmA2, &full[stage], (kb - p.kb_split) * 64, m0);
            }
          }
        }
        __syncwarp();
        if (p.conv) {
          if (lane < p.nrows && mine) {
            if constexpr (TWO)
              tma2_load_5d(sa + lane * p.bw * 128, &tmA, &full[stage], kc * 64, cw * p.cstride + p.taps[tap][0],
                           ch * p.cstride + p.taps[tap][1], cf + p.taps[tap][2], cb);
            else
              tma_load_5d(sa + lane * p.bw * 128, &tmA, &full[stage], kc * 64, cw * p.cstride + p.taps[tap][0],
                          ch * p.cstride + p.taps[tap][1], cf + p.taps[tap][2], cb);
          }
          if (++kc == p.cpk) {
            kc = 0;
            ++tap;
          }
        }
        if (++stage == nstages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  };
  if (warp == 0) {
    produce(0, p.two_prod ? 2 : 1);
  } else if (warp == 3 && p.two_prod) {
    produce(1, 2);
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The WHOLE warp runs the loop and one elected lane executes the tcgen05 instructions.  Under `if (lane == 0)`
    // the loop is divergent code: addresses and descriptors live in vector registers and every MMA pays a chain of
    // R2UR moves - tools/ubench/mma.cu: 116 clocks per issued MMA that way (166 for N = 256), while the tensor pipe
    // needs 80 for an N = 160 instruction and a lean converged loop issues at exactly the pipe's rate (64 / 80 / 128
    // clocks for N = 128 / 160 / 256 = 8190 FLOP/clk/SM).  The mbarrier probe of the NEXT stage is issued before this
    // stage's MMAs (a try_wait costs ~90 clocks even on a completed phase; its predicate is only consumed one
    // k-block later).
    if (cta_rank == 0) {  // pairs: only the leader CTA issues (for both CTAs)
      const bool leader = elect_one();
      constexpr uint32_t idesc = make_idesc_f16(Cfg::MMA_N, false, TWO ? 256 : 128);
      const uint32_t smem_base = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      int it = 0;  // tiles done by this CTA pair (ROT3 buffer rotation)
      bool ready = false;  // full[stage] already seen complete (probed one k-block ahead)
      for (int item = 0; item < n_items; ++item, ++it) {
        const bool is_tail = item >= n_full_items;
        const int kb_begin = is_tail ? tail_kb0 : 0, kb_end = is_tail ? tail_kb1 : p.num_kb;
        uint32_t d_tmem, d_tmem2 = 0;
        if constexpr (Cfg::ROT3) {
          const int h0 = 2 * it, h1 = h0 + 1;  // half-tile sequence numbers; buffer = h % 3, its use count = h / 3
          mbar_wait(&tempty[h0 % 3], ((h0 / 3) & 1) ^ 1, 2);
          mbar_wait(&tempty[h1 % 3], ((h1 / 3) & 1) ^ 1, 2);
          d_tmem = tmem_base + (h0 % 3) * Cfg::MMA_N;
          d_tmem2 = tmem_base + (h1 % 3) * Cfg::MMA_N;
        } else {
          mbar_wait(&tempty[as], aphase ^ 1, 2);
          d_tmem = tmem_base + as * Cfg::ACC_STRIDE;
        }
        tc_fence_after();
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          if (!ready) mbar_wait(&full[stage], phase, 3);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == nstages) {
            nstage = 0;
            nphase ^= 1;
          }
          ready = mbar_test_wait(&full[nstage], nphase);  // non-blocking; a "not yet" only costs the blocking wait above
          if (leader) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 1024, 0);
              const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 1024, 0);
              if constexpr (TWO) {
                umma2_f16(d_tmem, da, db, idesc, ((kb - kb_begin) | k) != 0 ? 1u : 0u);
                if constexpr (Cfg::MMAS == 2)
                  umma2_f16(d_tmem2, da, make_smem_desc_sw128(b_addr + Cfg::B_BOX_ROWS * 128 + k * 32, 1024, 0),
                            idesc, ((kb - kb_begin) | k) != 0 ? 1u : 0u);
              } else {
                umma_f16(d_tmem, da, db, idesc, ((kb - kb_begin) | k) != 0 ? 1u : 0u);
              }
            }
            if constexpr (TWO)
              umma2_commit(&empty[stage]);
            else
              umma_commit(&empty[stage]);  // smem slot reusable once these MMAs retire
          }
          __syncwarp();
          stage = nstage;
          phase = nphase;
        }
        if (leader) {
          if constexpr (Cfg::ROT3)
            umma2_commit(&tfull[it & 1]);
          else if constexpr (TWO)
            umma2_commit(&tfull[as]);
          else
            umma_commit(&tfull[as]);  // accumulator complete
        }
        __syncwarp();
        if (++as == Cfg::NACC) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ epilogue DMA lane (p.epi_dma)
    // Round R of this CTA (tile R / ROUNDS, columns (R % ROUNDS) * SW ...) uses staging tile R & 1.  For every round:
    // wait until the epilogue warps have staged it, store it, and - once the store has read the tile - hand the tile
    // to round R + 2, with that round's residual columns already travelling into it.
    if (p.epi_dma && lane == 0) {
      const int n_rounds = n_full_items * Cfg::ROUNDS;  // the split-K tail item does not go through the staging tiles
      uint8_t* sbuf[2] = {reinterpret_cast<uint8_t*>(sC), smem + (Cfg::STAGES - Cfg::STEAL) * Cfg::STAGE_BYTES};
      auto coords = [&](int R, int& m_base, int& nout0) {
        const int w = work_first + (R / Cfg::ROUNDS) * work_stride;
        m_base = m_tile_of(w) * Cfg::BM;
        nout0 = (w % p.n_tiles) * Cfg::NOUT + (R % Cfg::ROUNDS) * Cfg::SW;
      };
      auto prepare = [&](int R) {  // staging tile R & 1 is free: start round R's residual loads or just say so
        int m_base, nout0;
        coords(R, m_base, nout0);
        if (p.tma_r1 && m_base < p.M) {
          int nb = 0;
#pragma unroll
          for (int b = 0; b < Cfg::SW / 32; ++b) nb += (nout0 + b * 32 < p.n_store) ? 1 : 0;
          mbar_expect_tx(&stg_ready[R & 1], nb * 8192);
#pragma unroll
          for (int b = 0; b < Cfg::SW / 32; ++b)
            if (nout0 + b * 32 < p.n_store) tma_load_2d(sbuf[R & 1] + b * 8192, &tmR, &stg_ready[R & 1], nout0 + b * 32, m_base);
        } else {
          mbar_arrive(&stg_ready[R & 1]);
        }
      };
      for (int R = 0; R < 2 && R < n_rounds; ++R) prepare(R);
      for (int R = 0; R < n_rounds; ++R) {
        int m_base, nout0;
        coords(R, m_base, nout0);
        mbar_wait(&stg_full[R & 1], (R >> 1) & 1, 6);
        if (m_base < p.M) {
#pragma unroll
          for (int b = 0; b < Cfg::SW / 32; ++b)
            if (nout0 + b * 32 < p.n_store) tma_store_2d(&tmD, sbuf[R & 1] + b * 8192, nout0 + b * 32, m_base);
          bulk_commit_group();
        }
        if (R + 2 < n_rounds) {
          bulk_wait_group_read<0>();
          prepare(R + 2);
        }
      }
      bulk_wait_group<0>();  // the staging tiles must outlive the last store
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    // Per tile and per round of SW output columns: (1) coalesced copy of the R1 residual columns into the
    // staging buffer, (2) accumulator rows from TMEM, epilogue math, fp16 result back into the staging
    // buffer (the TMEM stage is released after the last round), (3) coalesced copy of the staging buffer
    // to global.  Thread = accumulator row; the two warps of a lane quarter take alternate column chunks.
    const int we = warp & 3;          // the TMEM lane quarter this warp may read (warp % 4)
    constexpr int EW = Cfg::EPI_WARPS / Cfg::GROUPS, ET = EW * 32, WPQ = EW / 4;  // per group; WPQ warps share a lane quarter
    const int grp = Cfg::GROUPS == 2 ? (warp - 4) / EW : 0;  // epilogue group: takes tiles grp, grp + GROUPS, ...
    const int half = ((warp - 4) % EW) >> 2;  // which of the quarter's WPQ warps: takes column chunks half, half + WPQ, ...
    const int row = we * 32 + lane;
    const int et = (threadIdx.x - 128) % ET;  // 0..ET-1 within the group
    sC += grp * (Cfg::C_BYTES / 2);
    sBias += grp * BN;
    constexpr int SW = Cfg::SW;
    constexpr int VPR = SW / 8;         // 16-byte vectors per staged row
    constexpr int CW = GEGLU ? 8 : 16;  // columns per chunk
    constexpr int NCH = SW / CW / WPQ;  // chunks per warp and round
    constexpr int NV = Cfg::BM * VPR / ET;  // staged vectors per thread
    static_assert(Cfg::BM * VPR % ET == 0 && (SW / CW) % WPQ == 0, "epilogue work split");
    auto epi_bar = [&] { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(ET) : "memory"); };
    // TMA-store mode: the staging tile is SW/32 boxes of [128 rows][32 columns] (8 KB each) in the 64-byte swizzle
    // of tmD: 16-byte piece j of row r sits at r * 64 + ((j ^ ((r >> 1) & 3)) << 4), so the row-per-thread accesses
    // of a quarter warp cover 128 contiguous bytes (no bank conflicts) and one thread per group stores the tile.
    const bool tst = p.tma_store != 0;
    uint8_t* sCb = reinterpret_cast<uint8_t*>(sC);
    auto stage_vec = [&](int r, int v) -> uint4* {  // 16-byte vector v of staged row r
      if (tst) return reinterpret_cast<uint4*>(sCb + (v >> 2) * 8192 + r * 64 + (((v & 3) ^ ((r >> 1) & 3)) << 4));
      return reinterpret_cast<uint4*>(sC + r * Cfg::C_PITCH + v * 8);
    };
    if (p.epi_dma) {
      // ---- DMA-lane mode (one group): no named barrier inside a tile, no wait for stores or residual loads
      if constexpr (!GEGLU && Cfg::GROUPS == 1) {
        uint8_t* sbuf[2] = {sCb, smem + (Cfg::STAGES - Cfg::STEAL) * Cfg::STAGE_BYTES};
        int as = 0, it = 0, rc = 0;
        uint32_t aphase = 0;
        for (int w = work_first; w < work_total; w += work_stride, ++it) {
          const int m_tile = m_tile_of(w);
          const int n_tile = w % p.n_tiles;
          const int m_base = m_tile * Cfg::BM;
          const int m = m_base + row;
          const bool m_ok = m < p.M;
          __half* sB = sBias + (it & 1) * BN;  // two bias tiles: tile it + 1 may be staged while stragglers finish tile it
          if (et < BN / 8) {
            uint4 bv = make_uint4(0, 0, 0, 0);
            if (p.bias != nullptr) bv = *reinterpret_cast<const uint4*>(p.bias + n_tile * BN + et * 8);
            *reinterpret_cast<uint4*>(sB + et * 8) = bv;
          }
          const __half* rv_row = nullptr;
          if (p.rowvec != nullptr && m_ok) {
            const int rr = ((m / p.rv_hw) / p.rv_div) % p.rv_mod;
            rv_row = p.rowvec + static_cast<long long>(rr) * p.rv_ld + n_tile * BN;
          }
          const uint32_t taddr_base = tmem_base + (static_cast<uint32_t>(we * 32) << 16);
          epi_bar();  // bias tile visible; everybody has left tile it - 1, so its bias tile may be rewritten by tile it + 1
#pragma unroll 1
          for (int rd = 0; rd < Cfg::ROUNDS; ++rd, ++rc) {
            const int col0 = rd * SW;
            const int rot_buf = (2 * it + rd) % 3;
            const uint32_t taddr = Cfg::ROT3 ? taddr_base + rot_buf * Cfg::MMA_N - col0 : taddr_base + as * Cfg::ACC_STRIDE;
            const int nout0 = n_tile * Cfg::NOUT + col0;
            uint8_t* sb = sbuf[rc & 1];
            auto svec = [&](int v) -> uint4* {  // 16-byte vector v of this thread's staged row (64-byte swizzle of tmD / tmR)
              return reinterpret_cast<uint4*>(sb + (v >> 2) * 8192 + row * 64 + (((v & 3) ^ ((row >> 1) & 3)) << 4));
            };
            uint4 rvv[NCH][CW / 8];
            if (rv_row != nullptr) {
#pragma unroll
              for (int ci = 0; ci < NCH; ++ci) {
                const int c = half + WPQ * ci;
#pragma unroll
                for (int hlf = 0; hlf < CW / 8; ++hlf)
                  rvv[ci][hlf] = (nout0 + c * CW < p.n_store)
                                     ? *reinterpret_cast<const uint4*>(rv_row + col0 + c * CW + hlf * 8)
                                     : make_uint4(0, 0, 0, 0);
              }
            }
            mbar_wait(&stg_ready[rc & 1], (rc >> 1) & 1, 7);  // staging tile free, residual columns landed
            if (rd == 0) {
              if constexpr (Cfg::ROT3)
                mbar_wait(&tfull[it & 1], (it >> 1) & 1, 4);
              else
                mbar_wait(&tfull[as], aphase, 4);
              tc_fence_after();
            }
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {
              const int c = half + WPQ * ci;
              uint32_t v[CW];
              tmem_ld_x16(taddr + col0 + c * CW, v);
              tmem_ld_wait();
              const int nout = nout0 + c * CW;
              float y[CW];
#pragma unroll
              for (int j = 0; j < CW; ++j) y[j] = __uint_as_float(v[j]);
              {
                float b8[8];
#pragma unroll
                for (int hlf = 0; hlf < CW / 8; ++hlf) {
                  load8(sB + col0 + c * CW + hlf * 8, b8);
#pragma unroll
                  for (int j = 0; j < 8; ++j) y[hlf * 8 + j] += b8[j];
                }
              }
              if (rv_row != nullptr) {
#pragma unroll
                for (int hlf = 0; hlf < CW / 8; ++hlf) {
                  const __half2* h2 = reinterpret_cast<const __half2*>(&rvv[ci][hlf]);
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(h2[j]);
                    y[hlf * 8 + 2 * j] += f.x;
                    y[hlf * 8 + 2 * j + 1] += f.y;
                  }
                }
              }
#pragma unroll
              for (int j = 0; j < CW; ++j) y[j] *= p.alpha;
              if (p.R1 != nullptr) {
                float b8[8];
#pragma unroll
                for (int hlf = 0; hlf < CW / 8; ++hlf) {
                  load8(reinterpret_cast<const __half*>(svec((c * CW) / 8 + hlf)), b8);
#pragma unroll
                  for (int j = 0; j < 8; ++j) y[hlf * 8 + j] += p.beta1 * b8[j];
                }
              }
              if (p.R2 != nullptr && m_ok && nout < p.n_store) {
                float b8[8];
#pragma unroll
                for (int hlf = 0; hlf < CW / 8; ++hlf) {
                  load8(p.R2 + static_cast<long long>(m) * p.ldr2 + nout + hlf * 8, b8);
#pragma unroll
                  for (int j = 0; j < 8; ++j) y[hlf * 8 + j] += p.beta2 * b8[j];
                }
              }
              __align__(16) __half o[CW];
#pragma unroll
              for (int j = 0; j < CW; j += 2) *reinterpret_cast<__half2*>(&o[j]) = __floats2half2_rn(y[j], y[j + 1]);
              const uint4* o4 = reinterpret_cast<const uint4*>(o);
#pragma unroll
              for (int q = 0; q < CW / 8; ++q) *svec((c * CW) / 8 + q) = o4[q];
            }
            fence_proxy_async_smem();  // staged results -> visible to the TMA engine
            if constexpr (Cfg::ROT3) {
              tc_fence_before();
              mbar_arrive_cluster(&tempty[rot_buf], 0);
            } else if (rd == Cfg::ROUNDS - 1) {
              tc_fence_before();
              if constexpr (TWO)
                mbar_arrive_cluster(&tempty[as], 0);
              else
                mbar_arrive(&tempty[as]);
            }
            mbar_arrive(&stg_full[rc & 1]);  // the DMA lane stores the tile
          }
          if (++as == Cfg::NACC) {
            as = 0;
            aphase ^= 1;
          }
        }
      }
    } else {
    uint32_t r1_phase = 0;
    int as = Cfg::GROUPS == 2 ? grp : 0;  // with two groups each owns one TMEM stage
    uint32_t aphase = 0;
    int it = grp;
    for (int w = work_first + grp * work_stride; w < work_total; w += Cfg::GROUPS * work_stride, it += Cfg::GROUPS) {
      const int m_tile = m_tile_of(w);
      const int n_tile = w % p.n_tiles;
      const int m_base = m_tile * Cfg::BM;
      const int m = m_base + row;
      const bool m_ok = m < p.M;
      // bias of this tile's BN weight rows -> smem (a global load per chunk stalled the epilogue ~500
      // cycles each, several times the K = 320 main loop)
      if (et < BN / 8) {
        uint4 bv = make_uint4(0, 0, 0, 0);
        if (p.bias != nullptr) bv = *reinterpret_cast<const uint4*>(p.bias + n_tile * BN + et * 8);
        *reinterpret_cast<uint4*>(sBias + et * 8) = bv;
      }
      const __half* rv_row = nullptr;
      if constexpr (!GEGLU) {
        if (p.rowvec != nullptr && m_ok) {
          const int rr = ((m / p.rv_hw) / p.rv_div) % p.rv_mod;
          rv_row = p.rowvec + static_cast<long long>(rr) * p.rv_ld + n_tile * BN;
        }
      }
      const uint32_t taddr_base = tmem_base + (static_cast<uint32_t>(we * 32) << 16);
#pragma unroll 1
      for (int rd = 0; rd < Cfg::ROUNDS; ++rd) {
        const int col0 = rd * SW;                       // first column of this round within the tile
        // TMEM address of this round's accumulator columns (ROT3: each 160-column half has its own buffer)
        const int rot_buf = (2 * it + rd) % 3;
        const uint32_t taddr = Cfg::ROT3 ? taddr_base + rot_buf * Cfg::MMA_N - col0
                                         : taddr_base + as * Cfg::ACC_STRIDE;
        const int nout0 = n_tile * Cfg::NOUT + col0;    // ... and in the output matrix
        const bool vec_ok = (p.ldd & 7) == 0 && nout0 + SW <= p.n_store;
        // per-image row vector of this thread's row and chunks -> registers, before any wait
        uint4 rvv[NCH][CW / 8];
        if constexpr (!GEGLU) {
          if (rv_row != nullptr) {
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {
              const int c = half + WPQ * ci;
#pragma unroll
              for (int hlf = 0; hlf < CW / 8; ++hlf)
                rvv[ci][hlf] = (nout0 + c * CW < p.n_store)
                                   ? *reinterpret_cast<const uint4*>(rv_row + col0 + c * CW + hlf * 8)
                                   : make_uint4(0, 0, 0, 0);
            }
          }
          if (p.R1 != nullptr && p.tma_r1) {
            // residual tile by TMA: thread 0 waits until the previous round's tensor store has read the staging tile,
            // then loads the boxes of this round into it; everybody waits on the mbarrier before the chunk loop
            if (et == 0 && m_base < p.M) {
              bulk_wait_group_read<0>();
              int nb = 0;
#pragma unroll
              for (int b = 0; b < SW / 32; ++b) nb += (nout0 + b * 32 < p.n_store) ? 1 : 0;
              mbar_expect_tx(&r1_full[grp], nb * 8192);
#pragma unroll
              for (int b = 0; b < SW / 32; ++b)
                if (nout0 + b * 32 < p.n_store) tma_load_2d(sCb + b * 8192, &tmR, &r1_full[grp], nout0 + b * 32, m_base);
            }
          } else if (p.R1 != nullptr) {
            const bool r_vec = (p.ldr1 & 7) == 0 && nout0 + SW <= p.n_store;
            uint4 val[NV];  // all loads in flight at once
#pragma unroll
            for (int k = 0; k < NV; ++k) {
              const int i = et + k * ET;
              // TMA-store mode walks box by box, four threads per 64-byte row piece (matches the swizzled layout)
              const int r = tst ? (i & 511) >> 2 : i / VPR, v = tst ? ((i >> 9) << 2) | (i & 3) : i - r * VPR;
              val[k] = make_uint4(0, 0, 0, 0);
              if (m_base + r < p.M) {
                const __half* src = p.R1 + static_cast<long long>(m_base + r) * p.ldr1 + nout0 + v * 8;
                if (r_vec || ((p.ldr1 & 7) == 0 && nout0 + v * 8 + 8 <= p.n_store)) {
                  val[k] = *reinterpret_cast<const uint4*>(src);
                } else if (nout0 + v * 8 >= p.n_store) {
                  // vector entirely beyond the stored columns (N padded to the tile width)
                } else {
                  __half tmp[8];
                  for (int j = 0; j < 8; ++j) tmp[j] = (nout0 + v * 8 + j < p.n_store) ? src[j] : __float2half(0.f);
                  val[k] = *reinterpret_cast<uint4*>(tmp);
                }
              }
            }
            if (tst) {
              if (et == 0) bulk_wait_group_read<0>();  // the previous round's tensor store has read the staging tile
              epi_bar();
            }
#pragma unroll
            for (int k = 0; k < NV; ++k) {
              const int i = et + k * ET;
              const int r = tst ? (i & 511) >> 2 : i / VPR, v = tst ? ((i >> 9) << 2) | (i & 3) : i - r * VPR;
              *stage_vec(r, v) = val[k];
            }
          }
        }
        if (tst && (GEGLU || p.R1 == nullptr) && et == 0) bulk_wait_group_read<0>();  // ... staging tile is free
        epi_bar();
        if constexpr (!GEGLU) {
          if (p.R1 != nullptr && p.tma_r1 && m_base < p.M) {
            mbar_wait(&r1_full[grp], r1_phase, 5);
            r1_phase ^= 1;
          }
        }
        if (rd == 0) {
          if constexpr (Cfg::ROT3)
            mbar_wait(&tfull[it & 1], (it >> 1) & 1, 4);
          else
            mbar_wait(&tfull[as], aphase, 4);
          tc_fence_after();
        }
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const int c = half + WPQ * ci;
          uint32_t v[CW];
          uint32_t g[CW];
          if constexpr (GEGLU) {
            tmem_ld_x8(taddr + col0 + c * CW, v);
            tmem_ld_x8(taddr + BN / 2 + col0 + c * CW, g);
          } else {
            tmem_ld_x16(taddr + col0 + c * CW, v);
          }
          tmem_ld_wait();
          const int nout = nout0 + c * CW;  // output column
          float y[CW];
#pragma unroll
          for (int j = 0; j < CW; ++j) y[j] = __uint_as_float(v[j]);
          {
            float b8[8];
#pragma unroll
            for (int hlf = 0; hlf < CW / 8; ++hlf) {
              load8(sBias + col0 + c * CW + hlf * 8, b8);
#pragma unroll
              for (int j = 0; j < 8; ++j) y[hlf * 8 + j] += b8[j];
            }
          }
          __align__(16) __half o[CW];
          if constexpr (GEGLU) {
            float gt[CW];
#pragma unroll
            for (int j = 0; j < CW; ++j) gt[j] = __uint_as_float(g[j]);
            {
              float b8[8];
              load8(sBias + BN / 2 + col0 + c * CW, b8);
#pragma unroll
              for (int j = 0; j < 8; ++j) gt[j] += b8[j];
            }
#pragma unroll
            for (int j = 0; j < CW; j += 2)
              *reinterpret_cast<__half2*>(&o[j]) = geglu_fp16x2(y[j], y[j + 1], gt[j], gt[j + 1]);
          } else {
            if (rv_row != nullptr) {
#pragma unroll
              for (int hlf = 0; hlf < CW / 8; ++hlf) {
                const __half2* h2 = reinterpret_cast<const __half2*>(&rvv[ci][hlf]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = __half22float2(h2[j]);
                  y[hlf * 8 + 2 * j] += f.x;
                  y[hlf * 8 + 2 * j + 1] += f.y;
                }
              }
            }
#pragma unroll
            for (int j = 0; j < CW; ++j) y[j] *= p.alpha;
            if (p.R1 != nullptr) {
              float b8[8];
#pragma unroll
              for (int hlf = 0; hlf < CW / 8; ++hlf) {
                load8(reinterpret_cast<const __half*>(stage_vec(row, (c * CW) / 8 + hlf)), b8);
#pragma unroll
                for (int j = 0; j < 8; ++j) y[hlf * 8 + j] += p.beta1 * b8[j];
              }
            }
            if (p.R2 != nullptr && m_ok && nout < p.n_store) {
              float b8[8];
#pragma unroll
              for (int hlf = 0; hlf < CW / 8; ++hlf) {
                load8(p.R2 + static_cast<long long>(m) * p.ldr2 + nout + hlf * 8, b8);
#pragma unroll
                for (int j = 0; j < 8; ++j) y[hlf * 8 + j] += p.beta2 * b8[j];
              }
            }
#pragma unroll
            for (int j = 0; j < CW; j += 2) *reinterpret_cast<__half2*>(&o[j]) = __floats2half2_rn(y[j], y[j + 1]);
          }
          const uint4* o4 = reinterpret_cast<const uint4*>(o);
#pragma unroll
          for (int q = 0; q < CW / 8; ++q) *stage_vec(row, (c * CW) / 8 + q) = o4[q];
        }
        if (tst) fence_proxy_async_smem();  // staged results -> visible to the TMA engine
        if constexpr (Cfg::ROT3) {
          tc_fence_before();
          mbar_arrive_cluster(&tempty[rot_buf], 0);  // this half's buffer is free for the tile after next... or next
        } else if (rd == Cfg::ROUNDS - 1) {
          tc_fence_before();
          if constexpr (TWO)
            mbar_arrive_cluster(&tempty[as], 0);
          else
            mbar_arrive(&tempty[as]);  // accumulator stage free: the next tile's MMAs may start
        }
        epi_bar();
        if (tst) {
          if (et == 0 && m_base < p.M) {
#pragma unroll
            for (int b = 0; b < SW / 32; ++b)
              if (nout0 + b * 32 < p.n_store) tma_store_2d(&tmD, sCb + b * 8192, nout0 + b * 32, m_base);
            bulk_commit_group();
          }
          continue;  // the staging tile is handed back by the wait at the top of the next round
        }
        for (int i = et; i < Cfg::BM * VPR; i += ET) {
          const int r = i / VPR, v = i - r * VPR;
          if (m_base + r < p.M) {
            const uint4 val = *reinterpret_cast<const uint4*>(sC + r * Cfg::C_PITCH + v * 8);
            __half* dst = p.D + out_row(p, m_base + r) * p.ldd + nout0 + v * 8;
            if (vec_ok || ((p.ldd & 7) == 0 && nout0 + v * 8 + 8 <= p.n_store)) {
              *reinterpret_cast<uint4*>(dst) = val;  // also the whole vectors of a tile that straddles n_store
            } else if (nout0 + v * 8 >= p.n_store) {
              // nothing to store
            } else {
              const __half* hv = reinterpret_cast<const __half*>(&val);
              for (int j = 0; j < 8; ++j)
                if (nout0 + v * 8 + j < p.n_store) dst[j] = hv[j];
            }
          }
        }
        epi_bar();  // staging buffer reusable
      }
      if constexpr (Cfg::GROUPS == 2) {
        aphase ^= 1;  // this group's stage completes once per tile of the group
      } else if (++as == Cfg::NACC) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (tst && et == 0) bulk_wait_group<0>();  // the staging tile must outlive the last tensor store
    }  // !p.epi_dma
    if constexpr (Cfg::ROT3) {
      if (tail_w >= 0) {
        // ---- split-K tail: this pair computed k-blocks [tail_kb0, tail_kb1) of tile tail_w
        const int S = p.splitk;
        const int it = n_full_items;  // sequence number of this tile on this pair (TMEM buffer rotation)
        const int m_base = m_tile_of(tail_w) * Cfg::BM;
        const int n_tile = tail_w % p.n_tiles;
        const int m = m_base + row;
        const uint32_t taddr_base = tmem_base + (static_cast<uint32_t>(we * 32) << 16);
        constexpr int NCHUNK = BN / 16;                       // 16-column chunks of the tile
        constexpr size_t PART = size_t(NCHUNK) * Cfg::BM * 16;  // floats of one CTA's partial, laid out [chunk][row][16]
        mbar_wait(&tfull[it & 1], (it >> 1) & 1, 4);
        tc_fence_after();
        float* mine = p.sk_ws + (size_t((tail_j * S + tail_s) * 2 + cta_rank)) * PART;
#pragma unroll 1
        for (int c = half; c < NCHUNK; c += WPQ) {
          const int col = c * 16;
          const uint32_t tcol = static_cast<uint32_t>(((2 * it + (col >= Cfg::MMA_N ? 1 : 0)) % 3) * Cfg::MMA_N +
                                                      (col >= Cfg::MMA_N ? col - Cfg::MMA_N : col));
          uint32_t v[16];
          tmem_ld_x16(taddr_base + tcol, v);
          tmem_ld_wait();
          uint4* dst = reinterpret_cast<uint4*>(mine + (size_t(c) * Cfg::BM + row) * 16);
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        __threadfence();  // partial visible device-wide before the arrival below
        epi_bar();
        unsigned* cnt = p.sk_cnt + (tail_j * 2 + cta_rank) * 2;  // [0] arrivals, [1] departures
        if (et == 0) {
          atomicAdd(cnt, 1u);
          while (*reinterpret_cast<volatile unsigned*>(cnt) < static_cast<unsigned>(S)) __nanosleep(64);
          __threadfence();
        }
        epi_bar();
        // every slice finishes the chunks c = slice, slice + S, ... ; the two warps of a lane quarter alternate
        int k_own = 0;
#pragma unroll 1
        for (int c = tail_s; c < NCHUNK; c += S, ++k_own) {
          if ((k_own % WPQ) != half) continue;
          float y[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) y[j] = 0.f;
          // four slices' vectors in flight at a time (one L2 round trip per group of four, not per slice); added in
          // slice order, so the result does not depend on arrival order
          for (int t0 = 0; t0 < S; t0 += 4) {
            float4 f[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int t = t0 + u < S ? t0 + u : S - 1;  // clamped loads are discarded below
              const float4* src = reinterpret_cast<const float4*>(
                  p.sk_ws + (size_t((tail_j * S + t) * 2 + cta_rank)) * PART + (size_t(c) * Cfg::BM + row) * 16);
#pragma unroll
              for (int q = 0; q < 4; ++q) f[u][q] = __ldcg(src + q);  // written by other SMs: bypass L1
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (t0 + u < S) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  y[4 * q] += f[u][q].x;
                  y[4 * q + 1] += f[u][q].y;
                  y[4 * q + 2] += f[u][q].z;
                  y[4 * q + 3] += f[u][q].w;
                }
              }
            }
          }
          const int col = c * 16;
          const int nout = n_tile * Cfg::NOUT + col;
          if (m < p.M && nout < p.n_store) {
            float b8[8];
            if (p.bias != nullptr) {
#pragma unroll
              for (int hlf = 0; hlf < 2; ++hlf) {
                load8(p.bias + n_tile * BN + col + hlf * 8, b8);
#pragma unroll
                for (int j = 0; j < 8; ++j) y[hlf * 8 + j] += b8[j];
              }
            }
            if (p.rowvec != nullptr) {
              const int rr = ((m / p.rv_hw) / p.rv_div) % p.rv_mod;
              const __half* rv = p.rowvec + static_cast<long long>(rr) * p.rv_ld + n_tile * BN + col;
#pragma unroll
              for (int hlf = 0; hlf < 2; ++hlf) {
                load8(rv + hlf * 8, b8);
#pragma unroll
                for (int j = 0; j < 8; ++j) y[hlf * 8 + j] += b8[j];
              }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) y[j] *= p.alpha;
            const bool full16 = nout + 16 <= p.n_store;
            if (p.R1 != nullptr) {
              const __half* r = p.R1 + static_cast<long long>(m) * p.ldr1 + nout;
              if (full16 && (p.ldr1 & 7) == 0) {
#pragma unroll
                for (int hlf = 0; hlf < 2; ++hlf) {
                  load8(r + hlf * 8, b8);
#pragma unroll
                  for (int j = 0; j < 8; ++j) y[hlf * 8 + j] += p.beta1 * b8[j];
                }
              } else {
                for (int j = 0; j < 16; ++j)
                  if (nout + j < p.n_store) y[j] += p.beta1 * __half2float(r[j]);
              }
            }
            if (p.R2 != nullptr) {
              const __half* r = p.R2 + static_cast<long long>(m) * p.ldr2 + nout;
              if (full16 && (p.ldr2 & 7) == 0) {
#pragma unroll
                for (int hlf = 0; hlf < 2; ++hlf) {
                  load8(r + hlf * 8, b8);
#pragma unroll
                  for (int j = 0; j < 8; ++j) y[hlf * 8 + j] += p.beta2 * b8[j];
                }
              } else {
                for (int j = 0; j < 16; ++j)
                  if (nout + j < p.n_store) y[j] += p.beta2 * __half2float(r[j]);
              }
            }
            __align__(16) __half o[16];
#pragma unroll
            for (int j = 0; j < 16; j += 2) *reinterpret_cast<__half2*>(&o[j]) = __floats2half2_rn(y[j], y[j + 1]);
            __half* dst = p.D + out_row(p, m) * p.ldd + nout;
            if (full16 && (p.ldd & 7) == 0) {
              reinterpret_cast<uint4*>(dst)[0] = reinterpret_cast<const uint4*>(o)[0];
              reinterpret_cast<uint4*>(dst)[1] = reinterpret_cast<const uint4*>(o)[1];
            } else {
              for (int j = 0; j < 16; ++j)
                if (nout + j < p.n_store) dst[j] = o[j];
            }
          }
        }
        // the last CTA through re-arms the counters for the next launch that uses this workspace
        epi_bar();
        if (et == 0) {
          if (atomicAdd(cnt + 1, 1u) == static_cast<unsigned>(S - 1)) {
            cnt[0] = 0u;
            cnt[1] = 0u;
            __threadfence();
          }
        }
      }
    }
  }

  tc_fence_before();
  if constexpr (TWO)
    cluster_sync_all();
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

// ----------------------------------------------------------------------------------------------
// Plain CUDA-core kernel with identical semantics (one thread per output). Bring-up cross-check only.
// ----------------------------------------------------------------------------------------------
struct SimtParams {
  int M, N, K, K1;
  const __half* A;
  long long lda;
  const __half* A2;
  long long lda2;
  int conv, cB, cF, cH, cW, cC, ntaps, cstride, cHin, cWin;
  int8_t taps[SVDPP_MAX_TAPS][4];
  const __half* Wt;
  long long ldw;
  GemmParams ep;  // epilogue fields reused
  int geglu;
  int bn;
};

__device__ __forceinline__ float simt_dot(const SimtParams& p, int m, int n) {
  float acc = 0.f;
  const __half* wrow = p.Wt + static_cast<long long>(n) * p.ldw;
  if (!p.conv) {
    for (int k = 0; k < p.K; ++k) {
      float a = (p.A2 != nullptr && k >= p.K1)
                    ? __half2float(p.A2[static_cast<long long>(m) * p.lda2 + (k - p.K1)])
                    : __half2float(p.A[static_cast<long long>(m) * p.lda + k]);
      acc += a * __half2float(wrow[k]);
    }
  } else {
    int w = m % p.cW;
    int r = m / p.cW;
    int h = r % p.cH;
    int img = r / p.cH;
    int f = img % p.cF;
    int b = img / p.cF;
    for (int t = 0; t < p.ntaps; ++t) {
      int ww = w * p.cstride + p.taps[t][0], hh = h * p.cstride + p.taps[t][1], ff = f + p.taps[t][2];
      if (ww < 0 || ww >= p.cWin || hh < 0 || hh >= p.cHin || ff < 0 || ff >= p.cF) continue;
      const __half* src = p.A + ((((static_cast<long long>(b) * p.cF + ff) * p.cHin + hh) * p.cWin) + ww) * p.cC;
      const __half* wk = wrow + t * p.cC;
      for (int c = 0; c < p.cC; ++c) acc += __half2float(src[c]) * __half2float(wk[c]);
    }
  }
  return acc;
}

__global__ void gemm_simt_kernel(const SimtParams p) {
  const int nout_total = p.geglu ? p.N / 2 : p.N;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(p.M) * nout_total) return;
  const int m = static_cast<int>(idx / nout_total);
  const int nout = static_cast<int>(idx % nout_total);
  if (nout >= p.ep.n_store) return;
  const GemmParams& e = p.ep;
  __half o;
  if (p.geglu) {
    const int half_bn = p.bn / 2;
    const int nt = nout / half_bn, j = nout % half_bn;
    const int nv = nt * p.bn + j, ng = nv + half_bn;
    float val = simt_dot(p, m, nv), gate = simt_dot(p, m, ng);
    if (e.bias) {
      val += __half2float(e.bias[nv]);
      gate += __half2float(e.bias[ng]);
    }
    o = geglu_fp16(val, gate);
  } else {
    float y = simt_dot(p, m, nout);
    if (e.bias) y += __half2float(e.bias[nout]);
    if (e.rowvec) {
      const int rr = ((m / e.rv_hw) / e.rv_div) % e.rv_mod;
      y += __half2float(e.rowvec[static_cast<long long>(rr) * e.rv_ld + nout]);
    }
    y *= e.alpha;
    if (e.R1) y += e.beta1 * __half2float(e.R1[static_cast<long long>(m) * e.ldr1 + nout]);
    if (e.R2) y += e.beta2 * __half2float(e.R2[static_cast<long long>(m) * e.ldr2 + nout]);
    o = __float2half_rn(y);
  }
  e.D[out_row(e, m) * e.ldd + nout] = o;
}

template <int BN, bool GEGLU, bool TWO, int EW_ = 0>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB,
                     const GemmParams& p_in, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, GEGLU, TWO, EW_>;
  static bool configured = false;
  auto kern = gemm_tc_kernel<BN, GEGLU, TWO, EW_>;
  if (!configured) {
    SVDPP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  GemmParams p = p_in;
  // Output through TMA tensor stores: rows of 32 columns (64 bytes) per box, 64-byte swizzle; the tensor map's
  // extents [n_store, M] clip the N padding and the M tail.  Not for the sub-pixel output map (rows scatter), for
  // row pitches that are not 16-byte multiples, or for the 80-column GEGLU round of the 160-wide tile.
  CUtensorMap tmD = tmB, tmR = tmB;
  p.tma_store = 0;
  p.tma_r1 = 0;
  p.epi_dma = 0;
  if (tuning().tma_store && Cfg::SW % 32 == 0 && p.up <= 1 && (p.ldd & 7) == 0 &&
      (reinterpret_cast<uintptr_t>(p.D) & 15) == 0) {
    static_assert(Cfg::SW % 32 != 0 || (Cfg::SW / 32) * 8192 <= Cfg::C_BYTES, "swizzled staging tile fits");
    uint64_t dims[2] = {static_cast<uint64_t>(p.n_store), static_cast<uint64_t>(p.M)};
    uint64_t str[1] = {static_cast<uint64_t>(p.ldd) * 2};
    uint32_t box[2] = {32, 128};
    if (encode_tmap_f16(&tmD, p.D, 2, dims, str, box, nullptr, 64)) return -5;
    p.tma_store = 1;
    if (tuning().tma_r1 && !GEGLU && p.R1 != nullptr && (p.ldr1 & 7) == 0 && (reinterpret_cast<uintptr_t>(p.R1) & 15) == 0) {
      uint64_t rstr[1] = {static_cast<uint64_t>(p.ldr1) * 2};
      if (encode_tmap_f16(&tmR, p.R1, 2, dims, rstr, box, nullptr, 64)) return -5;
      p.tma_r1 = 1;
    }
    // DMA-lane epilogue: measured (tools/ab_lib.py) +19.5 % on the level-0 qkv projection (256x256 pair tiles, K = 320,
    // 4 of 5 stages left), but 0.80-0.97x where the second staging tile costs TWO stages (160/320-wide tiles: 3 left)
    // and 0.98x at K = 640 - so by default only where one stage is given up and the main loop is at most 5 k-blocks.
    // epi_dma = 2 forces it wherever it is possible (tests).
    const bool dma_ok = !GEGLU && Cfg::GROUPS == 1 && (p.R1 == nullptr || p.tma_r1) && Cfg::STAGES - Cfg::STEAL >= 2;
    if (dma_ok && (tuning().epi_dma >= 2 || (tuning().epi_dma == 1 && Cfg::STEAL == 1 && p.num_kb <= tuning().epi_dma_max_kb)))
      p.epi_dma = 1;
  }
  p.two_prod = (tuning().two_prod && !p.epi_dma && p.num_kb >= tuning().two_prod_min_kb) ? 1 : 0;
  if constexpr (TWO) {
    const int pairs = ((p.m_tiles + 1) / 2) * p.n_tiles;
    int clusters = num_sms() / 2;
    if (pairs < clusters) clusters = pairs;
    p.splitk = 0;
    if constexpr (Cfg::ROT3) {
      // split-K tail: r tiles left for a last wave of `clusters` pairs -> each tile on S = clusters / r pairs
      const int full_waves = pairs / clusters, r = pairs % clusters;
      // measured (tools/tail_probe.py): the tail machinery (slice pipeline start, partial dump, cross-CTA wait, fix-up)
      // costs about 20-25 us whatever K is, an unsplit last wave one tile time (~0.5 us per k-block): K = 5760 at level 1
      // 434 -> 393 us, K = 2880 379 -> 369 us, K <= 2560 slower - hence the K threshold
      if (tuning().splitk && p.sk_ws != nullptr && r > 0 && full_waves >= 1 && p.num_kb >= tuning().splitk_min_total_kb) {
        int S = clusters / r;
        const int min_kb = tuning().splitk_min_kb > 0 ? tuning().splitk_min_kb : 1;
        if (S > p.num_kb / min_kb) S = p.num_kb / min_kb;
        if (S > BN / 16) S = BN / 16;
        const long long part_bytes = 2LL * (BN / 16) * Cfg::BM * 16 * 4;  // both CTAs of a pair
        if (S >= 2 && 4096 + static_cast<long long>(r) * S * part_bytes <= p.sk_ws_bytes &&
            static_cast<size_t>(r) * 4 * sizeof(unsigned) <= 4096) {
          p.splitk = S;
          // which tiles land in the split last wave depends on the walk order, and a split tile sums its k-slices in
          // another association than an unsplit one: keep the forward walk so results do not depend on "reverse"
          p.rev_m = 0;
          p.sk_r = r;
          p.sk_full = full_waves * clusters;
          p.sk_cnt = reinterpret_cast<unsigned*>(p.sk_ws);
          p.sk_ws = reinterpret_cast<float*>(reinterpret_cast<char*>(p.sk_ws) + 4096);
        }
      }
    }
    SVDPP_CUDA(launch_kernel(kern, dim3(2 * clusters), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, 2, tmA, tmA2, tmB,
                             tmD, tmR, p));
    return check_launch("gemm_tc_kernel<pair>");
  }
  const int total = p.m_tiles * p.n_tiles;
  const int grid = total < num_sms() ? total : num_sms();
  SVDPP_CUDA(launch_kernel(kern, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, 1, tmA, tmA2, tmB, tmD, tmR, p));
  return check_launch("gemm_tc_kernel");
}

}  // namespace svdpp

using namespace svdpp;

extern "C" int svdpp_gemm_f16(const svdpp_gemm_desc* d, int impl, svdpp_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SVDPP_CHECK_ARG(d != nullptr, "gemm: null descriptor");
  // impl 3: CTA pairs with 256-wide tiles (N must be a multiple of 256; GEGLU weights interleaved per 128)
  // impl 4: one CTA per 128x128 tile (N a multiple of 128; no GEGLU) - for small M, where 128x128 tiles fill the
  //         148 SMs' last wave better than 128x160 or 256x256
  // impl 5: impl 3 with 8 instead of 16 GEGLU epilogue warps (A/B measurements only)
  // impl 6: CTA pairs with 256x320 tiles, one accumulator stage (N a multiple of 320, no GEGLU)
  // impl 7: CTA pairs with 256x128 tiles (N a multiple of 128, no GEGLU): impl 4's width with half the B traffic per SM -
  //         for the N = 128 layers at very large M (the VAE decoder's full-resolution convolutions)
  const int BN = (impl == 3 || impl == 5) ? 256 : ((impl == 4 || impl == 7) ? 128 : (impl == 6 ? 320 : 160));
  SVDPP_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "gemm: bad shape M=%d N=%d K=%d", d->M, d->N, d->K);
  SVDPP_CHECK_ARG(d->K % 64 == 0, "gemm: K=%d must be a multiple of 64", d->K);
  SVDPP_CHECK_ARG(d->N % BN == 0, "gemm: N=%d must be a multiple of %d (pad the weight)", d->N, BN);
  SVDPP_CHECK_ARG(d->A && d->Wt && d->D, "gemm: null A/Wt/D");
  SVDPP_CHECK_ARG(!(d->geglu && (d->rowvec || d->R1 || d->R2)), "gemm: geglu epilogue takes bias only");
  const int nout_total = d->geglu ? d->N / 2 : d->N;
  const int n_store = d->n_store > 0 ? (d->n_store < nout_total ? d->n_store : nout_total) : nout_total;

  GemmParams p{};
  p.M = d->M;
  p.N = d->N;
  p.num_kb = d->K / 64;
  p.kb_split = p.num_kb;
  p.conv = d->conv;
  p.bias = static_cast<const __half*>(d->bias);
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
  p.n_store = n_store;
  p.sk_ws = static_cast<float*>(d->splitk_ws);
  p.sk_ws_bytes = d->splitk_ws != nullptr ? d->splitk_ws_bytes : 0;
  p.m_tiles = (d->M + 127) / 128;
  p.n_tiles = d->N / BN;
  p.rev_m = tuning().reverse == 1 ? 1 : 0;
  {
    // measured with the per-thread residual loads: K=320 +20 %, 640 +6 %, 1280 -3 %; with the residual on TMA loads the
    // prefetch no longer pays at K = 640 (0.97x) and costs 1-5 % beyond, so it is kept for K <= 320 only
    p.prefetch_r1 = (p.R1 != nullptr && p.num_kb <= tuning().r1_prefetch_max_kb) ? 1 : 0;
  }

  if (d->conv) {
    SVDPP_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= SVDPP_MAX_TAPS, "gemm: ntaps=%d", d->ntaps);
    SVDPP_CHECK_ARG(d->K == d->ntaps * d->cC, "gemm: conv K=%d != ntaps*C=%d", d->K, d->ntaps * d->cC);
    SVDPP_CHECK_ARG(static_cast<long long>(d->cB) * d->cF * d->cH * d->cW == d->M, "gemm: conv dims do not give M");
    SVDPP_CHECK_ARG(d->A2 == nullptr, "gemm: conv mode takes one source");
    for (int t = 0; t < d->ntaps; ++t)
      for (int j = 0; j < 4; ++j) p.taps[t][j] = d->taps[t][j];
    p.cF = d->cF;
    p.cH = d->cH;
    p.cW = d->cW;
    p.up = d->out_up > 1 ? d->out_up : 1;
    p.up_y = d->out_up_y;
    p.up_x = d->out_up_x;
    SVDPP_CHECK_ARG(p.up <= 2 && p.up_y >= 0 && p.up_y < p.up && p.up_x >= 0 && p.up_x < p.up, "gemm: bad sub-pixel output map");
    SVDPP_CHECK_ARG(p.up == 1 || !(d->R1 || d->R2), "gemm: the sub-pixel output map takes bias/rowvec epilogues only");
    p.cstride = d->conv_stride > 1 ? d->conv_stride : 1;
    SVDPP_CHECK_ARG(p.cstride <= 2, "gemm: conv_stride=%d unsupported", d->conv_stride);
    SVDPP_CHECK_ARG(p.cstride == 1 || (d->cHin >= d->cH && d->cWin >= d->cW), "gemm: strided conv needs cHin/cWin");
  }

  if (impl == 3 || impl == 5) {
    SVDPP_CHECK_ARG(d->N % 256 == 0, "gemm: impl 3 needs N %% 256 == 0 (N=%d)", d->N);
  }
  if (impl == 1) {
    SimtParams sp{};
    sp.M = d->M;
    sp.N = d->N;
    sp.K = d->K;
    sp.K1 = d->A2 ? d->K1 : d->K;
    sp.A = static_cast<const __half*>(d->A);
    sp.lda = d->lda;
    sp.A2 = static_cast<const __half*>(d->A2);
    sp.lda2 = d->lda2;
    sp.conv = d->conv;
    sp.cB = d->cB;
    sp.cF = d->cF;
    sp.cH = d->cH;
    sp.cW = d->cW;
    sp.cstride = d->conv_stride > 1 ? d->conv_stride : 1;
    sp.cHin = sp.cstride > 1 ? d->cHin : d->cH;
    sp.cWin = sp.cstride > 1 ? d->cWin : d->cW;
    sp.cC = d->cC;
    sp.ntaps = d->ntaps;
    for (int t = 0; t < SVDPP_MAX_TAPS; ++t)
      for (int j = 0; j < 4; ++j) sp.taps[t][j] = d->taps[t][j];
    sp.Wt = static_cast<const __half*>(d->Wt);
    sp.ldw = d->ldw;
    sp.ep = p;
    sp.geglu = d->geglu;
    sp.bn = BN;
    const long long total = static_cast<long long>(d->M) * nout_total;
    const int threads = 256;
    const long long blocks = (total + threads - 1) / threads;
    gemm_simt_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(sp);
    return check_launch("gemm_simt_kernel");
  }
  SVDPP_CHECK_ARG(impl == 0 || (impl >= 2 && impl <= 7), "gemm: unknown impl %d", impl);
  SVDPP_CHECK_ARG(!(impl == 7 && d->geglu), "gemm: impl 7 has no GEGLU epilogue");
  SVDPP_CHECK_ARG(!(impl == 6 && d->geglu), "gemm: impl 6 has no GEGLU epilogue");
  SVDPP_CHECK_ARG(!(impl == 4 && d->geglu), "gemm: impl 4 has no GEGLU epilogue");
  const bool two = impl == 2 || impl == 3 || impl == 5 || impl == 6 || impl == 7;  // CTA pairs (cta_group::2)

  CUtensorMap tmA, tmA2, tmB;
  if (!d->conv) {
    const int K1 = d->A2 ? d->K1 : d->K;
    SVDPP_CHECK_ARG(K1 > 0 && K1 % 64 == 0 && K1 <= d->K, "gemm: K1=%d must be a multiple of 64", K1);
    SVDPP_CHECK_ARG(d->lda % 8 == 0, "gemm: lda=%lld must be a multiple of 8", (long long)d->lda);
    p.kb_split = K1 / 64;
    uint64_t dims[2] = {static_cast<uint64_t>(K1), static_cast<uint64_t>(d->M)};
    uint64_t str[1] = {static_cast<uint64_t>(d->lda) * 2};
    uint32_t box[2] = {64, 128};
    if (encode_tmap_f16(&tmA, d->A, 2, dims, str, box)) return -5;
    if (d->A2) {
      SVDPP_CHECK_ARG(d->lda2 % 8 == 0, "gemm: lda2 must be a multiple of 8");
      uint64_t dims2[2] = {static_cast<uint64_t>(d->K - K1), static_cast<uint64_t>(d->M)};
      uint64_t str2[1] = {static_cast<uint64_t>(d->lda2) * 2};
      if (encode_tmap_f16(&tmA2, d->A2, 2, dims2, str2, box)) return -5;
    } else {
      tmA2 = tmA;
    }
  } else {
    const int W = d->cW;
    SVDPP_CHECK_ARG(d->cC % 64 == 0, "gemm: conv C=%d must be a multiple of 64", d->cC);
    SVDPP_CHECK_ARG((W >= 8 && 128 % W == 0) || W % 128 == 0,
                    "gemm: conv width %d not tileable by the window path (use svdpp_im2col_nhwc)", W);
    p.bw = W < 128 ? W : 128;
    p.nrows = 128 / p.bw;
    p.cpk = d->cC / 64;
    const int s = p.cstride;
    const int Win = s > 1 ? d->cWin : d->cW, Hin = s > 1 ? d->cHin : d->cH;
    uint64_t dims[5] = {static_cast<uint64_t>(d->cC), static_cast<uint64_t>(Win), static_cast<uint64_t>(Hin),
                        static_cast<uint64_t>(d->cF), static_cast<uint64_t>(d->cB)};
    uint64_t str[4];
    str[0] = static_cast<uint64_t>(d->cC) * 2;
    str[1] = str[0] * Win;
    str[2] = str[1] * Hin;
    str[3] = str[2] * d->cF;
    // strided windows: the box spans (bw - 1) * s + 1 input pixels and is traversed with element stride s, which
    // loads exactly bw pixels
    uint32_t box[5] = {64, static_cast<uint32_t>((p.bw - 1) * s + 1), 1, 1, 1};
    uint32_t estr[5] = {1, static_cast<uint32_t>(s), 1, 1, 1};
    SVDPP_CHECK_ARG(box[1] <= 256, "gemm: strided conv box too wide");
    if (encode_tmap_f16(&tmA, d->A, 5, dims, str, box, s > 1 ? estr : nullptr)) return -5;
    tmA2 = tmA;
  }
  {
    SVDPP_CHECK_ARG(d->ldw % 8 == 0, "gemm: ldw must be a multiple of 8");
    uint64_t dims[2] = {static_cast<uint64_t>(d->K), static_cast<uint64_t>(d->N)};
    uint64_t str[1] = {static_cast<uint64_t>(d->ldw) * 2};
    uint32_t box[2] = {64, static_cast<uint32_t>(impl == 6 ? 80 : (two ? BN / 2 : BN))};
    if (encode_tmap_f16(&tmB, d->Wt, 2, dims, str, box)) return -5;
  }
  if (impl == 4) return launch_tc<128, false, false>(tmA, tmA2, tmB, p, stream);
  if (impl == 7) return launch_tc<128, false, true>(tmA, tmA2, tmB, p, stream);
  if (impl == 6) return launch_tc<320, false, true>(tmA, tmA2, tmB, p, stream);
  if (impl == 5 && d->geglu) return launch_tc<256, true, true, 8>(tmA, tmA2, tmB, p, stream);
  if (impl == 5) return launch_tc<256, false, true>(tmA, tmA2, tmB, p, stream);
  if (impl == 3) {
    // 16 epilogue warps in two independent groups pay off only where the epilogue outweighs the main loop
    // (K = 320: 863 -> 957 TFLOP/s; K >= 640: no gain)
    if (d->geglu && d->K > 384) return launch_tc<256, true, true, 8>(tmA, tmA2, tmB, p, stream);
    if (d->geglu) return launch_tc<256, true, true>(tmA, tmA2, tmB, p, stream);
    return launch_tc<256, false, true>(tmA, tmA2, tmB, p, stream);
  }
  if (two) {
    if (d->geglu) return launch_tc<160, true, true>(tmA, tmA2, tmB, p, stream);
    return launch_tc<160, false, true>(tmA, tmA2, tmB, p, stream);
  }
  if (d->geglu) return launch_tc<160, true, false>(tmA, tmA2, tmB, p, stream);
  return launch_tc<160, false, false>(tmA, tmA2, tmB, p, stream);
}
