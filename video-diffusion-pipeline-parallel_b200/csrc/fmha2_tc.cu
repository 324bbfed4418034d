// svdpp_attn_spatial_f16, impl 2: spatial self-attention, head_dim 64, two 128-query tiles per CTA.
//
// One CTA per (256 queries, head, image), one CTA per SM, key blocks of 128:
//   warps 0..3   softmax of query tile 0      thread = query row = TMEM lane: the whole S row (128 fp32) is read
//   warps 4..7   softmax of query tile 1      ONCE into registers (the S buffer is released right there), online
//                                             softmax in the log2 domain, P goes back into TMEM as packed fp16
//   warp 8       TMA producer: Q tiles once, then K_j / V_j (128 keys x 64) through 3-slot rings
//   warps 9, 10  MMA issuers, one per tile (warp 9 also owns the TMEM allocation): S_t = Q_t K_j^T (A, B from
//                smem) as soon as the softmax warps have taken S_t(j-1) into registers, and O_t += P_t V_j with
//                P_t read straight from TMEM (tcgen05.mma with the A operand in tensor memory) - no smem round
//                trip, no proxy fence, no swizzled st.shared in the softmax loop
// TMEM (512 columns): S0 [0,128)  S1 [128,256)  O0 [256,320)  O1 [320,384)  P0 [384,448)  P1 [448,512).
// P has its own columns so that S_t(j+1) is computed WHILE the softmax warps exponentiate block j: they never
// wait for the tensor core in steady state, and the rows' exponentials keep the MUFU pipe (16 ex2/clk/SM, the
// bound of a head_dim-64 FMHA) busy.
// O stays in TMEM for the whole key loop and is rescaled lazily: the reference maximum of a row is raised only
// when a block's maximum exceeds it by more than 2^8, so probabilities stay <= 256 (exact in fp16).
#include <cuda_fp16.h>
#include <math_constants.h>

#include "common.h"
#include "ptx.cuh"

namespace svdpp {

struct Attn2Params {
  int S, n_kv;
  int q_off, k_off, v_off;
  float scale_log2;
  __half* out;
  long long ldo;
  int stagger;  // clocks by which query tile 1 starts behind tile 0 (softmax phases of the two warps of a scheduler interleave)
  unsigned int* trace;  // debug: phase timestamps of softmax warps 0 and 4 of CTA (0,0,0), [2][n_kv][8] (NULL: off)
};
unsigned int* g_attn_trace = nullptr;  // shared with fmha3_tc.cu

constexpr int A2_BQ = 128;                    // rows per query tile (two tiles per CTA)
constexpr int A2_BK = 128;                    // keys per block
constexpr int A2_Q_BYTES = A2_BQ * 64 * 2;    // 16 KB
constexpr int A2_KV_BYTES = A2_BK * 64 * 2;   // 16 KB
constexpr int A2_STAGES = 3;
constexpr int A2_SMEM_BYTES = 2 * A2_Q_BYTES + 2 * A2_STAGES * A2_KV_BYTES + 1024 /*align*/ + 256 /*barriers*/ +
                              2 * 2 * 2 * 128 * 4 /*row-max exchange*/;
// SPLIT = threads per query row: 1 -> 8 softmax warps (thread = row, 128 columns), 2 -> 16 softmax warps (two
// threads per row, 64 columns each, row maximum exchanged through smem + a 64-thread named barrier): four
// instead of two softmax warps per scheduler hide each other's TMEM-load / max / store phases, so the MUFU
// pipe idles less.
template <int SPLIT>
struct A2Cfg {
  static constexpr int SM_WARPS = 8 * SPLIT;        // softmax warps
  static constexpr int TMA_WARP = SM_WARPS;
  static constexpr int MMA_WARP0 = SM_WARPS + 1;    // + tile index
  static constexpr int THREADS = (SM_WARPS + 3) * 32;
  static constexpr int COLS = A2_BK / SPLIT;        // S columns per softmax thread
};

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 lanes x K) is read from tensor memory, two fp16 per column
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ uint64_t pack_f2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack_f2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Measured dead end: evaluating every fourth (or eighth) exponential with a degree-3 polynomial on the FMA pipe
// (poly3_exp2) to relieve the MUFU pipe was SLOWER (774 vs 799 TFLOP/s; with the packed-arithmetic loop 750 vs 805):
// the softmax warps run out of issue slots before the MUFU pipe (73 % busy) saturates.
// QP = 1 (SPLIT == 1 only): quarter-pipelined softmax, see the softmax branch.  POLY = n > 0: one group of four
// exponentials in every n is evaluated on the FMA pipe (poly3_exp2) instead of MUFU.
template <int SPLIT, int QP, int POLY>
__global__ void __launch_bounds__(A2Cfg<SPLIT>::THREADS, 1)
attn_spatial2_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                        const Attn2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [2][128][64]
  uint8_t* sK = sQ + 2 * A2_Q_BYTES;                    // [STAGES][128][64]
  uint8_t* sV = sK + A2_STAGES * A2_KV_BYTES;           // [STAGES][128][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + A2_STAGES * A2_KV_BYTES);
  uint64_t* q_full = bars;                   // [1]
  uint64_t* k_full = q_full + 1;             // [STAGES]
  uint64_t* k_empty = k_full + A2_STAGES;    // [STAGES]
  uint64_t* v_full = k_empty + A2_STAGES;    // [STAGES]
  uint64_t* v_empty = v_full + A2_STAGES;    // [STAGES]
  uint64_t* s_full = v_empty + A2_STAGES;    // [2]  S_t(j) complete
  uint64_t* p_ready = s_full + 2;            // [2]  P_t(j) stored (128 arrivals)
  uint64_t* pv_done = p_ready + 2;           // [2]  O_t += P_t(j) V_j retired
  uint64_t* s_free = pv_done + 2;            // [2]  S_t(j) is in registers (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);
  float* xch = reinterpret_cast<float*>(tmem_slot + 2);  // [2 parity][2 tiles][2 halves][128 rows] (SPLIT == 2)
  using Cfg = A2Cfg<SPLIT>;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * A2_BQ);
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const int row_base = img * p.S;  // first token row of this image in the qkv matrix

  if (warp == Cfg::TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < A2_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 2);  // both tiles' MMA warps
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 2);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_ready[t], 128 * SPLIT);
      mbar_init(&pv_done[t], 1);
      mbar_init(&s_free[t], 128 * SPLIT);
    }
    fence_mbar_init();
  }
  if (warp == Cfg::MMA_WARP0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();  // programmatic dependent launch: see ptx.cuh
  pdl_wait();
  const bool trace_on = p.trace != nullptr && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && lane == 0 &&
                        warp < Cfg::SM_WARPS && (warp & 3) == 0 && (SPLIT == 1 || (warp >> 2 & 1) == 0);
  unsigned int* const trace = trace_on ? p.trace + (warp / (4 * SPLIT)) * (p.n_kv * 8) : nullptr;
#define SVDPP_TR(j, ev)                                          \
  do {                                                           \
    if (trace_on) trace[(j) * 8 + (ev)] = static_cast<unsigned int>(clock64()); \
  } while (0)

  if (warp == Cfg::TMA_WARP) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * A2_Q_BYTES);
      tma_load_2d(sQ, &tmQ, q_full, p.q_off + head * 64, row_base + q0);
      tma_load_2d(sQ + A2_Q_BYTES, &tmQ, q_full, p.q_off + head * 64, row_base + q0 + A2_BQ);
      for (int j = 0; j < p.n_kv; ++j) {
        const int s = j % A2_STAGES;
        const uint32_t ph = (j / A2_STAGES) & 1;
        mbar_wait(&k_empty[s], ph ^ 1, 41);
        mbar_expect_tx(&k_full[s], A2_KV_BYTES);
        tma_load_2d(sK + s * A2_KV_BYTES, &tmKV, &k_full[s], p.k_off + head * 64, row_base + j * A2_BK);
        mbar_wait(&v_empty[s], ph ^ 1, 42);
        mbar_expect_tx(&v_full[s], A2_KV_BYTES);
        tma_load_2d(sV + s * A2_KV_BYTES, &tmKV, &v_full[s], p.v_off + head * 64, row_base + j * A2_BK);
      }
    }
  } else if (warp >= Cfg::MMA_WARP0) {
    // ------------------------------------------------------------------ MMA issuers (one per query tile)
    if (lane == 0) {
      const int t = warp - Cfg::MMA_WARP0;
      constexpr uint32_t idesc_s = make_idesc_f16(A2_BK, false);  // S: N = 128 keys, K-major B
      constexpr uint32_t idesc_o = make_idesc_f16(64, true);      // O: N = 64 dims, V read MN-major
      const uint32_t q_addr = smem_u32(sQ + t * A2_Q_BYTES);
      const uint32_t d_s = tmem_base + t * 128;
      const uint32_t d_o = tmem_base + 256 + t * 64;
      const uint32_t a_p = tmem_base + 384 + t * 64;
      auto issue_s = [&](int jj) {  // S_t = Q_t K_jj^T
        const int s = jj % A2_STAGES;
        mbar_wait(&k_full[s], (jj / A2_STAGES) & 1, 44);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + s * A2_KV_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(d_s, make_smem_desc_sw128(q_addr + k * 32, 1024, 0), make_smem_desc_sw128(k_addr + k * 32, 1024, 0),
                   idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[t]);
        umma_commit(&k_empty[s]);
      };
      mbar_wait(q_full, 0, 43);
      if (t == 1 && p.stagger > 0) {
        // Phase offset between the two query tiles.  Warp w (tile 0) and warp w + 4 (tile 1) share a scheduler and its
        // MUFU lanes; started together they stay in lock-step (both in the load / max / store phases at the same time,
        // MUFU idle).  An offset, once there, persists: whoever is alone in its exp phase runs at full MUFU rate.
        const long long c0 = clock64();
        while (clock64() - c0 < p.stagger) {
        }
      }
      issue_s(0);
      for (int j = 0; j < p.n_kv; ++j) {
        const int s = j % A2_STAGES;
        if (j + 1 < p.n_kv) {
          mbar_wait(&s_free[t], j & 1, 46);  // S_t(j) is in the softmax warps' registers
          issue_s(j + 1);
        }
        mbar_wait(&v_full[s], (j / A2_STAGES) & 1, 45);
        mbar_wait(&p_ready[t], j & 1, 47);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sV + s * A2_KV_BYTES);
#pragma unroll
        for (int k = 0; k < A2_BK / 16; ++k)  // O_t += P_t V_j, P_t (two fp16 per column) from TMEM
          umma_f16_ts(d_o, a_p + k * 8, make_smem_desc_sw128(v_addr + k * 2048, 1024, 8192), idesc_o,
                      (j | k) != 0 ? 1u : 0u);
        umma_commit(&pv_done[t]);
        umma_commit(&v_empty[s]);
      }
    }
  } else if constexpr (QP == 1) {
    // ------------------------------------------------------------------ softmax, quarter-pipelined (SPLIT == 1)
    // The 128 scores of a row are treated as four online-softmax blocks of 32 keys: the row maximum of quarter q + 1
    // (FMNMX3, ALU pipe) has no dependence on the exponentials of quarter q (MUFU), so ptxas interleaves the two in one
    // basic block and the max pass disappears from the critical path (it was a serial 64-deep chain, 15 % of the
    // softmax warps' samples in ncu).  Each quarter's probabilities go to TMEM as soon as they are packed.
    static_assert(SPLIT == 1, "quarter pipeline is written for one thread per row");
    const int t = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tmem_S = tmem_base + t * 128 + lane_sel;
    const uint32_t tmem_O = tmem_base + 256 + t * 64 + lane_sel;
    const uint32_t tmem_P = tmem_base + 384 + t * 64 + lane_sel;
    const uint64_t scale2 = pack_f2(p.scale_log2, p.scale_log2);
    float m_used = -CUDART_INF_F;
    float l_run = 0.f;
    auto qmax = [](const uint32_t(&x)[32]) {
      float a = fmax3(__uint_as_float(x[0]), __uint_as_float(x[1]), __uint_as_float(x[2]));
      float b = fmax3(__uint_as_float(x[3]), __uint_as_float(x[4]), __uint_as_float(x[5]));
      float c = fmax3(__uint_as_float(x[6]), __uint_as_float(x[7]), __uint_as_float(x[8]));
      float d = fmax3(__uint_as_float(x[9]), __uint_as_float(x[10]), __uint_as_float(x[11]));
      a = fmax3(a, __uint_as_float(x[12]), __uint_as_float(x[13]));
      b = fmax3(b, __uint_as_float(x[14]), __uint_as_float(x[15]));
      c = fmax3(c, __uint_as_float(x[16]), __uint_as_float(x[17]));
      d = fmax3(d, __uint_as_float(x[18]), __uint_as_float(x[19]));
      a = fmax3(a, __uint_as_float(x[20]), __uint_as_float(x[21]));
      b = fmax3(b, __uint_as_float(x[22]), __uint_as_float(x[23]));
      c = fmax3(c, __uint_as_float(x[24]), __uint_as_float(x[25]));
      d = fmax3(d, __uint_as_float(x[26]), __uint_as_float(x[27]));
      a = fmax3(a, __uint_as_float(x[28]), __uint_as_float(x[29]));
      b = fmax3(b, __uint_as_float(x[30]), __uint_as_float(x[31]));
      return fmaxf(fmaxf(a, b), fmaxf(c, d));
    };
    for (int j = 0; j < p.n_kv; ++j) {
      const int valid = p.S - j * A2_BK;  // my columns [0, valid) of this block are real keys
      SVDPP_TR(j, 0);
      mbar_wait(&s_full[t], j & 1, 49);
      tc_fence_after();
      SVDPP_TR(j, 1);
      uint32_t v[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_x32(tmem_S + c * 32, v[c]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[t]);  // the tensor core may overwrite S_t with block j+1 now
      SVDPP_TR(j, 2);
      if (valid < A2_BK) {      // warp-uniform: only the last key block of an image can be partial
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= valid) v[c][i] = 0xff800000u;  // -inf
      }
      float mq = qmax(v[0]);
      uint32_t pk[4][16];
      constexpr int FIRST_ST = 2;  // P(j) is first stored after this quarter: PV(j-1), issued at the end of block j-1,
                                   // has long retired by then (waiting for it after quarter 0 stalled the warp)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float m_blk = mq * p.scale_log2;
        bool raise = false;
        if (j == 0 && q == 0)
          m_used = m_blk;
        else
          raise = m_blk > m_used + 8.0f;
        if (__any_sync(0xffffffffu, raise)) {
          // rare: the reference maximum moves.  Everything accumulated under the old one is rescaled: O (every MMA that
          // has touched it must have retired), the row sum, and the quarters of P(j) already stored for the pending PV.
          if (j > 0) {
            mbar_wait(&pv_done[t], (j - 1) & 1, 50);
            tc_fence_after();
          }
          const float m_new = raise ? m_blk : m_used;
          const float alpha = fast_exp2(m_used - m_new);
          m_used = m_new;
          l_run *= alpha;
          if (j > 0) {
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              uint32_t o[16];
              tmem_ld_x16(tmem_O + c * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_x16(tmem_O + c * 16, o);
            }
          }
          if (q > 0) {  // quarters of P(j) computed under the old maximum: in registers up to FIRST_ST, in TMEM after
            const __half2 a2 = __float2half2_rn(alpha);
            if (q <= FIRST_ST) {
#pragma unroll
              for (int c = 0; c < q; ++c)
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  __half2 h = __hmul2(*reinterpret_cast<__half2*>(&pk[c][i]), a2);
                  pk[c][i] = *reinterpret_cast<uint32_t*>(&h);
                }
            } else {
              tmem_st_wait();
#pragma unroll 1
              for (int c = 0; c < q; ++c) {
                uint32_t o[16];
                tmem_ld_x16(tmem_P + c * 16, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  __half2 h = __hmul2(*reinterpret_cast<__half2*>(&o[i]), a2);
                  o[i] = *reinterpret_cast<uint32_t*>(&h);
                }
                tmem_st_x16(tmem_P + c * 16, o);
              }
            }
          }
          tmem_st_wait();
        }
        if (q == 0) SVDPP_TR(j, 3);
        // the next quarter's maximum: independent of this quarter's exponentials, scheduled in between them
        if (q < 3) mq = qmax(v[q + 1 < 4 ? q + 1 : 3]);
        const float neg_m = -m_used;
        const uint64_t negm2 = pack_f2(neg_m, neg_m);
        uint64_t lsa = 0ull, lsb = 0ull;  // (0.f, 0.f)
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const uint64_t xa = ffma2(pack_f2(__uint_as_float(v[q][2 * i]), __uint_as_float(v[q][2 * i + 1])), scale2, negm2);
          const uint64_t xb = ffma2(pack_f2(__uint_as_float(v[q][2 * i + 2]), __uint_as_float(v[q][2 * i + 3])), scale2, negm2);
          float x0, x1, x2, x3;
          unpack_f2(xa, x0, x1);
          unpack_f2(xb, x2, x3);
          float p0, p1, p2, p3;
          if (POLY > 0 && (i / 2) % (POLY > 0 ? POLY : 1) == (POLY > 0 ? POLY : 1) - 1) {
            p0 = poly3_exp2(x0), p1 = poly3_exp2(x1), p2 = poly3_exp2(x2), p3 = poly3_exp2(x3);
          } else {
            p0 = fast_exp2(x0), p1 = fast_exp2(x1), p2 = fast_exp2(x2), p3 = fast_exp2(x3);
          }
          lsa = fadd2(lsa, pack_f2(p0, p1));
          lsb = fadd2(lsb, pack_f2(p2, p3));
          pk[q][i] = pack_half2(p0, p1);
          pk[q][i + 1] = pack_half2(p2, p3);
        }
        float a0, a1, b0, b1;
        unpack_f2(lsa, a0, a1);
        unpack_f2(lsb, b0, b1);
        l_run += (a0 + a1) + (b0 + b1);
        if (q == 0) SVDPP_TR(j, 4);
        if (q == FIRST_ST) {
          if (j > 0) {  // P_t(j-1) must have been consumed before its columns are rewritten
            mbar_wait(&pv_done[t], (j - 1) & 1, 52);
            tc_fence_after();
          }
          SVDPP_TR(j, 5);
#pragma unroll
          for (int c = 0; c <= FIRST_ST; ++c) tmem_st_x16(tmem_P + c * 16, pk[c]);
        } else if (q > FIRST_ST) {
          tmem_st_x16(tmem_P + q * 16, pk[q]);
        }
      }
      SVDPP_TR(j, 6);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_ready[t]);
      SVDPP_TR(j, 7);
    }
    mbar_wait(&pv_done[t], (p.n_kv - 1) & 1, 51);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int q = q0 + t * A2_BQ + r;
    __half* dst = p.out + static_cast<long long>(row_base + q) * p.ldo + head * 64;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld_x32(tmem_O + c * 32, o);
      tmem_ld_wait();
      if (q < p.S) {
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          packed[i] = pack_half2(__uint_as_float(o[2 * i]) * inv_l, __uint_as_float(o[2 * i + 1]) * inv_l);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd)
          *reinterpret_cast<uint4*>(dst + c * 32 + qd * 8) =
              make_uint4(packed[4 * qd], packed[4 * qd + 1], packed[4 * qd + 2], packed[4 * qd + 3]);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax
    constexpr int COLS = Cfg::COLS;
    const int t = warp / (4 * SPLIT);              // query tile
    const int hf = SPLIT == 2 ? (warp >> 2) & 1 : 0;  // which half of the 128 key columns this thread owns
    const int r = (warp & 3) * 32 + lane;          // row in tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tmem_S = tmem_base + t * 128 + hf * COLS + lane_sel;
    const uint32_t tmem_O = tmem_base + 256 + t * 64 + hf * (64 / SPLIT) + lane_sel;
    const uint32_t tmem_P = tmem_base + 384 + t * 64 + hf * (COLS / 2) + lane_sel;
    const int pair_bar = 1 + t * 4 + (warp & 3);   // named barrier of the two warps sharing these rows
    float m_used = -CUDART_INF_F;
    float l_run = 0.f;
    for (int j = 0; j < p.n_kv; ++j) {
      const int valid = p.S - j * A2_BK - hf * COLS;  // my columns [0, valid) of this block are real keys
      SVDPP_TR(j, 0);
      mbar_wait(&s_full[t], j & 1, 49);
      tc_fence_after();
      SVDPP_TR(j, 1);
      uint32_t v[COLS];
      {
        uint32_t(*v4)[32] = reinterpret_cast<uint32_t(*)[32]>(v);
#pragma unroll
        for (int c = 0; c < COLS / 32; ++c) tmem_ld_x32(tmem_S + c * 32, v4[c]);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(&s_free[t]);  // the tensor core may overwrite S_t with block j+1 now
      SVDPP_TR(j, 2);
      if (valid < COLS) {  // warp-uniform: only the last key block of an image can be partial
#pragma unroll
        for (int i = 0; i < COLS; ++i)
          if (i >= valid) v[i] = 0xff800000u;  // -inf
      }
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int i = 0; i < COLS / 2; ++i) mx = fmax3(mx, __uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
      if constexpr (SPLIT == 2) {
        float* slot = xch + ((j & 1) * 2 + t) * 256;  // parity double buffer: a slot is rewritten two blocks later
        slot[hf * 128 + r] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        mx = fmaxf(mx, slot[(hf ^ 1) * 128 + r]);
      }
      const float m_blk = mx * p.scale_log2;
      bool raise = false;
      if (j == 0)
        m_used = m_blk;
      else
        raise = m_blk > m_used + 8.0f;
      if (j > 0 && __any_sync(0xffffffffu, raise)) {
        // every MMA that has touched O so far must have retired before O is rewritten
        mbar_wait(&pv_done[t], (j - 1) & 1, 50);
        tc_fence_after();
        const float m_new = raise ? m_blk : m_used;
        const float alpha = fast_exp2(m_used - m_new);
        m_used = m_new;
        l_run *= alpha;
#pragma unroll 1
        for (int c = 0; c < 4 / SPLIT; ++c) {  // 16 columns at a time: this rare path must not cost the main loop registers
          uint32_t o[16];
          tmem_ld_x16(tmem_O + c * 16, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_x16(tmem_O + c * 16, o);
        }
        tmem_st_wait();
      }
      // probabilities -> packed fp16 -> this tile's P columns
      SVDPP_TR(j, 3);
      const float neg_m = -m_used;
      uint32_t pk[COLS / 2];
      if constexpr (SPLIT == 1) {
        // packed f32x2 arithmetic (FFMA2 / FADD2) halves the issue slots of the scale-and-shift and of the row
        // sum: 752 -> 807 TFLOP/s at S = 9216.  (With two threads per row the 64-bit register pairs push the
        // 96-register budget into spills: 799 -> 717, so that variant keeps scalar arithmetic.)
        const uint64_t scale2 = pack_f2(p.scale_log2, p.scale_log2);
        const uint64_t negm2 = pack_f2(neg_m, neg_m);
        uint64_t lsa = 0ull, lsb = 0ull;  // (0.f, 0.f)
#pragma unroll
        for (int i = 0; i < COLS / 2; i += 2) {
          const uint64_t xa = ffma2(pack_f2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), scale2, negm2);
          const uint64_t xb = ffma2(pack_f2(__uint_as_float(v[2 * i + 2]), __uint_as_float(v[2 * i + 3])), scale2, negm2);
          float x0, x1, x2, x3;
          unpack_f2(xa, x0, x1);
          unpack_f2(xb, x2, x3);
          const float p0 = fast_exp2(x0), p1 = fast_exp2(x1), p2 = fast_exp2(x2), p3 = fast_exp2(x3);
          lsa = fadd2(lsa, pack_f2(p0, p1));
          lsb = fadd2(lsb, pack_f2(p2, p3));
          pk[i] = pack_half2(p0, p1);
          pk[i + 1] = pack_half2(p2, p3);
        }
        float a0, a1, b0, b1;
        unpack_f2(lsa, a0, a1);
        unpack_f2(lsb, b0, b1);
        l_run += (a0 + a1) + (b0 + b1);
      } else {
        float ls0 = 0.f, ls1 = 0.f, ls2 = 0.f, ls3 = 0.f;
#pragma unroll
        for (int i = 0; i < COLS / 2; i += 2) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(v[2 * i]), p.scale_log2, neg_m));
          const float p1 = fast_exp2(fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2, neg_m));
          const float p2 = fast_exp2(fmaf(__uint_as_float(v[2 * i + 2]), p.scale_log2, neg_m));
          const float p3 = fast_exp2(fmaf(__uint_as_float(v[2 * i + 3]), p.scale_log2, neg_m));
          ls0 += p0;
          ls1 += p1;
          ls2 += p2;
          ls3 += p3;
          pk[i] = pack_half2(p0, p1);
          pk[i + 1] = pack_half2(p2, p3);
        }
        l_run += (ls0 + ls1) + (ls2 + ls3);
      }
      SVDPP_TR(j, 4);
      if (j > 0) {  // P_t(j-1) must have been consumed (long since: it was issued a whole exp phase ago)
        mbar_wait(&pv_done[t], (j - 1) & 1, 52);
        tc_fence_after();
      }
      SVDPP_TR(j, 5);
      {
        const uint32_t(*pk2)[32] = reinterpret_cast<const uint32_t(*)[32]>(pk);
#pragma unroll
        for (int c = 0; c < COLS / 64; ++c) tmem_st_x32(tmem_P + c * 32, pk2[c]);
        SVDPP_TR(j, 6);
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(&p_ready[t]);
      SVDPP_TR(j, 7);
    }
    mbar_wait(&pv_done[t], (p.n_kv - 1) & 1, 51);
    tc_fence_after();
    if constexpr (SPLIT == 2) {  // row sum = the two halves' sums (same reference maximum on both sides)
      float* slot = xch + (((p.n_kv & 1) * 2) + t) * 256;
      slot[hf * 128 + r] = l_run;
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      l_run += slot[(hf ^ 1) * 128 + r];
    }
    const float inv_l = 1.0f / l_run;
    const int q = q0 + t * A2_BQ + r;
    __half* dst = p.out + static_cast<long long>(row_base + q) * p.ldo + head * 64 + hf * (64 / SPLIT);
#pragma unroll
    for (int c = 0; c < 2 / SPLIT; ++c) {  // each thread writes its 64 / SPLIT output dims
      uint32_t o[32];
      tmem_ld_x32(tmem_O + c * 32, o);
      tmem_ld_wait();
      if (q < p.S) {
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          packed[i] = pack_half2(__uint_as_float(o[2 * i]) * inv_l, __uint_as_float(o[2 * i + 1]) * inv_l);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd)
          *reinterpret_cast<uint4*>(dst + c * 32 + qd * 8) =
              make_uint4(packed[4 * qd], packed[4 * qd + 1], packed[4 * qd + 2], packed[4 * qd + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == Cfg::MMA_WARP0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int SPLIT, int QP, int POLY>
static int launch_a2(const CUtensorMap& tmQ, const CUtensorMap& tmKV, const Attn2Params& p, dim3 grid, cudaStream_t stream) {
  auto kern = attn_spatial2_tc_kernel<SPLIT, QP, POLY>;
  static bool configured = false;
  if (!configured) {
    SVDPP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, A2_SMEM_BYTES));
    configured = true;
  }
  SVDPP_CUDA(launch_kernel(kern, grid, dim3(A2Cfg<SPLIT>::THREADS), A2_SMEM_BYTES, stream, 1, tmQ, tmKV, p));
  return check_launch("attn_spatial2_tc_kernel");
}

// variant: 1 = thread per row, whole-row softmax (round 1); 2 = two threads per row; 4 = quarter-pipelined softmax;
// 5 / 6 = quarter-pipelined with every 8th / 4th group of exponentials on the FMA pipe
int launch_attn_spatial2(const svdpp_attn_desc* d, int variant, cudaStream_t stream) {
  SVDPP_CHECK_ARG(d->heads <= 65535 && d->n_img <= 65535, "attn: grid too large");
  Attn2Params p{};
  p.S = d->S;
  p.n_kv = (d->S + A2_BK - 1) / A2_BK;
  p.q_off = d->q_off;
  p.k_off = d->k_off;
  p.v_off = d->v_off;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.out = static_cast<__half*>(d->out);
  p.ldo = d->ldo;
  p.stagger = tuning().fmha_stagger;
  p.trace = g_attn_trace;
  CUtensorMap tmQ, tmKV;
  const long long rows = static_cast<long long>(d->n_img) * d->S;
  uint64_t dims[2] = {static_cast<uint64_t>(d->ld), static_cast<uint64_t>(rows)};
  uint64_t str[1] = {static_cast<uint64_t>(d->ld) * 2};
  uint32_t box_q[2] = {64, A2_BQ};
  uint32_t box_kv[2] = {64, A2_BK};
  if (encode_tmap_f16(&tmQ, d->qkv, 2, dims, str, box_q)) return -5;
  if (encode_tmap_f16(&tmKV, d->qkv, 2, dims, str, box_kv)) return -5;
  dim3 grid((d->S + 2 * A2_BQ - 1) / (2 * A2_BQ), d->heads, d->n_img);
  switch (variant) {
    case 2: return launch_a2<2, 0, 0>(tmQ, tmKV, p, grid, stream);
    case 4: return launch_a2<1, 1, 0>(tmQ, tmKV, p, grid, stream);
    case 5: return launch_a2<1, 1, 8>(tmQ, tmKV, p, grid, stream);
    case 6: return launch_a2<1, 1, 4>(tmQ, tmKV, p, grid, stream);
    default: return launch_a2<1, 0, 0>(tmQ, tmKV, p, grid, stream);
  }
}

}  // namespace svdpp

// Debug hook (tools/attn_trace.py): device buffer of >= 2 * n_kv * 8 uint32 that the two-tile FMHA fills with clock
// stamps of the softmax phases of warps 0 and 4 of its first CTA; NULL switches the tracing off again.
extern "C" int svdpp_debug_attn_trace(void* dev_buffer) {
  svdpp::g_attn_trace = static_cast<unsigned int*>(dev_buffer);
  return 0;
}
