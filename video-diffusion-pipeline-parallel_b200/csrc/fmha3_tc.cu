// svdpp_attn_spatial_f16, impl 7: spatial self-attention, head_dim 64, two 128-query tiles per CTA, 128-key blocks,
// the two tiles' exponential phases in strict alternation ("ping-pong").
//
// What bounds a head_dim-64 FMHA is the MUFU pipe (one ex2 per score, 16 / clk / SM = one warp instruction per 8
// clocks and scheduler).  Measurements that shaped this kernel (tools/ubench/mix.cu, tools/ubench/mma.cu,
// tools/attn_trace.py, all on B200):
//   * the softmax instruction mix (2 MUFU.EX2 + FFMA2 + FADD2 + FMNMX3 + F2FP per pair of scores) runs at the MUFU
//     rate alone (16.1 clk per pair and scheduler): nothing else in the loop competes for that pipe;
//   * a MUFU instruction blocks its warp until the pipe accepts it, so a warp cannot queue exponentials and do its
//     barrier work meanwhile: the fixed latencies of a block boundary (an mbarrier try_wait costs ~90 clocks even when
//     the phase completed long ago, the tcgen05.ld round trip, tcgen05.wait::st, the arrivals, the row maximum)
//     are MUFU-idle time of that warp and only ANOTHER warp of the scheduler can fill them;
//   * two free-running warps per scheduler (fmha2_tc.cu: warp w of tile 0 and warp w + 4 of tile 1) do not fill each
//     other's gaps: sharing the pipe pulls them into phase (while both are in their exponential phase each gets half
//     the rate, so they finish together and then sit in the non-MUFU phases together); the trace shows ~2900 clocks
//     per 128-key block of which 2 x 1024 are MUFU work; a start offset decays within a few dozen blocks;
//   * key blocks of 64 with S double-buffered in registers and TMEM (every boundary latency behind "pre-issued"
//     exponentials) were SLOWER (645 vs 778 TFLOP/s): twice as many boundaries, and see the second point.
// Hence: the warps of a scheduler take turns.  Warp w (tile 0) and warp w + 4 (tile 1) hand the MUFU pipe to each
// other through a pair of named barriers (bar.sync / bar.arrive, 64 threads): a warp does everything that needs no
// MUFU (wait for S, TMEM load, row maximum, wait for the previous PV product) while its partner exponentiates, then
// runs its 128 exponentials alone at the full rate.  For that a single warp has to keep the pipe busy by itself:
// the exponentials are issued in batches of 16, in place, and consumed (row sum, fp16 pack) one batch later, so no
// instruction waits for a MUFU result that was issued just before it (as far as ptxas lets it: at any optimisation level
// it re-schedules the consumers to one pair behind their producers, volatile asm or not, so a lone warp reaches ~75 % of
// the pipe's rate - tools/ubench/mix.cu - and the turns overlap by design).  The hand-over is signalled a few batches
// before the end of the phase (the partner's wake-up latency overlaps the tail).
// The MMA issuer warps run CONVERGED with one elected lane executing the tcgen05 instructions: under
// `if (lane == 0) { loop }` the descriptors live in vector registers and every MMA pays a chain of R2UR moves
// (116 clocks per issued MMA against 77 with a converged warp; the tensor pipe needs 64 for N = 128).
//
// One CTA per (256 queries, head, image), one CTA per SM:
//   warps 0..3 / 4..7  softmax of query tile 0 / 1 (thread = query row = TMEM lane)
//   warp 8             TMA producer: Q tiles once, K_j / V_j (128 keys x 64) through 3-slot rings
//   warps 9, 10        MMA issuers, one per tile
// TMEM (512 columns): S0 [0,128)  S1 [128,256)  O0 [256,320)  O1 [320,384)  P0 [384,448)  P1 [448,512).
// O stays in TMEM and is rescaled lazily (reference maximum raised only when a block exceeds it by more than 2^8).
#include <cuda_fp16.h>
#include <math_constants.h>

#include "common.h"
#include "ptx.cuh"

namespace svdpp {

struct Attn3Params {
  int S, n_kv;
  int q_off, k_off, v_off;
  float scale_log2;
  __half* out;
  long long ldo;
  int handover;         // batches (of 16 exponentials) before the end of a phase at which the partner is released (0..7)
  unsigned int* trace;  // debug: phase timestamps of CTA (0,0,0): [softmax warp 0, softmax warp 4][n_kv][8]
};
extern unsigned int* g_attn_trace;  // fmha2_tc.cu

constexpr int A3_BQ = 128;
constexpr int A3_BK = 128;
constexpr int A3_Q_BYTES = A3_BQ * 64 * 2;   // 16 KB
constexpr int A3_KV_BYTES = A3_BK * 64 * 2;  // 16 KB
constexpr int A3_STAGES = 3;
constexpr int A3_SMEM_BYTES = 2 * A3_Q_BYTES + 2 * A3_STAGES * A3_KV_BYTES + 1024 /*align*/ + 512 /*barriers*/;
constexpr int A3_TMA_WARP = 8;
constexpr int A3_MMA_WARP0 = 9;
constexpr int A3_THREADS = 11 * 32;

__device__ __forceinline__ void umma_f16_ts3(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint64_t a3_pack_f2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void a3_unpack_f2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t a3_ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t a3_fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint32_t a3_pack_half2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %2, %1;" : "=r"(r) : "f"(a), "f"(b));  // low half = a
  return r;
}
// exp2 on the FMA pipe for a share of the scores (tuning "fmha_poly"): Cody-Waite split x = n + f with the magic-number
// add (n in the low mantissa bits), degree-3 minimax polynomial of 2^f on [-0.5, 0.5] (max relative error 7.5e-5, a
// sixth of the fp16 half-ulp P is rounded to anyway), n added into the exponent field.  8 FMA-pipe / ALU instructions
// that issue in the 8-clock shadow of the neighbouring MUFU.EX2 instructions.
__device__ __forceinline__ float a3_poly_exp2(float x) {
  x = fmaxf(x, -125.0f);                   // masked keys (-inf) and underflow: 2^-125, rounds to 0 in fp16
  const float xf = x + 12582912.0f;        // 1.5 * 2^23: round-to-nearest integer part lands in the low mantissa bits
  const float f = x - (xf - 12582912.0f);  // [-0.5, 0.5]
  float pz = fmaf(0.0551716648042202f, f, 0.2426111251115799f);
  pz = fmaf(pz, f, 0.6932609677314758f);
  pz = fmaf(pz, f, 0.9999280571937561f);
  return __uint_as_float(__float_as_uint(pz) + (__float_as_uint(xf) << 23));
}
// element e of a batch of 16 goes to the polynomial when NPOLY of 16 are asked for (spread evenly)
template <int NPOLY>
__device__ __forceinline__ constexpr bool a3_is_poly(int e) { return NPOLY > 0 && ((e * NPOLY) % 16) < NPOLY; }
template <int NPOLY>
__device__ __forceinline__ float a3_exp2_sel(float x, int e) {
  return a3_is_poly<NPOLY>(e) ? a3_poly_exp2(x) : fast_exp2(x);
}

// named barriers of a warp pair (64 threads): sync = wait for the partner's arrive
__device__ __forceinline__ void a3_bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void a3_bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

// maximum of 32 raw scores (16 FMNMX3 in four chains)
__device__ __forceinline__ float a3_max32(const uint32_t* x) {
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
}

template <int NPOLY>
__global__ void __launch_bounds__(A3_THREADS, 1)
attn_spatial3_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                        const Attn3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [2][128][64]
  uint8_t* sK = sQ + 2 * A3_Q_BYTES;                    // [STAGES][128][64]
  uint8_t* sV = sK + A3_STAGES * A3_KV_BYTES;           // [STAGES][128][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + A3_STAGES * A3_KV_BYTES);
  uint64_t* q_full = bars;                      // [1]
  uint64_t* k_full = q_full + 1;                // [STAGES]
  uint64_t* k_empty = k_full + A3_STAGES;       // [STAGES]
  uint64_t* v_full = k_empty + A3_STAGES;       // [STAGES]
  uint64_t* v_empty = v_full + A3_STAGES;       // [STAGES]
  uint64_t* s_full = v_empty + A3_STAGES;       // [2]  S_t(j) complete (MMA commit)
  uint64_t* s_free = s_full + 2;                // [2]  S_t(j) is in registers (4 warp arrivals)
  uint64_t* p_ready = s_free + 2;               // [2]  P_t(j) stored (4 warp arrivals)
  uint64_t* pv_done = p_ready + 2;              // [2]  O_t += P_t(j) V_j retired (MMA commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  // warp index through a shuffle: provably warp-uniform, so everything derived from it stays in uniform registers
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * A3_BQ);
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const int row_base = img * p.S;

  if (warp == A3_TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < A3_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 2);  // both tiles' MMA warps
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 2);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 4);
      mbar_init(&p_ready[t], 4);
      mbar_init(&pv_done[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == A3_MMA_WARP0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  const bool trace_on = p.trace != nullptr && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && lane == 0 && (warp == 0 || warp == 4);
  unsigned int* const trace = trace_on ? p.trace + (warp >> 2) * (p.n_kv * 8) : nullptr;
#define A3_TR(j, ev)                                                              \
  do {                                                                            \
    if (trace_on) trace[(j) * 8 + (ev)] = static_cast<unsigned int>(clock64());   \
  } while (0)

  if (warp == A3_TMA_WARP) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * A3_Q_BYTES);
      tma_load_2d(sQ, &tmQ, q_full, p.q_off + head * 64, row_base + q0);
      tma_load_2d(sQ + A3_Q_BYTES, &tmQ, q_full, p.q_off + head * 64, row_base + q0 + A3_BQ);
      for (int j = 0; j < p.n_kv; ++j) {
        const int s = j % A3_STAGES;
        const uint32_t ph = (j / A3_STAGES) & 1;
        mbar_wait(&k_empty[s], ph ^ 1, 61);
        mbar_expect_tx(&k_full[s], A3_KV_BYTES);
        tma_load_2d(sK + s * A3_KV_BYTES, &tmKV, &k_full[s], p.k_off + head * 64, row_base + j * A3_BK);
        mbar_wait(&v_empty[s], ph ^ 1, 62);
        mbar_expect_tx(&v_full[s], A3_KV_BYTES);
        tma_load_2d(sV + s * A3_KV_BYTES, &tmKV, &v_full[s], p.v_off + head * 64, row_base + j * A3_BK);
      }
    }
  } else if (warp >= A3_MMA_WARP0) {
    // ------------------------------------------------------------------ MMA issuers (one per query tile), converged
    const int t = warp - A3_MMA_WARP0;
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc_f16(A3_BK, false);  // S: N = 128 keys, K-major B
    constexpr uint32_t idesc_o = make_idesc_f16(64, true);      // O: N = 64 dims, V read MN-major
    const uint32_t q_addr = smem_u32(sQ + t * A3_Q_BYTES);
    const uint32_t d_s = tmem_base + t * 128;
    const uint32_t d_o = tmem_base + 256 + t * 64;
    const uint32_t a_p = tmem_base + 384 + t * 64;
    auto issue_s = [&](int jj) {  // S_t = Q_t K_jj^T
      const int s = jj % A3_STAGES;
      mbar_wait(&k_full[s], (jj / A3_STAGES) & 1, 64);
      tc_fence_after();
      const uint32_t k_addr = smem_u32(sK + s * A3_KV_BYTES);
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(d_s, make_smem_desc_sw128(q_addr + k * 32, 1024, 0), make_smem_desc_sw128(k_addr + k * 32, 1024, 0),
                   idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[t]);
        umma_commit(&k_empty[s]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0, 63);
    issue_s(0);
    for (int j = 0; j < p.n_kv; ++j) {
      const int s = j % A3_STAGES;
      if (j + 1 < p.n_kv) {
        mbar_wait(&s_free[t], j & 1, 66);  // S_t(j) is in the softmax warps' registers
        tc_fence_after();
        issue_s(j + 1);
      }
      mbar_wait(&v_full[s], (j / A3_STAGES) & 1, 65);
      mbar_wait(&p_ready[t], j & 1, 67);
      tc_fence_after();
      const uint32_t v_addr = smem_u32(sV + s * A3_KV_BYTES);
      if (leader) {
#pragma unroll
        for (int k = 0; k < A3_BK / 16; ++k)  // O_t += P_t V_j, P_t (two fp16 per column) from TMEM
          umma_f16_ts3(d_o, a_p + k * 8, make_smem_desc_sw128(v_addr + k * 2048, 1024, 8192), idesc_o,
                       (j | k) != 0 ? 1u : 0u);
        umma_commit(&pv_done[t]);
        umma_commit(&v_empty[s]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax (thread = query row)
    const int t = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tmem_S = tmem_base + t * 128 + lane_sel;
    const uint32_t tmem_O = tmem_base + 256 + t * 64 + lane_sel;
    const uint32_t tmem_P = tmem_base + 384 + t * 64 + lane_sel;
    // MUFU turn-taking of the pair (warp w, warp w + 4): barrier `mine` is arrived on by the partner when its
    // exponential phase ends and waited on by me before mine starts; `theirs` the other way round
    const int bar_mine = 1 + (warp & 3) * 2 + t;
    const int bar_theirs = 1 + (warp & 3) * 2 + (t ^ 1);
    const uint64_t scale2 = a3_pack_f2(p.scale_log2, p.scale_log2);
    const int n_kv = p.n_kv;
    const int handover = p.handover;
    float m_used = -CUDART_INF_F;
    float l_run = 0.f;
    if (t == 1) a3_bar_arrive(bar_theirs);  // tile 0 goes first

    for (int j = 0; j < n_kv; ++j) {
      // ---- no MUFU work (runs while the partner exponentiates): S(j) into registers, row maximum, P buffer free
      const int valid = p.S - j * A3_BK;  // my columns [0, valid) of this block are real keys
      A3_TR(j, 0);
      mbar_wait(&s_full[t], j & 1, 69);
      tc_fence_after();
      A3_TR(j, 1);
      uint32_t v[128];
      {
        uint32_t(*v4)[32] = reinterpret_cast<uint32_t(*)[32]>(v);
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_x32(tmem_S + c * 32, v4[c]);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[t]);  // the tensor core may overwrite S_t with block j+1 now
      A3_TR(j, 2);
      if (valid < A3_BK) {  // warp-uniform: only the last key block of an image can be partial
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= valid) v[i] = 0xff800000u;  // -inf
      }
      const float mx = fmaxf(fmaxf(a3_max32(v), a3_max32(v + 32)), fmaxf(a3_max32(v + 64), a3_max32(v + 96)));
      const float m_blk = mx * p.scale_log2;
      bool raise = false;
      if (j == 0)
        m_used = m_blk;
      else
        raise = m_blk > m_used + 8.0f;
      if (j > 0) {  // PV(j-1) retired: P_t may be rewritten, O_t may be rescaled
        mbar_wait(&pv_done[t], (j - 1) & 1, 70);
        tc_fence_after();
      }
      if (j > 0 && __any_sync(0xffffffffu, raise)) {  // rare: the reference maximum moves
        const float m_new = raise ? m_blk : m_used;
        const float alpha = fast_exp2(m_used - m_new);
        m_used = m_new;
        l_run *= alpha;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t o[16];
          tmem_ld_x16(tmem_O + c * 16, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_x16(tmem_O + c * 16, o);
        }
        tmem_st_wait();
      }
      const float neg_m = -m_used;
      const uint64_t negm2 = a3_pack_f2(neg_m, neg_m);
      uint64_t lsa = 0ull, lsb = 0ull;
      A3_TR(j, 3);
      // ---- my turn on the MUFU pipe
      a3_bar_sync(bar_mine);
      A3_TR(j, 4);
      uint32_t pk[16];
#pragma unroll
      for (int q = 0; q <= 8; ++q) {
        if (q < 8) {  // issue batch q: 16 exponentials, in place
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const int c = q * 16 + i;
            const uint64_t xa = a3_ffma2(a3_pack_f2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])), scale2, negm2);
            const uint64_t xb = a3_ffma2(a3_pack_f2(__uint_as_float(v[c + 2]), __uint_as_float(v[c + 3])), scale2, negm2);
            float x0, x1, x2, x3;
            a3_unpack_f2(xa, x0, x1);
            a3_unpack_f2(xb, x2, x3);
            v[c] = __float_as_uint(a3_exp2_sel<NPOLY>(x0, i));
            v[c + 1] = __float_as_uint(a3_exp2_sel<NPOLY>(x1, i + 1));
            v[c + 2] = __float_as_uint(a3_exp2_sel<NPOLY>(x2, i + 2));
            v[c + 3] = __float_as_uint(a3_exp2_sel<NPOLY>(x3, i + 3));
          }
        }
        if (q == 8 - handover) {  // the partner may start: its wake-up overlaps my last batches
          if (!(t == 1 && j == n_kv - 1)) a3_bar_arrive(bar_theirs);  // (tile 0 waits n_kv times: initial arrive + n_kv - 1)
        }
        if (q > 0) {  // consume batch q - 1: row sum, fp16 pack, P columns [(q-1)*8, (q-1)*8 + 8)
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const int c = (q - 1) * 16 + i;
            const float p0 = __uint_as_float(v[c]), p1 = __uint_as_float(v[c + 1]);
            const float p2 = __uint_as_float(v[c + 2]), p3 = __uint_as_float(v[c + 3]);
            lsa = a3_fadd2(lsa, a3_pack_f2(p0, p1));
            lsb = a3_fadd2(lsb, a3_pack_f2(p2, p3));
            pk[((q - 1) & 1) * 8 + i / 2] = a3_pack_half2(p0, p1);
            pk[((q - 1) & 1) * 8 + i / 2 + 1] = a3_pack_half2(p2, p3);
          }
          if (((q - 1) & 1) == 1) tmem_st_x16(tmem_P + ((q - 1) >> 1) * 16, pk);
        }
      }
      A3_TR(j, 5);
      float a0, a1, b0, b1;
      a3_unpack_f2(lsa, a0, a1);
      a3_unpack_f2(lsb, b0, b1);
      l_run += (a0 + a1) + (b0 + b1);
      tmem_st_wait();
      A3_TR(j, 6);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[t]);
      A3_TR(j, 7);
    }
    mbar_wait(&pv_done[t], (n_kv - 1) & 1, 71);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int q = q0 + t * A3_BQ + r;
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
          packed[i] = a3_pack_half2(__uint_as_float(o[2 * i]) * inv_l, __uint_as_float(o[2 * i + 1]) * inv_l);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd)
          *reinterpret_cast<uint4*>(dst + c * 32 + qd * 8) =
              make_uint4(packed[4 * qd], packed[4 * qd + 1], packed[4 * qd + 2], packed[4 * qd + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == A3_MMA_WARP0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int NPOLY>
static int launch_attn3(dim3 grid, cudaStream_t stream, const CUtensorMap& tmQ, const CUtensorMap& tmKV, const Attn3Params& p) {
  static bool configured = false;
  if (!configured) {
    SVDPP_CUDA(cudaFuncSetAttribute(attn_spatial3_tc_kernel<NPOLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, A3_SMEM_BYTES));
    configured = true;
  }
  SVDPP_CUDA(launch_kernel(attn_spatial3_tc_kernel<NPOLY>, grid, dim3(A3_THREADS), A3_SMEM_BYTES, stream, 1, tmQ, tmKV, p));
  return check_launch("attn_spatial3_tc_kernel");
}

int launch_attn_spatial3(const svdpp_attn_desc* d, cudaStream_t stream) {
  SVDPP_CHECK_ARG(d->heads <= 65535 && d->n_img <= 65535, "attn: grid too large");
  Attn3Params p{};
  p.S = d->S;
  p.n_kv = (d->S + A3_BK - 1) / A3_BK;
  p.q_off = d->q_off;
  p.k_off = d->k_off;
  p.v_off = d->v_off;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.out = static_cast<__half*>(d->out);
  p.ldo = d->ldo;
  const int h = tuning().fmha_handover;
  p.handover = h < 0 ? 0 : (h > 7 ? 7 : h);
  p.trace = g_attn_trace;
  CUtensorMap tmQ, tmKV;
  const long long rows = static_cast<long long>(d->n_img) * d->S;
  uint64_t dims[2] = {static_cast<uint64_t>(d->ld), static_cast<uint64_t>(rows)};
  uint64_t str[1] = {static_cast<uint64_t>(d->ld) * 2};
  uint32_t box_q[2] = {64, A3_BQ};
  uint32_t box_kv[2] = {64, A3_BK};
  if (encode_tmap_f16(&tmQ, d->qkv, 2, dims, str, box_q)) return -5;
  if (encode_tmap_f16(&tmKV, d->qkv, 2, dims, str, box_kv)) return -5;
  dim3 grid((d->S + 2 * A3_BQ - 1) / (2 * A3_BQ), d->heads, d->n_img);
  switch (tuning().fmha_poly) {  // exponentials per batch of 16 evaluated on the FMA pipe
    case 0: return launch_attn3<0>(grid, stream, tmQ, tmKV, p);
    case 2: return launch_attn3<2>(grid, stream, tmQ, tmKV, p);
    case 3: return launch_attn3<3>(grid, stream, tmQ, tmKV, p);
    case 4: return launch_attn3<4>(grid, stream, tmQ, tmKV, p);
    case 5: return launch_attn3<5>(grid, stream, tmQ, tmKV, p);
    case 6: return launch_attn3<6>(grid, stream, tmQ, tmKV, p);
    case 8: return launch_attn3<8>(grid, stream, tmQ, tmKV, p);
    default: break;
  }
  set_error("attn: fmha_poly must be one of 0, 2, 3, 4, 5, 6, 8 (exponentials of 16 on the FMA pipe)");
  return -1;
}

}  // namespace svdpp
