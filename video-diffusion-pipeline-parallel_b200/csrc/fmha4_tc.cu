// svdpp_attn_spatial_f16, impl 8: the ping-pong FMHA of fmha3_tc.cu with TWO threads per query row (16 softmax warps).
//
// fmha3_tc.cu lets the two warps of a scheduler take turns on the MUFU pipe, but a lone warp only reaches ~75 % of the
// pipe's rate (ptxas keeps the consumers of an exponential one pair behind it; tools/ubench/mix.cu: 21.1 instead of 16.0
// clocks per pair with one warp per scheduler, 16.06 with two).  Here every scheduler holds FOUR softmax warps: the two
// half-row warps of tile 0 (columns [0, 64) and [64, 128) of the same 32 rows) and the two of tile 1.  The two half-row
// warps of a tile exponentiate TOGETHER - two warps feed the pipe at its full rate - while the other tile's two warps do
// their non-MUFU work; then the tiles swap (one named barrier of 128 threads per lane quarter and direction).  The same
// barrier publishes the half-row maxima that the two threads of a row exchange through shared memory.
// Everything else (TMEM layout, P in tensor memory, lazy rescale, converged MMA issuers, batched in-place exponentials) is
// fmha3_tc.cu's.
//   warps 0..7 / 8..15  softmax of query tile 0 / 1: warp = tile * 8 + half * 4 + lane quarter
//   warp 16             TMA producer        warps 17, 18   MMA issuers, one per tile
#include <cuda_fp16.h>
#include <math_constants.h>

#include "common.h"
#include "ptx.cuh"

namespace svdpp {

struct Attn4Params {
  int S, n_kv;
  int q_off, k_off, v_off;
  float scale_log2;
  __half* out;
  long long ldo;
  int handover;         // batches (of 16 exponentials, 4 per turn) before the end of a turn at which the other tile is released (0..3)
  unsigned int* trace;  // debug: phase timestamps of CTA (0,0,0): [softmax warp 0, softmax warp 4][n_kv][8]
};
extern unsigned int* g_attn_trace;  // fmha2_tc.cu

constexpr int A4_BQ = 128;
constexpr int A4_BK = 128;
constexpr int A4_Q_BYTES = A4_BQ * 64 * 2;   // 16 KB
constexpr int A4_KV_BYTES = A4_BK * 64 * 2;  // 16 KB
constexpr int A4_STAGES = 3;
constexpr int A4_SMEM_BYTES = 2 * A4_Q_BYTES + 2 * A4_STAGES * A4_KV_BYTES + 1024 /*align*/ + 512 /*barriers*/ +
                              2 * 2 * 2 * 128 * 4 /*half-row maximum / sum exchange*/;
constexpr int A4_TMA_WARP = 16;
constexpr int A4_MMA_WARP0 = 17;
constexpr int A4_THREADS = 19 * 32;

__device__ __forceinline__ void umma_f16_ts4(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint64_t a4_pack_f2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void a4_unpack_f2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t a4_ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t a4_fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint32_t a4_pack_half2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %2, %1;" : "=r"(r) : "f"(a), "f"(b));  // low half = a
  return r;
}
// named barriers of a warp pair (64 threads): sync = wait for the partner's arrive
__device__ __forceinline__ void a4_bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void a4_bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

// maximum of 32 raw scores (16 FMNMX3 in four chains)
__device__ __forceinline__ float a4_max32(const uint32_t* x) {
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

__global__ void __launch_bounds__(A4_THREADS, 1)
attn_spatial4_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                        const Attn4Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [2][128][64]
  uint8_t* sK = sQ + 2 * A4_Q_BYTES;                    // [STAGES][128][64]
  uint8_t* sV = sK + A4_STAGES * A4_KV_BYTES;           // [STAGES][128][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + A4_STAGES * A4_KV_BYTES);
  uint64_t* q_full = bars;                      // [1]
  uint64_t* k_full = q_full + 1;                // [STAGES]
  uint64_t* k_empty = k_full + A4_STAGES;       // [STAGES]
  uint64_t* v_full = k_empty + A4_STAGES;       // [STAGES]
  uint64_t* v_empty = v_full + A4_STAGES;       // [STAGES]
  uint64_t* s_full = v_empty + A4_STAGES;       // [2]  S_t(j) complete (MMA commit)
  uint64_t* s_free = s_full + 2;                // [2]  S_t(j) is in registers (4 warp arrivals)
  uint64_t* p_ready = s_free + 2;               // [2]  P_t(j) stored (4 warp arrivals)
  uint64_t* pv_done = p_ready + 2;              // [2]  O_t += P_t(j) V_j retired (MMA commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
  float* xch = reinterpret_cast<float*>(bars + 64);  // [2 parity][2 tiles][2 halves][128 rows]

  // warp index through a shuffle: provably warp-uniform, so everything derived from it stays in uniform registers
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * A4_BQ);
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const int row_base = img * p.S;

  if (warp == A4_TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < A4_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 2);  // both tiles' MMA warps
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 2);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 8);
      mbar_init(&p_ready[t], 8);
      mbar_init(&pv_done[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == A4_MMA_WARP0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  const bool trace_on = p.trace != nullptr && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && lane == 0 && (warp == 0 || warp == 8);
  unsigned int* const trace = trace_on ? p.trace + (warp >> 3) * (p.n_kv * 8) : nullptr;
#define A4_TR(j, ev)                                                              \
  do {                                                                            \
    if (trace_on) trace[(j) * 8 + (ev)] = static_cast<unsigned int>(clock64());   \
  } while (0)

  if (warp == A4_TMA_WARP) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * A4_Q_BYTES);
      tma_load_2d(sQ, &tmQ, q_full, p.q_off + head * 64, row_base + q0);
      tma_load_2d(sQ + A4_Q_BYTES, &tmQ, q_full, p.q_off + head * 64, row_base + q0 + A4_BQ);
      for (int j = 0; j < p.n_kv; ++j) {
        const int s = j % A4_STAGES;
        const uint32_t ph = (j / A4_STAGES) & 1;
        mbar_wait(&k_empty[s], ph ^ 1, 61);
        mbar_expect_tx(&k_full[s], A4_KV_BYTES);
        tma_load_2d(sK + s * A4_KV_BYTES, &tmKV, &k_full[s], p.k_off + head * 64, row_base + j * A4_BK);
        mbar_wait(&v_empty[s], ph ^ 1, 62);
        mbar_expect_tx(&v_full[s], A4_KV_BYTES);
        tma_load_2d(sV + s * A4_KV_BYTES, &tmKV, &v_full[s], p.v_off + head * 64, row_base + j * A4_BK);
      }
    }
  } else if (warp >= A4_MMA_WARP0) {
    // ------------------------------------------------------------------ MMA issuers (one per query tile), converged
    const int t = warp - A4_MMA_WARP0;
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc_f16(A4_BK, false);  // S: N = 128 keys, K-major B
    constexpr uint32_t idesc_o = make_idesc_f16(64, true);      // O: N = 64 dims, V read MN-major
    const uint32_t q_addr = smem_u32(sQ + t * A4_Q_BYTES);
    const uint32_t d_s = tmem_base + t * 128;
    const uint32_t d_o = tmem_base + 256 + t * 64;
    const uint32_t a_p = tmem_base + 384 + t * 64;
    auto issue_s = [&](int jj) {  // S_t = Q_t K_jj^T
      const int s = jj % A4_STAGES;
      mbar_wait(&k_full[s], (jj / A4_STAGES) & 1, 64);
      tc_fence_after();
      const uint32_t k_addr = smem_u32(sK + s * A4_KV_BYTES);
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
      const int s = j % A4_STAGES;
      if (j + 1 < p.n_kv) {
        mbar_wait(&s_free[t], j & 1, 66);  // S_t(j) is in the softmax warps' registers
        tc_fence_after();
        issue_s(j + 1);
      }
      mbar_wait(&v_full[s], (j / A4_STAGES) & 1, 65);
      mbar_wait(&p_ready[t], j & 1, 67);
      tc_fence_after();
      const uint32_t v_addr = smem_u32(sV + s * A4_KV_BYTES);
      if (leader) {
#pragma unroll
        for (int k = 0; k < A4_BK / 16; ++k)  // O_t += P_t V_j, P_t (two fp16 per column) from TMEM
          umma_f16_ts4(d_o, a_p + k * 8, make_smem_desc_sw128(v_addr + k * 2048, 1024, 8192), idesc_o,
                       (j | k) != 0 ? 1u : 0u);
        umma_commit(&pv_done[t]);
        umma_commit(&v_empty[s]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax (two threads per query row)
    const int t = warp >> 3;
    const int hf = (warp >> 2) & 1;               // which 64 of the 128 key columns of a block this thread owns
    const int qr = warp & 3;                      // TMEM lane quarter
    const int r = qr * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(qr * 32) << 16;
    const uint32_t tmem_S = tmem_base + t * 128 + hf * 64 + lane_sel;
    const uint32_t tmem_O = tmem_base + 256 + t * 64 + hf * 32 + lane_sel;
    const uint32_t tmem_P = tmem_base + 384 + t * 64 + hf * 32 + lane_sel;
    // turn barriers (128 threads: the two half-row warps of the tile whose turn starts sync, the two of the other tile arrive)
    const int bar_mine = 1 + qr * 2 + t;
    const int bar_theirs = 1 + qr * 2 + (t ^ 1);
    auto turn_sync = [&](int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); };
    auto turn_arrive = [&](int id) { asm volatile("bar.arrive %0, 128;" ::"r"(id) : "memory"); };
    const uint64_t scale2 = a4_pack_f2(p.scale_log2, p.scale_log2);
    const int n_kv = p.n_kv;
    const int handover = p.handover > 3 ? 3 : p.handover;
    float m_used = -CUDART_INF_F;
    float l_run = 0.f;
    if (t == 1) turn_arrive(bar_theirs);  // tile 0 goes first

    for (int j = 0; j < n_kv; ++j) {
      // ---- no MUFU work (runs while the other tile exponentiates): my half of S(j) into registers, half-row maximum
      const int valid = p.S - j * A4_BK - hf * 64;  // my columns [0, valid) of this block are real keys
      A4_TR(j, 0);
      mbar_wait(&s_full[t], j & 1, 69);
      tc_fence_after();
      A4_TR(j, 1);
      uint32_t v[64];
      {
        uint32_t(*v2)[32] = reinterpret_cast<uint32_t(*)[32]>(v);
        tmem_ld_x32(tmem_S, v2[0]);
        tmem_ld_x32(tmem_S + 32, v2[1]);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[t]);  // (8 warp arrivals) the tensor core may overwrite S_t with block j+1
      A4_TR(j, 2);
      if (valid < 64) {  // warp-uniform: only the last key block of an image can be partial
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= valid) v[i] = 0xff800000u;  // -inf
      }
      float mx = fmaxf(a4_max32(v), a4_max32(v + 32));
      float* slot = xch + ((j & 1) * 2 + t) * 256;  // parity double buffer: a slot is rewritten two blocks later
      slot[hf * 128 + r] = mx;
      if (j > 0) {  // PV(j-1) retired: P_t may be rewritten, O_t may be rescaled
        mbar_wait(&pv_done[t], (j - 1) & 1, 70);
        tc_fence_after();
      }
      A4_TR(j, 3);
      // ---- my tile's turn on the MUFU pipe; the barrier also publishes the other half's maximum
      turn_sync(bar_mine);
      A4_TR(j, 4);
      mx = fmaxf(mx, slot[(hf ^ 1) * 128 + r]);
      const float m_blk = mx * p.scale_log2;
      bool raise = false;
      if (j == 0)
        m_used = m_blk;
      else
        raise = m_blk > m_used + 8.0f;
      if (j > 0 && __any_sync(0xffffffffu, raise)) {  // rare: the reference maximum moves (both halves of a row agree)
        const float m_new = raise ? m_blk : m_used;
        const float alpha = fast_exp2(m_used - m_new);
        m_used = m_new;
        l_run *= alpha;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
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
      const uint64_t negm2 = a4_pack_f2(neg_m, neg_m);
      uint64_t lsa = 0ull, lsb = 0ull;
      uint32_t pk[16];
#pragma unroll
      for (int q = 0; q <= 4; ++q) {
        if (q < 4) {  // issue batch q: 16 exponentials, in place
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const int c = q * 16 + i;
            const uint64_t xa = a4_ffma2(a4_pack_f2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])), scale2, negm2);
            const uint64_t xb = a4_ffma2(a4_pack_f2(__uint_as_float(v[c + 2]), __uint_as_float(v[c + 3])), scale2, negm2);
            float x0, x1, x2, x3;
            a4_unpack_f2(xa, x0, x1);
            a4_unpack_f2(xb, x2, x3);
            v[c] = __float_as_uint(fast_exp2(x0));
            v[c + 1] = __float_as_uint(fast_exp2(x1));
            v[c + 2] = __float_as_uint(fast_exp2(x2));
            v[c + 3] = __float_as_uint(fast_exp2(x3));
          }
        }
        if (q == 4 - handover) {  // the other tile may start: its wake-up overlaps my last batches
          if (!(t == 1 && j == n_kv - 1)) turn_arrive(bar_theirs);  // (tile 0 waits n_kv times: initial arrive + n_kv - 1)
        }
        if (q > 0) {  // consume batch q - 1: row sum, fp16 pack, P columns [(q-1)*8, +8) of my half
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const int c = (q - 1) * 16 + i;
            const float p0 = __uint_as_float(v[c]), p1 = __uint_as_float(v[c + 1]);
            const float p2 = __uint_as_float(v[c + 2]), p3 = __uint_as_float(v[c + 3]);
            lsa = a4_fadd2(lsa, a4_pack_f2(p0, p1));
            lsb = a4_fadd2(lsb, a4_pack_f2(p2, p3));
            pk[((q - 1) & 1) * 8 + i / 2] = a4_pack_half2(p0, p1);
            pk[((q - 1) & 1) * 8 + i / 2 + 1] = a4_pack_half2(p2, p3);
          }
          if (((q - 1) & 1) == 1) tmem_st_x16(tmem_P + ((q - 1) >> 1) * 16, pk);
        }
      }
      A4_TR(j, 5);
      float a0, a1, b0, b1;
      a4_unpack_f2(lsa, a0, a1);
      a4_unpack_f2(lsb, b0, b1);
      l_run += (a0 + a1) + (b0 + b1);
      tmem_st_wait();
      A4_TR(j, 6);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[t]);  // (8 warp arrivals)
      A4_TR(j, 7);
    }
    // row sum = the two halves' sums (same reference maximum on both sides)
    {
      float* slot = xch + ((n_kv & 1) * 2 + t) * 256;
      slot[hf * 128 + r] = l_run;
      asm volatile("bar.sync 9, 512;" ::: "memory");  // all 16 softmax warps
      l_run += slot[(hf ^ 1) * 128 + r];
    }
    mbar_wait(&pv_done[t], (n_kv - 1) & 1, 71);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int q = q0 + t * A4_BQ + r;
    __half* dst = p.out + static_cast<long long>(row_base + q) * p.ldo + head * 64 + hf * 32;
    {
      uint32_t o[32];
      tmem_ld_x32(tmem_O, o);
      tmem_ld_wait();
      if (q < p.S) {
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          packed[i] = a4_pack_half2(__uint_as_float(o[2 * i]) * inv_l, __uint_as_float(o[2 * i + 1]) * inv_l);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd)
          *reinterpret_cast<uint4*>(dst + qd * 8) = make_uint4(packed[4 * qd], packed[4 * qd + 1], packed[4 * qd + 2], packed[4 * qd + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == A4_MMA_WARP0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_attn_spatial4(const svdpp_attn_desc* d, cudaStream_t stream) {
  SVDPP_CHECK_ARG(d->heads <= 65535 && d->n_img <= 65535, "attn: grid too large");
  Attn4Params p{};
  p.S = d->S;
  p.n_kv = (d->S + A4_BK - 1) / A4_BK;
  p.q_off = d->q_off;
  p.k_off = d->k_off;
  p.v_off = d->v_off;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.out = static_cast<__half*>(d->out);
  p.ldo = d->ldo;
  const int h = tuning().fmha_handover_split;
  p.handover = h < 0 ? 0 : (h > 3 ? 3 : h);
  p.trace = g_attn_trace;
  CUtensorMap tmQ, tmKV;
  const long long rows = static_cast<long long>(d->n_img) * d->S;
  uint64_t dims[2] = {static_cast<uint64_t>(d->ld), static_cast<uint64_t>(rows)};
  uint64_t str[1] = {static_cast<uint64_t>(d->ld) * 2};
  uint32_t box_q[2] = {64, A4_BQ};
  uint32_t box_kv[2] = {64, A4_BK};
  if (encode_tmap_f16(&tmQ, d->qkv, 2, dims, str, box_q)) return -5;
  if (encode_tmap_f16(&tmKV, d->qkv, 2, dims, str, box_kv)) return -5;
  dim3 grid((d->S + 2 * A4_BQ - 1) / (2 * A4_BQ), d->heads, d->n_img);
  static bool configured = false;
  if (!configured) {
    SVDPP_CUDA(cudaFuncSetAttribute(attn_spatial4_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A4_SMEM_BYTES));
    configured = true;
  }
  SVDPP_CUDA(launch_kernel(attn_spatial4_tc_kernel, grid, dim3(A4_THREADS), A4_SMEM_BYTES, stream, 1, tmQ, tmKV, p));
  return check_launch("attn_spatial4_tc_kernel");
}

}  // namespace svdpp
