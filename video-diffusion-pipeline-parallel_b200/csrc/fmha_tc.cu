// Self-attention kernels, head_dim 64, fp16 in / fp32 softmax / fp16 out, no mask.
//
// svdpp_attn_spatial_f16 (tcgen05): one CTA per (128-query tile, head, image); 2 CTAs per SM.
//   warps 0..3  softmax: thread = query row; S row read from TMEM, online softmax in the log2
//               domain, P written as fp16 into 128B-swizzled smem; O accumulates in TMEM and is
//               rescaled lazily (only when the row maximum grows by more than 2^8)
//   warp 4      TMA producer: Q once, then K_j / V_j blocks of 64 keys through 4-slot rings
//   warp 5      TMEM allocator + MMA issuer: S_j = Q K_j^T (double buffered) and O += P_j V_j (V read
//               MN-major straight from its row-major tile) into TMEM
// (svdpp_attn_temporal_f16, the F <= 32 frame attention, lives in attn_temporal.cu.)
#include <cuda_fp16.h>
#include <math_constants.h>

#include "common.h"
#include "ptx.cuh"

namespace svdpp {

struct AttnParams {
  int S, n_kv;
  int q_off, k_off, v_off;
  float scale_log2;
  __half* out;
  long long ldo;
};

constexpr int ATT_Q_BYTES = 128 * 64 * 2;   // 128 queries x 64 dims
constexpr int ATT_KV_BYTES = 64 * 64 * 2;   // 64 keys x 64 dims
constexpr int ATT_KV_STAGES = 4;
constexpr int ATT_P_BYTES = 128 * 64 * 2;   // 128 queries x 64 keys (one swizzle atom column)
constexpr int ATT_BKV = 64;
constexpr int ATT_SMEM_BYTES = ATT_Q_BYTES + 2 * ATT_KV_STAGES * ATT_KV_BYTES + 2 * ATT_P_BYTES + 256;

// Pipeline per CTA (key blocks of 64, j = 0..n_kv-1):
//   MMA thread :  S_0, S_1 | wait P_j -> O += P_j V_j ; S_{j+2} -> S buffer j&1 | ...
//   softmax    :  wait S_j -> row max -> (rare) rescale O -> wait PV_{j-2} -> P_j = exp2(S_j - m) -> signal
// S and P are double buffered, so the tensor core computes S_{j+1} and P_{j-1}V_{j-1} while the
// softmax warps work on block j; two CTAs per SM fill the remaining gaps.
__global__ void __launch_bounds__(192, 2)
attn_spatial_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                       const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_Q_BYTES;
  uint8_t* sV = sK + ATT_KV_STAGES * ATT_KV_BYTES;
  uint8_t* sP = sV + ATT_KV_STAGES * ATT_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * ATT_P_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;                      // [4]
  uint64_t* k_empty = k_full + ATT_KV_STAGES;       // [4]
  uint64_t* v_full = k_empty + ATT_KV_STAGES;       // [4]
  uint64_t* v_empty = v_full + ATT_KV_STAGES;       // [4]
  uint64_t* s_full = v_empty + ATT_KV_STAGES;       // [2]
  uint64_t* p_ready = s_full + 2;                   // [2]
  uint64_t* pv_done = p_ready + 2;                  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const int row_base = img * p.S;  // first token row of this image in the qkv matrix

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("svdpp: attention smem base not 1024-aligned\n");
    __trap();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < ATT_KV_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_ready[s], 128);
      mbar_init(&pv_done[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 128;  // S buffers: columns [0,64) and [64,128); O: [128,192)
  pdl_launch_dependents();  // programmatic dependent launch: see ptx.cuh
  pdl_wait();

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, ATT_Q_BYTES);
      tma_load_2d(sQ, &tmQ, q_full, p.q_off + head * 64, row_base + q0);
      for (int j = 0; j < p.n_kv; ++j) {
        const int s = j % ATT_KV_STAGES;
        const uint32_t ph = (j / ATT_KV_STAGES) & 1;
        mbar_wait(&k_empty[s], ph ^ 1, 11);
        mbar_expect_tx(&k_full[s], ATT_KV_BYTES);
        tma_load_2d(sK + s * ATT_KV_BYTES, &tmKV, &k_full[s], p.k_off + head * 64, row_base + j * ATT_BKV);
        mbar_wait(&v_empty[s], ph ^ 1, 12);
        mbar_expect_tx(&v_full[s], ATT_KV_BYTES);
        tma_load_2d(sV + s * ATT_KV_BYTES, &tmKV, &v_full[s], p.v_off + head * 64, row_base + j * ATT_BKV);
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(64, false);  // S: N = 64 keys, K-major B
      constexpr uint32_t idesc_o = make_idesc_f16(64, true);   // O: N = 64 dims, V read MN-major
      const uint32_t q_addr = smem_u32(sQ);
      auto issue_s = [&](int jj) {
        const int s = jj % ATT_KV_STAGES;
        mbar_wait(&k_full[s], (jj / ATT_KV_STAGES) & 1, 13);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + s * ATT_KV_BYTES);
        const uint32_t d_s = tmem_base + (jj & 1) * 64;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(d_s, make_smem_desc_sw128(q_addr + k * 32, 1024, 0), make_smem_desc_sw128(k_addr + k * 32, 1024, 0),
                   idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&k_empty[s]);
        umma_commit(&s_full[jj & 1]);
      };
      mbar_wait(q_full, 0, 14);
      issue_s(0);
      if (p.n_kv > 1) issue_s(1);
      for (int j = 0; j < p.n_kv; ++j) {
        const int b = j & 1;
        const int s = j % ATT_KV_STAGES;
        mbar_wait(&p_ready[b], (j >> 1) & 1, 15);  // P_j in smem, S_j fully consumed
        mbar_wait(&v_full[s], (j / ATT_KV_STAGES) & 1, 16);
        tc_fence_after();
        const uint32_t p_addr = smem_u32(sP + b * ATT_P_BYTES);
        const uint32_t v_addr = smem_u32(sV + s * ATT_KV_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_O, make_smem_desc_sw128(p_addr + k * 32, 1024, 0),
                   make_smem_desc_sw128(v_addr + k * 2048, 1024, 8192), idesc_o, (j | k) != 0 ? 1u : 0u);
        umma_commit(&v_empty[s]);
        umma_commit(&pv_done[b]);
        if (j + 2 < p.n_kv) issue_s(j + 2);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax (warps 0..3)
    const int r = threadIdx.x;  // query row in tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    // O stays in TMEM and is accumulated by the MMAs.  The softmax reference maximum m_used is only
    // raised when a block's row maximum exceeds it by more than 2^8 (log2 domain), so probabilities
    // stay <= 256 (exact in fp16/fp32) and the O row needs rescaling only on those rare raises.
    float m_used = -CUDART_INF_F;
    float l_run = 0.f;
    for (int j = 0; j < p.n_kv; ++j) {
      const int b = j & 1;
      const uint32_t tmem_S = tmem_base + b * 64 + lane_sel;
      const int valid = p.S - j * ATT_BKV;  // keys [0, valid) of this block are real
      const bool tail = valid < ATT_BKV;    // warp-uniform: only the last key block of an image can be partial
      mbar_wait(&s_full[b], (j >> 1) & 1, 17);
      tc_fence_after();
      // pass 1: row maximum
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_x32(tmem_S + c * 32, v);
        tmem_ld_wait();
        if (tail) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= valid) v[i] = 0xff800000u;  // -inf
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) mx = fmax3(mx, __uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
      }
      const float m_blk = mx * p.scale_log2;
      bool raise = false;
      if (j == 0)
        m_used = m_blk;
      else
        raise = m_blk > m_used + 8.0f;
      if (j > 0 && __any_sync(0xffffffffu, raise)) {
        // every MMA that has touched O so far must have retired before O is rewritten
        mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1, 18);
        tc_fence_after();
        const float m_new = raise ? m_blk : m_used;
        const float alpha = fast_exp2(m_used - m_new);
        m_used = m_new;
        l_run *= alpha;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_x32(tmem_O + lane_sel + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st_x32(tmem_O + lane_sel + c * 32, v);
        }
        tmem_st_wait();
      }
      if (j >= 2) mbar_wait(&pv_done[b], ((j - 2) >> 1) & 1, 19);  // P buffer b free: P_{j-2} V_{j-2} retired
      // pass 2: probabilities -> fp16 -> swizzled smem
      float lsum = 0.f;
      const float neg_m = -m_used;
      uint8_t* prow = sP + b * ATT_P_BYTES + r * 128;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_x32(tmem_S + c * 32, v);
        tmem_ld_wait();
        if (tail) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= valid) v[i] = 0xff800000u;  // exp2(-inf) = 0
        }
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float x0 = fmaf(__uint_as_float(v[2 * i]), p.scale_log2, neg_m);
          const float x1 = fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2, neg_m);
          const float p0 = fast_exp2(x0);
          const float p1 = fast_exp2(x1);  // (poly_exp2 for a quarter of these was measured 7 % slower: issue-bound)
          lsum += p0 + p1;
          __half2 h = __floats2half2_rn(p0, p1);
          packed[i] = *reinterpret_cast<uint32_t*>(&h);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (c * 4 + q) ^ (r & 7);
          *reinterpret_cast<uint4*>(prow + chunk * 16) =
              make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
        }
      }
      l_run += lsum;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_ready[b]);
    }
    mbar_wait(&pv_done[(p.n_kv - 1) & 1], ((p.n_kv - 1) >> 1) & 1, 20);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int q = q0 + r;
    __half* dst = p.out + static_cast<long long>(row_base + q) * p.ldo + head * 64;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_x32(tmem_O + lane_sel + c * 32, v);
      tmem_ld_wait();
      if (q < p.S) {
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          __half2 h = __floats2half2_rn(__uint_as_float(v[2 * i]) * inv_l, __uint_as_float(v[2 * i + 1]) * inv_l);
          packed[i] = *reinterpret_cast<uint32_t*>(&h);
        }
#pragma unroll
        for (int qd = 0; qd < 4; ++qd)
          *reinterpret_cast<uint4*>(dst + c * 32 + qd * 8) =
              make_uint4(packed[4 * qd], packed[4 * qd + 1], packed[4 * qd + 2], packed[4 * qd + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// Bring-up cross-check: one thread per (image, head, query), online softmax on CUDA cores.
__global__ void attn_spatial_simt_kernel(const __half* qkv, long long ld, AttnParams p, int heads, int n_img) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n_img) * heads * p.S;
  if (idx >= total) return;
  const int q = static_cast<int>(idx % p.S);
  const int head = static_cast<int>((idx / p.S) % heads);
  const int img = static_cast<int>(idx / (static_cast<long long>(p.S) * heads));
  const long long row0 = static_cast<long long>(img) * p.S;
  const __half* qp = qkv + (row0 + q) * ld + p.q_off + head * 64;
  float qf[64], o[64];
  for (int i = 0; i < 64; ++i) {
    qf[i] = __half2float(qp[i]);
    o[i] = 0.f;
  }
  float m = -CUDART_INF_F, l = 0.f;
  for (int k = 0; k < p.S; ++k) {
    const __half* kp = qkv + (row0 + k) * ld + p.k_off + head * 64;
    const __half* vp = qkv + (row0 + k) * ld + p.v_off + head * 64;
    float s = 0.f;
    for (int i = 0; i < 64; ++i) s += qf[i] * __half2float(kp[i]);
    s *= p.scale_log2;
    const float mn = fmaxf(m, s);
    const float a = exp2f(m - mn);
    const float pr = exp2f(s - mn);
    l = l * a + pr;
    const float prh = __half2float(__float2half_rn(pr));
    for (int i = 0; i < 64; ++i) o[i] = o[i] * a + prh * __half2float(vp[i]);
    m = mn;
  }
  __half* dst = p.out + (row0 + q) * p.ldo + head * 64;
  for (int i = 0; i < 64; ++i) dst[i] = __float2half_rn(o[i] / l);
}

}  // namespace svdpp

namespace svdpp {
int launch_attn_spatial2(const svdpp_attn_desc* d, int variant, cudaStream_t stream);  // fmha2_tc.cu
int launch_attn_spatial3(const svdpp_attn_desc* d, cudaStream_t stream);               // fmha3_tc.cu
int launch_attn_spatial4(const svdpp_attn_desc* d, cudaStream_t stream);               // fmha4_tc.cu
}

using namespace svdpp;

extern "C" int svdpp_attn_spatial_f16(const svdpp_attn_desc* d, int impl, svdpp_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SVDPP_CHECK_ARG(d && d->qkv && d->out, "attn: null pointers");
  SVDPP_CHECK_ARG(d->n_img > 0 && d->S > 0 && d->heads > 0, "attn: bad shape");
  SVDPP_CHECK_ARG(d->ld % 8 == 0 && d->ldo % 8 == 0, "attn: pitches must be multiples of 8");
  SVDPP_CHECK_ARG(d->q_off % 8 == 0 && d->k_off % 8 == 0 && d->v_off % 8 == 0, "attn: offsets must be multiples of 8");
  AttnParams p{};
  p.S = d->S;
  p.n_kv = (d->S + ATT_BKV - 1) / ATT_BKV;
  p.q_off = d->q_off;
  p.k_off = d->k_off;
  p.v_off = d->v_off;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.out = static_cast<__half*>(d->out);
  p.ldo = d->ldo;
  if (impl == 2) return launch_attn_spatial2(d, 1, stream);
  if (impl == 3) return launch_attn_spatial2(d, 2, stream);
  if (impl >= 4 && impl <= 6) return launch_attn_spatial2(d, impl, stream);
  if (impl == 7) return launch_attn_spatial3(d, stream);
  if (impl == 8) return launch_attn_spatial4(d, stream);
  if (impl == 1) {
    const long long total = static_cast<long long>(d->n_img) * d->heads * d->S;
    attn_spatial_simt_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, stream>>>(
        static_cast<const __half*>(d->qkv), d->ld, p, d->heads, d->n_img);
    return check_launch("attn_spatial_simt_kernel");
  }
  SVDPP_CHECK_ARG(impl == 0, "attn: unknown impl %d", impl);
  SVDPP_CHECK_ARG(d->heads <= 65535 && d->n_img <= 65535, "attn: grid too large");
  CUtensorMap tmQ, tmKV;
  const long long rows = static_cast<long long>(d->n_img) * d->S;
  uint64_t dims[2] = {static_cast<uint64_t>(d->ld), static_cast<uint64_t>(rows)};
  uint64_t str[1] = {static_cast<uint64_t>(d->ld) * 2};
  uint32_t box_q[2] = {64, 128};
  uint32_t box_kv[2] = {64, ATT_BKV};
  if (encode_tmap_f16(&tmQ, d->qkv, 2, dims, str, box_q)) return -5;
  if (encode_tmap_f16(&tmKV, d->qkv, 2, dims, str, box_kv)) return -5;
  static bool configured = false;
  if (!configured) {
    SVDPP_CUDA(cudaFuncSetAttribute(attn_spatial_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
    configured = true;
  }
  dim3 grid((d->S + 127) / 128, d->heads, d->n_img);
  SVDPP_CUDA(launch_kernel(attn_spatial_tc_kernel, grid, dim3(192), ATT_SMEM_BYTES, stream, 1, tmQ, tmKV, p));
  return check_launch("attn_spatial_tc_kernel");
}
