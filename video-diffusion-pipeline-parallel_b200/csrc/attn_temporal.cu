// svdpp_attn_temporal_f16: self-attention over the F <= 32 frames of every (batch, pixel, head), head_dim 64.
//
// Replaces attn1 of diffusers' TemporalBasicTransformerBlock (reached from svd_unet.py:389-395).  The
// reference transposes the activation to [B*H*W, F, C] first; here token (b, f, p) stays at row
// (b*F + f)*HW + p of the channels-last qkv matrix and the kernel gathers the F rows of a unit itself.
//
// HBM-bound (reads q, k, v once = 3*F*128 B per unit, writes F*128 B; 0.1 % of the UNet's FLOPs), so the
// design is a streaming one: each warp owns a private ring of STAGES smem slots and an mbarrier per slot;
// its lane 0 gathers the F frame rows of a (pixel, head) with ONE TMA tensor load per q/k/v (4-D map
// [B][F][HW][cols], box 64 cols x 32 frames, 128-byte swizzle, frames >= F zero-filled) STAGES-1 units
// ahead, and writes the result back with one TMA tensor store.  No block-level barrier anywhere.  The two
// small products run on the tensor cores (mma.sync m16n8k16, fragments via ldmatrix from the swizzled
// tiles), so the math is a few hundred cycles per unit.
#include <cuda_fp16.h>
#include <math_constants.h>
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace svdpp {

constexpr int TA_TILE_BYTES = 32 * 128;            // 32 frame rows (padded) x 64 fp16
constexpr int TA_SLOT_BYTES = 3 * TA_TILE_BYTES;   // q, k, v

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// byte offset of 16-byte chunk `ch` of row `r` inside a 1024-aligned tile: the TMA 128-byte swizzle
// (chunk ^ row%8), which also makes ldmatrix and the staged output writes bank-conflict free
__device__ __forceinline__ uint32_t ta_swz(int r, int ch) { return r * 128 + ((ch ^ (r & 7)) << 4); }

struct TAttnParams {
  int q_off, k_off, v_off;
  int F, HW, heads;
  long long units;  // B * HW * heads
  float scale_log2;
};

// tmIn : qkv as a 4-D tensor [B][F][HW][ld], box 64 columns x 1 pixel x 32 frames: one TMA load gathers the
//        (up to 32) frame rows of one pixel and head into a swizzled [32][64] tile, zero-filling frames >= F.
// tmOut: the output matrix the same way; the store clips frames >= F.
template <int TA_WARPS, int TA_STAGES, int CTAS_PER_SM>
__global__ void __launch_bounds__(TA_WARPS * 32, CTAS_PER_SM)
attn_temporal_mma_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOut,
                         const TAttnParams p) {
  extern __shared__ uint8_t ta_smem_raw[];
  uint8_t* ta_smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ta_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(ta_smem + TA_WARPS * TA_STAGES * TA_SLOT_BYTES);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* wbase = ta_smem + warp * (TA_STAGES * TA_SLOT_BYTES);
  const uint32_t wbase_u = smem_u32(wbase);
  uint64_t* full = bars + warp * TA_STAGES;
  const int F = p.F;

  if (lane == 0) {
    if (warp == 0) {
      tma_prefetch_desc(&tmIn);
      tma_prefetch_desc(&tmOut);
    }
    for (int s = 0; s < TA_STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncwarp();
  pdl_launch_dependents();  // programmatic dependent launch: see ptx.cuh
  pdl_wait();

  const long long stride = static_cast<long long>(gridDim.x) * TA_WARPS;
  long long unit = static_cast<long long>(blockIdx.x) * TA_WARPS + warp;

  auto issue = [&](long long u, int slot) {  // lane 0 only
    if (u < p.units) {
      const int head = static_cast<int>(u % p.heads);
      const long long bp = u / p.heads;
      const int pix = static_cast<int>(bp % p.HW);
      const int b = static_cast<int>(bp / p.HW);
      uint8_t* dst = wbase + slot * TA_SLOT_BYTES;
      mbar_expect_tx(&full[slot], TA_SLOT_BYTES);
      tma_load_4d(dst, &tmIn, &full[slot], p.q_off + head * 64, pix, 0, b);
      tma_load_4d(dst + TA_TILE_BYTES, &tmIn, &full[slot], p.k_off + head * 64, pix, 0, b);
      tma_load_4d(dst + 2 * TA_TILE_BYTES, &tmIn, &full[slot], p.v_off + head * 64, pix, 0, b);
    }
  };

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < TA_STAGES - 1; ++s) issue(unit + s * stride, s);
  }

  const int g = lane >> 2, q = lane & 3;
  const int n_mt = (F + 15) >> 4;  // 16-row query tiles
  const int n_nt = (F + 7) >> 3;   // 8-key tiles
  int slot = 0;
  uint32_t phase = 0;
  for (; unit < p.units; unit += stride) {
    if (lane == 0) {
      int nslot = slot + TA_STAGES - 1;
      if (nslot >= TA_STAGES) nslot -= TA_STAGES;
      // the slot being refilled was last read by the TMA store issued TA_STAGES-1 iterations ago
      bulk_wait_read<TA_STAGES - 2>();
      issue(unit + (TA_STAGES - 1) * stride, nslot);
    }
    mbar_wait(&full[slot], phase, 30);
    const uint32_t sq = wbase_u + slot * TA_SLOT_BYTES;
    const uint32_t sk = sq + TA_TILE_BYTES, sv = sk + TA_TILE_BYTES;
    uint8_t* so = wbase + slot * TA_SLOT_BYTES;  // output staging = the Q tile

#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      if (mt < n_mt) {
        // ---- S = Q K^T for query rows [16 mt, 16 mt + 16)
        float s[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int j = 0; j < 4; ++j) s[nt][j] = 0.f;
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {  // head dims [32 kp, 32 kp + 32)
          uint32_t a0[4], a1[4];
          const int arow = 16 * mt + (lane & 7) + ((lane >> 3) & 1) * 8;
          ldsm_x4(sq + ta_swz(arow, 4 * kp + (lane >> 4)), a0);
          ldsm_x4(sq + ta_swz(arow, 4 * kp + 2 + (lane >> 4)), a1);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            if (nt < n_nt) {
              uint32_t kb[4];
              ldsm_x4(sk + ta_swz(8 * nt + (lane & 7), 4 * kp + (lane >> 3)), kb);
              mma_16816(s[nt], a0, kb[0], kb[1]);
              mma_16816(s[nt], a1, kb[2], kb[3]);
            }
          }
        }
        // ---- softmax over the F valid keys (log2 domain), rows g and g + 8 of this tile
        float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = 8 * nt + 2 * q + (j & 1);
            s[nt][j] = col < F ? s[nt][j] * p.scale_log2 : -CUDART_INF_F;
          }
          mx0 = fmax3(mx0, s[nt][0], s[nt][1]);
          mx1 = fmax3(mx1, s[nt][2], s[nt][3]);
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        float l0 = 0.f, l1 = 0.f;
        uint32_t pa[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const float p0 = fast_exp2(s[nt][0] - mx0), p1 = fast_exp2(s[nt][1] - mx0);
          const float p2 = fast_exp2(s[nt][2] - mx1), p3 = fast_exp2(s[nt][3] - mx1);
          l0 += p0 + p1;
          l1 += p2 + p3;
          pa[nt][0] = pack_h2(p0, p1);
          pa[nt][1] = pack_h2(p2, p3);
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        // ---- O = P V
        float o[8][4];
#pragma unroll
        for (int dt = 0; dt < 8; ++dt)
#pragma unroll
          for (int j = 0; j < 4; ++j) o[dt][j] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {  // keys [16 kk, 16 kk + 16)
          if (kk < n_mt) {
            const uint32_t a[4] = {pa[2 * kk][0], pa[2 * kk][1], pa[2 * kk + 1][0], pa[2 * kk + 1][1]};
            const int vrow = 16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {
              uint32_t vb[4];
              ldsm_x4_t(sv + ta_swz(vrow, 2 * dp + (lane >> 4)), vb);
              mma_16816(o[2 * dp], a, vb[0], vb[1]);
              mma_16816(o[2 * dp + 1], a, vb[2], vb[3]);
            }
          }
        }
        // ---- normalise, fp16, into the Q tile (its rows of this m-tile are consumed; rows >= F are clipped
        //      by the store, so they may hold anything)
        const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
        const int r0 = 16 * mt + g, r1 = r0 + 8;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
          *reinterpret_cast<uint32_t*>(so + ta_swz(r0, dt) + q * 4) = pack_h2(o[dt][0] * inv0, o[dt][1] * inv0);
          *reinterpret_cast<uint32_t*>(so + ta_swz(r1, dt) + q * 4) = pack_h2(o[dt][2] * inv1, o[dt][3] * inv1);
        }
      }
    }
    fence_proxy_async_smem();  // generic-proxy writes of O -> visible to the TMA store
    __syncwarp();
    if (lane == 0) {
      const int head = static_cast<int>(unit % p.heads);
      const long long bp = unit / p.heads;
      tma_store_4d(&tmOut, so, head * 64, static_cast<int>(bp % p.HW), 0, static_cast<int>(bp / p.HW));
      bulk_commit();
    }
    if (++slot == TA_STAGES) {
      slot = 0;
      phase ^= 1;
    }
  }
  if (lane == 0) bulk_wait_read<0>();  // smem must outlive the last store's read
}

template <int WARPS, int STAGES, int CTAS_PER_SM>
static int launch_ta(const CUtensorMap& tmIn, const CUtensorMap& tmOut, const TAttnParams& p, cudaStream_t stream) {
  static_assert(STAGES >= 2, "need a slot to prefetch into");
  constexpr int smem = WARPS * STAGES * TA_SLOT_BYTES + 1024 /*align*/ + WARPS * STAGES * 8 /*barriers*/;
  auto kern = attn_temporal_mma_kernel<WARPS, STAGES, CTAS_PER_SM>;
  static bool configured = false;
  if (!configured) {
    SVDPP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  long long blocks = (p.units + WARPS - 1) / WARPS;
  if (blocks > static_cast<long long>(num_sms()) * CTAS_PER_SM) blocks = static_cast<long long>(num_sms()) * CTAS_PER_SM;
  SVDPP_CUDA(launch_kernel(kern, dim3(static_cast<unsigned>(blocks)), dim3(WARPS * 32), smem, stream, 1, tmIn, tmOut, p));
  return check_launch("attn_temporal_mma_kernel");
}

}  // namespace svdpp

using namespace svdpp;

extern "C" int svdpp_attn_temporal_f16(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off,
                                        void* out, int64_t ldo, int32_t B, int32_t F, int32_t HW, int32_t heads,
                                        float scale, svdpp_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SVDPP_CHECK_ARG(qkv && out, "attn_temporal: null pointers");
  SVDPP_CHECK_ARG(F >= 1 && F <= 32, "attn_temporal: F=%d must be in [1,32]", F);
  SVDPP_CHECK_ARG(B >= 1 && HW >= 1 && heads >= 1, "attn_temporal: bad shape");
  SVDPP_CHECK_ARG(ld % 8 == 0 && ldo % 8 == 0 && q_off % 8 == 0 && k_off % 8 == 0 && v_off % 8 == 0,
                  "attn_temporal: pitches/offsets must be multiples of 8");
  TAttnParams p{};
  p.q_off = q_off;
  p.k_off = k_off;
  p.v_off = v_off;
  p.F = F;
  p.HW = HW;
  p.heads = heads;
  p.units = static_cast<long long>(B) * HW * heads;
  p.scale_log2 = scale * 1.4426950408889634f;
  CUtensorMap tmIn, tmOut;
  {
    const uint32_t box[4] = {64, 1, 32, 1};
    uint64_t dims[4] = {static_cast<uint64_t>(ld), static_cast<uint64_t>(HW), static_cast<uint64_t>(F),
                        static_cast<uint64_t>(B)};
    uint64_t str[3] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(ld) * 2 * HW,
                       static_cast<uint64_t>(ld) * 2 * HW * F};
    if (encode_tmap_f16(&tmIn, qkv, 4, dims, str, box)) return -5;
    dims[0] = static_cast<uint64_t>(ldo);
    str[0] = static_cast<uint64_t>(ldo) * 2;
    str[1] = str[0] * HW;
    str[2] = str[1] * F;
    if (encode_tmap_f16(&tmOut, out, 4, dims, str, box)) return -5;
  }
  // (warps per CTA, ring slots per warp, CTAs per SM): 12 KB per slot; variant 0 is the measured best,
  // SVDPP_TA_VARIANT selects the others for experiments
  static const int variant = [] {
    const char* e = getenv("SVDPP_TA_VARIANT");
    return e ? atoi(e) : 0;
  }();
  if (variant == 1) return launch_ta<6, 3, 1>(tmIn, tmOut, p, stream);
  if (variant == 2) return launch_ta<4, 2, 2>(tmIn, tmOut, p, stream);
  if (variant == 3) return launch_ta<4, 3, 1>(tmIn, tmOut, p, stream);
  if (variant == 4) return launch_ta<3, 3, 2>(tmIn, tmOut, p, stream);
  return launch_ta<8, 2, 1>(tmIn, tmOut, p, stream);
}
