// Bandwidth kernels of the VAE front / back end (AutoencoderKLTemporalDecoder around the
// denoising loop, reference scripts/generate_video_demo.py:92-195).  Everything contraction-shaped in the VAE runs on
// the tcgen05 GEMM / implicit-GEMM conv of gemm_tc.cu; what is left is here:
//   softmax_rows_kernel   the single-head, head_dim-512 attention of the VAE mid blocks is two GEMMs (Q K^T, P V) around
//                         a row softmax: one block per row, logits fp16 in, probabilities fp16 out (in place), fp32 maths
//   transpose_kernel      V [S, C] -> V^T [C, S] so that P V is a GEMM with a K-major B operand
//   time_conv_out_kernel  the decoder's last op, Conv3d (3,1,1) over frames on 3 channels, fused with the change from
//                         channels-last [B*F, H*W, 3] to the caller's [B*F, 3, H, W]
//   frames_to_bytes_kernel the output format of the run: decoded frames [3, F, H, W] in [-1, 1] -> packed RGB bytes [F, H, W, 3]
//                         (the reference's ((x + 1) / 2 * 255).clamp(0, 255).to(uint8), bit for bit) and / or indices into a
//                         fixed 6 x 7 x 6 colour cube with 4 x 4 ordered dither, so that the GIF writer on the host only
//                         has to LZW-pack (its own adaptive palette search costs 5 s per 25-frame 576 x 1024 video)
//   attn_small_kernel     whole attention of a short sequence (the CLIP image encoder: 257 tokens, 16 heads of width 80) in
//                         one launch per layer: K and V of one (image, head) resident in shared memory, fp32 CUDA-core
//                         maths (10 GFLOP over the 32 layers - launch count, not arithmetic, was the cost of the
//                         GEMM -> softmax -> transpose -> GEMM chain per head: 2048 launches per image)
#include <cuda_fp16.h>

#include "common.h"
#include "ptx.cuh"

namespace svdpp {

// one block (256 threads) per row; the row (<= 16384 logits) is held in registers between the passes
__global__ void __launch_bounds__(256) softmax_rows_kernel(__half* __restrict__ x, long long ld, int n, int n_valid,
                                                           float scale_log2) {
  pdl_launch_dependents();
  pdl_wait();
  __half* row = x + static_cast<long long>(blockIdx.x) * ld;
  constexpr int MAXV = 8;                     // 8 x (256 threads x 8 halves) = 16384 columns
  uint4 v[MAXV];
  float mx = -3.0e38f;
  const int nvec = n >> 3;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = threadIdx.x + k * 256;
    if (vi < nvec) {
      v[k] = *reinterpret_cast<const uint4*>(row + vi * 8);
      const __half2* h = reinterpret_cast<const __half2*>(&v[k]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        if (vi * 8 + 2 * j < n_valid) mx = fmaxf(mx, f.x);          // columns >= n_valid are padding keys
        if (vi * 8 + 2 * j + 1 < n_valid) mx = fmaxf(mx, f.y);
      }
    }
  }
  __shared__ float red[8];
  __shared__ float bc;
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    bc = m;
  }
  __syncthreads();
  mx = bc;
  float sum = 0.f;
  float e[MAXV][8];
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = threadIdx.x + k * 256;
    if (vi < nvec) {
      const __half2* h = reinterpret_cast<const __half2*>(&v[k]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        e[k][2 * j] = vi * 8 + 2 * j < n_valid ? fast_exp2((f.x - mx) * scale_log2) : 0.f;
        e[k][2 * j + 1] = vi * 8 + 2 * j + 1 < n_valid ? fast_exp2((f.y - mx) * scale_log2) : 0.f;
        sum += e[k][2 * j] + e[k][2 * j + 1];
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    bc = 1.0f / s;
  }
  __syncthreads();
  const float inv = bc;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = threadIdx.x + k * 256;
    if (vi < nvec) {
      uint4 o;
      __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(e[k][2 * j] * inv, e[k][2 * j + 1] * inv);
      *reinterpret_cast<uint4*>(row + vi * 8) = o;
    }
  }
}

// [R, C] (row pitch ld_in) -> [C, R] (row pitch ld_out), 32 x 32 tiles through padded shared memory
__global__ void transpose_kernel(const __half* __restrict__ in, long long ld_in, __half* __restrict__ out, long long ld_out,
                                 int R, int C) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __half tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[i][threadIdx.x] = in[static_cast<long long>(r) * ld_in + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) out[static_cast<long long>(c) * ld_out + r] = tile[threadIdx.x][i];
  }
}

// x: channels-last [B, F, HW, xc >= 3] fp16 (the first 3 channels are read); w [3 (co), 3 (ci), 3 (kt)] fp16, bias [3];
// out [B*F, 3, HW] (fp16 or fp32)
template <typename OutT>
__global__ void time_conv_out_kernel(const __half* __restrict__ x, int xc, const __half* __restrict__ w,
                                     const __half* __restrict__ bias, OutT* __restrict__ out, int B, int F, long long HW) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sw[27], sb[3];
  if (threadIdx.x < 27) sw[threadIdx.x] = __half2float(w[threadIdx.x]);
  if (threadIdx.x < 3) sb[threadIdx.x] = __half2float(bias[threadIdx.x]);
  __syncthreads();
  const long long total = static_cast<long long>(B) * F * HW;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = idx % HW;
    const long long bf = idx / HW;
    const int f = static_cast<int>(bf % F);
    float acc[3] = {sb[0], sb[1], sb[2]};
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) {
      const int ff = f + kt - 1;
      if (ff < 0 || ff >= F) continue;
      const __half* px = x + ((bf + (kt - 1)) * HW + p) * xc;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float v = __half2float(px[ci]);
#pragma unroll
        for (int co = 0; co < 3; ++co) acc[co] = fmaf(v, sw[(co * 3 + ci) * 3 + kt], acc[co]);
      }
    }
#pragma unroll
    for (int co = 0; co < 3; ++co) out[(bf * 3 + co) * HW + p] = static_cast<OutT>(acc[co]);
  }
}

// Attention over a short sequence.  Grid (ceil(S_pad / 32), heads, images), 256 threads: warp w owns queries 4w .. 4w + 3 of
// the block's 32.  K / V of the (image, head) sit in shared memory as fp16 pairs with an ODD row pitch in words, so that
// lanes reading consecutive keys (scores) or consecutive channel pairs (P V) never collide on a bank.
//   scores   lane l holds keys l, l + 32, ... (KPL per lane) of its warp's four queries in registers, fp32
//   softmax  warp shuffles; probabilities normalised in fp32 and parked in shared memory as one float4 (4 queries) per key
//   P V      lane l owns channel pairs l and l + 32; one broadcast float4 + two V words per key
// Rows [S, S_pad) of an image (padding tokens) and columns [head_dim, out_head_stride) of a head are written as zeros.
template <int KPL>
__global__ void __launch_bounds__(256)
attn_small_kernel(const __half* __restrict__ qkv, long long ld, int q_off, int k_off, int v_off, int head_stride,
                  __half* __restrict__ out, long long ldo, int out_head_stride, int S, int S_pad, int hd, float scale_log2) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ uint32_t sm_small[];
  const int hw = hd >> 1;                    // words (fp16 pairs) per head row
  const int kw = hw | 1;                     // odd row pitch of K / V in shared memory
  uint32_t* sK = sm_small;                   // [S][kw]
  uint32_t* sV = sK + S * kw;                // [S][kw]
  uint32_t* sQ = sm_small + ((2 * S * kw + 3) & ~3);   // [32][hw], 16-byte aligned
  float4* sP = reinterpret_cast<float4*>(sQ + 32 * hw);  // [8 warps][KPL * 32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 32, head = blockIdx.y;
  const long long row_base = static_cast<long long>(blockIdx.z) * S_pad;
  const int ohw = out_head_stride >> 1;
  __half* const out_head = out + head * out_head_stride;

  if (q0 >= S) {  // padding tokens only: zeros
    for (int idx = threadIdx.x; idx < 32 * ohw; idx += 256) {
      const int r = idx / ohw, c = idx % ohw;
      if (q0 + r < S_pad) *reinterpret_cast<uint32_t*>(out_head + (row_base + q0 + r) * ldo + 2 * c) = 0u;
    }
    return;
  }
  const int v8 = hd >> 3;  // 16-byte vectors per head row
  for (int idx = threadIdx.x; idx < S * v8; idx += 256) {
    const int r = idx / v8, c = idx % v8;
    const __half* src = qkv + (row_base + r) * ld + head * head_stride + c * 8;
    const uint4 k4 = *reinterpret_cast<const uint4*>(src + k_off);
    const uint4 v4 = *reinterpret_cast<const uint4*>(src + v_off);
    uint32_t* dk = sK + r * kw + c * 4;
    uint32_t* dv = sV + r * kw + c * 4;
    dk[0] = k4.x; dk[1] = k4.y; dk[2] = k4.z; dk[3] = k4.w;
    dv[0] = v4.x; dv[1] = v4.y; dv[2] = v4.z; dv[3] = v4.w;
  }
  for (int idx = threadIdx.x; idx < 32 * v8; idx += 256) {
    const int r = idx / v8, c = idx % v8;
    uint4 q4 = make_uint4(0u, 0u, 0u, 0u);
    if (q0 + r < S) q4 = *reinterpret_cast<const uint4*>(qkv + (row_base + q0 + r) * ld + q_off + head * head_stride + c * 8);
    *reinterpret_cast<uint4*>(sQ + r * hw + c * 4) = q4;
  }
  __syncthreads();

  // ---- scores of queries 4w .. 4w + 3 against keys lane + 32 i
  float acc[4][KPL];
  int krow[KPL];
#pragma unroll
  for (int i = 0; i < KPL; ++i) {
    const int k = lane + 32 * i;
    krow[i] = (k < S ? k : S - 1) * kw;  // keys >= S are masked below; read a valid row meanwhile
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) acc[qi][i] = 0.f;
  }
  const uint32_t* qrow = sQ + (4 * warp) * hw;
  for (int dp = 0; dp < hw; ++dp) {
    float2 qf[4];
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) qf[qi] = __half22float2(*reinterpret_cast<const __half2*>(qrow + qi * hw + dp));
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      const float2 kf = __half22float2(*reinterpret_cast<const __half2*>(sK + krow[i] + dp));
#pragma unroll
      for (int qi = 0; qi < 4; ++qi) acc[qi][i] = fmaf(qf[qi].x, kf.x, fmaf(qf[qi].y, kf.y, acc[qi][i]));
    }
  }
  // ---- softmax (fp32), probabilities to shared memory
  float4* myP = sP + warp * (KPL * 32);
#pragma unroll
  for (int qi = 0; qi < 4; ++qi) {
    float mx = -3.0e38f;
#pragma unroll
    for (int i = 0; i < KPL; ++i)
      if (lane + 32 * i < S) mx = fmaxf(mx, acc[qi][i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      acc[qi][i] = lane + 32 * i < S ? fast_exp2((acc[qi][i] - mx) * scale_log2) : 0.f;
      sum += acc[qi][i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int i = 0; i < KPL; ++i) acc[qi][i] *= inv;
  }
#pragma unroll
  for (int i = 0; i < KPL; ++i) myP[lane + 32 * i] = make_float4(acc[0][i], acc[1][i], acc[2][i], acc[3][i]);
  __syncwarp();
  // ---- P V: lane owns channel pairs lane and lane + 32
  const int dp0 = lane < hw ? lane : 0, dp1 = lane + 32 < hw ? lane + 32 : 0;
  float2 o0[4], o1[4];
#pragma unroll
  for (int qi = 0; qi < 4; ++qi) o0[qi] = o1[qi] = make_float2(0.f, 0.f);
  for (int k = 0; k < S; ++k) {
    const float4 pr = myP[k];
    const float2 va = __half22float2(*reinterpret_cast<const __half2*>(sV + k * kw + dp0));
    const float2 vb = __half22float2(*reinterpret_cast<const __half2*>(sV + k * kw + dp1));
    const float pq[4] = {pr.x, pr.y, pr.z, pr.w};
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
      o0[qi].x = fmaf(pq[qi], va.x, o0[qi].x);
      o0[qi].y = fmaf(pq[qi], va.y, o0[qi].y);
      o1[qi].x = fmaf(pq[qi], vb.x, o1[qi].x);
      o1[qi].y = fmaf(pq[qi], vb.y, o1[qi].y);
    }
  }
#pragma unroll
  for (int qi = 0; qi < 4; ++qi) {
    const int q = q0 + 4 * warp + qi;
    if (q >= S_pad) continue;
    __half* dst = out_head + (row_base + q) * ldo;
    const bool real = q < S;
    if (lane < ohw)
      *reinterpret_cast<__half2*>(dst + 2 * lane) = (real && lane < hw) ? __floats2half2_rn(o0[qi].x, o0[qi].y) : __half2();
    if (lane + 32 < ohw)
      *reinterpret_cast<__half2*>(dst + 2 * (lane + 32)) =
          (real && lane + 32 < hw) ? __floats2half2_rn(o1[qi].x, o1[qi].y) : __half2();
  }
}

template <int KPL>
static int launch_attn_small(dim3 grid, size_t smem, cudaStream_t stream, const __half* qkv, long long ld, int q_off, int k_off,
                             int v_off, int head_stride, __half* out, long long ldo, int out_head_stride, int S, int S_pad, int hd,
                             float scale_log2) {
  static size_t configured = 0;
  if (smem > configured) {
    SVDPP_CUDA(cudaFuncSetAttribute(attn_small_kernel<KPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  SVDPP_CUDA(launch_kernel(attn_small_kernel<KPL>, grid, dim3(256), smem, stream, 1, qkv, ld, q_off, k_off, v_off, head_stride, out,
                           ldo, out_head_stride, S, S_pad, hd, scale_log2));
  return check_launch("attn_small_kernel");
}

// ------------------------------------------------------------------------------------ frames -> bytes
__device__ __forceinline__ float4 load_quad(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load_quad(const __half* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
// ((x + 1) / 2 * 255).clamp(0, 255).to(uint8): every operation rounded on its own, as torch's elementwise kernels
__device__ __forceinline__ unsigned to_byte(float x) {
  float v = __fmul_rn(__fmul_rn(__fadd_rn(x, 1.0f), 0.5f), 255.0f);
  v = fminf(fmaxf(v, 0.0f), 255.0f);          // NaN -> 0
  return static_cast<unsigned>(v);            // truncation
}
// level of byte v in a cube axis of L levels: floor(v * (L - 1) / 255 + threshold), threshold in (0, 1)
__device__ __forceinline__ unsigned cube_level(unsigned v, int L, float thr) {
  const float t = __fadd_rn(__fmul_rn(static_cast<float>(v), static_cast<float>(L - 1) / 255.0f), thr);
  const int q = static_cast<int>(t);
  return static_cast<unsigned>(q > L - 1 ? L - 1 : q);
}
__constant__ float kBayer4[16] = {0.5f / 16,  8.5f / 16,  2.5f / 16,  10.5f / 16, 12.5f / 16, 4.5f / 16, 14.5f / 16, 6.5f / 16,
                                  3.5f / 16,  11.5f / 16, 1.5f / 16,  9.5f / 16,  15.5f / 16, 7.5f / 16, 13.5f / 16, 5.5f / 16};

// x: element (c, f, p) at x[c * sc + f * sf + p], p = y * W + col, W % 4 == 0.  One thread per 4 pixels of a row: three
// 16-byte (fp32) / 8-byte (fp16) loads, 12 bytes of RGB and 4 bytes of indices out, all aligned.  HBM-bound.
template <typename T>
__global__ void __launch_bounds__(256) frames_to_bytes_kernel(const T* __restrict__ x, long long sc, long long sf, int F,
                                                              long long HW, int W, unsigned char* __restrict__ rgb,
                                                              unsigned char* __restrict__ idx, int dither) {
  pdl_launch_dependents();
  pdl_wait();
  const long long qpf = HW >> 2, quads = qpf * F;
  for (long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < quads;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long f = q / qpf, p = (q - f * qpf) << 2;
    const T* src = x + f * sf + p;
    const float4 c0 = load_quad(src), c1 = load_quad(src + sc), c2 = load_quad(src + 2 * sc);
    const float r[4] = {c0.x, c0.y, c0.z, c0.w}, g[4] = {c1.x, c1.y, c1.z, c1.w}, b[4] = {c2.x, c2.y, c2.z, c2.w};
    unsigned R[4], G[4], B[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      R[j] = to_byte(r[j]);
      G[j] = to_byte(g[j]);
      B[j] = to_byte(b[j]);
    }
    const long long o = f * HW + p;
    if (rgb != nullptr) {
      uint3 w;
      w.x = R[0] | (G[0] << 8) | (B[0] << 16) | (R[1] << 24);
      w.y = G[1] | (B[1] << 8) | (R[2] << 16) | (G[2] << 24);
      w.z = B[2] | (R[3] << 8) | (G[3] << 16) | (B[3] << 24);
      unsigned* dst = reinterpret_cast<unsigned*>(rgb + o * 3);
      dst[0] = w.x;
      dst[1] = w.y;
      dst[2] = w.z;
    }
    if (idx != nullptr) {
      const int y = static_cast<int>(p / W);        // the quad starts at a column that is a multiple of 4
      unsigned packed = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float thr = dither ? kBayer4[(y & 3) * 4 + j] : 0.5f;
        packed |= (cube_level(R[j], 6, thr) * 42u + cube_level(G[j], 7, thr) * 6u + cube_level(B[j], 6, thr)) << (8 * j);
      }
      *reinterpret_cast<unsigned*>(idx + o) = packed;
    }
  }
}

static inline unsigned vae_grid_for(long long n, int threads, int max_blocks = 148 * 16) {
  long long b = (n + threads - 1) / threads;
  if (b > max_blocks) b = max_blocks;
  if (b < 1) b = 1;
  return static_cast<unsigned>(b);
}

}  // namespace svdpp

using namespace svdpp;

extern "C" int svdpp_softmax_rows(void* x, int64_t ld, int32_t rows, int32_t n, int32_t n_valid, float scale,
                                  svdpp_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SVDPP_CHECK_ARG(x != nullptr && rows > 0 && n > 0, "softmax_rows: bad arguments");
  if (n_valid <= 0 || n_valid > n) n_valid = n;
  SVDPP_CHECK_ARG(n % 8 == 0 && ld % 8 == 0 && n <= 16384, "softmax_rows: n=%d must be a multiple of 8 and <= 16384", n);
  launch_kernel(softmax_rows_kernel, dim3(rows), dim3(256), 0, stream, 1, static_cast<__half*>(x), static_cast<long long>(ld), n,
                n_valid, scale * 1.4426950408889634f);
  return check_launch("softmax_rows_kernel");
}

extern "C" int svdpp_attn_small_f16(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, int32_t head_stride,
                                    void* out, int64_t ldo, int32_t out_head_stride, int32_t n_img, int32_t S, int32_t S_pad,
                                    int32_t heads, int32_t head_dim, float scale, svdpp_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SVDPP_CHECK_ARG(qkv && out && n_img > 0 && heads > 0 && S > 0 && S_pad >= S, "attn_small: bad arguments");
  SVDPP_CHECK_ARG(S <= 512, "attn_small: S=%d exceeds 512 keys (longer sequences: svdpp_attn_spatial_f16)", S);
  SVDPP_CHECK_ARG(head_dim >= 8 && head_dim <= 128 && head_dim % 8 == 0, "attn_small: head_dim=%d must be a multiple of 8, <= 128",
                  head_dim);
  SVDPP_CHECK_ARG(head_stride >= head_dim && out_head_stride >= head_dim && out_head_stride <= 128 && out_head_stride % 2 == 0,
                  "attn_small: head strides must cover head_dim (output stride <= 128, even)");
  SVDPP_CHECK_ARG(ld % 8 == 0 && q_off % 8 == 0 && k_off % 8 == 0 && v_off % 8 == 0 && head_stride % 8 == 0 && ldo % 2 == 0,
                  "attn_small: pitches and offsets must keep 16-byte alignment");
  SVDPP_CHECK_ARG(heads <= 65535 && n_img <= 65535, "attn_small: grid too large");
  const int hw = head_dim / 2, kw = hw | 1;
  const int kpl = S <= 288 ? 9 : 16;
  const size_t words = ((2 * static_cast<size_t>(S) * kw + 3) & ~static_cast<size_t>(3)) + 32 * hw + static_cast<size_t>(8) * kpl * 32 * 4;
  const size_t smem = words * 4;
  SVDPP_CHECK_ARG(smem <= 227 * 1024, "attn_small: S=%d x head_dim=%d needs %zu bytes of shared memory", S, head_dim, smem);
  const dim3 grid((S_pad + 31) / 32, heads, n_img);
  const float sl2 = scale * 1.4426950408889634f;
  if (kpl == 9)
    return launch_attn_small<9>(grid, smem, stream, static_cast<const __half*>(qkv), ld, q_off, k_off, v_off, head_stride,
                                static_cast<__half*>(out), ldo, out_head_stride, S, S_pad, head_dim, sl2);
  return launch_attn_small<16>(grid, smem, stream, static_cast<const __half*>(qkv), ld, q_off, k_off, v_off, head_stride,
                               static_cast<__half*>(out), ldo, out_head_stride, S, S_pad, head_dim, sl2);
}

extern "C" int svdpp_transpose_f16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int32_t R, int32_t C,
                                   svdpp_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SVDPP_CHECK_ARG(in && out && R > 0 && C > 0, "transpose: bad arguments");
  launch_kernel(transpose_kernel, dim3((C + 31) / 32, (R + 31) / 32), dim3(32, 8), 0, stream, 1, static_cast<const __half*>(in),
                static_cast<long long>(ld_in), static_cast<__half*>(out), static_cast<long long>(ld_out), R, C);
  return check_launch("transpose_kernel");
}

extern "C" int svdpp_time_conv_out(const void* x, int32_t x_channels, const void* w, const void* bias, void* out,
                                   int32_t out_fp32, int32_t B, int32_t F, int64_t HW, svdpp_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SVDPP_CHECK_ARG(x && w && bias && out && B > 0 && F > 0 && HW > 0 && x_channels >= 3, "time_conv_out: bad arguments");
  const long long total = static_cast<long long>(B) * F * HW;
  if (out_fp32)
    launch_kernel(time_conv_out_kernel<float>, dim3(vae_grid_for(total, 256)), dim3(256), 0, stream, 1, static_cast<const __half*>(x),
                  x_channels, static_cast<const __half*>(w), static_cast<const __half*>(bias), static_cast<float*>(out), B, F,
                  static_cast<long long>(HW));
  else
    launch_kernel(time_conv_out_kernel<__half>, dim3(vae_grid_for(total, 256)), dim3(256), 0, stream, 1, static_cast<const __half*>(x),
                  x_channels, static_cast<const __half*>(w), static_cast<const __half*>(bias), static_cast<__half*>(out), B, F,
                  static_cast<long long>(HW));
  return check_launch("time_conv_out_kernel");
}

extern "C" int svdpp_frames_to_bytes(const void* frames, int32_t is_fp32, int64_t stride_c, int64_t stride_f, int32_t F,
                                     int32_t H, int32_t W, void* rgb, void* idx, int32_t dither, svdpp_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SVDPP_CHECK_ARG(frames != nullptr && (rgb != nullptr || idx != nullptr), "frames_to_bytes: null pointer");
  SVDPP_CHECK_ARG(F > 0 && H > 0 && W > 0 && W % 4 == 0, "frames_to_bytes: W=%d must be a positive multiple of 4", W);
  SVDPP_CHECK_ARG(stride_c % 4 == 0 && stride_f % 4 == 0, "frames_to_bytes: strides must be multiples of 4 elements");
  const long long HW = static_cast<long long>(H) * W;
  const unsigned grid = vae_grid_for(HW / 4 * F, 256);
  if (is_fp32)
    launch_kernel(frames_to_bytes_kernel<float>, dim3(grid), dim3(256), 0, stream, 1, static_cast<const float*>(frames),
                  static_cast<long long>(stride_c), static_cast<long long>(stride_f), F, HW, W, static_cast<unsigned char*>(rgb),
                  static_cast<unsigned char*>(idx), dither);
  else
    launch_kernel(frames_to_bytes_kernel<__half>, dim3(grid), dim3(256), 0, stream, 1, static_cast<const __half*>(frames),
                  static_cast<long long>(stride_c), static_cast<long long>(stride_f), F, HW, W, static_cast<unsigned char*>(rgb),
                  static_cast<unsigned char*>(idx), dither);
  return check_launch("frames_to_bytes_kernel");
}
