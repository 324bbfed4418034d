// Bandwidth kernels of the VAE front / back end (AutoencoderKLTemporalDecoder around the
// denoising loop, reference scripts/generate_video_demo.py:92-195).  Everything contraction-shaped in the VAE runs on
// the tcgen05 GEMM / implicit-GEMM conv of gemm_tc.cu; what is left is here:
//   softmax_rows_kernel   the single-head, head_dim-512 attention of the VAE mid blocks is two GEMMs (Q K^T, P V) around
//                         a row softmax: one block per row, logits fp16 in, probabilities fp16 out (in place), fp32 maths
//   transpose_kernel      V [S, C] -> V^T [C, S] so that P V is a GEMM with a K-major B operand
//   time_conv_out_kernel  the decoder's last op, Conv3d (3,1,1) over frames on 3 channels, fused with the change from
//                         channels-last [B*F, H*W, 3] to the caller's [B*F, 3, H, W]
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
