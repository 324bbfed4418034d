// svdpp_unet_*: the whole SVD UNetSpatioTemporalConditionModel forward pass (and one whole denoising step) behind a
// handle, so that the operator the reference calls at src/models/svd_unet.py:389-395 - and the step around it,
// :351-439 - is reachable through the C ABI alone.
//
// What lives here is host code only: weight packing at load time (a few layout kernels below) and the launch sequence
// of one forward (~740 launches of the kernels in gemm_tc.cu / fmha*_tc.cu / attn_temporal.cu / bandwidth.cu).  It is
// the same sequence models/native_unet.py issues through ctypes; tests require the two to agree bit for bit.
//
// Memory: packed weights are owned by the handle (cudaMalloc at load).  Activations live in a CALLER-provided
// workspace: svdpp_unet_workspace_bytes() replays the launch sequence of a forward without launching anything, with
// the same first-fit arena, and returns the high-water mark; the forward then can never run out of space.  Buffers are
// reference-counted handles (released when the last user goes out of scope, like the tensors of the Python
// orchestration); every launch is on one stream, so a released buffer may be handed out again immediately.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.h"

namespace svdpp {

// ------------------------------------------------------------------------------------------ layout kernels (load time)
// dst [n_pad, k_pad] <- src [n, k] (row pitch src_ld), zero padded.  geglu_half > 0: the rows are the GEGLU projection
// [2*inner, k] (value rows, then gate rows) and are regrouped per tile of 2*half rows as [half value | half gate].
__global__ void pack_matrix_kernel(__half* dst, int n_pad, int k_pad, const __half* src, int n, int k, long long src_ld,
                                   int geglu_half, int inner) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(n_pad) * k_pad) return;
  const int r = static_cast<int>(idx / k_pad), c = static_cast<int>(idx % k_pad);
  int sr = r;
  bool ok = r < n;
  if (geglu_half > 0) {
    const int tile = r / (2 * geglu_half), within = r % (2 * geglu_half);
    const int j = tile * geglu_half + (within % geglu_half);
    ok = j < inner;
    sr = within < geglu_half ? j : inner + j;
  }
  dst[idx] = (ok && c < k) ? src[static_cast<long long>(sr) * src_ld + c] : __float2half(0.f);
}

// conv filter [Co, Ci, T] (T = kh*3+kw, or the temporal tap) -> [n_pad, k_pad] with K order (t, ci)
__global__ void pack_conv_kernel(__half* dst, int n_pad, int k_pad, const __half* src, int Co, int Ci, int T) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(n_pad) * k_pad) return;
  const int o = static_cast<int>(idx / k_pad), c = static_cast<int>(idx % k_pad);
  __half v = __float2half(0.f);
  if (o < Co && c < T * Ci) {
    const int t = c / Ci, ci = c % Ci;
    v = src[(static_cast<long long>(o) * Ci + ci) * T + t];
  }
  dst[idx] = v;
}

// "nearest 2x upsample, then Conv2d 3x3 pad 1" as four 2x2-tap convolutions on the low-resolution input, one per output
// parity (py, px): kernel rows / columns that land on the same input pixel are summed in fp32 (kh outer, kw inner - the
// order models/native_unet.py::subpixel_weight sums in) and rounded to fp16 once.  K order (ih, iw, ci).
__global__ void pack_subpixel_kernel(__half* dst, int n_pad, int k_pad, const __half* src, int Co, int Ci, int py, int px) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(n_pad) * k_pad) return;
  const int o = static_cast<int>(idx / k_pad), c = static_cast<int>(idx % k_pad);
  __half v = __float2half(0.f);
  if (o < Co && c < 4 * Ci) {
    const int part = c / Ci, ci = c % Ci;
    const int gy = part >> 1, gx = part & 1;  // which of the two input rows / columns this tap reads
    // parity 0: input offsets {-1: kernel row 0}, {0: rows 1, 2};  parity 1: {0: rows 0, 1}, {+1: row 2}
    int kh0, kh1, kw0, kw1;
    if (py == 0) { kh0 = gy == 0 ? 0 : 1; kh1 = gy == 0 ? 0 : 2; } else { kh0 = gy == 0 ? 0 : 2; kh1 = gy == 0 ? 1 : 2; }
    if (px == 0) { kw0 = gx == 0 ? 0 : 1; kw1 = gx == 0 ? 0 : 2; } else { kw0 = gx == 0 ? 0 : 2; kw1 = gx == 0 ? 1 : 2; }
    float acc = 0.f;
    for (int kh = kh0; kh <= kh1; ++kh)
      for (int kw = kw0; kw <= kw1; ++kw) acc += __half2float(src[((static_cast<long long>(o) * Ci + ci) * 3 + kh) * 3 + kw]);
    v = __float2half_rn(acc);
  }
  dst[idx] = v;
}

// the timestep of every sample as an fp32 device array (sinusoid_embed reads it); a kernel argument, not a host copy, so
// that a captured CUDA graph carries the value
__global__ void fill_f32_kernel(float* dst, float v, int n) {
  if (static_cast<int>(threadIdx.x) < n) dst[threadIdx.x] = v;
}

// ------------------------------------------------------------------------------------------ host-side model
struct Lin {
  __half* w = nullptr;
  __half* b = nullptr;
  int n = 0;       // true output width (after GEGLU halving for geglu)
  int n_pad = 0;   // rows of w
  int k_pad = 0;   // columns of w
  bool geglu = false;
  int impl = -1;   // forced tile shape, -1: choose per call
};
struct Norm {
  __half* g = nullptr;
  __half* b = nullptr;
};
struct Small {   // unpadded [n, k] weight + bias for the handful-of-rows linear kernel
  __half* w = nullptr;
  __half* b = nullptr;
  int n = 0, k = 0;
};
struct ResP {
  float eps = 1e-5f;
  Norm norm1, norm2, tnorm1, tnorm2;
  Lin conv1, conv2, tconv1, tconv2, shortcut;
  bool has_shortcut = false;
  float alpha = 0.5f;
  int temb_sp = 0, temb_t = 0, cout = 0;
};
struct TrP {
  int heads = 1;
  float eps = 1e-6f;
  Norm norm, norm1, norm3, t_norm_in, t_norm1, t_norm3;
  Lin proj_in, qkv1, out1, ff1, ff2, t_ffin1, t_ffin2, t_qkv, t_out1, t_ff1, t_ff2, proj_out;
  Lin ff1f, t_ffin1f, t_ff1f;  // the GEGLU projections packed for the fused feed-forward kernel (C <= 320 only; else empty)
  int ca_off = 0, ca_c = 0, t_ca_off = 0, t_ca_c = 0;
  Small pos1, pos2;
  float alpha = 0.5f;
  int id = 0;   // index into the frame-position cache
};
struct DownBlk {
  std::vector<ResP> res;
  std::vector<TrP> attn;
  bool has_down = false;
  Lin down;
};
struct UpBlk {
  std::vector<ResP> res;
  std::vector<TrP> attn;
  bool has_up = false;
  Lin up;          // 3x3 filter (fallback path)
  Lin up4[2][2];   // parity convolutions
};

struct Buf {   // one activation matrix [rows, cols] fp16 inside the workspace
  struct Arena* arena = nullptr;
  size_t off = 0, bytes = 0;
  __half* ptr = nullptr;
  long long rows = 0;
  int cols = 0;
  int hot = 0;   // which end of the matrix was written last (and is in L2): 0 = the last rows, 1 = the first rows
  ~Buf();
};
using T = std::shared_ptr<Buf>;

// first-fit arena over a caller buffer (or over nothing at all: the dry run only tracks the high-water mark)
struct Arena {
  uint8_t* base = nullptr;
  size_t cap = 0, top = 0, peak = 0;
  std::map<size_t, size_t> free_;   // offset -> bytes, coalesced
  bool overflow = false;
  size_t alloc(size_t bytes) {
    bytes = (bytes + 255) & ~static_cast<size_t>(255);
    for (auto it = free_.begin(); it != free_.end(); ++it)
      if (it->second >= bytes) {
        const size_t off = it->first, rest = it->second - bytes;
        free_.erase(it);
        if (rest) free_[off + bytes] = rest;
        return off;
      }
    const size_t off = top;
    top += bytes;
    peak = std::max(peak, top);
    if (base != nullptr && top > cap) overflow = true;
    return off;
  }
  void release(size_t off, size_t bytes) {
    bytes = (bytes + 255) & ~static_cast<size_t>(255);
    auto it = free_.emplace(off, bytes).first;
    auto nx = std::next(it);
    if (nx != free_.end() && it->first + it->second == nx->first) {
      it->second += nx->second;
      free_.erase(nx);
    }
    if (it != free_.begin()) {
      auto pv = std::prev(it);
      if (pv->first + pv->second == it->first) {
        pv->second += it->second;
        free_.erase(it);
        it = pv;
      }
    }
    if (it->first + it->second == top) {   // give the tail back so the mark reflects live data
      top = it->first;
      free_.erase(it);
    }
  }
};
Buf::~Buf() {
  if (arena) arena->release(off, bytes);
}

struct Epi {   // epilogue of one GEMM (see svdpp_gemm_desc)
  const __half* rowvec = nullptr;
  long long rv_ld = 0;
  int rv_hw = 1, rv_div = 1, rv_mod = 0;
  T r1, r2;
  float beta1 = 1.f, beta2 = 1.f, alpha = 1.f;
};

static const int8_t TAPS_3X3[9][4] = {{-1, -1, 0, 0}, {0, -1, 0, 0}, {1, -1, 0, 0}, {-1, 0, 0, 0}, {0, 0, 0, 0},
                                      {1, 0, 0, 0},   {-1, 1, 0, 0}, {0, 1, 0, 0}, {1, 1, 0, 0}};
static const int8_t TAPS_T3[3][4] = {{0, 0, -1, 0}, {0, 0, 0, 0}, {0, 0, 1, 0}};

constexpr size_t SPLITK_WS_BYTES = 4096 + static_cast<size_t>(74) * 2 * 20 * 128 * 16 * 4;

}  // namespace svdpp

using namespace svdpp;

struct svdpp_unet {
  svdpp_unet_config cfg{};
  int n_levels = 0;
  bool loaded = false;
  std::vector<void*> owned;          // device allocations of the packed weights
  size_t weight_bytes = 0;
  Lin conv_in, conv_out;
  Small time1, time2, add1, add2;
  std::vector<DownBlk> down;
  ResP mid_res[2];
  TrP mid_attn;
  std::vector<UpBlk> up;
  Norm norm_out;
  __half* temb_w = nullptr;
  __half* temb_b = nullptr;
  int temb_total = 0;
  __half* ca_wv = nullptr;
  svdpp_small_group* ca_table = nullptr;
  int ca_groups = 0, ca_total = 0, ca_max_n = 0;
  int n_transformers = 0;
  std::map<std::pair<int, int>, __half*> pos_cache;   // (transformer id, F) -> [F, C]
  int sms = 0;
  long long last_launches = 0;
  // load-time state
  std::unordered_map<std::string, svdpp_tensor_desc> sd;
  std::vector<std::pair<const __half*, long long>> temb_parts_w, temb_parts_b;   // (ptr, elements)
  std::vector<svdpp_small_group> ca_host;
  std::vector<const __half*> ca_wv_parts;
  std::vector<int> ca_wv_rows;
  std::string err;
};

namespace {

// ------------------------------------------------------------------------------------------ weight packing
struct Loader {
  svdpp_unet* u;
  cudaStream_t stream = nullptr;
  bool ok = true;

  void fail(const std::string& m) {
    if (ok) u->err = m;
    ok = false;
  }
  const svdpp_tensor_desc* get(const std::string& key, int ndim_min = 1) {
    auto it = u->sd.find(key);
    if (it == u->sd.end()) {
      fail("missing tensor '" + key + "'");
      return nullptr;
    }
    if (it->second.dtype != 0 || it->second.data == nullptr || it->second.ndim < ndim_min) {
      fail("tensor '" + key + "' must be a contiguous fp16 device tensor");
      return nullptr;
    }
    return &it->second;
  }
  bool has(const std::string& key) const { return u->sd.count(key) != 0; }
  __half* dmalloc(size_t elems) {
    void* p = nullptr;
    if (cudaMalloc(&p, std::max<size_t>(elems, 1) * sizeof(__half)) != cudaSuccess) {
      fail("cudaMalloc of a packed weight failed");
      return nullptr;
    }
    u->owned.push_back(p);
    u->weight_bytes += elems * sizeof(__half);
    return static_cast<__half*>(p);
  }
  static int ceil_to(int x, int m) { return (x + m - 1) / m * m; }
  static unsigned blocks(long long n) { return static_cast<unsigned>((n + 255) / 256); }

  __half* matrix(const __half* src, int n, int k, long long ld, int n_pad, int k_pad, int geglu_half = 0, int inner = 0) {
    __half* d = dmalloc(static_cast<size_t>(n_pad) * k_pad);
    if (!d) return nullptr;
    pack_matrix_kernel<<<blocks(static_cast<long long>(n_pad) * k_pad), 256, 0, stream>>>(d, n_pad, k_pad, src, n, k, ld,
                                                                                       geglu_half, inner);
    return d;
  }
  __half* vec(const std::string& key, int n_pad) {   // 1-D parameter, zero padded to n_pad (0: as is)
    const svdpp_tensor_desc* t = get(key);
    if (!t) return nullptr;
    const int n = static_cast<int>(t->shape[0]);
    return matrix(static_cast<const __half*>(t->data), n, 1, 1, n_pad > 0 ? n_pad : n, 1);
  }
  Norm norm(const std::string& p) { return Norm{vec(p + ".weight", 0), vec(p + ".bias", 0)}; }
  Small small(const std::string& p, bool bias = true) {
    Small s;
    const svdpp_tensor_desc* w = get(p + ".weight", 2);
    if (!w) return s;
    s.n = static_cast<int>(w->shape[0]);
    s.k = static_cast<int>(w->shape[1]);
    s.w = matrix(static_cast<const __half*>(w->data), s.n, s.k, s.k, s.n, s.k);
    if (bias) s.b = vec(p + ".bias", 0);
    return s;
  }
  Lin lin(const std::string& p, bool bias = true) {   // nn.Linear: rows padded to 160, columns to 64
    Lin l;
    const svdpp_tensor_desc* w = get(p + ".weight", 2);
    if (!w) return l;
    l.n = static_cast<int>(w->shape[0]);
    const int k = static_cast<int>(w->shape[1]);
    l.n_pad = ceil_to(l.n, 160);
    l.k_pad = ceil_to(k, 64);
    l.w = matrix(static_cast<const __half*>(w->data), l.n, k, k, l.n_pad, l.k_pad);
    if (bias) l.b = vec(p + ".bias", l.n_pad);
    return l;
  }
  Lin conv(const std::string& p, int taps, bool pad_cols) {   // Conv2d 3x3 (taps 9), Conv3d (3,1,1) (3), 1x1 (1)
    Lin l;
    const svdpp_tensor_desc* w = get(p + ".weight", 2);
    if (!w) return l;
    l.n = static_cast<int>(w->shape[0]);
    const int ci = static_cast<int>(w->shape[1]);
    l.n_pad = ceil_to(l.n, 160);
    l.k_pad = pad_cols ? ceil_to(taps * ci, 64) : taps * ci;
    l.w = dmalloc(static_cast<size_t>(l.n_pad) * l.k_pad);
    if (l.w)
      pack_conv_kernel<<<blocks(static_cast<long long>(l.n_pad) * l.k_pad), 256, 0, stream>>>(
          l.w, l.n_pad, l.k_pad, static_cast<const __half*>(w->data), l.n, ci, taps);
    l.b = vec(p + ".bias", l.n_pad);
    return l;
  }
  void conv_up(const std::string& p, Lin (&out)[2][2]) {
    const svdpp_tensor_desc* w = get(p + ".weight", 4);
    if (!w) return;
    const int co = static_cast<int>(w->shape[0]), ci = static_cast<int>(w->shape[1]);
    __half* b = vec(p + ".bias", ceil_to(co, 160));
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        Lin& l = out[py][px];
        l.n = co;
        l.n_pad = ceil_to(co, 160);
        l.k_pad = ceil_to(4 * ci, 64);
        l.b = b;
        l.w = dmalloc(static_cast<size_t>(l.n_pad) * l.k_pad);
        if (l.w)
          pack_subpixel_kernel<<<blocks(static_cast<long long>(l.n_pad) * l.k_pad), 256, 0, stream>>>(
              l.w, l.n_pad, l.k_pad, static_cast<const __half*>(w->data), co, ci, py, px);
      }
  }
  Lin geglu(const std::string& p) {
    Lin l;
    const svdpp_tensor_desc* w = get(p + ".weight", 2);
    const svdpp_tensor_desc* b = get(p + ".bias");
    if (!w || !b) return l;
    const int inner = static_cast<int>(w->shape[0]) / 2, k = static_cast<int>(w->shape[1]);
    int half = 80;
    if (u->cfg.gemm_impl == 3 && inner % 128 == 0) {   // packed for the 256-wide pair kernel
      half = 128;
      l.impl = 3;
    } else {
      l.impl = u->cfg.gemm_impl == 3 ? 0 : -1;
    }
    const int inner_pad = ceil_to(inner, half);
    l.n = inner;
    l.n_pad = 2 * inner_pad;
    l.k_pad = k;
    l.geglu = true;
    l.w = matrix(static_cast<const __half*>(w->data), 2 * inner, k, k, l.n_pad, k, half, inner);
    l.b = matrix(static_cast<const __half*>(b->data), 2 * inner, 1, 1, l.n_pad, 1, half, inner);
    return l;
  }
  // the same projection interleaved per 128 rows as [64 value | 64 gate] for svdpp_ff_geglu_f16 (C <= 320)
  Lin geglu_fused(const std::string& p) {
    Lin l;
    const svdpp_tensor_desc* w = get(p + ".weight", 2);
    const svdpp_tensor_desc* b = get(p + ".bias");
    if (!w || !b) return l;
    const int inner = static_cast<int>(w->shape[0]) / 2, k = static_cast<int>(w->shape[1]);
    if (k > 320 || k % 64 != 0 || inner != 4 * k) return l;
    l.n = inner;
    l.n_pad = 2 * inner;
    l.k_pad = k;
    l.geglu = true;
    l.w = matrix(static_cast<const __half*>(w->data), 2 * inner, k, k, l.n_pad, k, 64, inner);
    l.b = matrix(static_cast<const __half*>(b->data), 2 * inner, 1, 1, l.n_pad, 1, 64, inner);
    return l;
  }
  Lin fuse_qkv(const std::string& p) {
    Lin l;
    const svdpp_tensor_desc* q = get(p + ".to_q.weight", 2);
    const svdpp_tensor_desc* k = get(p + ".to_k.weight", 2);
    const svdpp_tensor_desc* v = get(p + ".to_v.weight", 2);
    if (!q || !k || !v) return l;
    const int c = static_cast<int>(q->shape[0]), kk = static_cast<int>(q->shape[1]);
    l.n = 3 * c;
    // 3C = 960 / 1920 are not multiples of 256: padding the rows by <= 7 % buys the 256x256 CTA-pair kernel
    const bool pad256 = u->cfg.gemm_impl == 3 && l.n % 256 != 0 && ceil_to(l.n, 256) <= 1.07 * l.n;
    l.n_pad = ceil_to(l.n, pad256 ? 256 : 160);
    l.k_pad = kk;
    l.w = dmalloc(static_cast<size_t>(l.n_pad) * kk);
    if (!l.w) return l;
    cudaMemsetAsync(l.w, 0, static_cast<size_t>(l.n_pad) * kk * sizeof(__half), stream);
    const svdpp_tensor_desc* parts[3] = {q, k, v};
    for (int i = 0; i < 3; ++i)
      cudaMemcpyAsync(l.w + static_cast<size_t>(i) * c * kk, parts[i]->data, static_cast<size_t>(c) * kk * sizeof(__half),
                      cudaMemcpyDeviceToDevice, stream);
    return l;
  }
  float alpha(const std::string& key) {   // AlphaBlender, image_only_indicator == 0: sigmoid(mix_factor), rounded to fp16
    auto it = u->sd.find(key);
    if (it == u->sd.end() || it->second.data == nullptr) {
      fail("missing tensor '" + key + "'");
      return 0.5f;
    }
    float x = 0.f;
    if (it->second.dtype == 1) {
      cudaMemcpy(&x, it->second.data, sizeof(float), cudaMemcpyDeviceToHost);
    } else {
      __half h;
      cudaMemcpy(&h, it->second.data, sizeof(__half), cudaMemcpyDeviceToHost);
      x = __half2float(h);
    }
    const float s = static_cast<float>(1.0 / (1.0 + exp(-static_cast<double>(x))));
    return __half2float(__float2half_rn(s));
  }
  int reg_temb(const std::string& p, int cout) {
    const svdpp_tensor_desc* w = get(p + ".weight", 2);
    const svdpp_tensor_desc* b = get(p + ".bias");
    const int off = u->temb_total;
    if (w && b) {
      u->temb_parts_w.emplace_back(static_cast<const __half*>(w->data), static_cast<long long>(w->shape[0]) * w->shape[1]);
      u->temb_parts_b.emplace_back(static_cast<const __half*>(b->data), static_cast<long long>(b->shape[0]));
    }
    u->temb_total += cout;
    return off;
  }
  void reg_cross(const std::string& p, int* off, int* c) {
    // cross-attention to the single CLIP token: to_v rows go into one stacked weight, to_out.0 into the grouped table
    const svdpp_tensor_desc* wv = get(p + ".to_v.weight", 2);
    Small o = small(p + ".to_out.0");
    *off = u->ca_total;
    *c = wv ? static_cast<int>(wv->shape[0]) : 0;
    if (!wv) return;
    u->ca_wv_parts.push_back(static_cast<const __half*>(wv->data));
    u->ca_wv_rows.push_back(*c);
    svdpp_small_group g{};
    g.W = o.w;
    g.bias = o.b;
    g.x_off = *off;
    g.y_off = *off;
    g.N = o.n;
    g.K = o.k;
    u->ca_host.push_back(g);
    u->ca_max_n = std::max(u->ca_max_n, o.n);
    u->ca_total += *c;
  }
  ResP resblock(const std::string& p, float eps) {
    ResP R;
    const std::string sp = p + ".spatial_res_block", tp = p + ".temporal_res_block";
    R.eps = eps;
    R.norm1 = norm(sp + ".norm1");
    R.conv1 = conv(sp + ".conv1", 9, true);
    R.norm2 = norm(sp + ".norm2");
    R.conv2 = conv(sp + ".conv2", 9, true);
    R.has_shortcut = has(sp + ".conv_shortcut.weight");
    if (R.has_shortcut) R.shortcut = conv(sp + ".conv_shortcut", 1, false);
    R.tnorm1 = norm(tp + ".norm1");
    R.tconv1 = conv(tp + ".conv1", 3, false);
    R.tnorm2 = norm(tp + ".norm2");
    R.tconv2 = conv(tp + ".conv2", 3, false);
    R.alpha = alpha(p + ".time_mixer.mix_factor");
    R.cout = R.conv1.n;
    R.temb_sp = reg_temb(sp + ".time_emb_proj", R.cout);
    R.temb_t = reg_temb(tp + ".time_emb_proj", R.cout);
    return R;
  }
  TrP transformer(const std::string& p, int heads) {
    TrP P;
    P.heads = heads;
    P.eps = u->cfg.eps_transformer;
    P.norm = norm(p + ".norm");
    P.proj_in = lin(p + ".proj_in");
    const std::string s = p + ".transformer_blocks.0", t = p + ".temporal_transformer_blocks.0";
    P.norm1 = norm(s + ".norm1");
    P.qkv1 = fuse_qkv(s + ".attn1");
    P.out1 = lin(s + ".attn1.to_out.0");
    reg_cross(s + ".attn2", &P.ca_off, &P.ca_c);
    P.norm3 = norm(s + ".norm3");
    P.ff1 = geglu(s + ".ff.net.0.proj");
    P.ff1f = geglu_fused(s + ".ff.net.0.proj");
    P.ff2 = lin(s + ".ff.net.2");
    P.t_norm_in = norm(t + ".norm_in");
    P.t_ffin1 = geglu(t + ".ff_in.net.0.proj");
    P.t_ffin1f = geglu_fused(t + ".ff_in.net.0.proj");
    P.t_ffin2 = lin(t + ".ff_in.net.2");
    P.t_norm1 = norm(t + ".norm1");
    P.t_qkv = fuse_qkv(t + ".attn1");
    P.t_out1 = lin(t + ".attn1.to_out.0");
    reg_cross(t + ".attn2", &P.t_ca_off, &P.t_ca_c);
    P.t_norm3 = norm(t + ".norm3");
    P.t_ff1 = geglu(t + ".ff.net.0.proj");
    P.t_ff1f = geglu_fused(t + ".ff.net.0.proj");
    P.t_ff2 = lin(t + ".ff.net.2");
    P.pos1 = small(p + ".time_pos_embed.linear_1");
    P.pos2 = small(p + ".time_pos_embed.linear_2");
    P.alpha = alpha(p + ".time_mixer.mix_factor");
    P.proj_out = lin(p + ".proj_out");
    P.id = u->n_transformers++;
    return P;
  }

  void build() {
    const svdpp_unet_config& c = u->cfg;
    const int nl = u->n_levels, L = c.layers_per_block;
    u->conv_in = conv("conv_in", 9, true);
    u->time1 = small("time_embedding.linear_1");
    u->time2 = small("time_embedding.linear_2");
    u->add1 = small("add_embedding.linear_1");
    u->add2 = small("add_embedding.linear_2");
    for (int i = 0; i < nl && ok; ++i) {
      DownBlk blk;
      const float eps = c.down_attn[i] ? c.eps_down_attn : c.eps_down;
      for (int j = 0; j < L; ++j) {
        blk.res.push_back(resblock("down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), eps));
        if (c.down_attn[i])
          blk.attn.push_back(transformer("down_blocks." + std::to_string(i) + ".attentions." + std::to_string(j),
                                         c.num_attention_heads[i]));
      }
      if (i != nl - 1) {
        blk.has_down = true;
        blk.down = conv("down_blocks." + std::to_string(i) + ".downsamplers.0.conv", 9, true);
      }
      u->down.push_back(std::move(blk));
    }
    u->mid_res[0] = resblock("mid_block.resnets.0", c.eps_mid);
    u->mid_res[1] = resblock("mid_block.resnets.1", c.eps_mid);
    u->mid_attn = transformer("mid_block.attentions.0", c.num_attention_heads[nl - 1]);
    for (int i = 0; i < nl && ok; ++i) {
      UpBlk blk;
      const int ri = nl - 1 - i;   // reversed level
      for (int j = 0; j < L + 1; ++j) {
        blk.res.push_back(resblock("up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), c.eps_up));
        if (c.down_attn[ri])
          blk.attn.push_back(transformer("up_blocks." + std::to_string(i) + ".attentions." + std::to_string(j),
                                         c.num_attention_heads[ri]));
      }
      if (i != nl - 1) {
        blk.has_up = true;
        blk.up = conv("up_blocks." + std::to_string(i) + ".upsamplers.0.conv", 9, true);
        conv_up("up_blocks." + std::to_string(i) + ".upsamplers.0.conv", blk.up4);
      }
      u->up.push_back(std::move(blk));
    }
    u->norm_out = norm("conv_norm_out");
    u->conv_out = conv("conv_out", 9, true);
    if (!ok) return;
    // every time_emb_proj of the network as one [sum cout, 1280] table (one launch per forward)
    long long w_elems = 0, b_elems = 0;
    for (auto& p : u->temb_parts_w) w_elems += p.second;
    for (auto& p : u->temb_parts_b) b_elems += p.second;
    u->temb_w = dmalloc(static_cast<size_t>(w_elems));
    u->temb_b = dmalloc(static_cast<size_t>(b_elems));
    long long wo = 0, bo = 0;
    for (size_t i = 0; i < u->temb_parts_w.size() && ok; ++i) {
      cudaMemcpyAsync(u->temb_w + wo, u->temb_parts_w[i].first, u->temb_parts_w[i].second * sizeof(__half),
                      cudaMemcpyDeviceToDevice, stream);
      cudaMemcpyAsync(u->temb_b + bo, u->temb_parts_b[i].first, u->temb_parts_b[i].second * sizeof(__half),
                      cudaMemcpyDeviceToDevice, stream);
      wo += u->temb_parts_w[i].second;
      bo += u->temb_parts_b[i].second;
    }
    // cross-attention: stacked to_v [sum C, ctx] and the device table of the grouped to_out launch
    const int ctx = c.cross_attention_dim;
    u->ca_wv = dmalloc(static_cast<size_t>(u->ca_total) * ctx);
    long long ro = 0;
    for (size_t i = 0; i < u->ca_wv_parts.size() && ok; ++i) {
      cudaMemcpyAsync(u->ca_wv + ro * ctx, u->ca_wv_parts[i], static_cast<size_t>(u->ca_wv_rows[i]) * ctx * sizeof(__half),
                      cudaMemcpyDeviceToDevice, stream);
      ro += u->ca_wv_rows[i];
    }
    u->ca_groups = static_cast<int>(u->ca_host.size());
    void* tbl = nullptr;
    if (cudaMalloc(&tbl, std::max<size_t>(1, u->ca_host.size()) * sizeof(svdpp_small_group)) != cudaSuccess) {
      fail("cudaMalloc of the cross-attention table failed");
      return;
    }
    u->owned.push_back(tbl);
    u->ca_table = static_cast<svdpp_small_group*>(tbl);
    cudaMemcpyAsync(tbl, u->ca_host.data(), u->ca_host.size() * sizeof(svdpp_small_group), cudaMemcpyHostToDevice, stream);
    if (cudaStreamSynchronize(stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) fail("weight packing failed on the device");
  }
};

// ------------------------------------------------------------------------------------------ one forward
struct Exec {
  svdpp_unet* u;
  Arena arena;
  bool dry;            // account only: no launches, no pointers
  cudaStream_t stream;
  void* gn_ws = nullptr;
  size_t gn_ws_bytes = 0;
  void* sk_ws = nullptr;
  int rc = 0;          // first error
  long long launches = 0;

  T newbuf(long long rows, int cols) {
    auto b = std::make_shared<Buf>();
    b->bytes = static_cast<size_t>(rows) * cols * sizeof(__half);
    b->off = arena.alloc(b->bytes);
    b->arena = &arena;
    b->rows = rows;
    b->cols = cols;
    b->ptr = dry ? nullptr : reinterpret_cast<__half*>(arena.base + b->off);
    return b;
  }
  void note(int e, int n = 1) {
    if (e != 0 && rc == 0) rc = e;
    launches += n;
  }
  bool live() const { return !dry && rc == 0 && !arena.overflow; }

  static bool window_path_ok(int W, int C) { return C % 64 == 0 && ((W >= 8 && 128 % W == 0) || W % 128 == 0); }

  // "zigzag": a kernel that streams over the rows of its input starts at the end its producer wrote LAST - with
  // activations of 150-600 MB against 126 MB of L2 that end is still cached, the other one is not - and leaves its own
  // output hot at the end where it stops.  Results do not depend on the order (tests run both).
  int rev_for(const T& a) const { return (tuning().zigzag && a && a->hot == 0) ? 1 : 0; }

  // tile shape for one GEMM: the candidate with the smallest estimated time = waves over the SMs x tile area / relative
  // per-SM speed (256x256 pairs 1.2, 256x320 pairs 1.15, 128x160 1.0, 128x128 0.8); same rule as NativeUNet._impl
  int pick_impl(const Lin& l, long long M) const {
    if (l.impl >= 0) return l.impl;
    if (u->cfg.gemm_impl != 3) return u->cfg.gemm_impl;
    static const bool no_bn128 = getenv("SVDPP_NO_BN128") != nullptr, no_pair320 = getenv("SVDPP_NO_PAIR320") != nullptr;
    const int N = l.n_pad, K = l.k_pad, sms = u->sms;
    const long long mt = (M + 127) / 128;
    const struct { int impl, bn; bool pair; double speed; } cand[4] = {{3, 256, true, 1.2}, {6, 320, true, 1.15},
                                                                      {0, 160, false, 1.0}, {4, 128, false, 0.8}};
    int best = 0;
    double best_t = -1.0;
    for (const auto& c : cand) {
      if (N % c.bn || (c.impl == 4 && no_bn128)) continue;
      if (c.impl == 6 && (K < 960 || no_pair320)) continue;
      const long long work = c.pair ? ((mt + 1) / 2) * (N / c.bn) : mt * (N / c.bn);
      const long long slots = c.pair ? sms / 2 : sms;
      const double t = static_cast<double>((work + slots - 1) / slots) * 128.0 * c.bn / c.speed;
      if (best_t < 0 || t < best_t) {
        best = c.impl;
        best_t = t;
      }
    }
    return best;
  }

  void gemm_into(T out, const __half* a_ptr, long long lda, long long M, const Lin& l, const Epi& e, const T& a2 = nullptr,
                 int k1 = 0, const int* conv_dims = nullptr, const int8_t (*taps)[4] = nullptr, int ntaps = 0, int impl = -1,
                 int conv_stride = 1, int in_h = 0, int in_w = 0, const int* out_up = nullptr, int rev = 0) {
    launches += 1;
    out->hot = rev;
    if (!live()) return;
    svdpp_gemm_desc d{};
    d.M = static_cast<int32_t>(M);
    d.N = l.n_pad;
    d.K = l.k_pad;
    d.A = a_ptr;
    d.lda = lda;
    if (a2) {
      d.A2 = a2->ptr;
      d.lda2 = a2->cols;
      d.K1 = k1;
    }
    if (conv_dims) {
      d.conv = 1;
      d.cB = conv_dims[0]; d.cF = conv_dims[1]; d.cH = conv_dims[2]; d.cW = conv_dims[3]; d.cC = conv_dims[4];
      d.ntaps = ntaps;
      for (int t = 0; t < ntaps; ++t)
        for (int j = 0; j < 4; ++j) d.taps[t][j] = taps[t][j];
      d.lda = conv_dims[4];
      if (conv_stride > 1) {
        d.conv_stride = conv_stride;
        d.cHin = in_h;
        d.cWin = in_w;
      }
    }
    d.Wt = l.w;
    d.ldw = l.k_pad;
    d.bias = l.b;
    d.rowvec = e.rowvec;
    d.rv_ld = e.rv_ld;
    d.rv_hw = e.rv_hw; d.rv_div = e.rv_div; d.rv_mod = e.rv_mod;
    if (e.r1) { d.R1 = e.r1->ptr; d.ldr1 = e.r1->cols; d.beta1 = e.beta1; }
    if (e.r2) { d.R2 = e.r2->ptr; d.ldr2 = e.r2->cols; d.beta2 = e.beta2; }
    d.alpha = e.alpha;
    d.geglu = l.geglu ? 1 : 0;
    d.D = out->ptr;
    d.ldd = out->cols;
    d.n_store = l.n;
    if (out_up) { d.out_up = out_up[0]; d.out_up_y = out_up[1]; d.out_up_x = out_up[2]; }
    const int im = impl >= 0 ? impl : pick_impl(l, M);
    if (im == 6) {
      d.splitk_ws = sk_ws;
      d.splitk_ws_bytes = static_cast<int64_t>(SPLITK_WS_BYTES);
    }
    tuning().reverse = rev;
    const int r = svdpp_gemm_f16(&d, im, stream);
    tuning().reverse = 0;
    if (r != 0 && rc == 0) rc = r;
  }
  // feed-forward: GEGLU projection l1 then l2 with epilogue e - as ONE kernel (svdpp_ff_geglu_f16) when the "ff_fused"
  // switch is on and the block is narrow enough (l1f packed), else as two GEMMs through a [rows, 4C] intermediate
  T feed_forward(const T& a, const Lin& l1, const Lin& l1f, const Lin& l2, const Epi& e) {
    if (!live() || !(tuning().ff_fused && l1f.w != nullptr)) {  // (the dry run that sizes the arena always plans the larger path)
      T f1 = linear(a, l1);
      return linear(f1, l2, e);
    }
    T out = newbuf(a->rows, l2.n);
    launches += 1;
    out->hot = 0;
    svdpp_ff_desc d{};
    d.M = static_cast<int32_t>(a->rows);
    d.C = l1f.k_pad;
    d.X = a->ptr;
    d.ldx = a->cols;
    d.W1 = l1f.w;
    d.b1 = l1f.b;
    d.W2 = l2.w;
    d.ldw2 = l2.k_pad;
    d.w2_rows = l2.n_pad;
    d.b2 = l2.b;
    d.rowvec = e.rowvec;
    d.rv_ld = e.rv_ld;
    d.rv_hw = e.rv_hw; d.rv_div = e.rv_div; d.rv_mod = e.rv_mod;
    if (e.r1) { d.R1 = e.r1->ptr; d.ldr1 = e.r1->cols; d.beta1 = e.beta1; }
    if (e.r2) { d.R2 = e.r2->ptr; d.ldr2 = e.r2->cols; d.beta2 = e.beta2; }
    d.alpha = e.alpha;
    d.D = out->ptr;
    d.ldd = out->cols;
    const int r = svdpp_ff_geglu_f16(&d, stream);
    if (r != 0 && rc == 0) rc = r;
    return out;
  }
  T linear(const T& a, const Lin& l, const Epi& e = Epi(), const T& a2 = nullptr) {
    T out = newbuf(a->rows, l.n);
    gemm_into(out, a->ptr, a->cols, a->rows, l, e, a2, a2 ? a->cols : 0, nullptr, nullptr, 0, -1, 1, 0, 0, nullptr, rev_for(a));
    return out;
  }
  T conv(const T& a, const Lin& l, int B, int F, int H, int W, int C, const int8_t (*taps)[4], int ntaps, const Epi& e = Epi()) {
    const long long M = static_cast<long long>(B) * F * H * W;
    T out = newbuf(M, l.n);
    if (window_path_ok(W, C)) {
      const int dims[5] = {B, F, H, W, C};
      gemm_into(out, a->ptr, C, M, l, e, nullptr, 0, dims, taps, ntaps, -1, 1, 0, 0, nullptr, rev_for(a));
      return out;
    }
    T cols = newbuf(M, ntaps * C);
    launches += 1;
    if (live())
      note(svdpp_im2col_nhwc(a->ptr, cols->ptr, cols->cols, B, F, H, W, C, H, W, 1, ntaps, &taps[0][0], stream), 0);
    gemm_into(out, cols->ptr, cols->cols, M, l, e);
    return out;
  }
  T gn(const T& x1, const Norm& n, int n_img, int HW, float eps, bool silu = true, const T& x2 = nullptr, int fps = 1) {
    const int C = x1->cols + (x2 ? x2->cols : 0);
    T out = newbuf(x1->rows, C);
    launches += 3;
    // input hot at its end: statistics from the end, apply forward (mode 1); hot at its start: statistics forward, apply
    // from the end (mode 2); the output is hot where the apply pass stops
    const int mode = tuning().zigzag ? (x1->hot == 0 ? 1 : 2) : 0;
    out->hot = mode == 2 ? 1 : 0;
    if (live()) {
      tuning().reverse = mode;
      note(svdpp_groupnorm_silu(x1->ptr, x1->cols, x2 ? x2->ptr : nullptr, x2 ? x2->cols : 0, n.g, n.b, out->ptr, n_img, HW,
                                fps, eps, silu ? 1 : 0, gn_ws, gn_ws_bytes, stream), 0);
      tuning().reverse = 0;
    }
    return out;
  }
  T layernorm(const T& x, const Norm& n, const __half* addvec = nullptr, int add_hw = 1, int add_mod = 1) {
    T out = newbuf(x->rows, x->cols);
    launches += 1;
    const int rev = rev_for(x);
    out->hot = rev;
    if (live()) {
      tuning().reverse = rev;
      note(svdpp_layernorm(x->ptr, x->cols, addvec, add_hw, add_mod, n.g, n.b, out->ptr, out->cols, static_cast<int>(x->rows),
                           x->cols, 1e-5f, stream), 0);
      tuning().reverse = 0;
    }
    return out;
  }
  void small_into(__half* y, long long ldy, const __half* x, const __half* x_add, long long ldx, int R, const __half* w, int N,
                  int K, const __half* b, int act_in, int act_out) {
    launches += 1;
    if (live()) note(svdpp_linear_small(x, x_add, ldx, w, K, b, y, ldy, R, N, K, act_in, act_out, stream), 0);
  }
  T small_mlp(const T& x, const Small& l1, const Small& l2, const __half* x_add = nullptr) {
    T h = newbuf(x->rows, l1.n);
    small_into(h->ptr, h->cols, x->ptr, x_add, x->cols, static_cast<int>(x->rows), l1.w, l1.n, l1.k, l1.b, 0, 1);
    T y = newbuf(x->rows, l2.n);
    small_into(y->ptr, y->cols, h->ptr, nullptr, h->cols, static_cast<int>(h->rows), l2.w, l2.n, l2.k, l2.b, 0, 0);
    return y;
  }

  T resblock(const T& x, const T& skip, const ResP& P, const T& tembs, int B, int F, int H, int W) {
    const int HW = H * W, n_img = B * F, cout = P.cout;
    const int cin = x->cols + (skip ? skip->cols : 0);
    T h1, r;
    {
      T a = gn(x, P.norm1, n_img, HW, P.eps, true, skip);
      Epi e;
      e.rowvec = dry ? nullptr : tembs->ptr + P.temb_sp;
      e.rv_ld = tembs->cols; e.rv_hw = HW; e.rv_div = F;
      if (dry) e.rowvec = nullptr;
      h1 = conv(a, P.conv1, B, F, H, W, cin, TAPS_3X3, 9, e);
    }
    T xs;
    {
      T a2 = gn(h1, P.norm2, n_img, HW, P.eps);
      h1.reset();
      r = P.has_shortcut ? linear(x, P.shortcut, Epi(), skip) : x;
      Epi e;
      e.r1 = r;
      xs = conv(a2, P.conv2, B, F, H, W, cout, TAPS_3X3, 9, e);
    }
    r.reset();
    T t2;
    {
      T t1 = gn(xs, P.tnorm1, n_img, HW, P.eps, true, nullptr, F);
      Epi e;
      e.rowvec = dry ? nullptr : tembs->ptr + P.temb_t;
      e.rv_ld = tembs->cols; e.rv_hw = HW; e.rv_div = F;
      t2 = conv(t1, P.tconv1, B, F, H, W, cout, TAPS_T3, 3, e);
    }
    T t3 = gn(t2, P.tnorm2, n_img, HW, P.eps, true, nullptr, F);
    t2.reset();
    // blend: alpha*xs + (1-alpha)*(xs + h) = xs + (1-alpha)*h
    Epi e;
    e.alpha = 1.0f - P.alpha;
    e.r1 = xs;
    return conv(t3, P.tconv2, B, F, H, W, cout, TAPS_T3, 3, e);
  }

  const __half* pos_embed(const TrP& P, int F, int C) {
    if (dry) return nullptr;
    auto key = std::make_pair(P.id, F);
    auto it = u->pos_cache.find(key);
    if (it != u->pos_cache.end()) return it->second;
    // first forward with this frame count: sinusoid of the frame indices -> MLP, kept for the life of the handle
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cs);
    if (cs != cudaStreamCaptureStatusNone) {
      set_error("unet: the first forward with a new frame count fills the frame-position cache and cannot be captured in a "
                "CUDA graph; run it eagerly once");
      if (rc == 0) rc = -6;
      return nullptr;
    }
    void* p = nullptr;
    if (cudaMalloc(&p, static_cast<size_t>(F) * C * sizeof(__half)) != cudaSuccess) {
      set_error("unet: cudaMalloc of the frame-position cache failed");
      if (rc == 0) rc = -2;
      return nullptr;
    }
    u->owned.push_back(p);
    T s = newbuf(F, C);
    launches += 1;
    if (live()) note(svdpp_sinusoid_embed(nullptr, 2, F, F, C, s->ptr, stream), 0);
    T h = newbuf(F, P.pos1.n);
    small_into(h->ptr, h->cols, s->ptr, nullptr, s->cols, F, P.pos1.w, P.pos1.n, P.pos1.k, P.pos1.b, 0, 1);
    small_into(static_cast<__half*>(p), C, h->ptr, nullptr, h->cols, F, P.pos2.w, P.pos2.n, P.pos2.k, P.pos2.b, 0, 0);
    u->pos_cache[key] = static_cast<__half*>(p);
    return static_cast<__half*>(p);
  }

  T transformer(const T& x, const TrP& P, const T& cvs, int B, int F, int H, int W) {
    const int HW = H * W, n_img = B * F, C = x->cols, heads = P.heads;
    const long long M = x->rows;
    const float scale = 1.0f / sqrtf(static_cast<float>(C / heads));
    const float a = P.alpha;
    T h0;
    {
      T g = gn(x, P.norm, n_img, HW, P.eps, false);
      h0 = linear(g, P.proj_in);
    }
    // --- spatial block
    T h2;
    {
      T att = newbuf(M, C);
      {
        T n1 = layernorm(h0, P.norm1);
        T qkv = linear(n1, P.qkv1);
        n1.reset();
        launches += 1;
        if (live()) {
          svdpp_attn_desc d{};
          d.qkv = qkv->ptr; d.ld = qkv->cols; d.q_off = 0; d.k_off = C; d.v_off = 2 * C;
          d.out = att->ptr; d.ldo = att->cols; d.n_img = n_img; d.S = HW; d.heads = heads; d.scale = scale;
          const int impl = u->cfg.attn_impl >= 0 ? u->cfg.attn_impl : (HW >= 1024 ? u->cfg.attn_impl_long : 0);
          note(svdpp_attn_spatial_f16(&d, impl, stream), 0);
        }
      }
      Epi e;
      e.r1 = h0;
      e.rowvec = dry ? nullptr : cvs->ptr + P.ca_off;
      e.rv_ld = cvs->cols; e.rv_hw = HW; e.rv_div = F;
      h2 = linear(att, P.out1, e);
    }
    h0.reset();
    T hs;
    {
      T n3 = layernorm(h2, P.norm3);
      Epi e;
      e.r1 = h2;
      hs = feed_forward(n3, P.ff1, P.ff1f, P.ff2, e);
    }
    h2.reset();
    // --- temporal block on (hs + frame-position embedding); token (b,f,p) stays at row (b*F+f)*HW+p
    const __half* pos = pos_embed(P, F, C);
    T t1;
    {
      T nin = layernorm(hs, P.t_norm_in, pos, HW, F);
      Epi e;
      e.r1 = hs;
      e.rowvec = pos; e.rv_ld = C; e.rv_hw = HW; e.rv_div = 1; e.rv_mod = F;
      t1 = feed_forward(nin, P.t_ffin1, P.t_ffin1f, P.t_ffin2, e);
    }
    T t2;
    {
      T attt = newbuf(M, C);
      {
        T n1t = layernorm(t1, P.t_norm1);
        T qkvt = linear(n1t, P.t_qkv);
        n1t.reset();
        launches += 1;
        if (live())
          note(svdpp_attn_temporal_f16(qkvt->ptr, qkvt->cols, 0, C, 2 * C, attt->ptr, attt->cols, B, F, HW, heads, scale, stream), 0);
      }
      Epi e;
      e.r1 = t1;
      e.rowvec = dry ? nullptr : cvs->ptr + P.t_ca_off;
      e.rv_ld = cvs->cols; e.rv_hw = HW; e.rv_div = F;
      t2 = linear(attt, P.t_out1, e);
    }
    t1.reset();
    T hb;
    {
      T n3t = layernorm(t2, P.t_norm3);
      // blend fused into the last temporal GEMM: a*hs + (1-a)*(ff + t2)
      Epi e;
      e.alpha = 1.0f - a;
      e.r1 = t2; e.beta1 = 1.0f - a;
      e.r2 = hs; e.beta2 = a;
      hb = feed_forward(n3t, P.t_ff1, P.t_ff1f, P.t_ff2, e);
    }
    t2.reset();
    hs.reset();
    Epi e;
    e.r1 = x;
    return linear(hb, P.proj_out, e);
  }

  // x_in: channels-last [B*F*H*W, in_channels]; returns the channels-last prediction [B*F*H*W, out_channels] in `out`
  void forward_nhwc(const __half* x_in_ptr, const float* t_host_vals, const __half* enc, const __half* ids, __half* out_ptr,
                    int B, int F, int H, int W, __half* t_scratch_unused = nullptr) {
    (void)t_scratch_unused;
    const svdpp_unet_config& c = u->cfg;
    const int boc0 = c.block_out_channels[0];
    const long long M0 = static_cast<long long>(B) * F * H * W;
    // --- embeddings
    T e_t, e_a;
    {
      T s_t = newbuf(B, boc0);
      launches += 1;
      if (live()) note(svdpp_sinusoid_embed(t_host_vals, 0, 0, B, boc0, s_t->ptr, stream), 0);
      e_t = small_mlp(s_t, u->time1, u->time2);
      const int ad = c.addition_time_embed_dim;
      T s_a = newbuf(static_cast<long long>(B) * 3, ad);
      launches += 1;
      if (live()) note(svdpp_sinusoid_embed(ids, 1, 0, B * 3, ad, s_a->ptr, stream), 0);
      s_a->rows = B;            // view [B, 3*ad]
      s_a->cols = 3 * ad;
      e_a = small_mlp(s_a, u->add1, u->add2);
    }
    T tembs = newbuf(B, u->temb_total);   // every time_emb_proj(silu(emb)) of the network at once
    small_into(tembs->ptr, tembs->cols, e_t->ptr, e_a->ptr, e_t->cols, B, u->temb_w, u->temb_total, e_t->cols, u->temb_b, 1, 0);
    e_t.reset();
    e_a.reset();
    // cross-attention with one context token: every block adds to_out(to_v(ctx)); all 32 vectors in two launches
    T cvs = newbuf(B, u->ca_total);
    {
      T h = newbuf(B, u->ca_total);
      small_into(h->ptr, h->cols, enc, nullptr, c.cross_attention_dim, B, u->ca_wv, u->ca_total, c.cross_attention_dim, nullptr, 0, 0);
      launches += 1;
      if (live())
        note(svdpp_linear_small_grouped(h->ptr, h->cols, u->ca_table, u->ca_groups, u->ca_max_n, cvs->ptr, cvs->cols, B, stream), 0);
    }
    // --- conv_in (8 channels: gather the 3x3 windows, K padded 72 -> 128)
    T x;
    {
      T cols = newbuf(M0, u->conv_in.k_pad);
      launches += 1;
      if (live())
        note(svdpp_im2col_nhwc(x_in_ptr, cols->ptr, cols->cols, B, F, H, W, c.in_channels, H, W, 1, 9, &TAPS_3X3[0][0], stream), 0);
      x = newbuf(M0, u->conv_in.n);
      gemm_into(x, cols->ptr, cols->cols, M0, u->conv_in, Epi());
    }
    std::vector<T> skips;
    skips.push_back(x);
    int h = H, w = W;
    for (auto& blk : u->down) {
      for (size_t j = 0; j < blk.res.size(); ++j) {
        x = resblock(x, nullptr, blk.res[j], tembs, B, F, h, w);
        if (!blk.attn.empty()) x = transformer(x, blk.attn[j], cvs, B, F, h, w);
        skips.push_back(x);
      }
      if (blk.has_down) {
        const int C = x->cols, ho = (h + 1) / 2, wo = (w + 1) / 2;
        const long long Mo = static_cast<long long>(B) * F * ho * wo;
        T y = newbuf(Mo, blk.down.n);
        if (window_path_ok(wo, C)) {
          // Conv2d 3x3 stride 2 pad 1 as strided TMA windows (element stride 2 along W, rows 2*ho + dh)
          const int dims[5] = {B, F, ho, wo, C};
          gemm_into(y, x->ptr, C, Mo, blk.down, Epi(), nullptr, 0, dims, TAPS_3X3, 9, -1, 2, h, w, nullptr, rev_for(x));
        } else {
          T cols = newbuf(Mo, 9 * C);
          launches += 1;
          if (live())
            note(svdpp_im2col_nhwc(x->ptr, cols->ptr, cols->cols, B, F, h, w, C, ho, wo, 2, 9, &TAPS_3X3[0][0], stream), 0);
          gemm_into(y, cols->ptr, cols->cols, Mo, blk.down, Epi());
        }
        x = y;
        h = ho;
        w = wo;
        skips.push_back(x);
      }
    }
    x = resblock(x, nullptr, u->mid_res[0], tembs, B, F, h, w);
    x = transformer(x, u->mid_attn, cvs, B, F, h, w);
    x = resblock(x, nullptr, u->mid_res[1], tembs, B, F, h, w);
    static const bool no_subpixel = getenv("SVDPP_NO_SUBPIXEL") != nullptr;
    for (auto& blk : u->up) {
      for (size_t j = 0; j < blk.res.size(); ++j) {
        T skip = skips.back();
        skips.pop_back();
        x = resblock(x, skip, blk.res[j], tembs, B, F, h, w);
        skip.reset();
        if (!blk.attn.empty()) x = transformer(x, blk.attn[j], cvs, B, F, h, w);
      }
      if (blk.has_up) {
        const int C = x->cols;
        if (window_path_ok(w, C) && !no_subpixel) {
          // nearest 2x + 3x3 conv as four 2x2-tap convs on the low-resolution input; each scatters to one output parity
          T out = newbuf(static_cast<long long>(B) * F * 4 * h * w, blk.up.n);
          const int dims[5] = {B, F, h, w, C};
          const int up_rev = rev_for(x);
          for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
              // K order (ih, iw, c): input row offsets {-1, 0} for parity 0, {0, +1} for parity 1; same along W
              int8_t taps[4][4];
              int n = 0;
              for (int gy = 0; gy < 2; ++gy)
                for (int gx = 0; gx < 2; ++gx) {
                  taps[n][0] = static_cast<int8_t>((px == 0 ? -1 : 0) + gx);
                  taps[n][1] = static_cast<int8_t>((py == 0 ? -1 : 0) + gy);
                  taps[n][2] = 0;
                  taps[n][3] = 0;
                  ++n;
                }
              const int up[3] = {2, py, px};
              gemm_into(out, x->ptr, C, static_cast<long long>(B) * F * h * w, blk.up4[py][px], Epi(), nullptr, 0, dims, taps, 4,
                        -1, 1, 0, 0, up, up_rev);
            }
          h *= 2;
          w *= 2;
          x = out;
        } else {
          T upb = newbuf(static_cast<long long>(B) * F * 4 * h * w, C);
          launches += 1;
          if (live()) note(svdpp_upsample2x_nhwc(x->ptr, upb->ptr, B * F, h, w, C, stream), 0);
          h *= 2;
          w *= 2;
          x = conv(upb, blk.up, B, F, h, w, C, TAPS_3X3, 9);
        }
      }
    }
    {
      T a = gn(x, u->norm_out, B * F, h * w, c.eps_out);
      x.reset();
      // conv_out writes straight into the caller's buffer
      auto outb = std::make_shared<Buf>();
      outb->ptr = out_ptr;
      outb->rows = M0;
      outb->cols = c.out_channels;
      const int C = a->cols;
      if (window_path_ok(w, C)) {
        const int dims[5] = {B, F, h, w, C};
        gemm_into(outb, a->ptr, C, M0, u->conv_out, Epi(), nullptr, 0, dims, TAPS_3X3, 9, -1, 1, 0, 0, nullptr, rev_for(a));
      } else {
        T cols = newbuf(M0, 9 * C);
        launches += 1;
        if (live())
          note(svdpp_im2col_nhwc(a->ptr, cols->ptr, cols->cols, B, F, h, w, C, h, w, 1, 9, &TAPS_3X3[0][0], stream), 0);
        gemm_into(outb, cols->ptr, cols->cols, M0, u->conv_out, Epi());
      }
    }
  }
};

// workspace layout: [GroupNorm scratch | split-K scratch | timestep values (fp32, 256 B) | extra fixed region | arena]
struct Layout {
  size_t gn_off = 0, gn_bytes = 0, sk_off = 0, t_off = 0, fixed_off = 0, fixed_bytes = 0, arena_off = 0;
};
static Layout make_layout(int B, int F, int H, int W, size_t fixed_bytes) {
  Layout L;
  auto up = [](size_t x) { return (x + 255) & ~static_cast<size_t>(255); };
  L.gn_bytes = up(svdpp_groupnorm_workspace_bytes(B * F, H * W));
  L.sk_off = L.gn_off + L.gn_bytes;
  L.t_off = L.sk_off + up(SPLITK_WS_BYTES);
  L.fixed_off = L.t_off + 256;
  L.fixed_bytes = up(fixed_bytes);
  L.arena_off = L.fixed_off + L.fixed_bytes;
  return L;
}

static size_t plan(svdpp_unet* u, int B, int F, int H, int W) {
  Exec ex{u};
  ex.dry = true;
  ex.stream = nullptr;
  ex.forward_nhwc(nullptr, nullptr, nullptr, nullptr, nullptr, B, F, H, W);
  // the first forward with a new frame count also builds the frame-position vectors through the arena
  size_t extra = 0;
  for (auto& blk : u->down)
    for (auto& P : blk.attn) extra = std::max(extra, static_cast<size_t>(F) * (P.pos1.k + P.pos1.n) * sizeof(__half) + 1024);
  return ex.arena.peak + extra;
}

static int run_forward(svdpp_unet* u, const __half* x_in, float timestep, const __half* enc, const __half* ids, __half* out,
                       uint8_t* ws, size_t ws_bytes, const Layout& L, int B, int F, int H, int W, cudaStream_t stream,
                       long long* launches) {
  if (ws_bytes < L.arena_off) {
    set_error("unet: workspace too small (%zu bytes)", ws_bytes);
    return -1;
  }
  // counters of the GroupNorm statistics kernel and of the split-K tail must be zero before their first use
  SVDPP_CUDA(cudaMemsetAsync(ws + L.gn_off, 0, 32768, stream));
  SVDPP_CUDA(cudaMemsetAsync(ws + L.sk_off, 0, 4096, stream));
  if (B > 64) {
    set_error("unet: batch %d > 64", B);
    return -1;
  }
  fill_f32_kernel<<<1, 64, 0, stream>>>(reinterpret_cast<float*>(ws + L.t_off), timestep, B);
  if (int e = check_launch("fill_f32_kernel")) return e;
  Exec ex{u};
  ex.dry = false;
  ex.stream = stream;
  ex.arena.base = ws + L.arena_off;
  ex.arena.cap = ws_bytes - L.arena_off;
  ex.gn_ws = ws + L.gn_off;
  ex.gn_ws_bytes = L.gn_bytes;
  ex.sk_ws = ws + L.sk_off;
  ex.forward_nhwc(x_in, reinterpret_cast<const float*>(ws + L.t_off), enc, ids, out, B, F, H, W);
  if (launches) *launches += ex.launches + 3;
  if (ex.arena.overflow) {
    set_error("unet: workspace too small: the activation arena needs %zu bytes, %zu given (svdpp_unet_workspace_bytes)",
              ex.arena.peak, ex.arena.cap);
    return -1;
  }
  return ex.rc;
}

static bool check_shape(const svdpp_unet* u, int B, int F, int H, int W) {
  if (u == nullptr || !u->loaded) {
    set_error("unet: handle has no weights (svdpp_unet_load_weights)");
    return false;
  }
  if (B < 1 || F < 1 || H < 1 || W < 1 || F > 32) {
    set_error("unet: bad shape B=%d F=%d H=%d W=%d (F <= 32)", B, F, H, W);
    return false;
  }
  return true;
}

}  // namespace

// ------------------------------------------------------------------------------------------ C ABI
extern "C" {

int svdpp_unet_create(svdpp_unet** out, const svdpp_unet_config* cfg) {
  SVDPP_CHECK_ARG(out != nullptr && cfg != nullptr, "unet_create: null argument");
  SVDPP_CHECK_ARG(cfg->n_levels >= 1 && cfg->n_levels <= SVDPP_UNET_MAX_LEVELS, "unet_create: n_levels=%d", cfg->n_levels);
  SVDPP_CHECK_ARG(cfg->layers_per_block >= 1 && cfg->in_channels % 8 == 0 && cfg->in_channels > 0 && cfg->out_channels > 0,
                  "unet_create: bad channel / layer counts");
  for (int i = 0; i < cfg->n_levels; ++i) {
    SVDPP_CHECK_ARG(cfg->block_out_channels[i] % 64 == 0 && cfg->block_out_channels[i] > 0,
                    "unet_create: block_out_channels[%d]=%d must be a multiple of 64", i, cfg->block_out_channels[i]);
    SVDPP_CHECK_ARG(!cfg->down_attn[i] || (cfg->num_attention_heads[i] > 0 &&
                                           cfg->block_out_channels[i] == 64 * cfg->num_attention_heads[i]),
                    "unet_create: level %d needs head_dim 64 (channels = 64 * heads)", i);
  }
  svdpp_unet* u = new svdpp_unet();
  u->cfg = *cfg;
  u->n_levels = cfg->n_levels;
  if (u->cfg.attn_impl_long <= 0) u->cfg.attn_impl_long = 7;  // ping-pong FMHA (fmha3_tc.cu)
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      n <= 0) {
    delete u;
    set_error("unet_create: no CUDA device");
    return -2;
  }
  u->sms = n;
  *out = u;
  return 0;
}

int svdpp_unet_load_weights(svdpp_unet* u, const svdpp_tensor_desc* tensors, int n) {
  SVDPP_CHECK_ARG(u != nullptr && tensors != nullptr && n > 0, "unet_load_weights: null argument");
  SVDPP_CHECK_ARG(!u->loaded, "unet_load_weights: weights already loaded (create a new handle)");
  for (int i = 0; i < n; ++i) {
    SVDPP_CHECK_ARG(tensors[i].name != nullptr, "unet_load_weights: tensor %d has no name", i);
    u->sd[tensors[i].name] = tensors[i];
  }
  Loader ld{u};
  ld.build();
  u->sd.clear();
  u->temb_parts_w.clear();
  u->temb_parts_b.clear();
  u->ca_wv_parts.clear();
  if (!ld.ok) {
    set_error("unet_load_weights: %s", u->err.c_str());
    return -1;
  }
  u->loaded = true;
  return 0;
}

size_t svdpp_unet_weight_bytes(const svdpp_unet* u) { return u ? u->weight_bytes : 0; }

size_t svdpp_unet_workspace_bytes(svdpp_unet* u, int B, int F, int H, int W) {
  if (!check_shape(u, B, F, H, W)) return 0;
  const size_t M = static_cast<size_t>(B) * F * H * W;
  // fixed region: the channels-last input and prediction of svdpp_unet_forward / svdpp_unet_step (batch 2B with guidance)
  const size_t fixed = 2 * M * (u->cfg.in_channels + u->cfg.out_channels) * sizeof(__half) + 1024;
  const Layout L = make_layout(2 * B, F, H, W, fixed);
  return L.arena_off + std::max(plan(u, B, F, H, W), plan(u, 2 * B, F, H, W));
}

long long svdpp_unet_last_launches(const svdpp_unet* u) { return u ? u->last_launches : 0; }

int svdpp_unet_forward_nhwc(svdpp_unet* u, const void* x_in, float timestep, const void* enc, const void* added_time_ids,
                            void* out, void* workspace, size_t ws_bytes, int B, int F, int H, int W, svdpp_stream stream_) {
  if (!check_shape(u, B, F, H, W)) return -1;
  SVDPP_CHECK_ARG(x_in && enc && added_time_ids && out && workspace, "unet_forward_nhwc: null pointer");
  const Layout L = make_layout(B, F, H, W, 0);
  u->last_launches = 0;
  return run_forward(u, static_cast<const __half*>(x_in), timestep, static_cast<const __half*>(enc),
                     static_cast<const __half*>(added_time_ids), static_cast<__half*>(out), static_cast<uint8_t*>(workspace),
                     ws_bytes, L, B, F, H, W, static_cast<cudaStream_t>(stream_), &u->last_launches);
}

int svdpp_unet_forward(svdpp_unet* u, const void* sample, float timestep, const void* enc, const void* added_time_ids,
                       void* out, void* workspace, size_t ws_bytes, int B, int F, int H, int W, svdpp_stream stream_) {
  if (!check_shape(u, B, F, H, W)) return -1;
  SVDPP_CHECK_ARG(sample && enc && added_time_ids && out && workspace, "unet_forward: null pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int Ci = u->cfg.in_channels, Co = u->cfg.out_channels;
  const size_t M = static_cast<size_t>(B) * F * H * W;
  const Layout L = make_layout(B, F, H, W, M * (Ci + Co) * sizeof(__half));
  SVDPP_CHECK_ARG(ws_bytes >= L.arena_off, "unet_forward: workspace too small");
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  __half* x_in = reinterpret_cast<__half*>(ws + L.fixed_off);
  __half* v = x_in + M * Ci;
  u->last_launches = 2;
  // [B, F, C, H, W] -> channels-last
  int r = svdpp_pack_unet_input(sample, static_cast<int64_t>(F) * Ci * H * W, static_cast<int64_t>(Ci) * H * W,
                                static_cast<int64_t>(H) * W, Ci, 1.0f, nullptr, 0, 0, 0, 0, x_in, 0, B, F, H, W, stream_);
  if (r != 0) return r;
  r = run_forward(u, x_in, timestep, static_cast<const __half*>(enc), static_cast<const __half*>(added_time_ids), v, ws,
                  ws_bytes, L, B, F, H, W, stream, &u->last_launches);
  if (r != 0) return r;
  return svdpp_nhwc_to_bfchw(v, out, B, F, Co, H, W, stream_);
}

int svdpp_unet_step(svdpp_unet* u, const void* latent, const void* image_latents, const void* uncond_image_latents,
                    const void* enc, const void* added_time_ids, const void* gs, float timestep, float in_div, float c_v,
                    float c_x, float sigma, float dt, void* out, void* workspace, size_t ws_bytes, int B, int F, int H, int W,
                    svdpp_stream stream_) {
  return svdpp_unet_step_handoff(u, latent, image_latents, uncond_image_latents, enc, added_time_ids, gs, timestep, in_div, c_v,
                                 c_x, sigma, dt, out, workspace, ws_bytes, B, F, H, W, nullptr, stream_);
}

int svdpp_unet_step_handoff(svdpp_unet* u, const void* latent, const void* image_latents, const void* uncond_image_latents,
                            const void* enc, const void* added_time_ids, const void* gs, float timestep, float in_div,
                            float c_v, float c_x, float sigma, float dt, void* out, void* workspace, size_t ws_bytes, int B,
                            int F, int H, int W, const svdpp_handoff* ho, svdpp_stream stream_) {
  if (!check_shape(u, B, F, H, W)) return -1;
  SVDPP_CHECK_ARG(latent && image_latents && enc && added_time_ids && out && workspace, "unet_step: null pointer");
  SVDPP_CHECK_ARG(uncond_image_latents == nullptr || gs != nullptr, "unet_step: guidance needs the per-frame scale vector");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool cfg = uncond_image_latents != nullptr;
  const int nb = cfg ? 2 * B : B;
  const int Cl = u->cfg.out_channels;                  // latent channels
  const int Ci = u->cfg.in_channels;
  SVDPP_CHECK_ARG(Ci == 2 * Cl, "unet_step: in_channels must be 2 x out_channels (latent | image latent)");
  const size_t M = static_cast<size_t>(B) * F * H * W;
  const Layout L = make_layout(nb, F, H, W, static_cast<size_t>(nb) * F * H * W * (Ci + Cl) * sizeof(__half));
  SVDPP_CHECK_ARG(ws_bytes >= L.arena_off, "unet_step: workspace too small");
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  __half* x_in = reinterpret_cast<__half*>(ws + L.fixed_off);
  __half* v = x_in + static_cast<size_t>(nb) * F * H * W * Ci;
  // element strides (b, f, c) of [B, C, F, H, W]
  const int64_t sb = static_cast<int64_t>(Cl) * F * H * W, sf = static_cast<int64_t>(H) * W, sc = static_cast<int64_t>(F) * H * W;
  u->last_launches = cfg ? 3 : 2;
  int r;
  if (cfg) {   // batch 2: unconditional half first (reference svd_unet.py:384-411 runs the two sequentially)
    r = svdpp_pack_unet_input(latent, sb, sf, sc, Cl, in_div, uncond_image_latents, sb, sf, sc, Cl, x_in, 0, B, F, H, W, stream_);
    if (r != 0) return r;
    r = svdpp_pack_unet_input(latent, sb, sf, sc, Cl, in_div, image_latents, sb, sf, sc, Cl, x_in + M * Ci, 0, B, F, H, W, stream_);
  } else {
    r = svdpp_pack_unet_input(latent, sb, sf, sc, Cl, in_div, image_latents, sb, sf, sc, Cl, x_in, 0, B, F, H, W, stream_);
  }
  if (r != 0) return r;
  r = run_forward(u, x_in, timestep, static_cast<const __half*>(enc), static_cast<const __half*>(added_time_ids), v, ws,
                  ws_bytes, L, nb, F, H, W, stream, &u->last_launches);
  if (r != 0) return r;
  return svdpp_euler_vpred_step_signal(latent, v, cfg ? v + M * Cl : nullptr, cfg ? gs : nullptr, 1, c_v, c_x, sigma, dt, out,
                                       B, Cl, F, H, W, ho, stream_);
}

void svdpp_unet_destroy(svdpp_unet* u) {
  if (u == nullptr) return;
  for (void* p : u->owned) cudaFree(p);
  delete u;
}

}  // extern "C"
