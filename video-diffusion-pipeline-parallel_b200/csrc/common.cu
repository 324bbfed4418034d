// Error state, device query and the tensor-map encoder (driver entry point fetched at run time so
// the library links without libcuda and loads on a machine that has no GPU).
#include "common.h"

#include <cudaTypedefs.h>
#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace svdpp {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

Tuning& tuning() {
  static Tuning t = [] {
    Tuning v;
    const char* e = getenv("SVDPP_NO_TMA_STORE");
    v.tma_store = (e != nullptr && e[0] != '0') ? 0 : 1;
    e = getenv("SVDPP_PDL");
    v.pdl = (e != nullptr && e[0] != '0') ? 1 : 0;
    e = getenv("SVDPP_NO_TMA_R1");
    v.tma_r1 = (e != nullptr && e[0] != '0') ? 0 : 1;
    e = getenv("SVDPP_EPI_DMA");
    v.epi_dma = e != nullptr ? atoi(e) : 1;
    e = getenv("SVDPP_TWO_PROD");
    v.two_prod = e != nullptr ? atoi(e) : 1;   // in-situ A/B (profiles/r2_ab_two_producers.json): 93.5 -> 91.1 ms per step, bit-identical
    v.two_prod_min_kb = 6;
    e = getenv("SVDPP_NO_SPLITK");
    v.splitk = (e != nullptr && e[0] != '0') ? 0 : 1;
    v.splitk_min_kb = 4;
    v.splitk_min_total_kb = 64;
    e = getenv("SVDPP_NO_R1_PREFETCH");
    v.r1_prefetch_max_kb = (e != nullptr && e[0] != '0') ? 0 : 5;
    e = getenv("SVDPP_EPI_DMA_MAX_KB");
    v.epi_dma_max_kb = e != nullptr ? atoi(e) : 5;
    e = getenv("SVDPP_FMHA_STAGGER");
    v.fmha_stagger = e != nullptr ? atoi(e) : 900;   // in-situ A/B (profiles/r2_ab_attn_insitu.json): 94.8 -> 92.9 ms per step with impl 4
    e = getenv("SVDPP_FMHA_HANDOVER");
    v.fmha_handover = e != nullptr ? atoi(e) : 2;
    e = getenv("SVDPP_FMHA_POLY");
    v.fmha_poly = e != nullptr ? atoi(e) : 0;
    e = getenv("SVDPP_FMHA_HANDOVER_SPLIT");
    v.fmha_handover_split = e != nullptr ? atoi(e) : 1;
    e = getenv("SVDPP_FF_FUSED");
    v.ff_fused = e != nullptr ? atoi(e) : 0;
    e = getenv("SVDPP_FF_PAIR");
    v.ff_pair = e != nullptr ? atoi(e) : 2;
    v.ff_dbg = 0;
    v.reverse = 0;
    v.reverse_gn_apply_same = 0;
    e = getenv("SVDPP_ZIGZAG");
    v.zigzag = e != nullptr ? atoi(e) : 0;
    return v;
  }();
  return t;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                    int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return -4;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = elem_strides ? elem_strides[i] : 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                  gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank=%d dims=[%llu,%llu,..] stride0=%llu box=[%u,%u,..] base=%p",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0, base);
    return -5;
  }
  return 0;
}

}  // namespace svdpp

extern "C" {

int svdpp_abi_version(void) { return SVDPP_ABI_VERSION; }

const char* svdpp_last_error(void) { return svdpp::g_err; }

static int* tuning_slot(const char* key) {
  if (key == nullptr) return nullptr;
  if (strcmp(key, "tma_store") == 0) return &svdpp::tuning().tma_store;
  if (strcmp(key, "pdl") == 0) return &svdpp::tuning().pdl;
  if (strcmp(key, "tma_r1") == 0) return &svdpp::tuning().tma_r1;
  if (strcmp(key, "two_prod") == 0) return &svdpp::tuning().two_prod;
  if (strcmp(key, "two_prod_min_kb") == 0) return &svdpp::tuning().two_prod_min_kb;
  if (strcmp(key, "epi_dma") == 0) return &svdpp::tuning().epi_dma;
  if (strcmp(key, "epi_dma_max_kb") == 0) return &svdpp::tuning().epi_dma_max_kb;
  if (strcmp(key, "splitk") == 0) return &svdpp::tuning().splitk;
  if (strcmp(key, "splitk_min_kb") == 0) return &svdpp::tuning().splitk_min_kb;
  if (strcmp(key, "r1_prefetch_max_kb") == 0) return &svdpp::tuning().r1_prefetch_max_kb;
  if (strcmp(key, "splitk_min_total_kb") == 0) return &svdpp::tuning().splitk_min_total_kb;
  if (strcmp(key, "fmha_stagger") == 0) return &svdpp::tuning().fmha_stagger;
  if (strcmp(key, "fmha_handover") == 0) return &svdpp::tuning().fmha_handover;
  if (strcmp(key, "fmha_poly") == 0) return &svdpp::tuning().fmha_poly;
  if (strcmp(key, "fmha_handover_split") == 0) return &svdpp::tuning().fmha_handover_split;
  if (strcmp(key, "ff_fused") == 0) return &svdpp::tuning().ff_fused;
  if (strcmp(key, "ff_dbg") == 0) return &svdpp::tuning().ff_dbg;
  if (strcmp(key, "ff_pair") == 0) return &svdpp::tuning().ff_pair;
  if (strcmp(key, "reverse") == 0) return &svdpp::tuning().reverse;
  if (strcmp(key, "reverse_gn_apply_same") == 0) return &svdpp::tuning().reverse_gn_apply_same;
  if (strcmp(key, "zigzag") == 0) return &svdpp::tuning().zigzag;
  return nullptr;
}

int svdpp_set_tuning(const char* key, int value) {
  int* s = tuning_slot(key);
  if (s == nullptr) {
    svdpp::set_error("unknown tuning key '%s'", key ? key : "(null)");
    return -1;
  }
  *s = value;
  return 0;
}

int svdpp_get_tuning(const char* key) {
  int* s = tuning_slot(key);
  return s ? *s : -1;
}

int svdpp_device_info(int* sm_major, int* sm_minor, int* n_sms) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    svdpp::set_error("no CUDA device");
    return -1;
  }
  int maj = 0, min = 0, n = 0;
  cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (sm_major) *sm_major = maj;
  if (sm_minor) *sm_minor = min;
  if (n_sms) *n_sms = n;
  return 0;
}

}  // extern "C"
