// Host-side helpers shared by the .cu files: error reporting and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <utility>

#include "../../include/svdpp.h"

namespace svdpp {

void set_error(const char* fmt, ...);

#define SVDPP_CHECK_ARG(cond, ...)      \
  do {                                  \
    if (!(cond)) {                      \
      ::svdpp::set_error(__VA_ARGS__);  \
      return -1;                        \
    }                                   \
  } while (0)

#define SVDPP_CUDA(expr)                                                              \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ::svdpp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -2;                                                                      \
    }                                                                                 \
  } while (0)

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return -3;
  }
  return 0;
}

int num_sms();

// Process-wide tuning switches (svdpp_set_tuning; defaults from the environment).
struct Tuning {
  int tma_store;  // GEMM output tiles through TMA tensor stores
  int pdl;        // programmatic dependent launch on every kernel
  int tma_r1;     // GEMM residual tile through TMA tensor loads (with tma_store)
  int epi_dma;    // GEMM: DMA-lane epilogue with two staging tiles for short main loops
  int epi_dma_max_kb;  // ... for K / 64 <= this
  int two_prod;        // GEMM: a second TMA producer warp (warp 3) takes every other k-block when K / 64 >= two_prod_min_kb
  int two_prod_min_kb;
  int splitk;          // GEMM impl 6: split-K tail when the descriptor carries a workspace
  int splitk_min_kb;   // ... fewest k-blocks per slice
  int r1_prefetch_max_kb;   // GEMM: producer warp prefetches the residual tile into L2 for K / 64 <= this (0: never)
  int splitk_min_total_kb;  // ... only for K / 64 >= this (the tail machinery costs ~20 us, a tile ~0.45 us per k-block)
  int fmha_stagger;         // two-tile FMHA: SM clocks by which query tile 1 starts behind tile 0 (0: together)
  int fmha_handover;        // ping-pong FMHA (impl 7): batches of 16 exponentials before the end of a turn at which the partner warp is released
  int fmha_poly;            // ping-pong FMHA (impl 7): exponentials per batch of 16 evaluated by a degree-3 polynomial on the FMA pipe (0, 2..6, 8)
  int fmha_handover_split;  // two-threads-per-row ping-pong FMHA (impl 8): the same, of the 4 batches of a turn
  int ff_fused;             // svdpp_unet_*: feed-forwards of blocks with C <= 320 through the fused kernel (svdpp_ff_geglu_f16)
  int ff_dbg;               // fused feed-forward: timing experiments (results are wrong when non-zero)
  int ff_pair;              // fused feed-forward: 0 single CTAs, 1 clusters of two sharing the weight stream through TMA multicast, 2 cta_group::2 pairs (default)
  int reverse;              // GEMM / LayerNorm / GroupNorm launches walk their rows from the end (L2-friendly after a forward producer)
  int reverse_gn_apply_same;  // GroupNorm apply pass in the same direction as the statistics pass (A/B only)
  int zigzag;               // svdpp_unet_*: alternate the traversal direction from producer to consumer
};
Tuning& tuning();

// Launch through cudaLaunchKernelEx so that the cluster shape and programmatic stream serialisation can be
// attached.  With PDL every kernel executes griddepcontrol.wait (pdl_wait() in ptx.cuh) before its first access to
// global memory, which orders it after the COMPLETION of the previous kernel: only prologues overlap.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = static_cast<unsigned>(cluster_x);
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (tuning().pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// fp16 tensor map, 128-byte swizzle (swizzle_bytes: 128 or 64), zero fill out of bounds. dims/box innermost first;
// strides_bytes has rank-1 entries (dimension 0 is contiguous).
// elem_strides (optional): traversal stride per dimension; a box of extent box[i] then loads
// ceil(box[i] / elem_strides[i]) elements (used for the stride-2 down-sampling convolutions).
int encode_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides = nullptr,
                    int swizzle_bytes = 128);

}  // namespace svdpp
