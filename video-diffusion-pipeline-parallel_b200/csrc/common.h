// Host-side helpers shared by the .cu files: error reporting and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

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

// fp16 tensor map, 128-byte swizzle, zero fill out of bounds. dims/box innermost first;
// strides_bytes has rank-1 entries (dimension 0 is contiguous).
// elem_strides (optional): traversal stride per dimension; a box of extent box[i] then loads
// ceil(box[i] / elem_strides[i]) elements (used for the stride-2 down-sampling convolutions).
int encode_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides = nullptr);

}  // namespace svdpp
