// Host-side plumbing shared by the C-ABI translation units: error strings, launch counting,
// TMA tensor-map encoding through the driver entry point (no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "../../include/clipa_b200.h"

namespace clipa {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();

// Encodes a 2-D bf16 tensor map: dim0 = contiguous (inner) extent, dim1 = outer extent,
// row pitch in bytes, box = {box0, box1}, 128-, 64- or 32-byte swizzle, OOB reads fill zeros / OOB
// writes are clipped.
// Returns CLIPA_OK or an error (message set).
int encode_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1,
                        uint64_t pitch_bytes, uint32_t box0, uint32_t box1, int swizzle_bytes = 128);

#define CLIPA_CHECK_CUDA(expr)                                                            \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::clipa::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                       \
      return CLIPA_ERR_CUDA;                                                              \
    }                                                                                     \
  } while (0)

#define CLIPA_REQUIRE(cond, code, ...)  \
  do {                                  \
    if (!(cond)) {                      \
      ::clipa::set_error(__VA_ARGS__);  \
      return (code);                    \
    }                                   \
  } while (0)

}  // namespace clipa
