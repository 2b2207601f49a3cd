#include "host_common.h"

#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>

namespace clipa {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1,
                        uint64_t pitch_bytes, uint32_t box0, uint32_t box1, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  CLIPA_REQUIRE(fn != nullptr, CLIPA_ERR_NO_DEVICE,
                "cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver?)");
  CLIPA_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, CLIPA_ERR_BAD_ARG,
                "TMA base pointer %p is not 16-byte aligned", base);
  CLIPA_REQUIRE((pitch_bytes & 15) == 0, CLIPA_ERR_BAD_ARG,
                "TMA row pitch %llu bytes is not a multiple of 16 (ld must be a multiple of 8)",
                (unsigned long long)pitch_bytes);
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 32   ? CU_TENSOR_MAP_SWIZZLE_32B
                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                        : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CLIPA_REQUIRE(r == CUDA_SUCCESS, CLIPA_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (CUresult %d; dims %llu x %llu pitch %llu box %u x %u)",
                (int)r, (unsigned long long)dim0, (unsigned long long)dim1,
                (unsigned long long)pitch_bytes, box0, box1);
  return CLIPA_OK;
}

}  // namespace clipa

extern "C" {
int clipa_abi_version(void) { return CLIPA_B200_ABI_VERSION; }
const char* clipa_last_error(void) { return clipa::g_err; }
int64_t clipa_launch_count(void) { return clipa::g_launches.load(std::memory_order_relaxed); }
}
