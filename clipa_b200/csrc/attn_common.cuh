// Helpers shared by the tcgen05 attention kernels (attention_tc.cu: one-tile problems, head_dim 64;
// attention_flash.cu: any sequence length, head_dim 64 / 80).
#pragma once
#include "ptx.cuh"

namespace clipa {

constexpr int kTcTileBytes = 128 * 128;                     // 128 rows x 128 B
constexpr int kTcPBytes = 2 * kTcTileBytes;                 // P: 128 rows x 128 keys bf16 = 2 swizzle atoms

__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ uint4 pack8_bf16(const float* f) {
  uint4 t;
  t.x = pack_bf16x2(f[0], f[1]);
  t.y = pack_bf16x2(f[2], f[3]);
  t.z = pack_bf16x2(f[4], f[5]);
  t.w = pack_bf16x2(f[6], f[7]);
  return t;
}
// byte offset of 16-byte chunk `chunk` (along the key axis) of row `row` in a [128 x 128-key] bf16
// tile stored as two 128B-swizzled atoms of 64 keys
__device__ __forceinline__ uint32_t p_tile_off(int row, int chunk) {
  return (chunk >> 3) * kTcTileBytes + row * 128 + (((chunk & 7) ^ (row & 7)) << 4);
}

}  // namespace clipa
