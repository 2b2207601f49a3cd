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

// Elementwise stage of the attention backward for 16 key columns of one query row:
//   P = 2^(S*scale_log2 - lse2),   dS = P * (dP - delta) * scale = P * (dP*scale + nds),  nds = -delta*scale
// on the packed-fp32 pipe (one FFMA2 for the exponent pair, one FFMA2 + one FMUL2 for the dS pair).  MASKED: columns
// outside [lo, hi) (padding, other samples of a packed tile, causal future) and rows that do not exist give P = dS = 0;
// chunks that lie inside the valid range take the predicate-free variant.
// Warp-uniform classification of one 16-key chunk (call with all 32 lanes): 0 = no lane needs the mask, 2 = no lane
// has a valid key in it (block-diagonal packed tiles, rows beyond the sequence), 1 = mixed.  Decided per lane, a warp
// whose rows straddle the end of the sequence (or two packed samples) executed BOTH variants of every chunk.
__device__ __forceinline__ int attn_chunk_kind(bool row_ok, int c, int lo, int hi) {
  const bool full = row_ok && c >= lo && c + 16 <= hi;
  const bool none = !row_ok || c >= hi || c + 16 <= lo;
  return __all_sync(0xffffffffu, full) ? 0 : (__all_sync(0xffffffffu, none) ? 2 : 1);
}

template <bool MASKED>
__device__ __forceinline__ void attn_bwd_chunk16(const uint32_t (&sv)[16], const uint32_t (&dv)[16], float scale_log2,
                                                 float lse2, float scale, float nds, int c, int lo, int hi, bool row_ok,
                                                 float (&pr)[16], float (&ds)[16]) {
  const float2 sl2 = splat2(scale_log2), nl2 = splat2(-lse2), sc2 = splat2(scale), nd2 = splat2(nds);
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const float2 t = fma2(make_float2(__uint_as_float(sv[j]), __uint_as_float(sv[j + 1])), sl2, nl2);
    float p0 = ex2_approx(t.x), p1 = ex2_approx(t.y);
    if (MASKED) {
      p0 = (row_ok && c + j >= lo && c + j < hi) ? p0 : 0.f;
      p1 = (row_ok && c + j + 1 >= lo && c + j + 1 < hi) ? p1 : 0.f;
    }
    const float2 d = fma2(make_float2(__uint_as_float(dv[j]), __uint_as_float(dv[j + 1])), sc2, nd2);
    const float2 g = mul2(make_float2(p0, p1), d);
    pr[j] = p0; pr[j + 1] = p1;
    ds[j] = g.x; ds[j + 1] = g.y;
  }
}

}  // namespace clipa
