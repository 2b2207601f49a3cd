// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] * B[N,K]^T)
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1      : MMA issuer     (one elected lane issues tcgen05.mma, accumulators in TMEM)
//   warps 2..9  : epilogue       (tcgen05.ld TMEM -> registers -> fused epilogue -> global)
//
// Tile 128 x BN x 64 (BN = 256 or 128), kStages-deep smem ring, two TMEM accumulator stages so the
// epilogue of tile i overlaps the MMAs of tile i+1.  Operands may be K-major or MN-major (the
// transposed view is consumed in place through the UMMA descriptor "major" bits, which is what
// makes dgrad / wgrad run without materialising any transpose).
#pragma once
#include "ptx.cuh"

namespace clipa {

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kNumEpiWarps = 8;
constexpr int kGemmThreads = 64 + 32 * kNumEpiWarps;

enum : int {
  EPI_STORE = 0,
  EPI_BIAS_ACT = 1,
  EPI_DACT = 2,
  EPI_ATOMIC_F32 = 3,
  EPI_LSE = 4,          // contrastive head: online row log-sum-exp, nothing stored per element
  EPI_SOFTMAX_GRAD = 5  // contrastive head: Pt = exp(s*acc - lse) - onehot
};

struct GemmParams {
  int M, N, K;
  int m_blocks, n_blocks, k_blocks;
  int n_per_chunk, n_chunks;  // chunk = consecutive n-blocks handled by one work item
  int split_k, kb_per_split;
  int num_items;
  // epilogue operands
  void* C;
  long long ldc;
  int c_f32;
  float alpha;
  const void* bias;
  int bias_f32;
  const __nv_bfloat16* residual;
  long long ldr;
  __nv_bfloat16* aux;
  long long ldaux;
  int act;
  int aux_deriv;  // BIAS_ACT: aux receives act'(f) instead of f;  DACT: aux already holds act'(f)
  int tma_store;  // bf16 outputs leave through per-warp smem staging + TMA store (tmap_c / tmap_aux)
  // contrastive head
  float scale_log2;  // logit_scale * log2(e)
  int label_offset;
  const float* lse;    // [M] natural-log lse (SOFTMAX_GRAD input)
  float* part_max;     // [n_chunks, M] (LSE)
  float* part_sum;     // [n_chunks, M]
  float* diag;         // [M] label logit, natural units
  float* dscale;       // scalar accumulator (SOFTMAX_GRAD)
  const float* scale_dev;  // if set: logit_scale is read from device memory (no host round trip), scale_log2 ignored
  int scale_out;           // SOFTMAX_GRAD: store scale * Pt (the gradient GEMMs then need no scale factor)
};

template <int BN>
struct GemmSmem {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = kStages * kStageBytes;
  static constexpr int kBiasOffset = kBarOffset + 256;   // 2 x 256 fp32 bias slices (per accumulator stage)
  static constexpr int kStoreOffset = ((kBiasOffset + 2048 + 1023) / 1024) * 1024;  // 8 warps x 2 KB staging
  static constexpr int kTotal = kStoreOffset + kNumEpiWarps * 2048 + 1024;  // + alignment slack
};

// ------------------------------------------------------------------------------------------------
// activation math (fp32)
// ------------------------------------------------------------------------------------------------
// erf-GELU is evaluated through the lower tail of the standard normal,
//     q(a) = Phi(-a) = 2^P(-a),   a = min(|x|, 6),
// with P a degree-6 minimax polynomial of log2 Phi(-a) on [0, 6] (|dP| <= 7.4e-5 in fp32 Horner
// form, i.e. a RELATIVE error of 5e-5 in q everywhere, 40x below the bf16 half-ulp of the result;
// beyond 6 the tail is < 1e-9 and the clamp costs nothing):
//     gelu(x)  = x * Phi(x)          = max(x, 0) - a * q
//     gelu'(x) = Phi(x) + x * phi(x) = 0.5 + copysign(0.5 - q, x) + x * 0.39894 * 2^(-x^2 log2e / 2)
// One MUFU.EX2 (two for the derivative) and six FFMA2 per PAIR of elements -- the pair arithmetic
// uses the packed-fp32 pipe (fma.rn.f32x2), so the 8 epilogue warps spend ~8 issue slots per
// element instead of the ~19 of the Abramowitz-Stegun rcp/exp form used before (which itself
// replaced libdevice erff: 417 TFLOP/s, epilogue-bound).
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_tanh(float x) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// q = Phi(-a) for a pair; `na` = -min(|x|, 6)
__device__ __forceinline__ float2 normal_tail2(float2 na) {
  float2 acc = fma2(splat2(2.343321648e-05f), na, splat2(6.197391776e-04f));
  acc = fma2(acc, na, splat2(7.260354701e-03f));
  acc = fma2(acc, na, splat2(5.141863227e-02f));
  acc = fma2(acc, na, splat2(-4.608635008e-01f));
  acc = fma2(acc, na, splat2(1.150480390e+00f));
  acc = fma2(acc, na, splat2(-1.0f));
  return make_float2(fast_ex2(acc.x), fast_ex2(acc.y));
}
__device__ __forceinline__ float2 neg_abs_clamped2(float2 x) {
  return make_float2(fmaxf(-fabsf(x.x), -6.0f), fmaxf(-fabsf(x.y), -6.0f));
}
__device__ __forceinline__ float2 gelu_erf_fwd2(float2 x) {
  const float2 na = neg_abs_clamped2(x);
  const float2 q = normal_tail2(na);
  return fma2(na, q, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
}
__device__ __forceinline__ float2 gelu_erf_bwd2(float2 x) {
  const float2 na = neg_abs_clamped2(x);
  const float2 q = normal_tail2(na);
  const float2 t = mul2(mul2(x, splat2(-0.72134752044448170f)), x);          // -x^2 * log2(e) / 2
  const float2 e = make_float2(fast_ex2(t.x), fast_ex2(t.y));
  const float2 h = fma2(q, splat2(-1.0f), splat2(0.5f));                       // 0.5 - q  (>= 0)
  const float2 hs = make_float2(__uint_as_float(__float_as_uint(h.x) | (__float_as_uint(x.x) & 0x80000000u)),
                                __uint_as_float(__float_as_uint(h.y) | (__float_as_uint(x.y) & 0x80000000u)));
  const float2 d = fma2(mul2(x, e), splat2(0.3989422804014327f), hs);
  return add2(d, splat2(0.5f));
}
// value and derivative of a pair, sharing the tail evaluation (the recompute pass of the MLP stores the derivative
// so that the dgrad x GELU' GEMM only multiplies: its epilogue was the slowest of the block, 79 % tensor-active)
__device__ __forceinline__ void gelu_erf_fwd_bwd2(float2 x, float2& g, float2& d) {
  const float2 na = neg_abs_clamped2(x);
  const float2 q = normal_tail2(na);
  g = fma2(na, q, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));   // bit-identical to gelu_erf_fwd2
  // phi(x) = 2^(-x^2 log2(e)/2 + log2(1/sqrt(2 pi)))
  const float2 t = fma2(mul2(x, splat2(-0.72134752044448170f)), x, splat2(-1.3257480647361593f));
  const float2 e = make_float2(fast_ex2(t.x), fast_ex2(t.y));
  const float2 h = fma2(q, splat2(-1.0f), splat2(0.5f));
  const float2 hs = make_float2(__uint_as_float(__float_as_uint(h.x) | (__float_as_uint(x.x) & 0x80000000u)),
                                __uint_as_float(__float_as_uint(h.y) | (__float_as_uint(x.y) & 0x80000000u)));
  d = fma2(x, e, add2(hs, splat2(0.5f)));                            // Phi(x) + x phi(x)
}
__device__ __forceinline__ float act_fwd(float x, int act) {
  if (act == 0) return gelu_erf_fwd2(make_float2(x, x)).x;
  if (act == 1) {
    const float u = 0.7978845608028654f * fmaf(0.044715f * x * x, x, x);
    return 0.5f * x * (1.0f + fast_tanh(u));
  }
  return x * fast_rcp(1.0f + fast_ex2(-1.702f * 1.4426950408889634f * x));
}
__device__ __forceinline__ float act_bwd(float x, int act) {
  if (act == 0) return gelu_erf_bwd2(make_float2(x, x)).x;
  if (act == 1) {
    const float x2 = x * x;
    const float u = 0.7978845608028654f * fmaf(0.044715f * x2, x, x);
    const float t = fast_tanh(u);
    const float du = 0.7978845608028654f * fmaf(3.0f * 0.044715f, x2, 1.0f);
    return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * du;
  }
  const float sg = fast_rcp(1.0f + fast_ex2(-1.702f * 1.4426950408889634f * x));
  return sg * (1.0f + 1.702f * x * (1.0f - sg));
}

__device__ __forceinline__ void load_bias32(const void* bias, int bias_f32, int col0, float (&b)[32]) {
  if (bias_f32) {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(bias) + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 t = __ldg(p + i);
      b[4 * i] = t.x; b[4 * i + 1] = t.y; b[4 * i + 2] = t.z; b[4 * i + 3] = t.w;
    }
  } else {
    const uint4* p = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(bias) + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 t = __ldg(p + i);
      b[8 * i] = bf16lo(t.x); b[8 * i + 1] = bf16hi(t.x);
      b[8 * i + 2] = bf16lo(t.y); b[8 * i + 3] = bf16hi(t.y);
      b[8 * i + 4] = bf16lo(t.z); b[8 * i + 5] = bf16hi(t.z);
      b[8 * i + 6] = bf16lo(t.w); b[8 * i + 7] = bf16hi(t.w);
    }
  }
}
__device__ __forceinline__ float load_bias1(const void* bias, int bias_f32, int col) {
  return bias_f32 ? __ldg(static_cast<const float*>(bias) + col)
                  : __bfloat162float(static_cast<const __nv_bfloat16*>(bias)[col]);
}
__device__ __forceinline__ void load_bf16x32(const __nv_bfloat16* p, float (&r)[32]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 t = __ldg(q + i);
    r[8 * i] = bf16lo(t.x); r[8 * i + 1] = bf16hi(t.x);
    r[8 * i + 2] = bf16lo(t.y); r[8 * i + 3] = bf16hi(t.y);
    r[8 * i + 4] = bf16lo(t.z); r[8 * i + 5] = bf16hi(t.z);
    r[8 * i + 6] = bf16lo(t.w); r[8 * i + 7] = bf16hi(t.w);
  }
}
__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* p, const float (&f)[32]) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 t;
    t.x = pack_bf16x2(f[8 * i], f[8 * i + 1]);
    t.y = pack_bf16x2(f[8 * i + 2], f[8 * i + 3]);
    t.z = pack_bf16x2(f[8 * i + 4], f[8 * i + 5]);
    t.w = pack_bf16x2(f[8 * i + 6], f[8 * i + 7]);
    q[i] = t;
  }
}
__device__ __forceinline__ void store_f32x32(float* p, const float (&f)[32]) {
  float4* q = reinterpret_cast<float4*>(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
}
__device__ __forceinline__ void red_add_f32x4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// The four store-type epilogues on one 32-column chunk of one output row (shared by the 1-CTA and
// 2-CTA kernels).  `sb` = this chunk's 32 bias values in shared memory (staged once per tile, so no
// global-load latency sits on the per-chunk critical path); `side` = the chunk's residual (STORE) or
// saved pre-activation (DACT) row segment, prefetched one chunk ahead by the caller.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack_bf16x32(const uint4 (&s)[4], float (&r)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r[8 * i] = bf16lo(s[i].x); r[8 * i + 1] = bf16hi(s[i].x);
    r[8 * i + 2] = bf16lo(s[i].y); r[8 * i + 3] = bf16hi(s[i].y);
    r[8 * i + 4] = bf16lo(s[i].z); r[8 * i + 5] = bf16hi(s[i].z);
    r[8 * i + 6] = bf16lo(s[i].w); r[8 * i + 7] = bf16hi(s[i].w);
  }
}
template <int EPI>
__device__ __forceinline__ const __nv_bfloat16* epi_side_ptr(const GemmParams& p, long long& ld) {
  if constexpr (EPI == EPI_STORE) { ld = p.ldr; return p.residual; }
  if constexpr (EPI == EPI_DACT) { ld = p.ldaux; return p.aux; }
  ld = 0;
  return nullptr;
}
// One 32-row x 32-column bf16 chunk: registers -> 64B-swizzled per-warp staging buffer -> TMA store.
// The TMA engine writes whole 64-byte row segments (and clips rows >= M / columns >= N), instead of
// 32 lanes issuing 16-byte stores to 32 different cache lines through the LSU.
// PENDING = 1 when the warp alternates between two staging buffers (the recompute pass writes two outputs per
// chunk): only the store issued two groups ago used this buffer, the most recent one may still be in flight.
template <int PENDING = 0>
__device__ __forceinline__ void chunk_store_tma(const CUtensorMap* tm, uint8_t* sbuf, const float (&f)[32],
                                                int col0, int row0) {
  const int lane = threadIdx.x & 31;
  if (lane == 0) tma_store_wait_read<PENDING>();  // the previous store from THIS buffer has finished reading it
  __syncwarp();
  uint8_t* rowp = sbuf + lane * 64;
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint4 t;
    t.x = pack_bf16x2(f[8 * k], f[8 * k + 1]);
    t.y = pack_bf16x2(f[8 * k + 2], f[8 * k + 3]);
    t.z = pack_bf16x2(f[8 * k + 4], f[8 * k + 5]);
    t.w = pack_bf16x2(f[8 * k + 6], f[8 * k + 7]);
    *reinterpret_cast<uint4*>(rowp + ((k ^ sw) << 4)) = t;
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(tm, sbuf, col0, row0);
    tma_store_commit();
  }
}

// same staging + TMA store for a chunk that is already packed to bf16 (t[k] = columns 8k .. 8k+7 of this lane's row)
template <int PENDING = 0>
__device__ __forceinline__ void chunk_store_tma_packed(const CUtensorMap* tm, uint8_t* sbuf, const uint4 (&t)[4], int col0,
                                                       int row0) {
  const int lane = threadIdx.x & 31;
  if (lane == 0) tma_store_wait_read<PENDING>();
  __syncwarp();
  uint8_t* rowp = sbuf + lane * 64;
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(rowp + ((k ^ sw) << 4)) = t[k];
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(tm, sbuf, col0, row0);
    tma_store_commit();
  }
}

template <int EPI>
__device__ __forceinline__ void epi_apply_store(const GemmParams& p, float (&f)[32], long long row, bool row_ok,
                                                int col0, const float* sb, const uint4 (&side)[4],
                                                const CUtensorMap* tm_c, const CUtensorMap* tm_aux, uint8_t* sbuf,
                                                int row0_warp, int aux_buf_off = 0) {
  const bool full = (col0 + 32 <= p.N);
  if constexpr (EPI == EPI_STORE) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] *= p.alpha;
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(sb + j);
        f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
      }
    }
    if (p.residual && row_ok) {
      if (full) {
        float rr[32];
        unpack_bf16x32(side, rr);
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] += rr[j];
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.N) f[j] += __bfloat162float(p.residual[row * p.ldr + col0 + j]);
      }
    }
    if (p.tma_store) {
      chunk_store_tma(tm_c, sbuf, f, col0, row0_warp);
    } else if (row_ok) {
      if (full) {
        if (p.c_f32) store_f32x32(static_cast<float*>(p.C) + row * p.ldc + col0, f);
        else store_bf16x32(static_cast<__nv_bfloat16*>(p.C) + row * p.ldc + col0, f);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = col0 + j;
          if (col < p.N) {
            if (p.c_f32) static_cast<float*>(p.C)[row * p.ldc + col] = f[j];
            else static_cast<__nv_bfloat16*>(p.C)[row * p.ldc + col] = __float2bfloat16(f[j]);
          }
        }
      }
    }
  } else if constexpr (EPI == EPI_BIAS_ACT) {
    // requires N % 32 == 0 (checked on the host); outputs always go through the TMA store
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(sb + j);
        const float2 lo = add2(make_float2(f[j], f[j + 1]), make_float2(b.x, b.y));
        const float2 hi = add2(make_float2(f[j + 2], f[j + 3]), make_float2(b.z, b.w));
        f[j] = lo.x; f[j + 1] = lo.y; f[j + 2] = hi.x; f[j + 3] = hi.y;
      }
    }
    if (p.aux && p.aux_deriv) {
      // recompute pass: aux <- act'(f) (bf16), C <- act(f); eight elements at a time so that only the packed
      // derivative (4 registers per group) lives next to the accumulator chunk
      uint4 dpk[4];
      if (p.act == 0) {   // branch outside the loops: one straight-line block of 16 independent pairs to interleave
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 d[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 g;
            gelu_erf_fwd_bwd2(make_float2(f[8 * k + 2 * j], f[8 * k + 2 * j + 1]), g, d[j]);
            f[8 * k + 2 * j] = g.x; f[8 * k + 2 * j + 1] = g.y;
          }
          dpk[k].x = pack_bf16x2(d[0].x, d[0].y); dpk[k].y = pack_bf16x2(d[1].x, d[1].y);
          dpk[k].z = pack_bf16x2(d[2].x, d[2].y); dpk[k].w = pack_bf16x2(d[3].x, d[3].y);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float dv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            dv[j] = act_bwd(f[8 * k + j], p.act);
            f[8 * k + j] = act_fwd(f[8 * k + j], p.act);
          }
          dpk[k].x = pack_bf16x2(dv[0], dv[1]); dpk[k].y = pack_bf16x2(dv[2], dv[3]);
          dpk[k].z = pack_bf16x2(dv[4], dv[5]); dpk[k].w = pack_bf16x2(dv[6], dv[7]);
        }
      }
      if (aux_buf_off) {  // two staging buffers per warp: neither store waits for the one issued just before it
        chunk_store_tma_packed<1>(tm_aux, sbuf + aux_buf_off, dpk, col0, row0_warp);
        chunk_store_tma<1>(tm_c, sbuf, f, col0, row0_warp);
        return;
      }
      chunk_store_tma_packed(tm_aux, sbuf, dpk, col0, row0_warp);
      chunk_store_tma(tm_c, sbuf, f, col0, row0_warp);
      return;
    }
    if (p.aux) {
      // the pre-activation is stored in bf16 and the activation is evaluated on the ROUNDED
      // value, so backward (which re-reads aux) sees the same operand
      chunk_store_tma(tm_aux, sbuf, f, col0, row0_warp);
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __bfloat162float(__float2bfloat16(f[j]));
    }
    if (p.act == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float2 g = gelu_erf_fwd2(make_float2(f[j], f[j + 1]));
        f[j] = g.x;
        f[j + 1] = g.y;
      }
    } else if (p.act == 1) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = act_fwd(f[j], 1);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = act_fwd(f[j], 2);
    }
    chunk_store_tma(tm_c, sbuf, f, col0, row0_warp);
  } else if constexpr (EPI == EPI_DACT) {
    float a[32];
    unpack_bf16x32(side, a);   // rows >= M carry zeros: their result is clipped by the TMA store
    if (p.aux_deriv) {         // aux already holds act'(f): one packed multiply per pair
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float2 g = mul2(make_float2(f[j], f[j + 1]), make_float2(a[j], a[j + 1]));
        f[j] = g.x;
        f[j + 1] = g.y;
      }
    } else if (p.act == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float2 g = mul2(make_float2(f[j], f[j + 1]), gelu_erf_bwd2(make_float2(a[j], a[j + 1])));
        f[j] = g.x;
        f[j + 1] = g.y;
      }
    } else if (p.act == 1) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] *= act_bwd(a[j], 1);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] *= act_bwd(a[j], 2);
    }
    chunk_store_tma(tm_c, sbuf, f, col0, row0_warp);
  } else if constexpr (EPI == EPI_ATOMIC_F32) {
    if (row_ok) {
      float* dst = static_cast<float*>(p.C) + row * p.ldc + col0;
      if (full) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          red_add_f32x4(dst + j, p.alpha * f[j], p.alpha * f[j + 1], p.alpha * f[j + 2], p.alpha * f[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.N) atomicAdd(dst + j, p.alpha * f[j]);
      }
    }
  }
}

// One accumulator stage's worth of store-type epilogue for one warp: `ncols` columns starting at
// tile-local column `col_local0`; TMEM loads and the side-operand loads run one chunk ahead.
// `release()` hands the accumulator stage back to the MMA issuer; it is called as soon as the LAST
// chunk sits in registers, so the next-but-one tile's MMAs start while this warp is still doing the
// math and the stores of its final chunk.
template <int EPI, class Release>
__device__ __forceinline__ void epi_run_store(const GemmParams& p, uint32_t t_warp, long long row, bool row_ok,
                                              int tile_col0, int col_local0, int ncols, const float* sbias_tile,
                                              const CUtensorMap* tm_c, const CUtensorMap* tm_aux, uint8_t* sbuf,
                                              Release release, int aux_buf_off = 0) {
  const int row0_warp = static_cast<int>(row) - static_cast<int>(threadIdx.x & 31);
  long long side_ld;
  const __nv_bfloat16* side_base = epi_side_ptr<EPI>(p, side_ld);
  const bool use_side = side_base != nullptr && row_ok;
  uint4 side_next[4] = {};
  uint32_t vnext[32];
  tmem_ld_32x32(t_warp, vnext);
  {
    const int col0 = tile_col0 + col_local0;
    if (use_side && col0 + 32 <= p.N) {
      const uint4* q = reinterpret_cast<const uint4*>(side_base + row * side_ld + col0);
#pragma unroll
      for (int i = 0; i < 4; ++i) side_next[i] = __ldg(q + i);
    }
  }
#pragma unroll 1
  for (int c = 0; c < ncols / 32; ++c) {
    const int col0 = tile_col0 + col_local0 + c * 32;
    float f[32];
    uint4 side[4];
    tmem_ld_wait_regs(vnext);
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(vnext[j]);
#pragma unroll
    for (int i = 0; i < 4; ++i) side[i] = side_next[i];
    if (c + 1 < ncols / 32) {
      tmem_ld_32x32(t_warp + (c + 1) * 32, vnext);
      const int coln = col0 + 32;
      if (use_side && coln + 32 <= p.N) {
        const uint4* q = reinterpret_cast<const uint4*>(side_base + row * side_ld + coln);
#pragma unroll
        for (int i = 0; i < 4; ++i) side_next[i] = __ldg(q + i);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) side_next[i] = make_uint4(0, 0, 0, 0);
      }
    } else {
      release();
    }
    if (col0 >= p.N) continue;  // warp-uniform
    epi_apply_store<EPI>(p, f, row, row_ok, col0, sbias_tile + col_local0 + c * 32, side, tm_c, tm_aux, sbuf,
                         row0_warp, aux_buf_off);
  }
}

// Stage the tile's bias slice (fp32) in shared memory: one value per epilogue thread.
__device__ __forceinline__ void epi_stage_bias(const GemmParams& p, float* sbias_tile, int tile_col0, int bn) {
  const int t = threadIdx.x - 64;  // epilogue threads are 64..319
  if (p.bias && t < bn) {
    const int col = tile_col0 + t;
    sbias_tile[t] = col < p.N ? load_bias1(p.bias, p.bias_f32, col) : 0.f;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(32 * kNumEpiWarps) : "memory");
}

// Per-thread (= per output row) running state of the contrastive-head LSE epilogue.
struct LseState {
  float m;  // running max (log2 domain)
  float l;  // running sum of 2^(t - m)
  float d;  // label logit (log2 domain), NaN-free sentinel below
  int found;
};

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_aux,
               const GemmParams p) {
  using S = GemmSmem<BN>;
  constexpr int kStages = S::kStages;
  constexpr uint32_t kTmemCols = 2 * BN;  // two accumulator stages
  constexpr uint32_t kIdesc = make_idesc_bf16(kBM, BN, A_MN, B_MN);

  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned tiles
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kNumEpiWarps);  // one elected arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int items_mn = p.m_blocks * p.n_chunks;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < p.num_items; w += gridDim.x) {
        const int split = w / items_mn;
        const int r = w - split * items_mn;
        const int m_blk = r / p.n_chunks;
        const int chunk = r - m_blk * p.n_chunks;
        const int n_begin = chunk * p.n_per_chunk;
        const int n_end = min(p.n_blocks, n_begin + p.n_per_chunk);
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(p.k_blocks, kb_begin + p.kb_per_split);
        for (int n_blk = n_begin; n_blk < n_end; ++n_blk) {
          for (int kb = kb_begin; kb < kb_end; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * S::kStageBytes;
            uint8_t* sb = sa + S::kABytes;
            mbar_expect_tx(&full_bar[stage], S::kStageBytes);
            if constexpr (!A_MN) {
              tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kBK, m_blk * kBM);
            } else {
#pragma unroll
              for (int a = 0; a < kBM / 64; ++a)
                tma_load_2d(sa + a * (kBK * 128), &tmap_a, &full_bar[stage], m_blk * kBM + a * 64,
                            kb * kBK);
            }
            if constexpr (!B_MN) {
              tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * kBK, n_blk * BN);
            } else {
#pragma unroll
              for (int a = 0; a < BN / 64; ++a)
                tma_load_2d(sb + a * (kBK * 128), &tmap_b, &full_bar[stage], n_blk * BN + a * 64,
                            kb * kBK);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    const IssueMode im = issue_mode(lane);
    if (im.in_loop) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < p.num_items; w += gridDim.x) {
        const int split = w / items_mn;
        const int r = w - split * items_mn;
        const int m_blk = r / p.n_chunks;
        const int chunk = r - m_blk * p.n_chunks;
        const int n_begin = chunk * p.n_per_chunk;
        const int n_end = min(p.n_blocks, n_begin + p.n_per_chunk);
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(p.k_blocks, kb_begin + p.kb_per_split);
        for (int n_blk = n_begin; n_blk < n_end; ++n_blk) {
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_addr = tmem_base + acc * BN;
          for (int kb = kb_begin; kb < kb_end; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
            const uint32_t sb = sa + S::kABytes;
            // K-major: rows of 128 B, 8-row groups 1024 B apart (SBO); k-step = +32 B.
            // MN-major: 64-element (128 B) MN atoms kBK*128 B apart (LBO), 8-k groups 1024 B apart
            //           (SBO); k-step (16 k-rows) = +2048 B.
            const uint64_t da = A_MN ? make_smem_desc_sw128(sa, kBK * 128, 1024)
                                     : make_smem_desc_sw128(sa, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(sb, kBK * 128, 1024)
                                     : make_smem_desc_sw128(sb, 16, 1024);
            constexpr uint32_t a_step = (A_MN ? 2048 : 32) >> 4;
            constexpr uint32_t b_step = (B_MN ? 2048 : 32) >> 4;
            if (im.issue) {
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k) {
                umma_bf16(d_addr, da + static_cast<uint64_t>(k * a_step),
                          db + static_cast<uint64_t>(k * b_step), kIdesc,
                          (kb > kb_begin || k > 0) ? 1u : 0u);
              }
              umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
            }
            im.sync();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          if (im.issue) umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
          im.sync();
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else {
    // ======================= epilogue warps =======================
    const int ew = warp - 2;
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int half = ew >> 2;        // which half of the BN columns
    constexpr int kColsPerWarp = BN / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < p.num_items; w += gridDim.x) {
      const int split = w / items_mn;
      const int r = w - split * items_mn;
      const int m_blk = r / p.n_chunks;
      const int chunk = r - m_blk * p.n_chunks;
      const int n_begin = chunk * p.n_per_chunk;
      const int n_end = min(p.n_blocks, n_begin + p.n_per_chunk);
      const int row = m_blk * kBM + q * 32 + lane;
      const bool row_ok = row < p.M;

      LseState st;
      st.m = -INFINITY; st.l = 0.f; st.d = 0.f; st.found = 0;
      float lse2_row = 0.f;
      float ds_acc = 0.f;
      float scale_log2 = p.scale_log2, out_mul = 1.f;
      if constexpr (EPI >= EPI_LSE) {
        if (p.scale_dev) scale_log2 = __ldg(p.scale_dev) * 1.4426950408889634f;
        if (p.scale_out) out_mul = scale_log2 * 0.69314718055994531f;
      }
      if constexpr (EPI == EPI_SOFTMAX_GRAD) lse2_row = row_ok ? p.lse[row] * 1.4426950408889634f : 0.f;
      const int label = row + p.label_offset;

      for (int n_blk = n_begin; n_blk < n_end; ++n_blk) {
        float* sbias_tile = reinterpret_cast<float*>(smem + S::kBiasOffset) + acc * 256;
        if constexpr (EPI <= EPI_ATOMIC_F32) epi_stage_bias(p, sbias_tile, n_blk * BN, BN);
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_warp = tmem_base + acc * BN + half * kColsPerWarp + (static_cast<uint32_t>(q * 32) << 16);
        if constexpr (EPI <= EPI_ATOMIC_F32) {
          epi_run_store<EPI>(p, t_warp, row, row_ok, n_blk * BN, half * kColsPerWarp, kColsPerWarp, sbias_tile,
                             &tmap_c, &tmap_aux, smem + S::kStoreOffset + ew * 2048, [&]() {
                               tc_fence_before();
                               __syncwarp();
                               if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                             });
        } else {
        uint32_t vnext[32];
        tmem_ld_32x32(t_warp, vnext);
#pragma unroll 1
        for (int c = 0; c < kColsPerWarp / 32; ++c) {
          const int col_local = half * kColsPerWarp + c * 32;
          const int col0 = n_blk * BN + col_local;
          float f[32];
          tmem_ld_wait_regs(vnext);
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(vnext[j]);
          if (c + 1 < kColsPerWarp / 32) tmem_ld_32x32(t_warp + (c + 1) * 32, vnext);  // prefetch next chunk
          if (col0 >= p.N) continue;  // warp-uniform
          const bool full = (col0 + 32 <= p.N);

          if constexpr (EPI == EPI_LSE) {
            // t = logit * log2(e); online (max, sum 2^(t-max)) per row
            float gmax = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              f[j] = (full || col0 + j < p.N) ? f[j] * scale_log2 : -INFINITY;
              gmax = fmaxf(gmax, f[j]);
            }
            const float m_new = fmaxf(st.m, gmax);
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) s += exp2f(f[j] - m_new);
            st.l = st.l * exp2f(st.m - m_new) + s;
            st.m = m_new;
            if (label >= col0 && label < col0 + 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j == label) st.d = f[j];
              st.found = 1;
            }
          } else if constexpr (EPI == EPI_SOFTMAX_GRAD) {
            if (row_ok) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float pt = exp2f(f[j] * scale_log2 - lse2_row) - ((col0 + j == label) ? 1.f : 0.f);
                const bool ok = full || (col0 + j < p.N);
                ds_acc += ok ? pt * f[j] : 0.f;
                f[j] = pt * out_mul;
              }
              __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.C) + row * p.ldc + col0;
              if (full) store_bf16x32(dst, f);
              else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) dst[j] = __float2bfloat16(f[j]);
              }
            }
          }
        }
          // all of this warp's TMEM reads for the stage are complete (wait::ld above)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }  // contrastive-head epilogues (the store-type epilogues release the stage themselves)
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }

      if constexpr (EPI == EPI_LSE) {
        if (row_ok) {
          // two warps (column halves) cover a row: slot = chunk*2 + half
          const long long slot = (static_cast<long long>(chunk) * 2 + half) * p.M + row;
          p.part_max[slot] = st.m;
          p.part_sum[slot] = st.l;
          if (st.found) p.diag[row] = st.d * 0.69314718055994531f;  // back to natural units
        }
      }
      if constexpr (EPI == EPI_SOFTMAX_GRAD) {
        float v = ds_acc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v != 0.f) atomicAdd(p.dscale, v);
      }
    }
  }

  // teardown
  if (warp >= 2 && lane == 0) tma_store_wait<0>();  // this warp's bulk stores have fully completed
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace clipa
