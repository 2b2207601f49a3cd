// Heads and tails of the two towers: the small HBM-bound passes around the transformer
// blocks that the reference runs as a dozen torch elementwise / indexing kernels.
//   clipa_preprocess_u8        uint8 image -> (x/255 - mean)/std -> bf16        training/train.py:191-197
//   clipa_patchify             [N,3,H,W] -> [N*grid, Kp] patch rows (conv1 with stride == kernel is a GEMM over
//                              these rows, open_clip/transformer.py:371,491-493; Kp = K rounded up to 8, zero filled)
//   clipa_assemble_tokens      [cls; patch tokens] + positional table -> [N, L, W]   transformer.py:495-499
//   clipa_embed_tokens         token_embedding[ids] + positional table -> [N, L, W]  model.py:245-247
//   clipa_pool_tokens          CLS / first / last / EOT-argmax row or token mean      transformer.py:509-529,
//                                                                                     model.py:251-262
//   clipa_l2_normalize         F.normalize(dim=-1), model.py:240,263; the output pointer may be the rank's slice
//                              of the all-gather buffer
// and their backward passes.  All are one pass over their operands: algorithmic bytes = inputs read once +
// outputs written once.
#include "host_common.h"
#include "ptx.cuh"

namespace clipa {

constexpr int kIoThreads = 256;
static inline unsigned io_blocks(long long work, int per_block = kIoThreads) {
  long long b = (work + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * 32;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

__device__ __forceinline__ float bf16_to_f(__nv_bfloat16 v) { return __bfloat162float(v); }

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kIoThreads)
preprocess_u8_kernel(const uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ out, long long total, int hw,
                     float3 mean, float3 inv_std) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  if ((hw & 3) == 0) {     // 4 pixels of one plane per thread
    const long long n4 = total >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const int c = (int)(((i << 2) / hw) % 3);
      const float m = c == 0 ? mean.x : (c == 1 ? mean.y : mean.z);
      const float s = c == 0 ? inv_std.x : (c == 1 ? inv_std.y : inv_std.z);
      const uchar4 v = reinterpret_cast<const uchar4*>(img)[i];
      uint2 o;
      o.x = pack_bf16x2((v.x * (1.f / 255.f) - m) * s, (v.y * (1.f / 255.f) - m) * s);
      o.y = pack_bf16x2((v.z * (1.f / 255.f) - m) * s, (v.w * (1.f / 255.f) - m) * s);
      reinterpret_cast<uint2*>(out)[i] = o;
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
      const int c = (int)((i / hw) % 3);
      const float m = c == 0 ? mean.x : (c == 1 ? mean.y : mean.z);
      const float s = c == 0 ? inv_std.x : (c == 1 ? inv_std.y : inv_std.z);
      out[i] = __float2bfloat16((img[i] * (1.f / 255.f) - m) * s);
    }
  }
}

// one thread per pair of consecutive patch-row elements (same image row: pw is even or the pair is split)
template <typename T>
__global__ void __launch_bounds__(kIoThreads)
patchify_kernel(const T* __restrict__ img, __nv_bfloat16* __restrict__ patches, long long rows, int H, int W, int ph,
                int pw, int gw, int gh, int K, int Kp) {
  const long long pairs = rows * (Kp >> 1);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += stride) {
    const long long row = i / (Kp >> 1);
    const int k0 = (int)(i - row * (Kp >> 1)) * 2;
    const int gx = (int)(row % gw);
    const long long t = row / gw;
    const int gy = (int)(t % gh);
    const long long n = t / gh;
    float v[2] = {0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = k0 + e;
      if (k < K) {
        const int c = k / (ph * pw), r = k - c * ph * pw, py = r / pw, px = r - py * pw;
        v[e] = (float)img[((n * 3 + c) * H + gy * ph + py) * (long long)W + gx * pw + px];
      }
    }
    reinterpret_cast<uint32_t*>(patches)[i] = pack_bf16x2(v[0], v[1]);
  }
}

// x[n, 0, :] = cls + pos[0];  x[n, 1+g, :] = tok[n*G+g, :] + pos[1+g, :]      (8 channels per thread)
__global__ void __launch_bounds__(kIoThreads)
assemble_tokens_kernel(const __nv_bfloat16* __restrict__ tok, const float* __restrict__ cls,
                       const float* __restrict__ pos, __nv_bfloat16* __restrict__ x, long long N, int L, int W) {
  const int wv = W >> 3;
  const long long total = N * L * wv;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int w8 = (int)(i % wv);
    const long long r = i / wv;
    const int l = (int)(r % L);
    const long long n = r / L;
    float f[8];
    if (l == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = cls[w8 * 8 + j];
    } else {
      const uint4 t = reinterpret_cast<const uint4*>(tok + (n * (L - 1) + l - 1) * W)[w8];
      f[0] = bf16lo(t.x); f[1] = bf16hi(t.x); f[2] = bf16lo(t.y); f[3] = bf16hi(t.y);
      f[4] = bf16lo(t.z); f[5] = bf16hi(t.z); f[6] = bf16lo(t.w); f[7] = bf16hi(t.w);
    }
    const float4 p0 = reinterpret_cast<const float4*>(pos + (long long)l * W)[w8 * 2];
    const float4 p1 = reinterpret_cast<const float4*>(pos + (long long)l * W)[w8 * 2 + 1];
    uint4 o;
    o.x = pack_bf16x2(f[0] + p0.x, f[1] + p0.y);
    o.y = pack_bf16x2(f[2] + p0.z, f[3] + p0.w);
    o.z = pack_bf16x2(f[4] + p1.x, f[5] + p1.y);
    o.w = pack_bf16x2(f[6] + p1.z, f[7] + p1.w);
    reinterpret_cast<uint4*>(x + r * W)[w8] = o;
  }
}

// dtok[n*G+g, :] = dx[n, 1+g, :]   (the cls / positional gradients are column sums of dx: clipa_colsum_accum)
__global__ void __launch_bounds__(kIoThreads)
assemble_tokens_bwd_kernel(const __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dtok, long long N, int L,
                           int W) {
  const int wv = W >> 3;
  const long long total = N * (L - 1) * wv;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int w8 = (int)(i % wv);
    const long long r = i / wv;
    const int g = (int)(r % (L - 1));
    const long long n = r / (L - 1);
    reinterpret_cast<uint4*>(dtok + r * W)[w8] = reinterpret_cast<const uint4*>(dx + (n * L + 1 + g) * W)[w8];
  }
}

// x[n, l, :] = table[ids[n, l], :] + pos[l, :]
__global__ void __launch_bounds__(kIoThreads)
embed_tokens_kernel(const long long* __restrict__ ids, const float* __restrict__ table, const float* __restrict__ pos,
                    __nv_bfloat16* __restrict__ x, long long rows, int L, int W, int V) {
  const int wv = W >> 3;
  const long long total = rows * wv;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int w8 = (int)(i % wv);
    const long long r = i / wv;
    const int l = (int)(r % L);
    long long id = ids[r];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    const float4* t = reinterpret_cast<const float4*>(table + id * W) + w8 * 2;
    const float4* p = reinterpret_cast<const float4*>(pos + (long long)l * W) + w8 * 2;
    const float4 t0 = t[0], t1 = t[1], p0 = p[0], p1 = p[1];
    uint4 o;
    o.x = pack_bf16x2(t0.x + p0.x, t0.y + p0.y);
    o.y = pack_bf16x2(t0.z + p0.z, t0.w + p0.w);
    o.z = pack_bf16x2(t1.x + p1.x, t1.y + p1.y);
    o.w = pack_bf16x2(t1.z + p1.z, t1.w + p1.w);
    reinterpret_cast<uint4*>(x + r * W)[w8] = o;
  }
}

__device__ __forceinline__ void red_add4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// dtable[ids[n, l], :] += dx[n, l, :]   (fp32 reductions at L2; rows of frequent tokens serialise there)
__global__ void __launch_bounds__(kIoThreads)
embed_tokens_bwd_kernel(const long long* __restrict__ ids, const __nv_bfloat16* __restrict__ dx,
                        float* __restrict__ dtable, long long rows, int W, int V) {
  const int wv = W >> 3;
  const long long total = rows * wv;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int w8 = (int)(i % wv);
    const long long r = i / wv;
    long long id = ids[r];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    const uint4 t = reinterpret_cast<const uint4*>(dx + r * W)[w8];
    float* d = dtable + id * W + w8 * 8;
    red_add4(d, bf16lo(t.x), bf16hi(t.x), bf16lo(t.y), bf16hi(t.y));
    red_add4(d + 4, bf16lo(t.z), bf16hi(t.z), bf16lo(t.w), bf16hi(t.w));
  }
}

enum PoolMode { kPoolFirst = 0, kPoolLast = 1, kPoolArgmaxId = 2, kPoolMeanAll = 3, kPoolMeanSkipFirst = 4 };

__device__ __forceinline__ int pool_index(int mode, const long long* ids, long long n, int L, int Lid) {
  if (mode == kPoolFirst) return 0;
  if (mode == kPoolLast) return L - 1;
  // first position of the largest token id (torch.argmax semantics; EOT is the largest BPE id)
  const long long* row = ids + n * Lid;
  long long best = row[0];
  int at = 0;
  for (int l = 1; l < Lid; ++l)
    if (row[l] > best) { best = row[l]; at = l; }
  return at;
}

// out[n, :] = x[n, idx, :]  or  mean over the tokens (fp32 sum, one rounding)
__global__ void __launch_bounds__(kIoThreads)
pool_tokens_kernel(const __nv_bfloat16* __restrict__ x, const long long* __restrict__ ids, __nv_bfloat16* __restrict__ out,
                   long long N, int L, int W, int mode) {
  const int wv = W >> 3;
  const long long total = N * wv;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int w8 = (int)(i % wv);
    const long long n = i / wv;
    if (mode <= kPoolArgmaxId) {
      const int l = pool_index(mode, ids, n, L, L);
      reinterpret_cast<uint4*>(out + n * W)[w8] = reinterpret_cast<const uint4*>(x + (n * L + l) * W)[w8];
    } else {
      const int l0 = mode == kPoolMeanSkipFirst ? 1 : 0;
      float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int l = l0; l < L; ++l) {
        const uint4 t = reinterpret_cast<const uint4*>(x + (n * L + l) * W)[w8];
        a[0] += bf16lo(t.x); a[1] += bf16hi(t.x); a[2] += bf16lo(t.y); a[3] += bf16hi(t.y);
        a[4] += bf16lo(t.z); a[5] += bf16hi(t.z); a[6] += bf16lo(t.w); a[7] += bf16hi(t.w);
      }
      const float inv = 1.f / (float)(L - l0);
      uint4 o;
      o.x = pack_bf16x2(a[0] * inv, a[1] * inv);
      o.y = pack_bf16x2(a[2] * inv, a[3] * inv);
      o.z = pack_bf16x2(a[4] * inv, a[5] * inv);
      o.w = pack_bf16x2(a[6] * inv, a[7] * inv);
      reinterpret_cast<uint4*>(out + n * W)[w8] = o;
    }
  }
}

// dx[n, l, :] = dout[n, :] at the pooled position (0 elsewhere), or dout[n, :] / count on every pooled token
__global__ void __launch_bounds__(kIoThreads)
pool_tokens_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const long long* __restrict__ ids,
                       __nv_bfloat16* __restrict__ dx, long long N, int L, int W, int mode) {
  const int wv = W >> 3;
  const long long total = N * L * wv;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int w8 = (int)(i % wv);
    const long long r = i / wv;
    const int l = (int)(r % L);
    const long long n = r / L;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (mode <= kPoolArgmaxId) {
      if (l == pool_index(mode, ids, n, L, L)) o = reinterpret_cast<const uint4*>(dout + n * W)[w8];
    } else {
      const int l0 = mode == kPoolMeanSkipFirst ? 1 : 0;
      if (l >= l0) {
        const uint4 t = reinterpret_cast<const uint4*>(dout + n * W)[w8];
        const float inv = 1.f / (float)(L - l0);
        o.x = pack_bf16x2(bf16lo(t.x) * inv, bf16hi(t.x) * inv);
        o.y = pack_bf16x2(bf16lo(t.y) * inv, bf16hi(t.y) * inv);
        o.z = pack_bf16x2(bf16lo(t.z) * inv, bf16hi(t.z) * inv);
        o.w = pack_bf16x2(bf16lo(t.w) * inv, bf16hi(t.w) * inv);
      }
    }
    reinterpret_cast<uint4*>(dx + r * W)[w8] = o;
  }
}

// one warp per row: y = x / max(||x||, eps); inv_norm saved for backward
__global__ void __launch_bounds__(kIoThreads)
l2_normalize_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, float* __restrict__ inv_norm,
                    long long rows, int E, long long ldy) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long wstride = ((long long)gridDim.x * blockDim.x) >> 5;
  const int ev = E >> 3;
  for (long long r = warp0; r < rows; r += wstride) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + r * E);
    float ss = 0.f;
    for (int v = lane; v < ev; v += 32) {
      const uint4 t = xr[v];
      ss += bf16lo(t.x) * bf16lo(t.x) + bf16hi(t.x) * bf16hi(t.x) + bf16lo(t.y) * bf16lo(t.y) + bf16hi(t.y) * bf16hi(t.y) +
            bf16lo(t.z) * bf16lo(t.z) + bf16hi(t.z) * bf16hi(t.z) + bf16lo(t.w) * bf16lo(t.w) + bf16hi(t.w) * bf16hi(t.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    if (lane == 0 && inv_norm) inv_norm[r] = inv;
    uint4* yr = reinterpret_cast<uint4*>(y + r * ldy);
    for (int v = lane; v < ev; v += 32) {
      const uint4 t = xr[v];
      uint4 o;
      o.x = pack_bf16x2(bf16lo(t.x) * inv, bf16hi(t.x) * inv);
      o.y = pack_bf16x2(bf16lo(t.y) * inv, bf16hi(t.y) * inv);
      o.z = pack_bf16x2(bf16lo(t.z) * inv, bf16hi(t.z) * inv);
      o.w = pack_bf16x2(bf16lo(t.w) * inv, bf16hi(t.w) * inv);
      yr[v] = o;
    }
  }
}

// dx = inv_norm * (dy - y * (y . dy)),  y = x * inv_norm recomputed from x (dy fp32 or bf16)
template <typename TG>
__global__ void __launch_bounds__(kIoThreads)
l2_normalize_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ inv_norm,
                        const TG* __restrict__ dy, __nv_bfloat16* __restrict__ dx, long long rows, int E) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long wstride = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < rows; r += wstride) {
    const float inv = inv_norm[r];
    float dot = 0.f;
    for (int e = lane; e < E; e += 32) dot += bf16_to_f(x[r * E + e]) * inv * (float)dy[r * E + e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    for (int e = lane; e < E; e += 32) {
      const float yv = bf16_to_f(x[r * E + e]) * inv;
      dx[r * E + e] = __float2bfloat16(inv * ((float)dy[r * E + e] - yv * dot));
    }
  }
}

}  // namespace clipa

using namespace clipa;

#define IO_LAUNCHED()                    \
  CLIPA_CHECK_CUDA(cudaGetLastError()); \
  count_launch();                        \
  return CLIPA_OK

extern "C" int clipa_preprocess_u8(const void* images_u8, void* out_bf16, int64_t n, int32_t height, int32_t width,
                                   const float* h_mean3, const float* h_std3, void* stream) {
  CLIPA_REQUIRE(images_u8 && out_bf16 && h_mean3 && h_std3, CLIPA_ERR_BAD_ARG, "preprocess_u8: null pointer");
  CLIPA_REQUIRE(n > 0 && height > 0 && width > 0, CLIPA_ERR_BAD_ARG, "preprocess_u8: bad dims");
  const long long total = (long long)n * 3 * height * width;
  const float3 mean = make_float3(h_mean3[0], h_mean3[1], h_mean3[2]);
  const float3 inv = make_float3(1.f / h_std3[0], 1.f / h_std3[1], 1.f / h_std3[2]);
  preprocess_u8_kernel<<<io_blocks(total / 4 + 1), kIoThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(images_u8), static_cast<__nv_bfloat16*>(out_bf16), total, height * width, mean, inv);
  IO_LAUNCHED();
}

extern "C" int clipa_patchify(const void* images, int32_t image_dtype, void* patches, int64_t n, int32_t height,
                              int32_t width, int32_t patch_h, int32_t patch_w, int32_t k_padded, void* stream) {
  CLIPA_REQUIRE(images && patches, CLIPA_ERR_BAD_ARG, "patchify: null pointer");
  CLIPA_REQUIRE(n > 0 && patch_h > 0 && patch_w > 0 && height >= patch_h && width >= patch_w, CLIPA_ERR_BAD_ARG,
                "patchify: bad dims");
  const int gh = height / patch_h, gw = width / patch_w, K = 3 * patch_h * patch_w;
  CLIPA_REQUIRE(k_padded >= K && k_padded % 8 == 0, CLIPA_ERR_BAD_ARG, "patchify: k_padded %d must be a multiple of 8 >= %d",
                k_padded, K);
  const long long rows = (long long)n * gh * gw;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned blocks = io_blocks(rows * (k_padded / 2));
  if (image_dtype == CLIPA_BF16)
    patchify_kernel<<<blocks, kIoThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(images),
                                                   static_cast<__nv_bfloat16*>(patches), rows, height, width, patch_h,
                                                   patch_w, gw, gh, K, k_padded);
  else
    patchify_kernel<<<blocks, kIoThreads, 0, s>>>(static_cast<const float*>(images), static_cast<__nv_bfloat16*>(patches),
                                                   rows, height, width, patch_h, patch_w, gw, gh, K, k_padded);
  IO_LAUNCHED();
}

extern "C" int clipa_assemble_tokens(const void* tok, const float* cls, const float* pos, void* x, int64_t n, int32_t L,
                                     int32_t W, void* stream) {
  CLIPA_REQUIRE(tok && cls && pos && x, CLIPA_ERR_BAD_ARG, "assemble_tokens: null pointer");
  CLIPA_REQUIRE(n > 0 && L > 1 && W > 0 && W % 8 == 0, CLIPA_ERR_UNSUPPORTED, "assemble_tokens: need W %% 8 == 0, L > 1");
  assemble_tokens_kernel<<<io_blocks((long long)n * L * (W / 8)), kIoThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(tok), cls, pos, static_cast<__nv_bfloat16*>(x), n, L, W);
  IO_LAUNCHED();
}

extern "C" int clipa_assemble_tokens_bwd(const void* dx, void* dtok, int64_t n, int32_t L, int32_t W, void* stream) {
  CLIPA_REQUIRE(dx && dtok, CLIPA_ERR_BAD_ARG, "assemble_tokens_bwd: null pointer");
  CLIPA_REQUIRE(n > 0 && L > 1 && W > 0 && W % 8 == 0, CLIPA_ERR_UNSUPPORTED, "assemble_tokens_bwd: need W %% 8 == 0, L > 1");
  assemble_tokens_bwd_kernel<<<io_blocks((long long)n * (L - 1) * (W / 8)), kIoThreads, 0,
                               static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(dx),
                                                                    static_cast<__nv_bfloat16*>(dtok), n, L, W);
  IO_LAUNCHED();
}

extern "C" int clipa_embed_tokens(const int64_t* ids, const float* table, const float* pos, void* x, int64_t n, int32_t L,
                                  int32_t W, int32_t vocab, void* stream) {
  CLIPA_REQUIRE(ids && table && pos && x, CLIPA_ERR_BAD_ARG, "embed_tokens: null pointer");
  CLIPA_REQUIRE(n > 0 && L > 0 && W > 0 && W % 8 == 0 && vocab > 0, CLIPA_ERR_UNSUPPORTED, "embed_tokens: need W %% 8 == 0");
  embed_tokens_kernel<<<io_blocks((long long)n * L * (W / 8)), kIoThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(ids), table, pos, static_cast<__nv_bfloat16*>(x), (long long)n * L, L, W, vocab);
  IO_LAUNCHED();
}

extern "C" int clipa_embed_tokens_bwd(const int64_t* ids, const void* dx, float* dtable, int64_t n, int32_t L, int32_t W,
                                      int32_t vocab, void* stream) {
  CLIPA_REQUIRE(ids && dx && dtable, CLIPA_ERR_BAD_ARG, "embed_tokens_bwd: null pointer");
  CLIPA_REQUIRE(n > 0 && L > 0 && W > 0 && W % 8 == 0 && vocab > 0, CLIPA_ERR_UNSUPPORTED, "embed_tokens_bwd: need W %% 8 == 0");
  embed_tokens_bwd_kernel<<<io_blocks((long long)n * L * (W / 8)), kIoThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(ids), static_cast<const __nv_bfloat16*>(dx), dtable, (long long)n * L, W, vocab);
  IO_LAUNCHED();
}

extern "C" int clipa_pool_tokens(const void* x, const int64_t* ids, void* out, int64_t n, int32_t L, int32_t W,
                                 int32_t mode, void* stream) {
  CLIPA_REQUIRE(x && out, CLIPA_ERR_BAD_ARG, "pool_tokens: null pointer");
  CLIPA_REQUIRE(mode >= 0 && mode <= 4 && (mode != kPoolArgmaxId || ids), CLIPA_ERR_BAD_ARG, "pool_tokens: bad mode %d", mode);
  CLIPA_REQUIRE(n > 0 && L > 0 && W % 8 == 0 && (mode != kPoolMeanSkipFirst || L > 1), CLIPA_ERR_UNSUPPORTED,
                "pool_tokens: need W %% 8 == 0");
  pool_tokens_kernel<<<io_blocks((long long)n * (W / 8)), kIoThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), reinterpret_cast<const long long*>(ids), static_cast<__nv_bfloat16*>(out), n, L,
      W, mode);
  IO_LAUNCHED();
}

extern "C" int clipa_pool_tokens_bwd(const void* dout, const int64_t* ids, void* dx, int64_t n, int32_t L, int32_t W,
                                     int32_t mode, void* stream) {
  CLIPA_REQUIRE(dout && dx, CLIPA_ERR_BAD_ARG, "pool_tokens_bwd: null pointer");
  CLIPA_REQUIRE(mode >= 0 && mode <= 4 && (mode != kPoolArgmaxId || ids), CLIPA_ERR_BAD_ARG, "pool_tokens_bwd: bad mode %d",
                mode);
  CLIPA_REQUIRE(n > 0 && L > 0 && W % 8 == 0 && (mode != kPoolMeanSkipFirst || L > 1), CLIPA_ERR_UNSUPPORTED,
                "pool_tokens_bwd: need W %% 8 == 0");
  pool_tokens_bwd_kernel<<<io_blocks((long long)n * L * (W / 8)), kIoThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dout), reinterpret_cast<const long long*>(ids), static_cast<__nv_bfloat16*>(dx), n,
      L, W, mode);
  IO_LAUNCHED();
}

extern "C" int clipa_l2_normalize(const void* x, void* y, int64_t ldy, float* inv_norm, int64_t rows, int32_t E,
                                  void* stream) {
  CLIPA_REQUIRE(x && y, CLIPA_ERR_BAD_ARG, "l2_normalize: null pointer");
  CLIPA_REQUIRE(rows > 0 && E > 0 && E % 8 == 0 && ldy % 8 == 0 && ldy >= E, CLIPA_ERR_UNSUPPORTED,
                "l2_normalize: need E %% 8 == 0 and ldy %% 8 == 0");
  l2_normalize_kernel<<<io_blocks(rows * 32), kIoThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), inv_norm, rows, E, ldy);
  IO_LAUNCHED();
}

extern "C" int clipa_l2_normalize_bwd(const void* x, const float* inv_norm, const void* dy, int32_t dy_dtype, void* dx,
                                      int64_t rows, int32_t E, void* stream) {
  CLIPA_REQUIRE(x && inv_norm && dy && dx, CLIPA_ERR_BAD_ARG, "l2_normalize_bwd: null pointer");
  CLIPA_REQUIRE(rows > 0 && E > 0, CLIPA_ERR_BAD_ARG, "l2_normalize_bwd: bad dims");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dy_dtype == CLIPA_F32)
    l2_normalize_bwd_kernel<<<io_blocks(rows * 32), kIoThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(x), inv_norm,
                                                                        static_cast<const float*>(dy),
                                                                        static_cast<__nv_bfloat16*>(dx), rows, E);
  else
    l2_normalize_bwd_kernel<<<io_blocks(rows * 32), kIoThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(x), inv_norm,
                                                                        static_cast<const __nv_bfloat16*>(dy),
                                                                        static_cast<__nv_bfloat16*>(dx), rows, E);
  IO_LAUNCHED();
}
