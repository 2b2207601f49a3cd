// fp32 parity path (BASELINE north_star: "within 1e-5 (fp32)").  tcgen05 has no fp32 MMA (TF32 is ~1e-3), so
// precision='fp32' runs the SAME step on CUDA-core kernels with fp32 storage and fp32 FMA arithmetic:
//   clipa_gemm_f32            every F.linear / `@` of the path (strided operands: transposes are free)
//   clipa_act_f32             nn.GELU / QuickGELU and their derivative
//   clipa_layernorm_f32_*     F.layer_norm
//   clipa_attention_f32_*     the SDPA inside nn.MultiheadAttention
//   clipa_colsum_f32          bias gradients
//   clipa_row_lse_f32 / clipa_softmax_grad_f32   the cross-entropy of ClipLoss over materialised logits
// Throughput is irrelevant here (one warp per row, scalar dot products); it exists so that the fp32 reference
// results are reproduced to round-off on the device, with the same C-ABI conventions as the bf16 path.
#include <cmath>

#include "host_common.h"

namespace clipa {

// ---------------------------------------------------------------------------------------------------------------
// GEMM: C[m,n] = epilogue(alpha * sum_k A[m*rsa + k*csa] * B[n*rsb + k*csb])
// ---------------------------------------------------------------------------------------------------------------
struct GemmF32Params {
  int M, N, K;
  const float* A; long long rsa, csa;
  const float* B; long long rsb, csb;
  float* C; long long ldc;
  float alpha;
  const float* bias;          // [N] or null
  const float* residual; long long ldr;   // [M,N] or null (added after bias)
  int accumulate;             // C += value instead of C = value
};

constexpr int kF32Tile = 64, kF32K = 16;

__global__ void __launch_bounds__(256)
gemm_f32_kernel(const GemmF32Params p) {
  __shared__ float As[kF32K][kF32Tile + 1];
  __shared__ float Bs[kF32K][kF32Tile + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * kF32Tile, n0 = blockIdx.x * kF32Tile;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < p.K; k0 += kF32K) {
    for (int e = threadIdx.x; e < kF32Tile * kF32K; e += 256) {
      // consecutive threads walk the contiguous direction of each operand
      int r, k;
      if (p.csa == 1) { k = e % kF32K; r = e / kF32K; } else { r = e % kF32Tile; k = e / kF32Tile; }
      As[k][r] = (m0 + r < p.M && k0 + k < p.K) ? p.A[(long long)(m0 + r) * p.rsa + (long long)(k0 + k) * p.csa] : 0.f;
      if (p.csb == 1) { k = e % kF32K; r = e / kF32K; } else { r = e % kF32Tile; k = e / kF32Tile; }
      Bs[k][r] = (n0 + r < p.N && k0 + k < p.K) ? p.B[(long long)(n0 + r) * p.rsb + (long long)(k0 + k) * p.csb] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kF32K; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = p.alpha * acc[i][j];
      if (p.bias) v += p.bias[n];
      if (p.residual) v += p.residual[(long long)m * p.ldr + n];
      float* c = p.C + (long long)m * p.ldc + n;
      *c = p.accumulate ? *c + v : v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// activations (open_clip/transformer.py:37-40, model.py:128-129)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_f32(float x, int act) {
  if (act == 0) return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
  if (act == 1) {
    const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
    return 0.5f * x * (1.f + tanhf(u));
  }
  return x / (1.f + expf(-1.702f * x));
}
__device__ __forceinline__ float dact_f32(float x, int act) {
  if (act == 0) return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
  if (act == 1) {
    const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
    const float t = tanhf(u);
    return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * 0.7978845608028654f * (1.f + 3.f * 0.044715f * x * x);
  }
  const float s = 1.f / (1.f + expf(-1.702f * x));
  return s + 1.702f * x * s * (1.f - s);
}
// mode 0: out = act(x);  mode 1: out = dy * act'(x)
__global__ void act_f32_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ out,
                               long long n, int act, int mode) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = mode == 0 ? act_f32(x[i], act) : dy[i] * dact_f32(x[i], act);
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm, one warp per row (two-pass statistics)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum_f32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void __launch_bounds__(256)
ln_f32_fwd_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b, float* __restrict__ y,
                  float* __restrict__ mean_out, float* __restrict__ rstd_out, long long rows, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float* xr = x + row * D;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) s += xr[d];
  const float mean = warp_sum_f32(s) / D;
  float v = 0.f;
  for (int d = lane; d < D; d += 32) { const float t = xr[d] - mean; v = fmaf(t, t, v); }
  const float rstd = rsqrtf(warp_sum_f32(v) / D + eps);
  for (int d = lane; d < D; d += 32) y[row * D + d] = (xr[d] - mean) * rstd * g[d] + b[d];
  if (lane == 0 && mean_out) { mean_out[row] = mean; rstd_out[row] = rstd; }
}

__global__ void __launch_bounds__(256)
ln_f32_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ g,
                  const float* __restrict__ mean_in, const float* __restrict__ rstd_in, float* __restrict__ dx,
                  float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows, int D) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float mean = mean_in[row], rstd = rstd_in[row];
  const float* xr = x + row * D;
  const float* dr = dy + row * D;
  float s1 = 0.f, s2 = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float xh = (xr[d] - mean) * rstd, gg = dr[d] * g[d];
    s1 += gg;
    s2 = fmaf(gg, xh, s2);
  }
  s1 = warp_sum_f32(s1) / D;
  s2 = warp_sum_f32(s2) / D;
  for (int d = lane; d < D; d += 32) {
    const float xh = (xr[d] - mean) * rstd, gg = dr[d] * g[d];
    dx[row * D + d] = rstd * (gg - s1 - xh * s2);
    atomicAdd(dgamma + d, dr[d] * xh);
    atomicAdd(dbeta + d, dr[d]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// attention core, one warp per query row (forward, dQ) / per key row (dK, dV); scores of the row in shared memory
// ---------------------------------------------------------------------------------------------------------------
constexpr int kAttF32Warps = 4;

__global__ void __launch_bounds__(kAttF32Warps * 32)
attn_f32_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out, float* __restrict__ lse, int batch, int L,
                    int H, int hd, int causal, float scale) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long r = (long long)blockIdx.x * kAttF32Warps + warp;      // (n, h, i)
  if (r >= (long long)batch * H * L) return;
  const int i = (int)(r % L);
  const int h = (int)((r / L) % H);
  const long long n = r / ((long long)L * H);
  const int D = H * hd;
  const long long pitch = 3LL * D;
  float* qs = sm + (size_t)warp * (hd + L);
  float* sc = qs + hd;
  const float* qrow = qkv + (n * L + i) * pitch + h * hd;
  for (int d = lane; d < hd; d += 32) qs[d] = qrow[d];
  __syncwarp();
  const int jend = causal ? i + 1 : L;
  float mx = -INFINITY;
  for (int j = lane; j < jend; j += 32) {
    const float* krow = qkv + (n * L + j) * pitch + D + h * hd;
    float s = 0.f;
    for (int d = 0; d < hd; ++d) s = fmaf(qs[d], krow[d], s);
    s *= scale;
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max_f32(mx);
  float sum = 0.f;
  for (int j = lane; j < jend; j += 32) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum_f32(sum);
  __syncwarp();
  const float inv = 1.f / sum;
  for (int d = lane; d < hd; d += 32) {
    float o = 0.f;
    for (int j = 0; j < jend; ++j) o = fmaf(sc[j], qkv[(n * L + j) * pitch + 2 * D + h * hd + d], o);
    out[(n * L + i) * (long long)D + h * hd + d] = o * inv;
  }
  if (lane == 0) lse[(n * H + h) * L + i] = mx + logf(sum);
}

// dQ row i: ds_j = p_j (dO_i . V_j - delta_i) * scale;  dQ_i = sum_j ds_j K_j
__global__ void __launch_bounds__(kAttF32Warps * 32)
attn_f32_bwd_q_kernel(const float* __restrict__ qkv, const float* __restrict__ out, const float* __restrict__ dout,
                      const float* __restrict__ lse, float* __restrict__ dqkv, int batch, int L, int H, int hd, int causal,
                      float scale) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long r = (long long)blockIdx.x * kAttF32Warps + warp;
  if (r >= (long long)batch * H * L) return;
  const int i = (int)(r % L);
  const int h = (int)((r / L) % H);
  const long long n = r / ((long long)L * H);
  const int D = H * hd;
  const long long pitch = 3LL * D;
  float* qs = sm + (size_t)warp * (2 * hd + L);
  float* dos = qs + hd;
  float* ds = dos + hd;
  const float* qrow = qkv + (n * L + i) * pitch + h * hd;
  const float* dorow = dout + (n * L + i) * (long long)D + h * hd;
  const float* orow = out + (n * L + i) * (long long)D + h * hd;
  float delta = 0.f;
  for (int d = lane; d < hd; d += 32) {
    qs[d] = qrow[d];
    dos[d] = dorow[d];
    delta = fmaf(dorow[d], orow[d], delta);
  }
  delta = warp_sum_f32(delta);
  __syncwarp();
  const float l = lse[(n * H + h) * L + i];
  const int jend = causal ? i + 1 : L;
  for (int j = lane; j < jend; j += 32) {
    const float* krow = qkv + (n * L + j) * pitch + D + h * hd;
    const float* vrow = krow + D;
    float s = 0.f, dp = 0.f;
    for (int d = 0; d < hd; ++d) {
      s = fmaf(qs[d], krow[d], s);
      dp = fmaf(dos[d], vrow[d], dp);
    }
    ds[j] = expf(s * scale - l) * (dp - delta) * scale;
  }
  __syncwarp();
  for (int d = lane; d < hd; d += 32) {
    float g = 0.f;
    for (int j = 0; j < jend; ++j) g = fmaf(ds[j], qkv[(n * L + j) * pitch + D + h * hd + d], g);
    dqkv[(n * L + i) * pitch + h * hd + d] = g;
  }
}

// dK, dV row j: over the queries i that attend to key j
__global__ void __launch_bounds__(kAttF32Warps * 32)
attn_f32_bwd_kv_kernel(const float* __restrict__ qkv, const float* __restrict__ out, const float* __restrict__ dout,
                       const float* __restrict__ lse, float* __restrict__ dqkv, int batch, int L, int H, int hd, int causal,
                       float scale) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long r = (long long)blockIdx.x * kAttF32Warps + warp;
  if (r >= (long long)batch * H * L) return;
  const int j = (int)(r % L);
  const int h = (int)((r / L) % H);
  const long long n = r / ((long long)L * H);
  const int D = H * hd;
  const long long pitch = 3LL * D;
  float* ks = sm + (size_t)warp * (2 * hd + 2 * L);
  float* vs = ks + hd;
  float* ps = vs + hd;
  float* ds = ps + L;
  const float* krow = qkv + (n * L + j) * pitch + D + h * hd;
  for (int d = lane; d < hd; d += 32) { ks[d] = krow[d]; vs[d] = krow[D + d]; }
  __syncwarp();
  const int ibeg = causal ? j : 0;
  for (int i = ibeg + lane; i < L; i += 32) {
    const float* qrow = qkv + (n * L + i) * pitch + h * hd;
    const float* dorow = dout + (n * L + i) * (long long)D + h * hd;
    const float* orow = out + (n * L + i) * (long long)D + h * hd;
    float s = 0.f, dp = 0.f, delta = 0.f;
    for (int d = 0; d < hd; ++d) {
      s = fmaf(qrow[d], ks[d], s);
      dp = fmaf(dorow[d], vs[d], dp);
      delta = fmaf(dorow[d], orow[d], delta);
    }
    const float pv = expf(s * scale - lse[(n * H + h) * L + i]);
    ps[i] = pv;
    ds[i] = pv * (dp - delta) * scale;
  }
  __syncwarp();
  for (int d = lane; d < hd; d += 32) {
    float gk = 0.f, gv = 0.f;
    for (int i = ibeg; i < L; ++i) {
      gk = fmaf(ds[i], qkv[(n * L + i) * pitch + h * hd + d], gk);
      gv = fmaf(ps[i], dout[(n * L + i) * (long long)D + h * hd + d], gv);
    }
    dqkv[(n * L + j) * pitch + D + h * hd + d] = gk;
    dqkv[(n * L + j) * pitch + 2 * D + h * hd + d] = gv;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// column sums, row log-sum-exp, softmax gradient
// ---------------------------------------------------------------------------------------------------------------
__global__ void colsum_f32_kernel(const float* __restrict__ x, long long ldx, float* __restrict__ out, long long rows, int N,
                                  long long rows_per_block) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float s = 0.f;
  for (long long r = r0; r < r1; ++r) s += x[r * ldx + n];
  atomicAdd(out + n, s);
}

// one warp per row of logits [M, N]: lse[m] = logsumexp_n logits[m, n]; diag[m] = logits[m, m + label_offset]
__global__ void __launch_bounds__(256)
row_lse_f32_kernel(const float* __restrict__ logits, long long ld, int M, int N, int label_offset, float* __restrict__ lse,
                   float* __restrict__ diag) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= M) return;
  const float* lr = logits + row * ld;
  float mx = -INFINITY;
  for (int n = lane; n < N; n += 32) mx = fmaxf(mx, lr[n]);
  mx = warp_max_f32(mx);
  float s = 0.f;
  for (int n = lane; n < N; n += 32) s += expf(lr[n] - mx);
  s = warp_sum_f32(s);
  if (lane == 0) {
    lse[row] = mx + logf(s);
    diag[row] = lr[row + label_offset];
  }
}

// pt[m, n] = exp(logits[m, n] - lse[m]) - [n == m + label_offset];  dscale += sum pt * logits / scale
__global__ void __launch_bounds__(256)
softmax_grad_f32_kernel(const float* __restrict__ logits, long long ld, int M, int N, int label_offset,
                        const float* __restrict__ lse, const float* __restrict__ scale, float* __restrict__ pt,
                        float* __restrict__ dscale) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= M) return;
  const float* lr = logits + row * ld;
  const float l = lse[row], inv_s = 1.f / scale[0];
  float acc = 0.f;
  for (int n = lane; n < N; n += 32) {
    const float v = expf(lr[n] - l) - (n == row + label_offset ? 1.f : 0.f);
    pt[row * (long long)N + n] = v;
    acc = fmaf(v, lr[n] * inv_s, acc);
  }
  acc = warp_sum_f32(acc);
  if (lane == 0) atomicAdd(dscale, acc);
}

// out[n] += sum_k v[k] * W[k*ld + n]  (vector x matrix; thread per column, k split over blockIdx.y)
__global__ void __launch_bounds__(128)
gemv_f32_kernel(const float* __restrict__ v, const float* __restrict__ W, long long ld, float* __restrict__ out, int K, int N) {
  const int n = blockIdx.x * 128 + threadIdx.x;
  const int k0 = blockIdx.y * 64, k1 = k0 + 64 < K ? k0 + 64 : K;
  if (n >= N) return;
  float acc = 0.f;
  for (int k = k0; k < k1; ++k) acc = fmaf(v[k], W[(long long)k * ld + n], acc);
  atomicAdd(out + n, acc);
}

static inline unsigned f32_blocks(long long threads, int per_block = 256) {
  long long b = (threads + per_block - 1) / per_block;
  return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace clipa

using namespace clipa;

#define F32_LAUNCHED()                  \
  CLIPA_CHECK_CUDA(cudaGetLastError()); \
  count_launch();                       \
  return CLIPA_OK

extern "C" int clipa_gemm_f32(int32_t M, int32_t N, int32_t K, const float* A, int64_t rsa, int64_t csa, const float* B,
                              int64_t rsb, int64_t csb, float* C, int64_t ldc, float alpha, const float* bias,
                              const float* residual, int64_t ldr, int32_t accumulate, void* stream) {
  CLIPA_REQUIRE(A && B && C, CLIPA_ERR_BAD_ARG, "gemm_f32: null pointer");
  CLIPA_REQUIRE(M > 0 && N > 0 && K > 0, CLIPA_ERR_BAD_ARG, "gemm_f32: bad dims %d %d %d", M, N, K);
  GemmF32Params p{M, N, K, A, rsa, csa, B, rsb, csb, C, ldc, alpha, bias, residual, ldr, accumulate};
  dim3 grid((N + kF32Tile - 1) / kF32Tile, (M + kF32Tile - 1) / kF32Tile);
  gemm_f32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  F32_LAUNCHED();
}

extern "C" int clipa_act_f32(const float* x, const float* dy, float* out, int64_t n, int32_t act, int32_t derivative,
                             void* stream) {
  CLIPA_REQUIRE(x && out && (!derivative || dy), CLIPA_ERR_BAD_ARG, "act_f32: null pointer");
  CLIPA_REQUIRE(n > 0 && act >= 0 && act <= 2, CLIPA_ERR_BAD_ARG, "act_f32: bad arguments");
  long long blocks = (n + 255) / 256;
  if (blocks > 65535) blocks = 65535;
  act_f32_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, dy, out, n, act, derivative ? 1 : 0);
  F32_LAUNCHED();
}

extern "C" int clipa_layernorm_f32_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean,
                                       float* rstd, int64_t rows, int32_t D, float eps, void* stream) {
  CLIPA_REQUIRE(x && gamma && beta && y && mean && rstd, CLIPA_ERR_BAD_ARG, "layernorm_f32_fwd: null pointer");
  CLIPA_REQUIRE(rows > 0 && D > 0, CLIPA_ERR_BAD_ARG, "layernorm_f32_fwd: bad dims");
  ln_f32_fwd_kernel<<<f32_blocks(rows * 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, gamma, beta, y, mean, rstd, rows,
                                                                                          D, eps);
  F32_LAUNCHED();
}

extern "C" int clipa_layernorm_f32_bwd(const float* dy, const float* x, const float* gamma, const float* mean,
                                       const float* rstd, float* dx, float* dgamma, float* dbeta, int64_t rows, int32_t D,
                                       void* stream) {
  CLIPA_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta, CLIPA_ERR_BAD_ARG, "layernorm_f32_bwd: null pointer");
  CLIPA_REQUIRE(rows > 0 && D > 0, CLIPA_ERR_BAD_ARG, "layernorm_f32_bwd: bad dims");
  ln_f32_bwd_kernel<<<f32_blocks(rows * 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, x, gamma, mean, rstd, dx, dgamma,
                                                                                          dbeta, rows, D);
  F32_LAUNCHED();
}

extern "C" int clipa_attention_f32_fwd(const float* qkv, float* out, float* lse, int32_t batch, int32_t L, int32_t heads,
                                       int32_t head_dim, int32_t causal, void* stream) {
  CLIPA_REQUIRE(qkv && out && lse, CLIPA_ERR_BAD_ARG, "attention_f32_fwd: null pointer");
  CLIPA_REQUIRE(batch > 0 && L > 0 && heads > 0 && head_dim > 0, CLIPA_ERR_BAD_ARG, "attention_f32_fwd: bad dims");
  const size_t smem = (size_t)kAttF32Warps * (head_dim + L) * sizeof(float);
  CLIPA_REQUIRE(smem <= 200 * 1024, CLIPA_ERR_UNSUPPORTED, "attention_f32_fwd: L=%d too long for the parity kernel", L);
  if (smem > 48 * 1024)
    CLIPA_CHECK_CUDA(cudaFuncSetAttribute(attn_f32_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long rows = (long long)batch * heads * L;
  attn_f32_fwd_kernel<<<f32_blocks(rows, kAttF32Warps), kAttF32Warps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      qkv, out, lse, batch, L, heads, head_dim, causal, 1.0f / sqrtf((float)head_dim));
  F32_LAUNCHED();
}

extern "C" int clipa_attention_f32_bwd(const float* qkv, const float* out, const float* dout, const float* lse, float* dqkv,
                                       int32_t batch, int32_t L, int32_t heads, int32_t head_dim, int32_t causal,
                                       void* stream) {
  CLIPA_REQUIRE(qkv && out && dout && lse && dqkv, CLIPA_ERR_BAD_ARG, "attention_f32_bwd: null pointer");
  CLIPA_REQUIRE(batch > 0 && L > 0 && heads > 0 && head_dim > 0, CLIPA_ERR_BAD_ARG, "attention_f32_bwd: bad dims");
  const size_t smem = (size_t)kAttF32Warps * (2 * head_dim + 2 * L) * sizeof(float);
  CLIPA_REQUIRE(smem <= 200 * 1024, CLIPA_ERR_UNSUPPORTED, "attention_f32_bwd: L=%d too long for the parity kernel", L);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (smem > 48 * 1024) {
    CLIPA_CHECK_CUDA(cudaFuncSetAttribute(attn_f32_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CLIPA_CHECK_CUDA(cudaFuncSetAttribute(attn_f32_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const long long rows = (long long)batch * heads * L;
  const float scale = 1.0f / sqrtf((float)head_dim);
  attn_f32_bwd_q_kernel<<<f32_blocks(rows, kAttF32Warps), kAttF32Warps * 32, smem, s>>>(qkv, out, dout, lse, dqkv, batch, L,
                                                                                        heads, head_dim, causal, scale);
  CLIPA_CHECK_CUDA(cudaGetLastError());
  count_launch();
  attn_f32_bwd_kv_kernel<<<f32_blocks(rows, kAttF32Warps), kAttF32Warps * 32, smem, s>>>(qkv, out, dout, lse, dqkv, batch, L,
                                                                                         heads, head_dim, causal, scale);
  F32_LAUNCHED();
}

extern "C" int clipa_gemv_f32_accum(const float* v, const float* W, int64_t ld, float* out, int32_t K, int32_t N,
                                    void* stream) {
  CLIPA_REQUIRE(v && W && out, CLIPA_ERR_BAD_ARG, "gemv_f32_accum: null pointer");
  CLIPA_REQUIRE(K > 0 && N > 0 && ld >= N, CLIPA_ERR_BAD_ARG, "gemv_f32_accum: bad dims");
  dim3 grid((N + 127) / 128, (K + 63) / 64);
  gemv_f32_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(v, W, ld, out, K, N);
  F32_LAUNCHED();
}

extern "C" int clipa_colsum_f32(const float* x, int64_t ldx, float* out, int64_t rows, int32_t N, void* stream) {
  CLIPA_REQUIRE(x && out, CLIPA_ERR_BAD_ARG, "colsum_f32: null pointer");
  CLIPA_REQUIRE(rows > 0 && N > 0, CLIPA_ERR_BAD_ARG, "colsum_f32: bad dims");
  const long long rpb = 256;
  dim3 grid((N + 127) / 128, (unsigned)((rows + rpb - 1) / rpb));
  colsum_f32_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, out, rows, N, rpb);
  F32_LAUNCHED();
}

extern "C" int clipa_row_lse_f32(const float* logits, int64_t ld, int32_t M, int32_t N, int32_t label_offset, float* lse,
                                 float* diag, void* stream) {
  CLIPA_REQUIRE(logits && lse && diag, CLIPA_ERR_BAD_ARG, "row_lse_f32: null pointer");
  CLIPA_REQUIRE(M > 0 && N > 0 && label_offset >= 0 && label_offset + M <= N, CLIPA_ERR_BAD_ARG, "row_lse_f32: bad dims");
  row_lse_f32_kernel<<<f32_blocks((long long)M * 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, ld, M, N,
                                                                                                 label_offset, lse, diag);
  F32_LAUNCHED();
}

extern "C" int clipa_softmax_grad_f32(const float* logits, int64_t ld, int32_t M, int32_t N, int32_t label_offset,
                                      const float* lse, const float* scale_dev, float* pt, float* dscale, void* stream) {
  CLIPA_REQUIRE(logits && lse && scale_dev && pt && dscale, CLIPA_ERR_BAD_ARG, "softmax_grad_f32: null pointer");
  CLIPA_REQUIRE(M > 0 && N > 0, CLIPA_ERR_BAD_ARG, "softmax_grad_f32: bad dims");
  softmax_grad_f32_kernel<<<f32_blocks((long long)M * 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, ld, M, N, label_offset, lse, scale_dev, pt, dscale);
  F32_LAUNCHED();
}
