// LayerNorm forward/backward (bf16 I/O, fp32 statistics and parameters), one warp per row,
// 16-byte vector loads, warp-shuffle reductions.  HBM-bound: forward moves 2 bytes in + 2 bytes out
// per element; backward 3 reads (+1 optional residual-grad read) + 1 write.
//
// Reference semantics: F.layer_norm via LayerNormFp32 (open_clip/transformer.py:19-26): the
// input is up-cast to fp32, normalised with biased variance and eps inside the rsqrt, scaled by
// the fp32 affine parameters and cast back to the input dtype.
#include "host_common.h"
#include "ptx.cuh"

namespace clipa {

constexpr int kLnWarps = 8;  // warps (rows in flight) per block

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& t, float (&f)[8]) {
  f[0] = bf16lo(t.x); f[1] = bf16hi(t.x); f[2] = bf16lo(t.y); f[3] = bf16hi(t.y);
  f[4] = bf16lo(t.z); f[5] = bf16hi(t.z); f[6] = bf16lo(t.w); f[7] = bf16hi(t.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 t;
  t.x = pack_bf16x2(f[0], f[1]); t.y = pack_bf16x2(f[2], f[3]);
  t.z = pack_bf16x2(f[4], f[5]); t.w = pack_bf16x2(f[6], f[7]);
  return t;
}

// VPL = 16-byte vectors per lane (D <= 256*VPL)
template <int VPL>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
              float* __restrict__ mean_out, float* __restrict__ rstd_out, long long rows, int D,
              float eps) {
  const int lane = threadIdx.x & 31;
  const int nvec = D >> 3;
  const long long warp_global = (long long)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  const long long warp_stride = (long long)gridDim.x * kLnWarps;
  const float inv_d = 1.0f / (float)D;
  for (long long row = warp_global; row < rows; row += warp_stride) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * D);
    float v[VPL][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        uint4 t = __ldg(xr + vi);
        unpack8(t, v[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i][j];
      }
    }
    const float mean = warp_sum(s) * inv_d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (lane + 32 * i < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i][j] - mean;
          sq += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) * inv_d + eps);
    uint4* yr = reinterpret_cast<uint4*>(y + row * D);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * vi);
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * vi + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * vi);
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * vi + 1);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
        yr[vi] = pack8(o);
      }
    }
    if (lane == 0 && mean_out) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
  }
}

// Backward.  With xhat = (x-mean)*rstd and g = dy*gamma:
//   dx = rstd * (g - mean_D(g) - xhat * mean_D(g*xhat)) (+ dres)
//   dgamma += sum_rows dy*xhat ;  dbeta += sum_rows dy
// Each warp keeps its dgamma/dbeta partials in registers over its grid-stride rows; the block
// reduces them through shared memory and issues one red.add per column per block.  To keep the
// register footprint low enough for 2 blocks/SM the row is held PACKED (bf16x2) between the
// statistics pass and the output pass, and gamma is read from shared memory.
template <int VPL>
__global__ void __launch_bounds__(kLnWarps * 32, 2)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
              const float* __restrict__ gamma, const float* __restrict__ mean_in,
              const float* __restrict__ rstd_in, const __nv_bfloat16* __restrict__ dres,
              __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
              long long rows, int D) {
  extern __shared__ float red[];  // [D] gamma, then [kLnWarps][2][D] reduction scratch
  float* sgam = red;
  float* scratch = red + D;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nvec = D >> 3;
  const long long warp_global = (long long)blockIdx.x * kLnWarps + warp;
  const long long warp_stride = (long long)gridDim.x * kLnWarps;
  const float inv_d = 1.0f / (float)D;

  for (int c = threadIdx.x; c < D; c += blockDim.x) sgam[c] = gamma[c];
  __syncthreads();

  float acc_g[VPL][8], acc_b[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc_g[i][j] = 0.f; acc_b[i][j] = 0.f; }
  }

  for (long long row = warp_global; row < rows; row += warp_stride) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * D);
    const uint4* dyr = reinterpret_cast<const uint4*>(dy + row * D);
    const float mean = mean_in[row];
    const float rstd = rstd_in[row];
    uint4 xp[VPL], dp[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        xp[i] = __ldg(xr + vi);
        dp[i] = __ldg(dyr + vi);
      }
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        float xv[8], dv[8];
        unpack8(xp[i], xv);
        unpack8(dp[i], dv);
        const float4 g0 = *reinterpret_cast<const float4*>(sgam + vi * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(sgam + vi * 8 + 4);
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mean) * rstd;
          const float g = dv[j] * gm[j];
          s1 += g;
          s2 = fmaf(g, xh, s2);
          acc_g[i][j] = fmaf(dv[j], xh, acc_g[i][j]);
          acc_b[i][j] += dv[j];
        }
      }
    }
    s1 = warp_sum(s1) * inv_d;
    s2 = warp_sum(s2) * inv_d;
    uint4* dxr = reinterpret_cast<uint4*>(dx + row * D);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        float xv[8], dv[8], o[8];
        unpack8(xp[i], xv);
        unpack8(dp[i], dv);
        const float4 g0 = *reinterpret_cast<const float4*>(sgam + vi * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(sgam + vi * 8 + 4);
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mean) * rstd;
          o[j] = rstd * (dv[j] * gm[j] - s1 - xh * s2);
        }
        if (dres) {
          float r[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(dres + row * D) + vi), r);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += r[j];
        }
        dxr[vi] = pack8(o);
      }
    }
  }

  // block reduction of the parameter-gradient partials
  float* my_g = scratch + (size_t)warp * 2 * D;
  float* my_b = my_g + D;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        my_g[vi * 8 + j] = acc_g[i][j];
        my_b[vi * 8 + j] = acc_b[i][j];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int w = 0; w < kLnWarps; ++w) {
      sg += scratch[(size_t)w * 2 * D + c];
      sb += scratch[(size_t)w * 2 * D + D + c];
    }
    atomicAdd(dgamma + c, sg);
    atomicAdd(dbeta + c, sb);
  }
}

// out[n] += sum over rows of x[row, n]  (bias gradients). Block = 8 row-lanes x 32 col-lanes,
// each thread owns 8 consecutive columns (16-byte loads, 512 B per warp per row).
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, float* __restrict__ out,
              long long rows, int N, long long rows_per_block) {
  __shared__ float red[8][256];
  const int cl = threadIdx.x & 31;
  const int rl = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cl) * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < N) {
    for (long long r = r0 + rl; r < r1; r += 8) {
      float v[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + r * ldx + col)), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][cl * 8 + j] = acc[j];
  __syncthreads();
  const int c = threadIdx.x;  // 256 columns per block
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][c];
  const int gc = blockIdx.x * 256 + c;
  if (gc < N) atomicAdd(out + gc, s);
}

template <int VPL>
static int launch_ln_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean,
                         float* rstd, long long rows, int D, float eps, cudaStream_t s) {
  long long blocks = (rows + kLnWarps - 1) / kLnWarps;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  ln_fwd_kernel<VPL><<<(unsigned)blocks, kLnWarps * 32, 0, s>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const float*>(gamma),
      static_cast<const float*>(beta), static_cast<__nv_bfloat16*>(y), mean, rstd, rows, D, eps);
  CLIPA_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return CLIPA_OK;
}

template <int VPL>
static int launch_ln_bwd(const void* dy, const void* x, const void* gamma, const float* mean,
                         const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta,
                         long long rows, int D, cudaStream_t s) {
  long long blocks = (rows + kLnWarps - 1) / kLnWarps;
  const long long cap = (long long)num_sms() * 2;
  if (blocks > cap) blocks = cap;
  const size_t smem = ((size_t)kLnWarps * 2 + 1) * D * sizeof(float);
  auto kern = ln_bwd_kernel<VPL>;
  if (smem > 48 * 1024)
    CLIPA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)blocks, kLnWarps * 32, smem, s>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x),
      static_cast<const float*>(gamma), mean, rstd, static_cast<const __nv_bfloat16*>(dres),
      static_cast<__nv_bfloat16*>(dx), dgamma, dbeta, rows, D);
  CLIPA_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return CLIPA_OK;
}

}  // namespace clipa

using namespace clipa;

extern "C" int clipa_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y,
                                   float* mean, float* rstd, int64_t rows, int32_t D, float eps,
                                   void* stream) {
  CLIPA_REQUIRE(x && gamma && beta && y, CLIPA_ERR_BAD_ARG, "layernorm_fwd: null pointer");
  CLIPA_REQUIRE((mean == nullptr) == (rstd == nullptr), CLIPA_ERR_BAD_ARG,
                "layernorm_fwd: mean and rstd must both be given or both NULL");
  CLIPA_REQUIRE(rows > 0 && D > 0 && D % 8 == 0 && D <= 2048, CLIPA_ERR_UNSUPPORTED,
                "layernorm_fwd: need rows>0 and D %% 8 == 0, D <= 2048 (rows=%lld D=%d)",
                (long long)rows, D);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int vpl = (D / 8 + 31) / 32;
  switch (vpl) {
    case 1: return launch_ln_fwd<1>(x, gamma, beta, y, mean, rstd, rows, D, eps, s);
    case 2: return launch_ln_fwd<2>(x, gamma, beta, y, mean, rstd, rows, D, eps, s);
    case 3: return launch_ln_fwd<3>(x, gamma, beta, y, mean, rstd, rows, D, eps, s);
    case 4: return launch_ln_fwd<4>(x, gamma, beta, y, mean, rstd, rows, D, eps, s);
    case 5: return launch_ln_fwd<5>(x, gamma, beta, y, mean, rstd, rows, D, eps, s);
    default: return launch_ln_fwd<8>(x, gamma, beta, y, mean, rstd, rows, D, eps, s);
  }
}

extern "C" int clipa_layernorm_bwd(const void* dy, const void* x, const void* gamma,
                                   const float* mean, const float* rstd, const void* dres, void* dx,
                                   float* dgamma, float* dbeta, int64_t rows, int32_t D,
                                   void* stream) {
  CLIPA_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta, CLIPA_ERR_BAD_ARG,
                "layernorm_bwd: null pointer");
  CLIPA_REQUIRE(rows > 0 && D > 0 && D % 8 == 0 && D <= 2048, CLIPA_ERR_UNSUPPORTED,
                "layernorm_bwd: need rows>0 and D %% 8 == 0, D <= 2048 (rows=%lld D=%d)",
                (long long)rows, D);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int vpl = (D / 8 + 31) / 32;
  switch (vpl) {
    case 1: return launch_ln_bwd<1>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, D, s);
    case 2: return launch_ln_bwd<2>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, D, s);
    case 3: return launch_ln_bwd<3>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, D, s);
    case 4: return launch_ln_bwd<4>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, D, s);
    case 5: return launch_ln_bwd<5>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, D, s);
    default: return launch_ln_bwd<8>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, D, s);
  }
}

extern "C" int clipa_colsum_accum(const void* x, int64_t ldx, float* out, int64_t rows, int32_t N,
                                  void* stream) {
  CLIPA_REQUIRE(x && out, CLIPA_ERR_BAD_ARG, "colsum: null pointer");
  CLIPA_REQUIRE(rows > 0 && N > 0 && N % 8 == 0 && ldx % 8 == 0, CLIPA_ERR_UNSUPPORTED,
                "colsum: need N %% 8 == 0 and ldx %% 8 == 0 (N=%d ldx=%lld)", N, (long long)ldx);
  const int col_blocks = (N + 255) / 256;
  long long row_blocks = ((long long)num_sms() * 8 + col_blocks - 1) / col_blocks;
  long long rpb = (rows + row_blocks - 1) / row_blocks;
  if (rpb < 64) rpb = 64;
  rpb = (rpb + 7) / 8 * 8;
  row_blocks = (rows + rpb - 1) / rpb;
  dim3 grid(col_blocks, (unsigned)row_blocks);
  colsum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, out, rows, N, rpb);
  CLIPA_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return CLIPA_OK;
}
