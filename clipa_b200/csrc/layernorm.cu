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

constexpr int kLnWarps = 8;      // forward / colsum: warps per block (2 blocks per SM)
constexpr int kLnBwdWarps = 12;  // backward: one 12-warp block per SM (170 registers per thread, no spills)

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

// ---- packed-pair helpers: a 16-byte vector of 8 bf16 <-> four float2; all row arithmetic runs on
// the packed-fp32 pipe (FADD2 / FMUL2 / FFMA2), which halves the issue slots per element -- with
// scalar fp32 math these kernels were issue-bound at ~55-60 % of the HBM roofline.
__device__ __forceinline__ float2 bf2_to_f2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ void unpack8p(const uint4& t, float2 (&f)[4]) {
  f[0] = bf2_to_f2(t.x); f[1] = bf2_to_f2(t.y); f[2] = bf2_to_f2(t.z); f[3] = bf2_to_f2(t.w);
}
__device__ __forceinline__ uint4 pack8p(const float2 (&f)[4]) {
  uint4 t;
  t.x = pack_bf16x2(f[0].x, f[0].y); t.y = pack_bf16x2(f[1].x, f[1].y);
  t.z = pack_bf16x2(f[2].x, f[2].y); t.w = pack_bf16x2(f[3].x, f[3].y);
  return t;
}
// Affine parameters in shared memory, split so that a lane's two float4 halves of its 8 columns sit
// in separate arrays (16-byte lane stride: conflict-free LDS.128):  lo[vi] = p[8vi..8vi+3],
// hi[vi] = p[8vi+4..8vi+7].
__device__ __forceinline__ void stage_param(float4* lo, float4* hi, const float* __restrict__ p, int nvec) {
  for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
    lo[vi] = __ldg(reinterpret_cast<const float4*>(p) + 2 * vi);
    hi[vi] = __ldg(reinterpret_cast<const float4*>(p) + 2 * vi + 1);
  }
}
__device__ __forceinline__ void load_param(const float4* lo, const float4* hi, int vi, float2 (&g)[4]) {
  const float4 a = lo[vi], b = hi[vi];
  g[0] = make_float2(a.x, a.y); g[1] = make_float2(a.z, a.w);
  g[2] = make_float2(b.x, b.y); g[3] = make_float2(b.z, b.w);
}

// Forward.  VPL = 16-byte vectors per lane (D <= 256*VPL).  A warp normalises TWO adjacent rows per
// iteration: eight 16-byte loads per lane are in flight before the first reduction, the two
// shuffle-reduction chains interleave, and every gamma/beta vector read from shared memory serves
// both rows.
template <int VPL, bool FULL>   // FULL: D == 256 * VPL, no column predicates
__global__ void __launch_bounds__(kLnWarps * 32, 2)
ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
              float* __restrict__ mean_out, float* __restrict__ rstd_out, long long rows, int D,
              float eps) {
  extern __shared__ float4 ln_smem4[];
  const int nvec = D >> 3;
  float4* g_lo = ln_smem4;
  float4* g_hi = g_lo + nvec;
  float4* b_lo = g_hi + nvec;
  float4* b_hi = b_lo + nvec;
  stage_param(g_lo, g_hi, gamma, nvec);
  stage_param(b_lo, b_hi, beta, nvec);
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  const long long warp_stride = (long long)gridDim.x * kLnWarps;
  const float inv_d = 1.0f / (float)D;
  for (long long pair = warp_global; 2 * pair < rows; pair += warp_stride) {
    const long long row0 = 2 * pair;
    const bool two = row0 + 1 < rows;
    const long long row1 = two ? row0 + 1 : row0;
    const uint4* xr0 = reinterpret_cast<const uint4*>(x + row0 * D);
    const uint4* xr1 = reinterpret_cast<const uint4*>(x + row1 * D);
    uint4 p0[VPL], p1[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + 32 * i;
      if (FULL || vi < nvec) {
        p0[i] = __ldg(xr0 + vi);
        p1[i] = __ldg(xr1 + vi);
      }
    }
    float2 v0[VPL][4], v1[VPL][4];
    float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (FULL || lane + 32 * i < nvec) {
        unpack8p(p0[i], v0[i]);
        unpack8p(p1[i], v1[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s0 = add2(s0, v0[i][j]);
          s1 = add2(s1, v1[i][j]);
        }
      }
    }
    float a0 = s0.x + s0.y, a1 = s1.x + s1.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    const float mean0 = a0 * inv_d, mean1 = a1 * inv_d;
    const float2 nm0 = splat2(-mean0), nm1 = splat2(-mean1);
    float2 q0 = make_float2(0.f, 0.f), q1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (FULL || lane + 32 * i < nvec) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v0[i][j] = add2(v0[i][j], nm0);
          v1[i][j] = add2(v1[i][j], nm1);
          q0 = fma2(v0[i][j], v0[i][j], q0);
          q1 = fma2(v1[i][j], v1[i][j], q1);
        }
      }
    }
    a0 = q0.x + q0.y;
    a1 = q1.x + q1.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    const float rstd0 = rsqrtf(a0 * inv_d + eps), rstd1 = rsqrtf(a1 * inv_d + eps);
    const float2 r0 = splat2(rstd0), r1 = splat2(rstd1);
    uint4* yr0 = reinterpret_cast<uint4*>(y + row0 * D);
    uint4* yr1 = reinterpret_cast<uint4*>(y + row1 * D);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + 32 * i;
      if (FULL || vi < nvec) {
        float2 g[4], b[4], o0[4], o1[4];
        load_param(g_lo, g_hi, vi, g);
        load_param(b_lo, b_hi, vi, b);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o0[j] = fma2(mul2(v0[i][j], r0), g[j], b[j]);
          o1[j] = fma2(mul2(v1[i][j], r1), g[j], b[j]);
        }
        yr0[vi] = pack8p(o0);
        if (two) yr1[vi] = pack8p(o1);
      }
    }
    if (lane == 0 && mean_out) {
      mean_out[row0] = mean0;
      rstd_out[row0] = rstd0;
      if (two) {
        mean_out[row1] = mean1;
        rstd_out[row1] = rstd1;
      }
    }
  }
}

// Backward.  With xhat = (x-mean)*rstd and g = dy*gamma:
//   dx = rstd * (g - mean_D(g) - xhat * mean_D(g*xhat)) (+ dres)
//   dgamma += sum_rows dy*xhat ;  dbeta += sum_rows dy
// Each warp keeps its dgamma/dbeta partials in registers over its grid-stride rows; the block
// reduces them through shared memory and issues one red.add per column per block.  The row is
// held PACKED (bf16x2) between the statistics pass and the output pass and gamma is read from
// shared memory; one 12-warp block per SM gives each thread 170 registers (64 of them partials).
// DXS: also accumulate the column sums of the OUTPUT dx into dxsum -- the bias gradient of the Linear layer
// whose output this LayerNorm's input stream received (d b_out = colsum(dx1), d b_proj of the previous block =
// colsum(dx)): the standalone column-sum pass over a tensor this kernel has just produced disappears.  The sums
// are kept per warp in shared memory (split float4 halves, conflict-free), not in registers.
template <int VPL, bool FULL, bool DXS>
__global__ void __launch_bounds__(kLnBwdWarps * 32, 1)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
              const float* __restrict__ gamma, const float* __restrict__ mean_in,
              const float* __restrict__ rstd_in, const __nv_bfloat16* __restrict__ dres,
              __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
              float* __restrict__ dxsum, long long rows, int D) {
  extern __shared__ float4 ln_smem4[];  // gamma (split halves), then [kLnBwdWarps][2][D] reduction scratch
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nvec = D >> 3;
  float4* g_lo = ln_smem4;
  float4* g_hi = g_lo + nvec;
  float* scratch = reinterpret_cast<float*>(g_hi + nvec);
  // [kLnBwdWarps][2][D] reduction scratch, then (DXS) [kLnBwdWarps][D] dx column sums as lo / hi float4 halves
  float4* sx_lo = reinterpret_cast<float4*>(scratch + (size_t)kLnBwdWarps * 2 * D) + (size_t)warp * 2 * nvec;
  float4* sx_hi = sx_lo + nvec;
  const long long warp_global = (long long)blockIdx.x * kLnBwdWarps + warp;
  const long long warp_stride = (long long)gridDim.x * kLnBwdWarps;
  const float inv_d = 1.0f / (float)D;

  stage_param(g_lo, g_hi, gamma, nvec);
  if (DXS) {
    for (int vi = lane; vi < nvec; vi += 32) {
      sx_lo[vi] = make_float4(0.f, 0.f, 0.f, 0.f);
      sx_hi[vi] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();

  float2 acc_g[VPL][4], acc_b[VPL][4];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc_g[i][j] = make_float2(0.f, 0.f); acc_b[i][j] = make_float2(0.f, 0.f); }
  }

  for (long long row = warp_global; row < rows; row += warp_stride) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * D);
    const uint4* dyr = reinterpret_cast<const uint4*>(dy + row * D);
    const uint4* rr = reinterpret_cast<const uint4*>(dres ? dres + row * D : x);
    const float mean = mean_in[row];
    const float rstd = rstd_in[row];
    uint4 xp[VPL], dp[VPL], rp[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + 32 * i;
      if (FULL || vi < nvec) {
        xp[i] = __ldg(xr + vi);
        dp[i] = __ldg(dyr + vi);
      }
    }
    const float2 r2 = splat2(rstd), nmr = splat2(-mean * rstd);
    float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + 32 * i;
      if (FULL || vi < nvec) {
        float2 xv[4], dv[4], gm[4];
        unpack8p(xp[i], xv);
        unpack8p(dp[i], dv);
        load_param(g_lo, g_hi, vi, gm);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 xh = fma2(xv[j], r2, nmr);
          const float2 g = mul2(dv[j], gm[j]);
          s1 = add2(s1, g);
          s2 = fma2(g, xh, s2);
          acc_g[i][j] = fma2(dv[j], xh, acc_g[i][j]);
          acc_b[i][j] = add2(acc_b[i][j], dv[j]);
        }
      }
    }
    if (dres) {   // issued ahead of the reductions; consumed in the output pass
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int vi = lane + 32 * i;
        if (FULL || vi < nvec) rp[i] = __ldg(rr + vi);
      }
    }
    float m1 = s1.x + s1.y, m2 = s2.x + s2.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m1 += __shfl_xor_sync(0xffffffffu, m1, o);
      m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    }
    // dx = rstd*g - rstd*m1 - xhat*(rstd*m2)
    const float2 c1 = splat2(-rstd * m1 * inv_d), c2 = splat2(-rstd * m2 * inv_d);
    uint4* dxr = reinterpret_cast<uint4*>(dx + row * D);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + 32 * i;
      if (FULL || vi < nvec) {
        float2 xv[4], dv[4], gm[4], o[4];
        unpack8p(xp[i], xv);
        unpack8p(dp[i], dv);
        load_param(g_lo, g_hi, vi, gm);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 xh = fma2(xv[j], r2, nmr);
          const float2 g = mul2(dv[j], gm[j]);
          o[j] = fma2(xh, c2, fma2(g, r2, c1));
        }
        if (dres) {
          float2 r[4];
          unpack8p(rp[i], r);
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = add2(o[j], r[j]);
        }
        dxr[vi] = pack8p(o);
        if (DXS) {
          float4 a = sx_lo[vi], b = sx_hi[vi];
          a.x += o[0].x; a.y += o[0].y; a.z += o[1].x; a.w += o[1].y;
          b.x += o[2].x; b.y += o[2].y; b.z += o[3].x; b.w += o[3].y;
          sx_lo[vi] = a;
          sx_hi[vi] = b;
        }
      }
    }
  }

  // block reduction of the parameter-gradient partials
  float* my_g = scratch + (size_t)warp * 2 * D;
  float* my_b = my_g + D;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + 32 * i;
    if (FULL || vi < nvec) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        *reinterpret_cast<float2*>(my_g + vi * 8 + 2 * j) = acc_g[i][j];
        *reinterpret_cast<float2*>(my_b + vi * 8 + 2 * j) = acc_b[i][j];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int w = 0; w < kLnBwdWarps; ++w) {
      sg += scratch[(size_t)w * 2 * D + c];
      sb += scratch[(size_t)w * 2 * D + D + c];
    }
    atomicAdd(dgamma + c, sg);
    atomicAdd(dbeta + c, sb);
    if (DXS) {
      const float* sx = scratch + (size_t)kLnBwdWarps * 2 * D;
      const int vi = c >> 3, r = c & 7;
      const int idx = r < 4 ? vi * 4 + r : nvec * 4 + vi * 4 + (r - 4);
      float sd = 0.f;
#pragma unroll
      for (int w = 0; w < kLnBwdWarps; ++w) sd += sx[(size_t)w * D + idx];
      atomicAdd(dxsum + c, sd);
    }
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
  long long blocks = ((rows + 1) / 2 + kLnWarps - 1) / kLnWarps;   // one warp per pair of rows
  const long long cap = (long long)num_sms() * 2;                   // persistent: 2 resident blocks per SM
  if (blocks > cap) blocks = cap;
  const size_t smem = (size_t)2 * D * sizeof(float);
  if (D == 256 * VPL)
    ln_fwd_kernel<VPL, true><<<(unsigned)blocks, kLnWarps * 32, smem, s>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const float*>(gamma),
        static_cast<const float*>(beta), static_cast<__nv_bfloat16*>(y), mean, rstd, rows, D, eps);
  else
    ln_fwd_kernel<VPL, false><<<(unsigned)blocks, kLnWarps * 32, smem, s>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const float*>(gamma),
        static_cast<const float*>(beta), static_cast<__nv_bfloat16*>(y), mean, rstd, rows, D, eps);
  CLIPA_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return CLIPA_OK;
}

int colsum_launch(const void* x, long long ldx, float* out, long long rows, int N, cudaStream_t s);

template <int VPL>
static int launch_ln_bwd(const void* dy, const void* x, const void* gamma, const float* mean,
                         const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta,
                         float* dxsum, long long rows, int D, cudaStream_t s) {
  long long blocks = (rows + kLnBwdWarps - 1) / kLnBwdWarps;
  const long long cap = (long long)num_sms();
  if (blocks > cap) blocks = cap;
  size_t smem = ((size_t)kLnBwdWarps * 2 + 1) * D * sizeof(float);
  const size_t smem_dxs = smem + (size_t)kLnBwdWarps * D * sizeof(float);
  const bool fuse = dxsum != nullptr && smem_dxs <= 227 * 1024;   // very wide rows: separate column-sum pass below
  if (fuse) smem = smem_dxs;
  const bool full = D == 256 * VPL;
  auto kern = fuse ? (full ? ln_bwd_kernel<VPL, true, true> : ln_bwd_kernel<VPL, false, true>)
                   : (full ? ln_bwd_kernel<VPL, true, false> : ln_bwd_kernel<VPL, false, false>);
  if (smem > 48 * 1024)
    CLIPA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)blocks, kLnBwdWarps * 32, smem, s>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x),
      static_cast<const float*>(gamma), mean, rstd, static_cast<const __nv_bfloat16*>(dres),
      static_cast<__nv_bfloat16*>(dx), dgamma, dbeta, dxsum, rows, D);
  CLIPA_CHECK_CUDA(cudaGetLastError());
  count_launch();
  if (dxsum != nullptr && !fuse) return colsum_launch(dx, D, dxsum, rows, D, s);
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
                                   float* dgamma, float* dbeta, float* dxsum, int64_t rows, int32_t D,
                                   void* stream) {
  CLIPA_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta, CLIPA_ERR_BAD_ARG,
                "layernorm_bwd: null pointer");
  CLIPA_REQUIRE(rows > 0 && D > 0 && D % 8 == 0 && D <= 2048, CLIPA_ERR_UNSUPPORTED,
                "layernorm_bwd: need rows>0 and D %% 8 == 0, D <= 2048 (rows=%lld D=%d)",
                (long long)rows, D);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int vpl = (D / 8 + 31) / 32;
  switch (vpl) {
    case 1: return launch_ln_bwd<1>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, dxsum, rows, D, s);
    case 2: return launch_ln_bwd<2>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, dxsum, rows, D, s);
    case 3: return launch_ln_bwd<3>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, dxsum, rows, D, s);
    case 4: return launch_ln_bwd<4>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, dxsum, rows, D, s);
    case 5: return launch_ln_bwd<5>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, dxsum, rows, D, s);
    default: return launch_ln_bwd<8>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, dxsum, rows, D, s);
  }
}

extern "C" int clipa_colsum_accum(const void* x, int64_t ldx, float* out, int64_t rows, int32_t N,
                                  void* stream) {
  CLIPA_REQUIRE(x && out, CLIPA_ERR_BAD_ARG, "colsum: null pointer");
  CLIPA_REQUIRE(rows > 0 && N > 0 && N % 8 == 0 && ldx % 8 == 0, CLIPA_ERR_UNSUPPORTED,
                "colsum: need N %% 8 == 0 and ldx %% 8 == 0 (N=%d ldx=%lld)", N, (long long)ldx);
  return clipa::colsum_launch(x, ldx, out, rows, N, static_cast<cudaStream_t>(stream));
}

int clipa::colsum_launch(const void* x, long long ldx, float* out, long long rows, int N, cudaStream_t stream) {
  const int col_blocks = (N + 255) / 256;
  long long row_blocks = ((long long)num_sms() * 8 + col_blocks - 1) / col_blocks;
  long long rpb = (rows + row_blocks - 1) / row_blocks;
  if (rpb < 64) rpb = 64;
  rpb = (rpb + 7) / 8 * 8;
  row_blocks = (rows + rpb - 1) / rpb;
  dim3 grid(col_blocks, (unsigned)row_blocks);
  colsum_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx, out, rows, N, rpb);
  CLIPA_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return CLIPA_OK;
}
