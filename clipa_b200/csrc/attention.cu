// Multi-head self-attention core for CLIPA's short sequences (L = 8..257 tokens, head_dim 64/80).
//
// One CTA per (sample, head): Q, K, V (and dO in backward) of that head are staged once in shared
// memory with cp.async (padded rows -> conflict-free ldmatrix), every warp owns 16-row tiles and
// runs the whole key range out of shared memory with online softmax (fp32) -- the score matrix
// never leaves registers.  Tensor work is mma.sync m16n8k16 (bf16 in, fp32 accumulate): at these
// sequence lengths the kernel is bound by HBM traffic (41 KB moved per 1.7 MFLOP at L=82), not by
// the tensor pipe; see DESIGN.md for the roofline.
//
// Reference semantics: F.scaled_dot_product_attention inside nn.MultiheadAttention
// (open_clip/transformer.py:234-236) with the optional causal additive mask of the text tower
// (open_clip/transformer.py:618-624), dropout 0.
#include "host_common.h"
#include "ptx.cuh"

namespace clipa {

constexpr int kAttnMaxWarps = 8;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Shared-memory tile of `rows` x HD bf16 with a 16-byte pad per row (conflict-free ldmatrix).
template <int HD>
struct Tile {
  static constexpr int kStride = HD + 8;  // elements
  __nv_bfloat16* p;
  __device__ __forceinline__ uint32_t addr(int row, int col) const {
    return smem_u32(p + row * kStride + col);
  }
};

// stage rows [0, L) of one head (global row pitch `pitch` elements) and zero rows [L, Lp)
template <int HD>
__device__ __forceinline__ void stage_tile(Tile<HD> t, const __nv_bfloat16* g, long long pitch,
                                           int L, int Lp) {
  constexpr int kChunks = HD / 8;
  for (int idx = threadIdx.x; idx < Lp * kChunks; idx += blockDim.x) {
    const int r = idx / kChunks, c = idx - r * kChunks;
    __nv_bfloat16* dst = t.p + r * Tile<HD>::kStride + c * 8;
    if (r < L) cp_async16(dst, g + (long long)r * pitch + c * 8);
    else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
  }
}

// A-operand fragments (16 rows x HD) for the rows starting at row0
template <int HD>
__device__ __forceinline__ void load_a_frags(const Tile<HD>& t, int row0, uint32_t (&f)[HD / 16][4]) {
  const int lane = threadIdx.x & 31;
  const int r = row0 + (lane & 7) + 8 * ((lane >> 3) & 1);
  const int c = 8 * (lane >> 4);
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) ldsm_x4(t.addr(r, kk * 16 + c), f[kk]);
}

// acc[nt] (16 x 8 per n-tile, 8 n-tiles = 64 columns j0..j0+63) += A(16 x HD) * T[j][:]^T
// i.e. B[k=d][n=j] = T[j][d]; used for Q K^T and dO V^T.
template <int HD>
__device__ __forceinline__ void mma_a_tT(const uint32_t (&a)[HD / 16][4], const Tile<HD>& t, int j0,
                                         int Lp, float (&acc)[8][4]) {
  const int lane = threadIdx.x & 31;
  const int jr = (lane & 7) + 8 * (lane >> 4);
  const int dc = 8 * ((lane >> 3) & 1);
#pragma unroll
  for (int np = 0; np < 4; ++np) {
    if (j0 + 16 * np < Lp) {
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        uint32_t b[4];
        ldsm_x4(t.addr(j0 + 16 * np + jr, kk * 16 + dc), b);
        mma_bf16(acc[2 * np], a[kk], b[0], b[1]);
        mma_bf16(acc[2 * np + 1], a[kk], b[2], b[3]);
      }
    }
  }
}

// out[dn] (16 x 8 per n-tile over HD) += P(16 x 64 as bf16 A-fragments pa[4][4]) * T[j0.., :]
// i.e. B[k=j][n=d] = T[j][d]; used for P V and dS K.
template <int HD>
__device__ __forceinline__ void mma_p_t(const uint32_t (&pa)[4][4], const Tile<HD>& t, int j0, int Lp,
                                        float (&out)[HD / 8][4]) {
  const int lane = threadIdx.x & 31;
  const int jr = (lane & 7) + 8 * ((lane >> 3) & 1);
  const int dc = 8 * (lane >> 4);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    if (j0 + 16 * kk < Lp) {
#pragma unroll
      for (int dp = 0; dp < HD / 16; ++dp) {
        uint32_t b[4];
        ldsm_x4_t(t.addr(j0 + 16 * kk + jr, dp * 16 + dc), b);
        mma_bf16(out[2 * dp], pa[kk], b[0], b[1]);
        mma_bf16(out[2 * dp + 1], pa[kk], b[2], b[3]);
      }
    }
  }
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                float* __restrict__ lse, int L, int H, int causal, float scale_log2) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int Lp = (L + 15) & ~15;
  const int D = H * HD;
  const int n = blockIdx.x / H, h = blockIdx.x - n * H;
  const long long pitch = 3LL * D;
  const __nv_bfloat16* base = qkv + (long long)n * L * pitch + h * HD;

  Tile<HD> sQ{reinterpret_cast<__nv_bfloat16*>(smem_attn)};
  Tile<HD> sK{sQ.p + Lp * Tile<HD>::kStride};
  Tile<HD> sV{sK.p + Lp * Tile<HD>::kStride};
  stage_tile<HD>(sQ, base, pitch, L, Lp);
  stage_tile<HD>(sK, base + D, pitch, L, Lp);
  stage_tile<HD>(sV, base + 2 * D, pitch, L, Lp);
  cp_async_wait_all();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;

  for (int row0 = warp * 16; row0 < Lp; row0 += nwarps * 16) {
    uint32_t qf[HD / 16][4];
    load_a_frags<HD>(sQ, row0, qf);
    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const int i0 = row0 + g, i1 = row0 + g + 8;

    for (int j0 = 0; j0 < Lp; j0 += 64) {
      if (causal && j0 > row0 + 15) break;  // whole chunk is above the diagonal
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
      mma_a_tT<HD>(qf, sK, j0, Lp, s);
      float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = j0 + nt * 8 + 2 * t + (e & 1);
          const int i = (e & 2) ? i1 : i0;
          const bool ok = (j < L) && !(causal && j > i);
          const float v = ok ? s[nt][e] * scale_log2 : -INFINITY;
          s[nt][e] = v;
          if (e & 2) cm1 = fmaxf(cm1, v); else cm0 = fmaxf(cm0, v);
        }
      }
      cm0 = quad_max(cm0);
      cm1 = quad_max(cm1);
      const float mn0 = fmaxf(m0, cm0), mn1 = fmaxf(m1, cm1);
      // rows that have seen no valid key yet keep m = -inf; guard the (-inf) - (-inf) case
      const float ms0 = (mn0 == -INFINITY) ? 0.f : mn0, ms1 = (mn1 == -INFINITY) ? 0.f : mn1;
      const float c0 = exp2f(m0 - ms0), c1 = exp2f(m1 - ms1);
      float rs0 = 0.f, rs1 = 0.f;
      uint32_t pa[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float p0 = exp2f(s[nt][0] - ms0), p1 = exp2f(s[nt][1] - ms0);
        const float p2 = exp2f(s[nt][2] - ms1), p3 = exp2f(s[nt][3] - ms1);
        rs0 += p0 + p1;
        rs1 += p2 + p3;
        pa[nt >> 1][(nt & 1) * 2] = pack_bf16x2(p0, p1);
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
      }
      l0 = l0 * c0 + quad_sum(rs0);
      l1 = l1 * c1 + quad_sum(rs1);
      m0 = mn0;
      m1 = mn1;
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
      mma_p_t<HD>(pa, sV, j0, Lp, o);
    }

    const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
    __nv_bfloat16* orow0 = out + ((long long)n * L + i0) * D + h * HD;
    __nv_bfloat16* orow1 = out + ((long long)n * L + i1) * D + h * HD;
#pragma unroll
    for (int dn = 0; dn < HD / 8; ++dn) {
      const int d = dn * 8 + 2 * t;
      if (i0 < L) *reinterpret_cast<uint32_t*>(orow0 + d) = pack_bf16x2(o[dn][0] * inv0, o[dn][1] * inv0);
      if (i1 < L) *reinterpret_cast<uint32_t*>(orow1 + d) = pack_bf16x2(o[dn][2] * inv1, o[dn][3] * inv1);
    }
    if (t == 0) {
      float* lrow = lse + ((long long)n * H + h) * L;
      if (i0 < L) lrow[i0] = (m0 + log2f(l0)) * 0.69314718055994531f;
      if (i1 < L) lrow[i1] = (m1 + log2f(l1)) * 0.69314718055994531f;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward.  With P = exp(S - lse), delta_i = sum_d dO_id O_id:
//   dV = P^T dO ; dP = dO V^T ; dS = P o (dP - delta) * scale ; dQ = dS K ; dK = dS^T Q
// Pass 1 is row-parallel (a warp owns 16 queries -> dQ); pass 2 is column-parallel (a warp owns
// 16 keys, recomputes the transposed tiles -> dK, dV).  No atomics, no global scratch.
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out,
                                const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                                __nv_bfloat16* __restrict__ dqkv, int L, int H, int causal, float scale) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int Lp = (L + 15) & ~15;
  const int D = H * HD;
  const int n = blockIdx.x / H, h = blockIdx.x - n * H;
  const long long pitch = 3LL * D;
  const __nv_bfloat16* base = qkv + (long long)n * L * pitch + h * HD;
  const __nv_bfloat16* obase = out + (long long)n * L * D + h * HD;
  const __nv_bfloat16* dobase = dout + (long long)n * L * D + h * HD;
  __nv_bfloat16* dbase = dqkv + (long long)n * L * pitch + h * HD;

  Tile<HD> sQ{reinterpret_cast<__nv_bfloat16*>(smem_attn)};
  Tile<HD> sK{sQ.p + Lp * Tile<HD>::kStride};
  Tile<HD> sV{sK.p + Lp * Tile<HD>::kStride};
  Tile<HD> sdO{sV.p + Lp * Tile<HD>::kStride};
  float* sLse2 = reinterpret_cast<float*>(sdO.p + Lp * Tile<HD>::kStride);
  float* sDelta = sLse2 + Lp;
  stage_tile<HD>(sQ, base, pitch, L, Lp);
  stage_tile<HD>(sK, base + D, pitch, L, Lp);
  stage_tile<HD>(sV, base + 2 * D, pitch, L, Lp);
  stage_tile<HD>(sdO, dobase, D, L, Lp);
  cp_async_wait_all();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const float scale_log2 = scale * 1.4426950408889634f;

  // delta_i and log2-domain lse_i
  for (int i = warp; i < Lp; i += nwarps) {
    float acc = 0.f;
    if (i < L && lane < HD / 8) {
      const uint4 ov = __ldg(reinterpret_cast<const uint4*>(obase + (long long)i * D) + lane);
      const uint4 dv = *reinterpret_cast<const uint4*>(sdO.p + i * Tile<HD>::kStride + lane * 8);
      acc = bf16lo(ov.x) * bf16lo(dv.x) + bf16hi(ov.x) * bf16hi(dv.x) +
            bf16lo(ov.y) * bf16lo(dv.y) + bf16hi(ov.y) * bf16hi(dv.y) +
            bf16lo(ov.z) * bf16lo(dv.z) + bf16hi(ov.z) * bf16hi(dv.z) +
            bf16lo(ov.w) * bf16lo(dv.w) + bf16hi(ov.w) * bf16hi(dv.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      sDelta[i] = acc;
      sLse2[i] = (i < L) ? lse[((long long)n * H + h) * L + i] * 1.4426950408889634f : 0.f;
    }
  }
  __syncthreads();

  // ---------------- pass 1: dQ (row-parallel) ----------------
  for (int row0 = warp * 16; row0 < Lp; row0 += nwarps * 16) {
    uint32_t qf[HD / 16][4], dof[HD / 16][4];
    load_a_frags<HD>(sQ, row0, qf);
    load_a_frags<HD>(sdO, row0, dof);
    float dq[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }
    const int i0 = row0 + g, i1 = row0 + g + 8;
    const float ls0 = sLse2[i0], ls1 = sLse2[i1];
    const float de0 = sDelta[i0], de1 = sDelta[i1];
    for (int j0 = 0; j0 < Lp; j0 += 64) {
      if (causal && j0 > row0 + 15) break;
      float s[8][4], dp[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
        dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
      }
      mma_a_tT<HD>(qf, sK, j0, Lp, s);
      mma_a_tT<HD>(dof, sV, j0, Lp, dp);
      uint32_t dsa[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        float ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = j0 + nt * 8 + 2 * t + (e & 1);
          const int i = (e & 2) ? i1 : i0;
          const bool ok = (j < L) && (i < L) && !(causal && j > i);
          const float p = ok ? exp2f(s[nt][e] * scale_log2 - ((e & 2) ? ls1 : ls0)) : 0.f;
          ds[e] = p * (dp[nt][e] - ((e & 2) ? de1 : de0)) * scale;
        }
        dsa[nt >> 1][(nt & 1) * 2] = pack_bf16x2(ds[0], ds[1]);
        dsa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
      }
      mma_p_t<HD>(dsa, sK, j0, Lp, dq);
    }
#pragma unroll
    for (int dn = 0; dn < HD / 8; ++dn) {
      const int d = dn * 8 + 2 * t;
      if (i0 < L) *reinterpret_cast<uint32_t*>(dbase + (long long)i0 * pitch + d) = pack_bf16x2(dq[dn][0], dq[dn][1]);
      if (i1 < L) *reinterpret_cast<uint32_t*>(dbase + (long long)i1 * pitch + d) = pack_bf16x2(dq[dn][2], dq[dn][3]);
    }
  }

  // ---------------- pass 2: dK, dV (column-parallel; tiles are transposed) ----------------
  for (int key0 = warp * 16; key0 < Lp; key0 += nwarps * 16) {
    uint32_t kf[HD / 16][4], vf[HD / 16][4];
    load_a_frags<HD>(sK, key0, kf);
    load_a_frags<HD>(sV, key0, vf);
    float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
      dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    const int j0r = key0 + g, j1r = key0 + g + 8;  // this thread's key rows
    for (int q0 = 0; q0 < Lp; q0 += 64) {
      if (causal && q0 + 63 < key0) continue;  // every query in the chunk precedes these keys
      float st[8][4], dpt[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
        dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
      }
      mma_a_tT<HD>(kf, sQ, q0, Lp, st);    // S^T  = K Q^T   (rows = keys, cols = queries)
      mma_a_tT<HD>(vf, sdO, q0, Lp, dpt);  // dP^T = V dO^T
      uint32_t pta[4][4], dsta[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        float pv[4], dsv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = q0 + nt * 8 + 2 * t + (e & 1);  // query index (column)
          const int j = (e & 2) ? j1r : j0r;            // key index (row)
          const bool ok = (i < L) && (j < L) && !(causal && j > i);
          const int ic = min(i, Lp - 1);
          const float p = ok ? exp2f(st[nt][e] * scale_log2 - sLse2[ic]) : 0.f;
          pv[e] = p;
          dsv[e] = p * (dpt[nt][e] - sDelta[ic]) * scale;
        }
        pta[nt >> 1][(nt & 1) * 2] = pack_bf16x2(pv[0], pv[1]);
        pta[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(pv[2], pv[3]);
        dsta[nt >> 1][(nt & 1) * 2] = pack_bf16x2(dsv[0], dsv[1]);
        dsta[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(dsv[2], dsv[3]);
      }
      mma_p_t<HD>(pta, sdO, q0, Lp, dv);  // dV += P^T dO
      mma_p_t<HD>(dsta, sQ, q0, Lp, dk);  // dK += dS^T Q
    }
#pragma unroll
    for (int dn = 0; dn < HD / 8; ++dn) {
      const int d = dn * 8 + 2 * t;
      if (j0r < L) {
        *reinterpret_cast<uint32_t*>(dbase + (long long)j0r * pitch + D + d) = pack_bf16x2(dk[dn][0], dk[dn][1]);
        *reinterpret_cast<uint32_t*>(dbase + (long long)j0r * pitch + 2 * D + d) = pack_bf16x2(dv[dn][0], dv[dn][1]);
      }
      if (j1r < L) {
        *reinterpret_cast<uint32_t*>(dbase + (long long)j1r * pitch + D + d) = pack_bf16x2(dk[dn][2], dk[dn][3]);
        *reinterpret_cast<uint32_t*>(dbase + (long long)j1r * pitch + 2 * D + d) = pack_bf16x2(dv[dn][2], dv[dn][3]);
      }
    }
  }
}

static int attn_warps(int L) {
  int w = (L + 15) / 16;
  return w > kAttnMaxWarps ? kAttnMaxWarps : w;
}

template <int HD>
static int launch_attn_fwd(const void* qkv, void* out, float* lse, int batch, int L, int H, int causal,
                           cudaStream_t s) {
  const int Lp = (L + 15) & ~15;
  const size_t smem = (size_t)3 * Lp * (HD + 8) * 2;
  CLIPA_REQUIRE(smem <= 227 * 1024, CLIPA_ERR_UNSUPPORTED,
                "attention_fwd: L=%d needs %zu B of shared memory (> 227 KB)", L, smem);
  auto kern = attn_fwd_kernel<HD>;
  if (smem > 48 * 1024)
    CLIPA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)HD);
  kern<<<batch * H, attn_warps(L) * 32, smem, s>>>(static_cast<const __nv_bfloat16*>(qkv),
                                                   static_cast<__nv_bfloat16*>(out), lse, L, H, causal,
                                                   scale_log2);
  CLIPA_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return CLIPA_OK;
}

template <int HD>
static int launch_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                           void* dqkv, int batch, int L, int H, int causal, cudaStream_t s) {
  const int Lp = (L + 15) & ~15;
  const size_t smem = (size_t)4 * Lp * (HD + 8) * 2 + (size_t)2 * Lp * sizeof(float);
  CLIPA_REQUIRE(smem <= 227 * 1024, CLIPA_ERR_UNSUPPORTED,
                "attention_bwd: L=%d needs %zu B of shared memory (> 227 KB)", L, smem);
  auto kern = attn_bwd_kernel<HD>;
  if (smem > 48 * 1024)
    CLIPA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const float scale = 1.0f / sqrtf((float)HD);
  kern<<<batch * H, attn_warps(L) * 32, smem, s>>>(
      static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(out),
      static_cast<const __nv_bfloat16*>(dout), lse, static_cast<__nv_bfloat16*>(dqkv), L, H, causal,
      scale);
  CLIPA_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return CLIPA_OK;
}

int attention_fwd_tc(const void* qkv, void* out, float* lse, int batch, int L, int H, int causal,
                     cudaStream_t stream);  // attention_tc.cu
int attention_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                     int batch, int L, int H, int causal, cudaStream_t stream);
int attention_fwd_flash(const void* qkv, void* out, float* lse, int batch, int L, int H, int hd, int causal,
                        cudaStream_t stream);  // attention_flash.cu
int attention_bwd_flash(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                        void* workspace, long long workspace_bytes, int batch, int L, int H, int hd, int causal,
                        cudaStream_t stream);
long long attention_bwd_flash_workspace(int batch, int L, int H, int hd);

// 0 = auto; 1 = mma.sync kernels wherever they take the shape; 2 = flash tcgen05 kernels for every
// head_dim 64 / 80 shape (also the one-tile shapes attention_tc.cu would take).  Tests and A/B runs only.
static int g_attn_mode = 0;

enum AttnPath { kPathTc, kPathFlash, kPathMma };
static AttnPath attn_path(int L, int hd) {
  const bool flash_ok = hd == 64 || hd == 80;
  if (g_attn_mode == 1) return kPathMma;
  if (g_attn_mode == 2 && flash_ok) return kPathFlash;
  if (hd == 64 && L <= 128) return kPathTc;       // one-tile kernels (packed short sequences, pipelined bwd)
  if (flash_ok) return kPathFlash;                // L > 128 (224 / 336 px fine-tune) and every ViT-H tower
  return kPathMma;                                // head_dim 96 / 128
}

}  // namespace clipa

using namespace clipa;

extern "C" int clipa_set_attention_mode(int mode) {
  CLIPA_REQUIRE(mode >= 0 && mode <= 2, CLIPA_ERR_BAD_ARG, "clipa_set_attention_mode: mode %d", mode);
  g_attn_mode = mode;
  return CLIPA_OK;
}

extern "C" int clipa_attention_fwd(const void* qkv, void* out, float* lse, int32_t batch, int32_t L,
                                   int32_t heads, int32_t head_dim, int32_t causal, void* stream) {
  CLIPA_REQUIRE(qkv && out && lse, CLIPA_ERR_BAD_ARG, "attention_fwd: null pointer");
  CLIPA_REQUIRE(batch > 0 && L > 0 && heads > 0, CLIPA_ERR_BAD_ARG, "attention_fwd: bad dims");
  CLIPA_REQUIRE((long long)batch * heads < (1LL << 31), CLIPA_ERR_UNSUPPORTED, "attention_fwd: grid too large");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (attn_path(L, head_dim)) {
    case kPathTc: return attention_fwd_tc(qkv, out, lse, batch, L, heads, causal, s);
    case kPathFlash: return attention_fwd_flash(qkv, out, lse, batch, L, heads, head_dim, causal, s);
    default: break;
  }
  switch (head_dim) {
    case 64: return launch_attn_fwd<64>(qkv, out, lse, batch, L, heads, causal, s);
    case 80: return launch_attn_fwd<80>(qkv, out, lse, batch, L, heads, causal, s);
    case 96: return launch_attn_fwd<96>(qkv, out, lse, batch, L, heads, causal, s);
    case 128: return launch_attn_fwd<128>(qkv, out, lse, batch, L, heads, causal, s);
    default:
      set_error("attention_fwd: head_dim %d not built (64, 80, 96, 128)", head_dim);
      return CLIPA_ERR_UNSUPPORTED;
  }
}

extern "C" int64_t clipa_attention_bwd_workspace(int32_t batch, int32_t L, int32_t heads, int32_t head_dim) {
  if (batch <= 0 || L <= 0 || heads <= 0) return 0;
  return attn_path(L, head_dim) == kPathFlash ? attention_bwd_flash_workspace(batch, L, heads, head_dim) : 0;
}

extern "C" int clipa_attention_bwd(const void* qkv, const void* out, const void* dout,
                                   const float* lse, void* dqkv, void* workspace, int64_t workspace_bytes,
                                   int32_t batch, int32_t L, int32_t heads, int32_t head_dim, int32_t causal,
                                   void* stream) {
  CLIPA_REQUIRE(qkv && out && dout && lse && dqkv, CLIPA_ERR_BAD_ARG, "attention_bwd: null pointer");
  CLIPA_REQUIRE(batch > 0 && L > 0 && heads > 0, CLIPA_ERR_BAD_ARG, "attention_bwd: bad dims");
  CLIPA_REQUIRE((long long)batch * heads < (1LL << 31), CLIPA_ERR_UNSUPPORTED, "attention_bwd: grid too large");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (attn_path(L, head_dim)) {
    case kPathTc: return attention_bwd_tc(qkv, out, dout, lse, dqkv, batch, L, heads, causal, s);
    case kPathFlash:
      return attention_bwd_flash(qkv, out, dout, lse, dqkv, workspace, workspace_bytes, batch, L, heads, head_dim,
                                 causal, s);
    default: break;
  }
  switch (head_dim) {
    case 64: return launch_attn_bwd<64>(qkv, out, dout, lse, dqkv, batch, L, heads, causal, s);
    case 80: return launch_attn_bwd<80>(qkv, out, dout, lse, dqkv, batch, L, heads, causal, s);
    case 96: return launch_attn_bwd<96>(qkv, out, dout, lse, dqkv, batch, L, heads, causal, s);
    case 128: return launch_attn_bwd<128>(qkv, out, dout, lse, dqkv, batch, L, heads, causal, s);
    default:
      set_error("attention_bwd: head_dim %d not built (64, 80, 96, 128)", head_dim);
      return CLIPA_ERR_UNSUPPORTED;
  }
}
