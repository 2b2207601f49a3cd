// tcgen05 attention for CLIPA's short sequences: head_dim 64, L <= 128 (one M=128 tile holds every
// query of a (sample, head) problem; configs 1, 2, 3, 5 and every text tower).  The mma.sync kernels
// in attention.cu cover longer sequences and other head widths.
//
// FORWARD -- persistent CTA, 10 warps, two problems in flight (S of problem i+2 is issued as soon as
// the softmax of problem i has left TMEM, so a softmax group never waits for its scores):
//   warp 0      TMA producer: Q, K, V head slices (L rows x 128 B, 128B-swizzled) -> 3-stage smem ring
//   warp 1      MMA issuer:   S = Q K^T  (tcgen05.mma M128 x N=ceil16(L) x K64, fp32 in TMEM)
//                             O = P V    (A = P from smem, B = V consumed MN-major in place)
//   warps 2-5   softmax/epilogue group for even problems   } one thread per query row:
//   warps 6-9   softmax/epilogue group for odd problems    } tcgen05.ld S row (once) -> mask, max,
//                                                            exp2, sum -> P (bf16) to swizzled smem;
//                                                            later tcgen05.ld O row -> 1/l -> 128 B store
// The score matrix lives only in TMEM/registers; per problem HBM traffic is the algorithmic
// minimum (Q, K, V read once, O written once).
//
// Short sequences are PACKED: G = floor(128 / L) consecutive samples of the same head share one
// M = 128 tile (text towers: 8 x 16 tokens, 16 x 8 tokens); the score tile is then block-diagonal --
// row i only attends to the columns of its own sample -- and every warp touches only the 32-column
// chunks its rows' blocks overlap, so the work per sample does not grow with the packing.
#include <type_traits>

#include "attn_common.cuh"
#include "host_common.h"
#include "ptx.cuh"

namespace clipa {

constexpr int kTcHd = 64;
constexpr int kTcThreads = 320;
constexpr int kTcStageBytes = 3 * kTcTileBytes;             // Q, K, V
// Q/K/V smem ring: 3 slots for one-sample tiles, 2 for packed tiles (which then keep 63 KB of L1 for
// their all-rows-active output stores; with the 3-slot footprint they ran 10 % slower)
constexpr int tc_ring(bool packed) { return packed ? 2 : 3; }
constexpr int tc_smem_bar(bool packed) { return tc_ring(packed) * kTcStageBytes + 2 * kTcPBytes; }
constexpr int tc_smem_total(bool packed) { return tc_smem_bar(packed) + 256 + 1024; }

struct AttnTcParams {
  __nv_bfloat16* out;
  float* lse;
  int L, H, batch;
  int G;     // samples packed per 128-row tile
  int GL;    // G * L rows / keys in use per tile
  int npad;  // ceil16(G * L)
  float scale_log2;
};

// Per-thread view of the packing: which key columns row `row` may attend to.
struct RowSpan {
  int lo, hi;     // valid key columns [lo, hi)
  int sample;     // sample index within the tile
  int token;      // token index within the sample
};
template <bool CAUSAL, bool PACKED>
__device__ __forceinline__ RowSpan row_span(int row, int L, int GL) {
  RowSpan r;
  if (!PACKED) {   // one sample per tile: everything folds to the unpacked constants
    r.sample = 0;
    r.lo = 0;
    r.token = row;
    r.hi = CAUSAL ? row + 1 : L;
    return r;
  }
  const int rr = min(row, GL - 1);
  r.sample = rr / L;
  r.lo = r.sample * L;
  r.token = rr - r.lo;
  r.hi = CAUSAL ? rr + 1 : r.lo + L;
  return r;
}

// NCH = 32-column chunks of the score row held in registers; PACKED = more than one sample per tile
template <int NCH, bool CAUSAL, bool PACKED>
__global__ void __launch_bounds__(kTcThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  constexpr int kRing = tc_ring(PACKED);
  constexpr int kTcSmemBar = tc_smem_bar(PACKED);
  uint8_t* p_base = smem + kRing * kTcStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTcSmemBar);
  uint64_t* full = bars;           // [3] TMA landed
  uint64_t* kv_empty = bars + 3;   // [3] smem stage free
  uint64_t* s_full = bars + 6;     // [2] S in TMEM
  uint64_t* p_full = bars + 8;     // [2] P in smem (and S consumed)
  uint64_t* o_full = bars + 10;    // [2] O in TMEM
  uint64_t* t_free = bars + 12;    // [2] TMEM stage + P buffer free (PACKED schedule only)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14);
  // Two schedules, chosen by measurement (profiles/r01_attn_fwd_schedules.log):
  //  * one sample per tile (ViT towers): 3-slot operand ring, S(i+2) is issued right behind PV(i) --
  //    the group finds its next scores in TMEM when it comes back from storing O(i)
  //    (L = 82: 222 -> 148 us);
  //  * packed tiles (text towers, 48 KB of operands per tile, HBM-latency-bound): 2-slot ring, S(i)
  //    issued once the group has read O(i-2); the eager schedule was 10-25 % slower here.

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L, H = p.H, D = H * kTcHd;
  const int tiles_per_head = (p.batch + p.G - 1) / p.G;
  const int total = tiles_per_head * H;

  // Stale rows [L,128) of the K/V tiles are multiplied by exactly-zero probabilities, so they must
  // never hold NaN/Inf bit patterns: zero everything once.
  for (int i = threadIdx.x; i < kTcSmemBar / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    for (int i = 0; i < kRing; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&t_free[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // number of problems this CTA owns: blockIdx.x, +gridDim.x, ...
  const int n_local = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < n_local; ++i) {
        const int s = i % kRing;
        const uint32_t ph = (i / kRing) & 1;
        const int prob = blockIdx.x + i * gridDim.x;
        const int n = (prob / H) * p.G, h = prob % H;   // first sample of the tile, head
        mbar_wait(&kv_empty[s], ph ^ 1);
        uint8_t* st = smem + s * kTcStageBytes;
        mbar_expect_tx(&full[s], 3u * (uint32_t)p.GL * 128u);
        tma_load_2d(st, &tmap_qkv, &full[s], h * kTcHd, n * L);
        tma_load_2d(st + kTcTileBytes, &tmap_qkv, &full[s], D + h * kTcHd, n * L);
        tma_load_2d(st + 2 * kTcTileBytes, &tmap_qkv, &full[s], 2 * D + h * kTcHd, n * L);
      }
    }
  } else if (warp == 1) {
    const IssueMode im = issue_mode(lane);
    if (im.in_loop) {
      const uint32_t idesc_s = make_idesc_bf16(128, (uint32_t)p.npad, false, false);
      const uint32_t idesc_o = make_idesc_bf16(128, kTcHd, false, true);
      const int ksteps = p.npad / 16;
      // TMEM stage / P buffer j & 1, smem ring slot j % kRing
      auto issue_s = [&](int j) {
        const int r = j % kRing;
        mbar_wait(&full[r], (j / kRing) & 1);
        if constexpr (PACKED) mbar_wait(&t_free[j & 1], ((j >> 1) & 1) ^ 1);   // group has read O(j-2)
        tc_fence_after();
        const uint32_t qa = smem_u32(smem + r * kTcStageBytes);
        const uint32_t ka = qa + kTcTileBytes;
        const uint32_t d_s = tmem_base + (j & 1) * 256;
#pragma unroll
        for (int k = 0; k < kTcHd / 16; ++k) {
          const uint64_t da = make_smem_desc_sw128(qa + k * 32, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(ka + k * 32, 16, 1024);
          if (im.issue) umma_bf16(d_s, da, db, idesc_s, k > 0 ? 1u : 0u);
        }
        if (im.issue) umma_commit(&s_full[j & 1]);
        im.sync();
      };
      auto issue_pv = [&](int i) {
        const int sj = i & 1;
        const int r = i % kRing;
        // P(i) written, S(i) consumed; the O columns of this TMEM stage were read out before the
        // group started on problem i, and P is not rewritten until the group has seen o_full(i)
        mbar_wait(&p_full[sj], (i >> 1) & 1);
        tc_fence_after();
        const uint32_t pa = smem_u32(p_base + sj * kTcPBytes);
        const uint32_t va = smem_u32(smem + r * kTcStageBytes + 2 * kTcTileBytes);
        const uint32_t d_o = tmem_base + sj * 256 + 128;
        for (int kk = 0; kk < ksteps; ++kk) {
          const uint64_t da = make_smem_desc_sw128(pa + (kk >> 2) * kTcTileBytes + (kk & 3) * 32, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(va + kk * 2048, 8192, 1024);
          if (im.issue) umma_bf16(d_o, da, db, idesc_o, kk > 0 ? 1u : 0u);
        }
        if (im.issue) umma_commit(&o_full[sj]);
        if (im.issue) umma_commit(&kv_empty[r]);
        im.sync();
      };
      if constexpr (PACKED) {
        for (int i = 0; i < n_local; ++i) {
          issue_s(i);
          if (i > 0) issue_pv(i - 1);
        }
        if (n_local > 0) issue_pv(n_local - 1);
      } else {
        if (n_local > 0) issue_s(0);
        if (n_local > 1) issue_s(1);
        for (int i = 0; i < n_local; ++i) {
          issue_pv(i);
          if (i + 2 < n_local) issue_s(i + 2);   // scores of the group's next problem, behind PV(i)
        }
      }
    }
  } else {
    const int grp = (warp - 2) >> 2;  // 0: even problems, 1: odd problems
    const int q = warp & 3;           // TMEM lane quarter
    const int row = q * 32 + lane;    // query index handled by this thread
    const bool warp_has_rows = q * 32 < p.GL;  // warp-uniform: all 32 rows beyond the tile's rows -> idle
    const RowSpan span = row_span<CAUSAL, PACKED>(row, L, p.GL);
    // 32-column chunks this WARP has to look at: from its first row's block to its last row's block
    const int c_first = PACKED ? (min(q * 32, p.GL - 1) / L * L) >> 5 : 0;
    const int c_last = PACKED ? (min(p.GL, (min(q * 32 + 31, p.GL - 1) / L + 1) * L) - 1) >> 5 : NCH - 1;
    uint8_t* pbuf = p_base + grp * kTcPBytes;
    const uint32_t t_s = tmem_base + grp * 256 + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t t_o = t_s + 128;
    for (int i = grp; i < n_local; i += 2) {
      const uint32_t ph = (i >> 1) & 1;
      const int prob = blockIdx.x + i * gridDim.x;
      const int n = (prob / H) * p.G, h = prob % H;
      const bool row_valid = row < p.GL && n + span.sample < p.batch;
      mbar_wait(&s_full[grp], ph);
      tc_fence_after();
      float l = 0.f, ms = 0.f;
      if (warp_has_rows) {
        // ---- single TMEM pass: the whole score row (NCH x 32 fp32) is pulled into registers once
        // (TMEM reads are the scarce resource: ~64 B/clk/SM), then max -> exp2 -> sum -> bf16 P.
        uint32_t v[NCH][32];
#pragma unroll
        for (int c = 0; c < NCH; ++c)
          if (c >= c_first && c <= c_last) tmem_ld_32x32(t_s + c * 32, v[c]);
        tmem_ld_wait();
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if (c >= c_first && c <= c_last) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int key = c * 32 + j;
              float x = __uint_as_float(v[c][j]);
              if ((PACKED && key < span.lo) || key >= span.hi) x = -INFINITY;
              v[c][j] = __float_as_uint(x);
              mx[j & 3] = fmaxf(mx[j & 3], x);
            }
          }
        }
        const float m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        ms = (m == -INFINITY) ? 0.f : m * p.scale_log2;
        float sm[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if (c >= c_first && c <= c_last) {   // chunks outside the warp's blocks stay zero in smem forever
            float pr[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              pr[j] = ex2_approx(fmaf(__uint_as_float(v[c][j]), p.scale_log2, -ms));  // exp2(-inf) = 0
              sm[j & 3] += pr[j];
            }
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8)
              *reinterpret_cast<uint4*>(pbuf + p_tile_off(row, c * 4 + g8)) = pack8_bf16(pr + 8 * g8);
          }
        }
        l = (sm[0] + sm[1]) + (sm[2] + sm[3]);
        fence_proxy_async_smem();  // generic-proxy P stores -> visible to the MMA (async proxy)
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[grp]);
      // ---- epilogue: O row / l -> global
      mbar_wait(&o_full[grp], ph);
      tc_fence_after();
      if (warp_has_rows) {
        const float inv = l > 0.f ? 1.f / l : 0.f;
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(t_o, o0);
        tmem_ld_32x32(t_o + 32, o1);
        tmem_ld_wait();
        if constexpr (PACKED) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_free[grp]);
        }
        if (row_valid) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + ((long long)n * L + row) * D + h * kTcHd);
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(o0[j]) * inv;
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) dst[g8] = pack8_bf16(f + 8 * g8);
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(o1[j]) * inv;
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) dst[4 + g8] = pack8_bf16(f + 8 * g8);
          p.lse[((long long)(n + span.sample) * H + h) * L + span.token] = (ms + log2f(l)) * 0.69314718055994531f;
        }
      } else if constexpr (PACKED) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_free[grp]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// host launcher; returns CLIPA_ERR_UNSUPPORTED when the shape is outside this kernel's envelope
int attention_fwd_tc(const void* qkv, void* out, float* lse, int batch, int L, int H, int causal,
                     cudaStream_t stream) {
  CLIPA_REQUIRE(L >= 1 && L <= 128, CLIPA_ERR_UNSUPPORTED, "attention_fwd_tc: L=%d > 128", L);
  const int D = H * kTcHd;
  CUtensorMap tm;
  // 2-D view of the packed qkv matrix: inner = 3D columns, outer = batch*L rows; box = 64 cols x L rows
  const int G = 128 / L;   // samples per tile
  const int GL = G * L;
  int rc = encode_tmap_2d_bf16(&tm, qkv, (uint64_t)3 * D, (uint64_t)batch * L, (uint64_t)3 * D * 2, kTcHd,
                               (uint32_t)GL);
  if (rc) return rc;
  AttnTcParams p;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.L = L; p.H = H; p.batch = batch;
  p.G = G; p.GL = GL;
  p.npad = (GL + 15) & ~15;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)kTcHd);
  long long total = (long long)((batch + G - 1) / G) * H;
  int grid = num_sms();
  if (grid > total) grid = (int)total;
  const int nch = (p.npad + 31) / 32;
  const int smem_bytes = tc_smem_total(G > 1);
  auto launch = [&](auto kern) -> int {
    CLIPA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    kern<<<grid, kTcThreads, smem_bytes, stream>>>(tm, p);
    return CLIPA_OK;
  };
  auto pick = [&](auto causal_c, auto packed_c) -> int {
    constexpr bool C = decltype(causal_c)::value, P = decltype(packed_c)::value;
    switch (nch) {
      case 1: return launch(attn_fwd_tc_kernel<1, C, P>);
      case 2: return launch(attn_fwd_tc_kernel<2, C, P>);
      case 3: return launch(attn_fwd_tc_kernel<3, C, P>);
      default: return launch(attn_fwd_tc_kernel<4, C, P>);
    }
  };
  using T = std::true_type;
  using F = std::false_type;
  const int lrc = causal ? (G > 1 ? pick(T{}, T{}) : pick(T{}, F{})) : (G > 1 ? pick(F{}, T{}) : pick(F{}, F{}));
  if (lrc) return lrc;
  CLIPA_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return CLIPA_OK;
}

// ================================================================================================
// BACKWARD (head_dim 64, L <= 128).  Per (sample, head), all on M = 128 tiles:
//   S  = Q K^T          dP = dO V^T                       (both K-major operands)
//   P  = exp(S*scale - lse),  dS = P o (dP - delta) * scale,  delta_i = sum_d dO_id O_id
//   dV = P^T dO         dK = dS^T Q                       (A = P / dS consumed MN-major = transposed
//   dQ = dS K                                              in place; B = dO / Q / K MN-major)
// P and dS are written once (bf16, 128B-swizzled [query][key] tiles) and serve as the A operand of
// three MMAs through the descriptor "major" bit -- no transposes, no atomics, no global scratch.
// TMEM: S 128 + dP 128 + dQ 64 + dK 64 + dV 64 = 448 of 512 columns (one problem in flight per CTA);
// the Q/K/V/dO/O smem ring is 2 stages deep so TMA of problem i+1 overlaps the math of problem i
// (O rides the ring too: delta is computed from shared memory, no global-load latency per problem).
//   warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = 8 worker warps: two per TMEM lane
//   quarter, each owning one half of the columns of a row for the elementwise stage and epilogue.
// ================================================================================================
constexpr int kBwTiles = 5;                                      // Q, K, V, dO, O
constexpr int kBwStageBytes = kBwTiles * kTcTileBytes;
constexpr int kBwPOff = 2 * kBwStageBytes;                      // P  (2 atoms)
constexpr int kBwDsOff = kBwPOff + kTcPBytes;                   // dS (2 atoms)
constexpr int kBwSmemBar = kBwDsOff + kTcPBytes;
constexpr int kBwSmemTotal = kBwSmemBar + 256 + 1024;
static_assert(kBwSmemTotal <= 227 * 1024, "attention backward shared memory budget");

struct AttnBwParams {
  const float* lse;
  __nv_bfloat16* dqkv;
  int L, H, batch;
  int G, GL;   // samples per tile, rows in use (see AttnTcParams)
  int npad;
  float scale;
};

// 32 fp32 columns of one TMEM row -> 32 bf16 (64 B) at dst + col0
__device__ __forceinline__ void store_row32_bf16(__nv_bfloat16* dst, uint32_t taddr, int col0, bool ok) {
  uint32_t v[32];
  tmem_ld_32x32(taddr + col0, v);
  tmem_ld_wait();
  if (ok) {
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    uint4* d4 = reinterpret_cast<uint4*>(dst + col0);
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) d4[g8] = pack8_bf16(f + 8 * g8);
  }
}

// dQ, dK, dV: 32 fp32 columns of one TMEM row each -> three 64-byte bf16 row segments; the three
// TMEM loads are in flight together (one wait)
__device__ __forceinline__ void store_grad_rows(__nv_bfloat16* drow, int D, uint32_t t_row, uint32_t col_dq,
                                                uint32_t col_dk, uint32_t col_dv, int c0, bool ok) {
  uint32_t a[32], b[32], c[32];
  tmem_ld_32x32(t_row + col_dq + c0, a);
  tmem_ld_32x32(t_row + col_dk + c0, b);
  tmem_ld_32x32(t_row + col_dv + c0, c);
  tmem_ld_wait();
  if (ok) {
    auto put = [&](__nv_bfloat16* dst, const uint32_t (&v)[32]) {
      uint4* d4 = reinterpret_cast<uint4*>(dst + c0);
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        uint4 t;
        t.x = pack_bf16x2(__uint_as_float(v[8 * g8]), __uint_as_float(v[8 * g8 + 1]));
        t.y = pack_bf16x2(__uint_as_float(v[8 * g8 + 2]), __uint_as_float(v[8 * g8 + 3]));
        t.z = pack_bf16x2(__uint_as_float(v[8 * g8 + 4]), __uint_as_float(v[8 * g8 + 5]));
        t.w = pack_bf16x2(__uint_as_float(v[8 * g8 + 6]), __uint_as_float(v[8 * g8 + 7]));
        d4[g8] = t;
      }
    };
    put(drow, a);
    put(drow + D, b);
    put(drow + 2 * D, c);
  }
}

template <bool CAUSAL, bool PACKED>
__global__ void __launch_bounds__(kTcThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                   const __grid_constant__ CUtensorMap tmap_o, const AttnBwParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* p_buf = smem + kBwPOff;
  uint8_t* ds_buf = smem + kBwDsOff;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBwSmemBar);
  uint64_t* full = bars;           // [2] TMA landed
  uint64_t* kv_empty = bars + 2;   // [2] smem stage free
  uint64_t* sdp_full = bars + 4;   // S and dP in TMEM
  uint64_t* pds_full = bars + 5;   // P and dS in smem (8 warp arrivals)
  uint64_t* grad_full = bars + 6;  // dQ, dK, dV in TMEM
  uint64_t* t_free = bars + 7;     // TMEM + P/dS buffers free (8 warp arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L, H = p.H, D = H * kTcHd;
  const int tiles_per_head = (p.batch + p.G - 1) / p.G;
  const int total = tiles_per_head * H;
  const long long pitch = 3LL * D;

  for (int i = threadIdx.x; i < kBwSmemBar / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_o);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 8);
    mbar_init(grad_full, 1);
    mbar_init(t_free, 8);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  constexpr uint32_t kColS = 0, kColDp = 128, kColDq = 256, kColDk = 320, kColDv = 384;

  const int n_local = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < n_local; ++i) {
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        const int prob = blockIdx.x + i * gridDim.x;
        const int n = (prob / H) * p.G, h = prob % H;
        mbar_wait(&kv_empty[s], ph ^ 1);
        uint8_t* st = smem + s * kBwStageBytes;
        mbar_expect_tx(&full[s], (uint32_t)kBwTiles * (uint32_t)p.GL * 128u);
        tma_load_2d(st, &tmap_qkv, &full[s], h * kTcHd, n * L);
        tma_load_2d(st + kTcTileBytes, &tmap_qkv, &full[s], D + h * kTcHd, n * L);
        tma_load_2d(st + 2 * kTcTileBytes, &tmap_qkv, &full[s], 2 * D + h * kTcHd, n * L);
        tma_load_2d(st + 3 * kTcTileBytes, &tmap_do, &full[s], h * kTcHd, n * L);
        tma_load_2d(st + 4 * kTcTileBytes, &tmap_o, &full[s], h * kTcHd, n * L);
      }
    }
  } else if (warp == 1) {
    const IssueMode im = issue_mode(lane);
    if (im.in_loop) {
      const uint32_t idesc_s = make_idesc_bf16(128, (uint32_t)p.npad, false, false);   // A K-major, B K-major
      const uint32_t idesc_t = make_idesc_bf16(128, kTcHd, true, true);                // A MN (P^T/dS^T), B MN
      const uint32_t idesc_q = make_idesc_bf16(128, kTcHd, false, true);               // A K (dS), B MN (K)
      const int ksteps = p.npad / 16;
      // S and dP of problem i+1 are issued right behind the gradient MMAs of problem i: their TMEM
      // columns are free as soon as the workers have turned S/dP(i) into P/dS(i) (pds_full), so the
      // workers find the next scores ready when they come back from storing dQ/dK/dV(i).
      auto issue_scores = [&](int i) {
        const int s = i & 1;
        mbar_wait(&full[s], (i >> 1) & 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(smem + s * kBwStageBytes);
        const uint32_t ka = qa + kTcTileBytes, va = qa + 2 * kTcTileBytes, doa = qa + 3 * kTcTileBytes;
#pragma unroll
        for (int k = 0; k < kTcHd / 16; ++k)
          if (im.issue) umma_bf16(tmem_base + kColS, make_smem_desc_sw128(qa + k * 32, 16, 1024),
                    make_smem_desc_sw128(ka + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < kTcHd / 16; ++k)
          if (im.issue) umma_bf16(tmem_base + kColDp, make_smem_desc_sw128(doa + k * 32, 16, 1024),
                    make_smem_desc_sw128(va + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        if (im.issue) umma_commit(sdp_full);
        im.sync();
      };
      if (n_local > 0) issue_scores(0);
      for (int i = 0; i < n_local; ++i) {
        const int s = i & 1;
        const uint32_t pi = i & 1;  // phase of the single-stage barriers
        const uint32_t qa = smem_u32(smem + s * kBwStageBytes);
        const uint32_t ka = qa + kTcTileBytes, doa = qa + 3 * kTcTileBytes;
        mbar_wait(pds_full, pi);      // P/dS(i) in smem; S/dP(i) consumed
        mbar_wait(t_free, pi ^ 1);    // dQ/dK/dV(i-1) read out of TMEM
        tc_fence_after();
        const uint32_t pa = smem_u32(p_buf), dsa = smem_u32(ds_buf);
        for (int kk = 0; kk < ksteps; ++kk) {
          // contraction over queries: A = P^T / dS^T (MN-major view of the [query][key] tiles:
          // key atoms 16 KB apart = LBO, 8-query groups 1 KB apart = SBO), B = dO / Q MN-major
          const uint64_t a_p = make_smem_desc_sw128(pa + kk * 2048, kTcTileBytes, 1024);
          const uint64_t a_ds = make_smem_desc_sw128(dsa + kk * 2048, kTcTileBytes, 1024);
          const uint64_t b_do = make_smem_desc_sw128(doa + kk * 2048, 8192, 1024);
          const uint64_t b_q = make_smem_desc_sw128(qa + kk * 2048, 8192, 1024);
          if (im.issue) umma_bf16(tmem_base + kColDv, a_p, b_do, idesc_t, kk > 0 ? 1u : 0u);
          if (im.issue) umma_bf16(tmem_base + kColDk, a_ds, b_q, idesc_t, kk > 0 ? 1u : 0u);
        }
        for (int kk = 0; kk < ksteps; ++kk) {
          // contraction over keys: A = dS K-major, B = K MN-major
          const uint64_t a_ds = make_smem_desc_sw128(dsa + (kk >> 2) * kTcTileBytes + (kk & 3) * 32, 16, 1024);
          const uint64_t b_k = make_smem_desc_sw128(ka + kk * 2048, 8192, 1024);
          if (im.issue) umma_bf16(tmem_base + kColDq, a_ds, b_k, idesc_q, kk > 0 ? 1u : 0u);
        }
        if (im.issue) umma_commit(grad_full);
        if (im.issue) umma_commit(&kv_empty[s]);
        im.sync();
        if (i + 1 < n_local) issue_scores(i + 1);
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;     // which half of the columns
    const int row = q * 32 + lane;        // query index (elementwise stage) / key index (dK, dV rows)
    const bool row_in_tile = row < p.GL;
    const RowSpan span = row_span<CAUSAL, PACKED>(row, L, p.GL);
    const bool warp_writes = q * 32 < p.npad;  // rows >= npad are never contracted: skip their P/dS
    const bool warp_stores = q * 32 < p.GL;    // rows beyond the tile's rows have no gradient rows to store
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float scale_log2 = p.scale * 1.4426950408889634f;
    // The two warps of a lane quarter split the key columns [0, npad) evenly at 16-column
    // granularity (npad = 96: 48 + 48); of its share a warp visits only the 16-column chunks its
    // rows' blocks overlap (block-diagonal tile when samples are packed).
    const int blk_lo = PACKED ? min(q * 32, p.GL - 1) / L * L : 0;
    const int blk_hi = PACKED ? min(p.GL, (min(q * 32 + 31, p.GL - 1) / L + 1) * L) : L;
    const int hsplit = ((p.npad + 31) >> 5) << 4;
    const int c_begin = max(half * hsplit, blk_lo & ~15);
    const int c_end = min(min(p.npad, half * hsplit + hsplit), (blk_hi + 15) & ~15);
    auto lse_index = [&](int prob) -> long long {   // lse is [batch, H, L]
      const int n = (prob / H) * p.G + span.sample, h = prob % H;
      return n < p.batch ? ((long long)n * H + h) * L + span.token : -1;
    };
    // the next problem's lse (one float per thread) is loaded a whole problem ahead and only
    // touched (scaled) where it is used, so the global-load latency never sits on the critical path
    float lse_raw = 0.f;
    if (n_local > 0 && row_in_tile) {
      const long long li = lse_index(blockIdx.x);
      if (li >= 0) lse_raw = p.lse[li];
    }
    for (int i = 0; i < n_local; ++i) {
      const int s = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      const uint32_t pi = i & 1;
      const int prob = blockIdx.x + i * gridDim.x;
      const int n = (prob / H) * p.G, h = prob % H;
      const bool row_ok = row_in_tile && n + span.sample < p.batch;
      // delta_i = sum_d dO_id * O_id from the swizzled smem tiles (TMA already landed them)
      mbar_wait(&full[s], ph);
      float delta = 0.f;
      if (row_ok && c_begin < c_end) {
        const uint8_t* dot = smem + s * kBwStageBytes + 3 * kTcTileBytes + row * 128;
        const uint8_t* ot = dot + kTcTileBytes;
        float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t off = ((c ^ (row & 7)) << 4);
          const uint4 a = *reinterpret_cast<const uint4*>(ot + off);
          const uint4 b = *reinterpret_cast<const uint4*>(dot + off);
          d4[0] = fmaf(bf16lo(a.x), bf16lo(b.x), fmaf(bf16hi(a.x), bf16hi(b.x), d4[0]));
          d4[1] = fmaf(bf16lo(a.y), bf16lo(b.y), fmaf(bf16hi(a.y), bf16hi(b.y), d4[1]));
          d4[2] = fmaf(bf16lo(a.z), bf16lo(b.z), fmaf(bf16hi(a.z), bf16hi(b.z), d4[2]));
          d4[3] = fmaf(bf16lo(a.w), bf16lo(b.w), fmaf(bf16hi(a.w), bf16hi(b.w), d4[3]));
        }
        delta = (d4[0] + d4[1]) + (d4[2] + d4[3]);
      }
      mbar_wait(sdp_full, pi);
      tc_fence_after();
      const float lse2 = lse_raw * 1.4426950408889634f;
      const float nds = -delta * p.scale;
      const int mlo = PACKED ? span.lo : 0;
      if (i + 1 < n_local && row_in_tile) {
        const long long li = lse_index(prob + (int)gridDim.x);
        lse_raw = li >= 0 ? p.lse[li] : 0.f;
      }
      if (warp_writes) {
        for (int c = c_begin; c < c_end; c += 16) {
          float pr[16], ds[16];
          const int kind = attn_chunk_kind(row_ok, c, mlo, span.hi);
          if (kind != 2) {
            uint32_t sv[16], dv[16];
            tmem_ld_32x16(t_row + kColS + c, sv);
            tmem_ld_32x16(t_row + kColDp + c, dv);
            tmem_ld_wait();
            if (kind == 0)
              attn_bwd_chunk16<false>(sv, dv, scale_log2, lse2, p.scale, nds, c, mlo, span.hi, row_ok, pr, ds);
            else
              attn_bwd_chunk16<true>(sv, dv, scale_log2, lse2, p.scale, nds, c, mlo, span.hi, row_ok, pr, ds);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) pr[j] = ds[j] = 0.f;
          }
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            const uint32_t off = p_tile_off(row, (c >> 3) + g8);
            *reinterpret_cast<uint4*>(p_buf + off) = pack8_bf16(pr + 8 * g8);
            *reinterpret_cast<uint4*>(ds_buf + off) = pack8_bf16(ds + 8 * g8);
          }
        }
        fence_proxy_async_smem();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      // ---- epilogue: this warp stores 32 of the 64 columns of each gradient row
      mbar_wait(grad_full, pi);
      tc_fence_after();
      if (warp_stores) {
        __nv_bfloat16* drow = p.dqkv + ((long long)n * L + row) * pitch + h * kTcHd;
        const int c0 = half * 32;
        store_grad_rows(drow, D, t_row, kColDq, kColDk, kColDv, c0, row_ok);   // rows of dQ, dK (row = key), dV
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_free);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ================================================================================================
// BACKWARD, software-pipelined variant for 64 < L <= 96 (one sample per tile; ViT towers at 65 / 82
// tokens).  Same math and warp roles as attn_bwd_tc_kernel, but the tiles are 96 rows tall
// (12 KB), which leaves room for TWO P/dS buffers next to the 2-stage operand ring, and the
// worker loop is skewed by one problem:
//     workers:  ... | P/dS(i+1) from S/dP(i+1) | store dQ/dK/dV(i) | P/dS(i+2) | store (i+1) | ...
//     tensor :  ... | S,dP(i+2) | dV,dK,dQ(i+1) | S,dP(i+3) | dV,dK,dQ(i+2) | ...
// so the gradient MMAs of problem i run while the workers are busy with the scores of problem
// i+1, and the scores of problem i+2 are in TMEM before the workers ask for them.  Neither the MMA
// latency nor the barrier round trips sit on the workers' critical path any more (in the
// unpipelined kernel the workers spent 28 % of their time waiting for the gradient MMAs and 9 % for
// the scores; ncu source view, profiles/).
// The M = 128 MMAs read 32 rows past the 96-row A tiles; those accumulator rows (TMEM lanes
// 96..127) are never read.
// ================================================================================================
constexpr int kPRows = 96;
constexpr int kPTile = kPRows * 128;                              // 12 KB
constexpr int kPStage = kBwTiles * kPTile;                        // Q, K, V, dO, O
constexpr int kPPds = 2 * kPTile;                                 // one P (or dS) buffer: 2 key atoms
// Layout: [P0 dS0 P1 dS1 | stage 0 | stage 1 | barriers].  The P/dS buffers come first so that the
// 32-row over-read of the last dS atom lands in the operand ring, not past the allocation.
constexpr int kPPOff = 0;                                         // P/dS buffer b at b * 2 * kPPds
constexpr int kPStageOff = 4 * kPPds;
constexpr int kPSmemBar = kPStageOff + 2 * kPStage;
constexpr int kPSmemTotal = kPSmemBar + 256 + 1024;
static_assert(kPSmemTotal <= 227 * 1024, "pipelined attention backward shared memory budget");

__device__ __forceinline__ uint32_t p96_tile_off(int row, int chunk) {
  return (chunk >> 3) * kPTile + row * 128 + (((chunk & 7) ^ (row & 7)) << 4);
}

template <bool CAUSAL>
__global__ void __launch_bounds__(kTcThreads, 1)
attn_bwd_tc_pipe_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                        const __grid_constant__ CUtensorMap tmap_o, const AttnBwParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPSmemBar);
  uint64_t* full = bars;           // [2] TMA landed
  uint64_t* kv_empty = bars + 2;   // [2] smem stage free
  uint64_t* sdp_full = bars + 4;   // S and dP in TMEM
  uint64_t* pds_full = bars + 5;   // [2] P and dS buffer b written (8 warp arrivals)
  uint64_t* grad_full = bars + 7;  // dQ, dK, dV in TMEM
  uint64_t* t_free = bars + 8;     // dQ/dK/dV read out of TMEM (8 warp arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L, H = p.H, D = H * kTcHd;
  const int total = p.batch * H;
  const long long pitch = 3LL * D;

  for (int i = threadIdx.x; i < kPSmemBar / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_o);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&pds_full[i], 8);
    }
    mbar_init(sdp_full, 1);
    mbar_init(grad_full, 1);
    mbar_init(t_free, 8);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  constexpr uint32_t kColS = 0, kColDp = 128, kColDq = 256, kColDk = 320, kColDv = 384;

  const int n_local = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < n_local; ++i) {
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        const int prob = blockIdx.x + i * gridDim.x;
        const int n = prob / H, h = prob - n * H;
        mbar_wait(&kv_empty[s], ph ^ 1);
        uint8_t* st = smem + kPStageOff + s * kPStage;
        mbar_expect_tx(&full[s], (uint32_t)kBwTiles * (uint32_t)L * 128u);
        tma_load_2d(st, &tmap_qkv, &full[s], h * kTcHd, n * L);
        tma_load_2d(st + kPTile, &tmap_qkv, &full[s], D + h * kTcHd, n * L);
        tma_load_2d(st + 2 * kPTile, &tmap_qkv, &full[s], 2 * D + h * kTcHd, n * L);
        tma_load_2d(st + 3 * kPTile, &tmap_do, &full[s], h * kTcHd, n * L);
        tma_load_2d(st + 4 * kPTile, &tmap_o, &full[s], h * kTcHd, n * L);
      }
    }
  } else if (warp == 1) {
    const IssueMode im = issue_mode(lane);
    if (im.in_loop) {
      const uint32_t idesc_s = make_idesc_bf16(128, (uint32_t)p.npad, false, false);   // A K-major, B K-major
      const uint32_t idesc_t = make_idesc_bf16(128, kTcHd, true, true);                // A MN (P^T/dS^T), B MN
      const uint32_t idesc_q = make_idesc_bf16(128, kTcHd, false, true);               // A K (dS), B MN (K)
      const int ksteps = p.npad / 16;
      auto issue_scores = [&](int i) {
        const int s = i & 1;
        mbar_wait(&full[s], (i >> 1) & 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(smem + kPStageOff + s * kPStage);
        const uint32_t ka = qa + kPTile, va = qa + 2 * kPTile, doa = qa + 3 * kPTile;
#pragma unroll
        for (int k = 0; k < kTcHd / 16; ++k)
          if (im.issue) umma_bf16(tmem_base + kColS, make_smem_desc_sw128(qa + k * 32, 16, 1024),
                    make_smem_desc_sw128(ka + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < kTcHd / 16; ++k)
          if (im.issue) umma_bf16(tmem_base + kColDp, make_smem_desc_sw128(doa + k * 32, 16, 1024),
                    make_smem_desc_sw128(va + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        if (im.issue) umma_commit(sdp_full);
        im.sync();
      };
      if (n_local > 0) issue_scores(0);
      for (int i = 0; i < n_local; ++i) {
        const int s = i & 1;
        mbar_wait(&pds_full[s], (i >> 1) & 1);   // P/dS(i) in buffer s; S/dP(i) consumed
        tc_fence_after();
        if (i + 1 < n_local) issue_scores(i + 1);
        mbar_wait(t_free, (i & 1) ^ 1);          // dQ/dK/dV(i-1) read out of TMEM
        tc_fence_after();
        const uint32_t qa = smem_u32(smem + kPStageOff + s * kPStage);
        const uint32_t ka = qa + kPTile, doa = qa + 3 * kPTile;
        const uint32_t pa = smem_u32(smem + kPPOff + s * 2 * kPPds), dsa = pa + kPPds;
        for (int kk = 0; kk < ksteps; ++kk) {
          // contraction over queries: A = P^T / dS^T (MN-major view of the [query][key] tiles: key
          // atoms one tile apart = LBO, 8-query groups 1 KB apart = SBO), B = dO / Q MN-major
          const uint64_t a_p = make_smem_desc_sw128(pa + kk * 2048, kPTile, 1024);
          const uint64_t a_ds = make_smem_desc_sw128(dsa + kk * 2048, kPTile, 1024);
          const uint64_t b_do = make_smem_desc_sw128(doa + kk * 2048, 8192, 1024);
          const uint64_t b_q = make_smem_desc_sw128(qa + kk * 2048, 8192, 1024);
          if (im.issue) umma_bf16(tmem_base + kColDv, a_p, b_do, idesc_t, kk > 0 ? 1u : 0u);
          if (im.issue) umma_bf16(tmem_base + kColDk, a_ds, b_q, idesc_t, kk > 0 ? 1u : 0u);
        }
        for (int kk = 0; kk < ksteps; ++kk) {
          // contraction over keys: A = dS K-major, B = K MN-major
          const uint64_t a_ds = make_smem_desc_sw128(dsa + (kk >> 2) * kPTile + (kk & 3) * 32, 16, 1024);
          const uint64_t b_k = make_smem_desc_sw128(ka + kk * 2048, 8192, 1024);
          if (im.issue) umma_bf16(tmem_base + kColDq, a_ds, b_k, idesc_q, kk > 0 ? 1u : 0u);
        }
        if (im.issue) umma_commit(grad_full);
        if (im.issue) umma_commit(&kv_empty[s]);
        im.sync();
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;     // which half of the columns
    const int row = q * 32 + lane;        // query index (scores stage) / key index (dK, dV rows)
    const bool row_ok = row < L;
    const bool warp_writes = q * 32 < p.npad;  // rows >= npad are never contracted: skip their P/dS
    const bool warp_stores = q * 32 < L;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float scale_log2 = p.scale * 1.4426950408889634f;
    const int hsplit = ((p.npad + 31) >> 5) << 4;          // 96 keys: 48 + 48
    const int c_begin = half * hsplit;
    const int c_end = min(p.npad, c_begin + hsplit);
    const int hi = CAUSAL ? row + 1 : L;                   // valid keys [0, hi)

    float lse_raw = 0.f;
    if (n_local > 0 && row_ok) lse_raw = p.lse[(long long)blockIdx.x * L + row];   // lse is [batch, H, L]

    // scores stage of problem j: S/dP (TMEM) -> P/dS (bf16, swizzled smem buffer j & 1)
    auto scores_stage = [&](int j) {
      const int s = j & 1;
      const int prob = blockIdx.x + j * gridDim.x;
      mbar_wait(&full[s], (j >> 1) & 1);
      float delta = 0.f;
      if (row_ok && c_begin < c_end) {   // delta = sum_d dO * O from the swizzled smem tiles
        const uint8_t* dot = smem + kPStageOff + s * kPStage + 3 * kPTile + row * 128;
        const uint8_t* ot = dot + kPTile;
        float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t off = ((c ^ (row & 7)) << 4);
          const uint4 a = *reinterpret_cast<const uint4*>(ot + off);
          const uint4 b = *reinterpret_cast<const uint4*>(dot + off);
          d4[0] = fmaf(bf16lo(a.x), bf16lo(b.x), fmaf(bf16hi(a.x), bf16hi(b.x), d4[0]));
          d4[1] = fmaf(bf16lo(a.y), bf16lo(b.y), fmaf(bf16hi(a.y), bf16hi(b.y), d4[1]));
          d4[2] = fmaf(bf16lo(a.z), bf16lo(b.z), fmaf(bf16hi(a.z), bf16hi(b.z), d4[2]));
          d4[3] = fmaf(bf16lo(a.w), bf16lo(b.w), fmaf(bf16hi(a.w), bf16hi(b.w), d4[3]));
        }
        delta = (d4[0] + d4[1]) + (d4[2] + d4[3]);
      }
      mbar_wait(sdp_full, j & 1);
      tc_fence_after();
      const float lse2 = lse_raw * 1.4426950408889634f;
      const float nds = -delta * p.scale;
      if (j + 1 < n_local && row_ok)   // next problem's lse: consumed one whole stage later
        lse_raw = p.lse[(long long)(prob + (int)gridDim.x) * L + row];
      if (warp_writes) {
        uint8_t* p_buf = smem + kPPOff + s * 2 * kPPds;
        uint8_t* ds_buf = p_buf + kPPds;
        for (int c = c_begin; c < c_end; c += 16) {
          float pr[16], ds[16];
          const int kind = attn_chunk_kind(row_ok, c, 0, hi);
          if (kind != 2) {
            uint32_t sv[16], dv[16];
            tmem_ld_32x16(t_row + kColS + c, sv);
            tmem_ld_32x16(t_row + kColDp + c, dv);
            tmem_ld_wait();
            if (kind == 0)
              attn_bwd_chunk16<false>(sv, dv, scale_log2, lse2, p.scale, nds, c, 0, hi, row_ok, pr, ds);
            else
              attn_bwd_chunk16<true>(sv, dv, scale_log2, lse2, p.scale, nds, c, 0, hi, row_ok, pr, ds);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) pr[j] = ds[j] = 0.f;
          }
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            const uint32_t off = p96_tile_off(row, (c >> 3) + g8);
            *reinterpret_cast<uint4*>(p_buf + off) = pack8_bf16(pr + 8 * g8);
            *reinterpret_cast<uint4*>(ds_buf + off) = pack8_bf16(ds + 8 * g8);
          }
        }
        fence_proxy_async_smem();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_full[s]);
    };

    if (n_local > 0) scores_stage(0);
    for (int i = 0; i < n_local; ++i) {
      if (i + 1 < n_local) scores_stage(i + 1);
      // ---- gradients of problem i: this warp stores 32 of the 64 columns of each row
      const int prob = blockIdx.x + i * gridDim.x;
      const int n = prob / H, h = prob - n * H;
      mbar_wait(grad_full, i & 1);
      tc_fence_after();
      if (warp_stores) {
        __nv_bfloat16* drow = p.dqkv + ((long long)n * L + row) * pitch + h * kTcHd;
        store_grad_rows(drow, D, t_row, kColDq, kColDk, kColDv, half * 32, row_ok);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_free);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

int attention_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                     int batch, int L, int H, int causal, cudaStream_t stream) {
  CLIPA_REQUIRE(L >= 1 && L <= 128, CLIPA_ERR_UNSUPPORTED, "attention_bwd_tc: L=%d > 128", L);
  const int D = H * kTcHd;
  CUtensorMap tq, td, to;
  if (L > 64 && L <= kPRows) {   // one sample per tile and it fits 96-row tiles: pipelined kernel
    int rc = encode_tmap_2d_bf16(&tq, qkv, (uint64_t)3 * D, (uint64_t)batch * L, (uint64_t)3 * D * 2, kTcHd,
                                 (uint32_t)L);
    if (rc) return rc;
    rc = encode_tmap_2d_bf16(&td, dout, (uint64_t)D, (uint64_t)batch * L, (uint64_t)D * 2, kTcHd, (uint32_t)L);
    if (rc) return rc;
    rc = encode_tmap_2d_bf16(&to, out, (uint64_t)D, (uint64_t)batch * L, (uint64_t)D * 2, kTcHd, (uint32_t)L);
    if (rc) return rc;
    AttnBwParams p;
    p.lse = lse;
    p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
    p.L = L; p.H = H; p.batch = batch;
    p.G = 1; p.GL = L;
    p.npad = (L + 15) & ~15;
    p.scale = 1.0f / sqrtf((float)kTcHd);
    long long total = (long long)batch * H;
    int grid = num_sms();
    if (grid > total) grid = (int)total;
    auto launch = [&](auto kern) -> int {
      CLIPA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmemTotal));
      kern<<<grid, kTcThreads, kPSmemTotal, stream>>>(tq, td, to, p);
      return CLIPA_OK;
    };
    const int lrc = causal ? launch(attn_bwd_tc_pipe_kernel<true>) : launch(attn_bwd_tc_pipe_kernel<false>);
    if (lrc) return lrc;
    CLIPA_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return CLIPA_OK;
  }
  const int G = 128 / L, GL = G * L;
  int rc = encode_tmap_2d_bf16(&tq, qkv, (uint64_t)3 * D, (uint64_t)batch * L, (uint64_t)3 * D * 2, kTcHd,
                               (uint32_t)GL);
  if (rc) return rc;
  rc = encode_tmap_2d_bf16(&td, dout, (uint64_t)D, (uint64_t)batch * L, (uint64_t)D * 2, kTcHd, (uint32_t)GL);
  if (rc) return rc;
  rc = encode_tmap_2d_bf16(&to, out, (uint64_t)D, (uint64_t)batch * L, (uint64_t)D * 2, kTcHd, (uint32_t)GL);
  if (rc) return rc;
  AttnBwParams p;
  p.lse = lse;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  p.L = L; p.H = H; p.batch = batch;
  p.G = G; p.GL = GL;
  p.npad = (GL + 15) & ~15;
  p.scale = 1.0f / sqrtf((float)kTcHd);
  long long total = (long long)((batch + G - 1) / G) * H;
  int grid = num_sms();
  if (grid > total) grid = (int)total;
  auto launch = [&](auto kern) -> int {
    CLIPA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwSmemTotal));
    kern<<<grid, kTcThreads, kBwSmemTotal, stream>>>(tq, td, to, p);
    return CLIPA_OK;
  };
  const int lrc = causal ? (G > 1 ? launch(attn_bwd_tc_kernel<true, true>) : launch(attn_bwd_tc_kernel<true, false>))
                         : (G > 1 ? launch(attn_bwd_tc_kernel<false, true>) : launch(attn_bwd_tc_kernel<false, false>));
  if (lrc) return lrc;
  CLIPA_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return CLIPA_OK;
}

}  // namespace clipa
