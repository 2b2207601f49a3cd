// tcgen05 attention for every sequence length and head_dim 64 / 80: the shapes attention_tc.cu's
// one-tile kernels do not take (L > 128: the 224-px fine-tune stage with 257 tokens, 336 px with 577;
// head_dim 80: every ViT-H/14 tower).  Reference call site: nn.MultiheadAttention inside
// ResidualAttentionBlock.attention, open_clip/transformer.py:209,234-236.
//
// A sequence is cut into T = ceil(L / R) tiles of Rt = ceil(L / T) rows, the same cut for queries and keys
// (R = 128 forward, 96 backward: 257 -> 3 x 86; 577 -> 5 x 116 / 7 x 83).  Short sequences (2 L <= R) are PACKED:
// G = R / L consecutive samples of a head share a tile and the score tile is block-diagonal (ViT-H/14 at 37
// tokens: 3 samples per forward tile, 2 per backward tile).  head_dim 80 = one 128B-swizzled operand tile of 64 columns plus a
// 32B-swizzled tile of 16 columns: K-major MMAs take the fifth k-step from the small tile, MN-major
// operands (V, and in backward dO / Q / K) add one N = 16 MMA next to the N = 64 one.
//
// FORWARD (flash, online softmax) -- persistent CTA, work item = (sample, head, query tile):
//   warp 0      TMA: Q of the item -> one of 2 Q buffers; K_j, V_j -> 2-slot rings
//   warp 1      MMA: S = Q K_j^T (M128 x N=npad, two TMEM buffers), O_j = P V_j (fresh accumulator per key tile)
//   warps 2-9   two threads per query row (one per half of the key columns / output columns): S half-row ->
//               registers (once), row max exchanged through smem, running max / sum, P (bf16) -> swizzled
//               smem; then the O_j half-row from TMEM: o = o * alpha + O_j in REGISTERS (no TMEM
//               read-modify-write, no correction warps); after the last key tile o / l -> HBM.
// HBM traffic per (sample, head): Q, O once; K, V once per query tile (T times, L2 hits).
#include <cstdlib>
#include <type_traits>

#include "attn_common.cuh"
#include "host_common.h"
#include "ptx.cuh"

// experiment switches of the forward softmax workers (A/B builds: CLIPA_B200_NVCC_FLAGS="-DCLIPA_FLASH_MASKMODE=0")
#ifndef CLIPA_FLASH_MASKMODE
#define CLIPA_FLASH_MASKMODE 2     // 0: per-element key test on every chunk of a masked warp; 1: warp-uniform chunk kinds (votes); 2: uniform, plain tiles only
#endif
#ifndef CLIPA_FLASH_MATHMODE
#define CLIPA_FLASH_MATHMODE 1     // 0: scalar FFMA / FADD; 1: packed fp32 (FFMA2 / FADD2)
#endif

namespace clipa {

// Warp roles.  CLIPA_FLASH_ROLES = 0: warp 0 = TMA, warp 1 = MMA issuer, warps 2-9 = workers (TMEM lane quarter
// = warp & 3): the two producer warps share SM sub-partitions 0 and 1 with the quarter-0 / quarter-1 workers.
// = 1: 12 warps, workers 0-7, TMA = warp 10, MMA = warp 11 (sub-partitions 2 and 3: with 86-row tiles quarter 3 has
// no rows, so the issuer gets a scheduler of its own); warps 8, 9 exit at once.
#ifndef CLIPA_FLASH_ROLES
#define CLIPA_FLASH_ROLES 0
#endif
#if CLIPA_FLASH_ROLES == 1
constexpr int kFlThreads = 384;
constexpr int kFlTmaWarp = 10, kFlMmaWarp = 11;
__device__ __forceinline__ bool fl_is_worker(int warp) { return warp < 8; }
__device__ __forceinline__ int fl_half(int warp) { return warp >> 2; }
#else
constexpr int kFlThreads = 320;
constexpr int kFlTmaWarp = 0, kFlMmaWarp = 1;
__device__ __forceinline__ bool fl_is_worker(int warp) { return warp >= 2; }
__device__ __forceinline__ int fl_half(int warp) { return (warp - 2) >> 2; }
#endif
constexpr int kFlMaxRowsFwd = 128;  // forward: rows (queries / keys) per tile (full 128-row operand tiles)
constexpr int kFlMaxRowsBwd = 96;   // backward: 96-row operand tiles (shared-memory budget)

template <int HD>
struct FlashTile {
  static_assert(HD == 64 || HD == 80, "head_dim 64 or 80");
  static constexpr int kRem = HD - 64;                      // columns of the 32B-swizzled remainder tile
  static constexpr int kMainBytes = kTcTileBytes;           // 128 rows x 128 B (128B swizzle)
  static constexpr int kRemBytes = kRem ? 128 * 32 : 0;     // 128 rows x 32 B  (32B swizzle)
  static constexpr int kOpBytes = kMainBytes + kRemBytes;   // one operand (Q, K, V, dO or O) of one tile
  static constexpr uint32_t kRowBytes = 2 * HD;             // bytes TMA delivers per row
};

struct FlashParams {
  __nv_bfloat16* out;
  float* lse;
  int L, H, batch;
  int T;       // tiles per sequence
  int Rt;      // rows per tile (G * L when several short sequences are packed into one tile)
  int npad;    // ceil16(Rt)
  int G;       // samples packed per tile (1 unless T == 1 and 2 * L <= tile rows): block-diagonal scores
  float scale_log2;
};

template <int HD>
struct FlashFwdSmem {
  using Tl = FlashTile<HD>;
  static constexpr int kQOff = 0;                        // 2 Q buffers (item parity)
  static constexpr int kKOff = 2 * Tl::kOpBytes;         // 2 K slots (step parity)
  static constexpr int kVOff = 4 * Tl::kOpBytes;         // 2 V slots (step parity)
  static constexpr int kPOff = 6 * Tl::kOpBytes;         // 2 P buffers (step parity)
  static constexpr int kBarOff = kPOff + 2 * kTcPBytes;
  static constexpr int kXOff = kBarOff + 256;            // row-max / row-sum exchange between the two column halves
  static constexpr int kXBytes = (2 * 2 * 128 + 2 * 128) * 4;
  static constexpr int kTotal = kXOff + kXBytes + 1024;
  static_assert(kTotal <= 227 * 1024, "flash attention forward shared memory budget");
};

// loads the [rows x HD] slice of one operand (main 64 columns + remainder) into an operand tile
template <int HD>
__device__ __forceinline__ void flash_load_op(uint8_t* dst, const CUtensorMap* tm_main, const CUtensorMap* tm_rem,
                                              uint64_t* bar, int col, int row) {
  tma_load_2d(dst, tm_main, bar, col, row);
  if constexpr (FlashTile<HD>::kRem != 0) tma_load_2d(dst + FlashTile<HD>::kMainBytes, tm_rem, bar, col + 64, row);
}

// S (+)= A B^T over head_dim, both operands K-major operand tiles
template <int HD>
__device__ __forceinline__ void flash_mma_kmajor(uint32_t d, uint32_t a, uint32_t b, uint32_t idesc, bool issue) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (issue) umma_bf16(d, make_smem_desc_sw128(a + k * 32, 16, 1024), make_smem_desc_sw128(b + k * 32, 16, 1024),
                         idesc, k > 0 ? 1u : 0u);
  if constexpr (FlashTile<HD>::kRem != 0)
    if (issue) umma_bf16(d, make_smem_desc_sw32(a + FlashTile<HD>::kMainBytes, 16, 256),
                         make_smem_desc_sw32(b + FlashTile<HD>::kMainBytes, 16, 256), idesc, 1u);
}

// D[128 x HD] (+)= A[128 x 16*ksteps] (K-major, 128B-swizzled [row][key] tile of 2 atoms) * B (operand tile
// consumed MN-major: rows = contraction index).  `idesc64` / `idesc16`: N = 64 / N = 16 descriptors.
template <int HD>
__device__ __forceinline__ void flash_mma_a_k_b_mn(uint32_t d, uint32_t a_tile, uint32_t b_op, int ksteps, uint32_t idesc64,
                                                   uint32_t idesc16, bool accumulate, bool issue) {
  for (int kk = 0; kk < ksteps; ++kk) {
    const uint64_t da = make_smem_desc_sw128(a_tile + (kk >> 2) * kTcTileBytes + (kk & 3) * 32, 16, 1024);
    const uint32_t acc = (accumulate || kk > 0) ? 1u : 0u;
    if (issue) umma_bf16(d, da, make_smem_desc_sw128(b_op + kk * 2048, 8192, 1024), idesc64, acc);
    if constexpr (FlashTile<HD>::kRem != 0)
      if (issue) umma_bf16(d + 64, da, make_smem_desc_sw32(b_op + FlashTile<HD>::kMainBytes + kk * 512, 256, 256),
                           idesc16, acc);
  }
}

// (item, key tile) position of a step; advance() returns true when it moved on to the next item
struct StepCursor {
  int it, j, T;
  int n, h, i;     // decoded item: sample, head, query tile
  __device__ __forceinline__ void init(int tiles) { it = 0; j = 0; T = tiles; n = h = i = 0; }
  __device__ __forceinline__ bool advance() {
    if (++j == T) { j = 0; ++it; return true; }
    return false;
  }
};

// named barrier shared by the two warps (column halves) of one TMEM lane quarter
__device__ __forceinline__ void pair_sync(int q) {
  asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
}

// Schedule (k = running step index over all (item, key tile) pairs of this CTA; S, P and O_j are double
// buffered by step parity -- with a single O_j buffer P V(k+1) had to wait for the workers to read O_j(k)
// and sat on their critical path: 20 % of the worker samples waited on o_full, ncu source view r02):
//   MMA    : S(0) S(1) | PV(0) S(2) | PV(1) S(3) | ...        PV(k) after p_full(k) and o_read(k-2)
//   workers: softmax(0) | softmax(1) O(0) | softmax(2) O(1) | ...
// so the score tile of step k+1 is already in TMEM when the workers finish step k, and P V of step k
// runs under the softmax of step k+1.  K_j rides 2 steps ahead of V_j in the TMA stream (its slot is
// free as soon as S has been computed).
// NCH = 16-column score chunks per thread (3: tiles up to 96 keys, 4: up to 128)
template <int HD, bool CAUSAL, int NCH>
__global__ void __launch_bounds__(kFlThreads, 1)
attn_fwd_flash_kernel(const __grid_constant__ CUtensorMap tmap_main, const __grid_constant__ CUtensorMap tmap_rem,
                      const FlashParams p) {
  using Tl = FlashTile<HD>;
  using Sm = FlashFwdSmem<HD>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Sm::kBarOff);
  uint64_t* k_full = bars;             // [2]
  uint64_t* k_empty = bars + 2;        // [2] S(k) done
  uint64_t* v_full = bars + 4;         // [2]
  uint64_t* v_empty = bars + 6;        // [2] PV(k) done
  uint64_t* q_full = bars + 8;         // [2]
  uint64_t* q_empty = bars + 10;       // [2] last S of the item done
  uint64_t* s_full = bars + 12;        // [2] S in TMEM
  uint64_t* p_full = bars + 14;        // [2] P in smem, S consumed (8 warp arrivals)
  uint64_t* o_full = bars + 16;        // [2] O_j in TMEM
  uint64_t* o_read = bars + 18;        // [2] O_j read out (8 warp arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 20);
  float* xmax = reinterpret_cast<float*>(smem + Sm::kXOff);   // [step parity][half][row]
  float* xsum = xmax + 2 * 2 * 128;                            // [half][row]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L, H = p.H, D = H * HD, T = p.T, Rt = p.Rt;
  const int G = p.G;
  const int total = ((p.batch + G - 1) / G) * H * T;

  // rows [Rt, 128) of the operand tiles are never written by TMA and are multiplied by exactly-zero
  // probabilities: they must not hold NaN/Inf patterns
  for (int i = threadIdx.x; i < Sm::kBarOff / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_main);
    if constexpr (Tl::kRem != 0) tma_prefetch_desc(&tmap_rem);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 8);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_read[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  constexpr uint32_t kColO = 256;      // S buffers at columns 0 and 128, O_j buffers at 256 and 384

  const int n_local = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int K = n_local * T;           // steps of this CTA
  // local item `it` -> (first sample n, head h, query tile i); the T query tiles of one (n, h) are consecutive
  // work items, i.e. run on neighbouring CTAs at the same time (their K/V re-reads hit L2)
  auto decode = [&](int it, int& n, int& h, int& i) {
    const int w = blockIdx.x + it * gridDim.x;
    i = w % T;
    const int nh = w / T;
    h = nh % H;
    n = (nh / H) * G;
  };

  if (warp == kFlTmaWarp) {
    if (lane == 0) {
      // step cursors (item, key tile) advance incrementally: no integer division per step
      StepCursor ck, cv;
      ck.init(T); cv.init(T);
      if (K > 0) decode(0, ck.n, ck.h, ck.i);
      cv.n = ck.n; cv.h = ck.h; cv.i = ck.i;
      auto load_k = [&](int k) {      // k == ck's step
        if (ck.j == 0) {
          const int b = ck.it & 1;
          mbar_wait(&q_empty[b], ((ck.it >> 1) & 1) ^ 1);
          mbar_expect_tx(&q_full[b], (uint32_t)Rt * Tl::kRowBytes);
          flash_load_op<HD>(smem + Sm::kQOff + b * Tl::kOpBytes, &tmap_main, &tmap_rem, &q_full[b], ck.h * HD,
                            ck.n * L + ck.i * Rt);
        }
        const int s = k & 1;
        mbar_wait(&k_empty[s], ((k >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[s], (uint32_t)Rt * Tl::kRowBytes);
        flash_load_op<HD>(smem + Sm::kKOff + s * Tl::kOpBytes, &tmap_main, &tmap_rem, &k_full[s], D + ck.h * HD,
                          ck.n * L + ck.j * Rt);
        if (ck.advance() && ck.it < n_local) decode(ck.it, ck.n, ck.h, ck.i);
      };
      auto load_v = [&](int k) {      // k == cv's step
        const int s = k & 1;
        mbar_wait(&v_empty[s], ((k >> 1) & 1) ^ 1);
        mbar_expect_tx(&v_full[s], (uint32_t)Rt * Tl::kRowBytes);
        flash_load_op<HD>(smem + Sm::kVOff + s * Tl::kOpBytes, &tmap_main, &tmap_rem, &v_full[s], 2 * D + cv.h * HD,
                          cv.n * L + cv.j * Rt);
        if (cv.advance() && cv.it < n_local) decode(cv.it, cv.n, cv.h, cv.i);
      };
      if (K > 0) load_k(0);
      if (K > 1) load_k(1);
      for (int k = 0; k < K; ++k) {
        if (k + 2 < K) load_k(k + 2);
        load_v(k);
      }
    }
  } else if (warp == kFlMmaWarp) {
    const IssueMode im = issue_mode(lane);
    if (im.in_loop) {
      const uint32_t idesc_s = make_idesc_bf16(128, (uint32_t)p.npad, false, false);
      const uint32_t idesc_o64 = make_idesc_bf16(128, 64, false, true);
      const uint32_t idesc_o16 = make_idesc_bf16(128, 16, false, true);
      const int ksteps = p.npad / 16;
      StepCursor cs;
      cs.init(T);
      auto issue_s = [&](int k) {     // k == cs's step
        const int it = cs.it, j = cs.j, s = k & 1, b = it & 1;
        cs.advance();
        mbar_wait(&k_full[s], (k >> 1) & 1);
        if (j == 0) mbar_wait(&q_full[b], (it >> 1) & 1);
        tc_fence_after();
        flash_mma_kmajor<HD>(tmem_base + s * 128, smem_u32(smem + Sm::kQOff + b * Tl::kOpBytes),
                             smem_u32(smem + Sm::kKOff + s * Tl::kOpBytes), idesc_s, im.issue);
        if (im.issue) umma_commit(&s_full[s]);
        if (im.issue) umma_commit(&k_empty[s]);
        if (j == T - 1 && im.issue) umma_commit(&q_empty[b]);
        im.sync();
      };
      if (K > 0) issue_s(0);
      if (K > 1) issue_s(1);
      for (int k = 0; k < K; ++k) {
        const int s = k & 1;
        mbar_wait(&p_full[s], (k >> 1) & 1);                  // P(k) written, S(k) consumed
        if (k > 1) mbar_wait(&o_read[s], ((k >> 1) & 1) ^ 1);  // O(k-2) read out of this TMEM buffer
        mbar_wait(&v_full[s], (k >> 1) & 1);
        tc_fence_after();
        flash_mma_a_k_b_mn<HD>(tmem_base + kColO + s * 128, smem_u32(smem + Sm::kPOff + s * kTcPBytes),
                               smem_u32(smem + Sm::kVOff + s * Tl::kOpBytes), ksteps, idesc_o64, idesc_o16, false,
                               im.issue);
        if (im.issue) umma_commit(&o_full[s]);
        if (im.issue) umma_commit(&v_empty[s]);
        im.sync();
        if (k + 2 < K) issue_s(k + 2);                // S buffer s is free: p_full(k) has been seen
      }
    }
  } else if (fl_is_worker(warp)) {
    const int q = warp & 3;
    const int half = fl_half(warp);       // column half of the score row / of the output row
    const int row = q * 32 + lane;
    const bool warp_active = q * 32 < Rt;  // warp-uniform and equal for the two halves of a quarter
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // The loop is instantiated twice: warps whose 32 rows lie beyond the tile only keep the barrier protocol going.
    // With a runtime `if (active)` around the math, the running output row became a phi of {updated, untouched}
    // that ptxas placed in the TMEM-load registers: 64 register moves per step in the active warps.
    auto run_workers = [&](auto act_tag) {
    constexpr bool active = decltype(act_tag)::value;
    const int hsplit = ((p.npad + 31) >> 5) << 4;                // 96 keys: 48 + 48
    const int c_begin = half * hsplit;
    const int c_end = min(p.npad, c_begin + hsplit);
    // packed tiles: this row's sample within the tile and its first key column
    const int psample = G > 1 ? min(row, Rt - 1) / L : 0;
    const int plo = psample * L;
    float m = -INFINITY, ms = 0.f, l = 0.f;
    float o_main[32], o_rem[16];
#pragma unroll
    for (int d = 0; d < 32; ++d) o_main[d] = 0.f;
#pragma unroll
    for (int d = 0; d < 16; ++d) o_rem[d] = 0.f;

    // scores of step k (cursor c) -> running max / partial sum, P(k) (bf16) into buffer k & 1; returns alpha(k)
    auto softmax_step = [&](int k, const StepCursor& c) -> float {
      const int s = k & 1;
      int lo = 0, hi;                                                 // valid keys [lo, hi) of this tile
      if (G > 1) {
        lo = plo;
        hi = CAUSAL ? min(plo + L, row + 1) : plo + L;
      } else {
        hi = min(Rt, L - c.j * Rt);
        if (CAUSAL) hi = min(hi, c.i * Rt + row - c.j * Rt + 1);
      }
      const bool need_mask = CAUSAL || G > 1 || hi < c_end;           // warp-uniform
      mbar_wait(&s_full[s], (k >> 1) & 1);
      tc_fence_after();
      float alpha = 0.f;
      if (active) {
        uint32_t v[NCH][16];
#pragma unroll
        for (int c3 = 0; c3 < NCH; ++c3)
          if (c_begin + 16 * c3 < c_end) tmem_ld_32x16(t_row + s * 128 + c_begin + 16 * c3, v[c3]);
        tmem_ld_wait();
        // per 16-column chunk, decided for the whole warp: 0 = every key valid, 1 = some masked, 2 = none valid.  Only the chunk
        // that straddles a boundary pays for the per-element test (2 ISETP + 2 FSEL each: with the test on every
        // element the masked half-row warps ran 1.6x the instructions of the others and set the pace of the tile).
        int kind[NCH];
#pragma unroll
        for (int c3 = 0; c3 < NCH; ++c3) {
          const int base = c_begin + 16 * c3;
#if CLIPA_FLASH_MASKMODE == 0
          kind[c3] = need_mask ? 1 : 0;                               // every chunk of a masked warp pays the test
#elif CLIPA_FLASH_MASKMODE == 2
          // plain (non-causal, unpacked) tiles: [0, hi) is the same for every row, so only the chunk that holds
          // `hi` is tested -- decided without votes, and never "skip", so the exp pass stays one straight block
          kind[c3] = (!CAUSAL && G == 1) ? ((need_mask && base + 16 > hi) ? 1 : 0) : (need_mask ? 1 : 0);
#else
          const bool full = !need_mask || (base >= lo && base + 16 <= hi);
          const bool none = need_mask && (base >= hi || base + 16 <= lo);
          kind[c3] = __all_sync(0xffffffffu, full) ? 0 : (__all_sync(0xffffffffu, none) ? 2 : 1);   // warp-uniform
#endif
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c3 = 0; c3 < NCH; ++c3) {
          if (c_begin + 16 * c3 < c_end && kind[c3] != 2) {
            if (kind[c3] == 1) {
              const int rel_lo = lo - (c_begin + 16 * c3), rel_hi = hi - (c_begin + 16 * c3);
#pragma unroll
              for (int jj = 0; jj < 16; ++jj)
                if (jj >= rel_hi || jj < rel_lo) v[c3][jj] = __float_as_uint(-INFINITY);
            }
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) mx[jj & 3] = fmaxf(mx[jj & 3], __uint_as_float(v[c3][jj]));
          }
        }
        const float mine = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        xmax[(s * 2 + half) * 128 + row] = mine;
        pair_sync(q);
        const float m_new = fmaxf(m, fmaxf(mine, xmax[(s * 2 + (half ^ 1)) * 128 + row]));
        const float ms_new = (m_new == -INFINITY) ? 0.f : m_new * p.scale_log2;
        alpha = (m == -INFINITY) ? 0.f : ex2_approx(ms - ms_new);
        m = m_new;
        ms = ms_new;
        float2 sm01 = make_float2(0.f, 0.f), sm23 = make_float2(0.f, 0.f);
        const float2 sc2 = splat2(p.scale_log2), nms2 = splat2(-ms);
        uint8_t* pbuf = smem + Sm::kPOff + s * kTcPBytes;
#pragma unroll
        for (int c3 = 0; c3 < NCH; ++c3) {
          if (c_begin + 16 * c3 < c_end) {
            float pr[16];
            if (kind[c3] != 2) {
#if CLIPA_FLASH_MATHMODE == 0
#pragma unroll
              for (int jj = 0; jj < 16; jj += 4) {
                pr[jj] = ex2_approx(fmaf(__uint_as_float(v[c3][jj]), p.scale_log2, -ms));
                pr[jj + 1] = ex2_approx(fmaf(__uint_as_float(v[c3][jj + 1]), p.scale_log2, -ms));
                pr[jj + 2] = ex2_approx(fmaf(__uint_as_float(v[c3][jj + 2]), p.scale_log2, -ms));
                pr[jj + 3] = ex2_approx(fmaf(__uint_as_float(v[c3][jj + 3]), p.scale_log2, -ms));
                sm01.x += pr[jj]; sm01.y += pr[jj + 1]; sm23.x += pr[jj + 2]; sm23.y += pr[jj + 3];
              }
#else
#pragma unroll
              for (int jj = 0; jj < 16; jj += 4) {     // packed fp32: one FFMA2 / FADD2 per pair; exp2(-inf) = 0
                const float2 t0 = fma2(make_float2(__uint_as_float(v[c3][jj]), __uint_as_float(v[c3][jj + 1])), sc2, nms2);
                const float2 t1 = fma2(make_float2(__uint_as_float(v[c3][jj + 2]), __uint_as_float(v[c3][jj + 3])), sc2, nms2);
                pr[jj] = ex2_approx(t0.x); pr[jj + 1] = ex2_approx(t0.y);
                pr[jj + 2] = ex2_approx(t1.x); pr[jj + 3] = ex2_approx(t1.y);
                sm01 = add2(sm01, make_float2(pr[jj], pr[jj + 1]));
                sm23 = add2(sm23, make_float2(pr[jj + 2], pr[jj + 3]));
              }
#endif
            } else {
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) pr[jj] = 0.f;
            }
#pragma unroll
            for (int g8 = 0; g8 < 2; ++g8)
              *reinterpret_cast<uint4*>(pbuf + p_tile_off(row, ((c_begin + 16 * c3) >> 3) + g8)) = pack8_bf16(pr + 8 * g8);
          }
        }
        const float2 sm = add2(sm01, sm23);
        l = fmaf(l, alpha, sm.x + sm.y);
        fence_proxy_async_smem();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[s]);
      return alpha;
    };

    StepCursor cn, co;      // cn: the step the softmax works on (one ahead), co: the step whose O_j is read
    cn.init(T); co.init(T);
    if (K > 0) decode(0, cn.n, cn.h, cn.i);
    co.n = cn.n; co.h = cn.h; co.i = cn.i;
    float alpha_cur = 0.f, alpha_next = 0.f;
    if (K > 0) {
      alpha_cur = softmax_step(0, cn);
      if (cn.advance() && cn.it < n_local) decode(cn.it, cn.n, cn.h, cn.i);
    }
    for (int k = 0; k < K; ++k) {
      const bool last = co.j == T - 1;
      const float l_fin = l, ms_fin = ms;      // state of step k's item before the next step touches it
      if (k + 1 < K) {
        if (last) { m = -INFINITY; ms = 0.f; l = 0.f; }   // step k+1 opens the next item
        alpha_next = softmax_step(k + 1, cn);
        if (cn.advance() && cn.it < n_local) decode(cn.it, cn.n, cn.h, cn.i);
      }
      // ---- O_j (fresh accumulator of key tile j) -> running output row in registers
      mbar_wait(&o_full[k & 1], (k >> 1) & 1);
      tc_fence_after();
      if (active) {
        const uint32_t t_o = t_row + kColO + (k & 1) * 128;
        uint32_t t[32];
        tmem_ld_32x32(t_o + half * 32, t);
        if (Tl::kRem != 0 && half == 0) {
          uint32_t t2[16];
          tmem_ld_32x16(t_o + 64, t2);
          tmem_ld_wait();
#pragma unroll
          for (int d = 0; d < 16; d += 2)
            fma2_acc(o_rem[d], o_rem[d + 1], alpha_cur, __uint_as_float(t2[d]), __uint_as_float(t2[d + 1]));
        } else {
          tmem_ld_wait();
        }
#pragma unroll
        for (int d = 0; d < 32; d += 2)
          fma2_acc(o_main[d], o_main[d + 1], alpha_cur, __uint_as_float(t[d]), __uint_as_float(t[d + 1]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_read[k & 1]);
      if (last && active) {
        const int n = co.n, h = co.h, i = co.i;
        xsum[half * 128 + row] = l_fin;
        pair_sync(q);
        const float lt = l_fin + xsum[(half ^ 1) * 128 + row];
        const bool row_valid = G > 1 ? (row < Rt && n + psample < p.batch) : row < min(Rt, L - i * Rt);
        if (row_valid) {
          const float inv = lt > 0.f ? 1.f / lt : 0.f;
          const long long grow = (long long)n * L + i * Rt + row;
          __nv_bfloat16* dst = p.out + grow * D + h * HD;
#pragma unroll
          for (int d = 0; d < 32; ++d) o_main[d] *= inv;
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) reinterpret_cast<uint4*>(dst + half * 32)[g8] = pack8_bf16(o_main + 8 * g8);
          if (Tl::kRem != 0 && half == 0) {
#pragma unroll
            for (int d = 0; d < 16; ++d) o_rem[d] *= inv;
#pragma unroll
            for (int g8 = 0; g8 < 2; ++g8) reinterpret_cast<uint4*>(dst + 64)[g8] = pack8_bf16(o_rem + 8 * g8);
          }
          if (half == 0)
            p.lse[((long long)(n + psample) * H + h) * L + i * Rt + row - plo] = (ms_fin + log2f(lt)) * 0.69314718055994531f;
        }
#pragma unroll
        for (int d = 0; d < 32; ++d) o_main[d] = 0.f;
#pragma unroll
        for (int d = 0; d < 16; ++d) o_rem[d] = 0.f;
      }
      alpha_cur = alpha_next;
      if (co.advance() && co.it < n_local) decode(co.it, co.n, co.h, co.i);
    }
    };   // run_workers
    if (warp_active) run_workers(std::true_type{}); else run_workers(std::false_type{});
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// Sequence -> tiles.  Long sequences: T equal tiles of Rt = ceil(L / T) rows.  Short ones (T == 1): G = max_rows / L
// consecutive samples of a head share one tile (rows of G samples are contiguous in the [batch*L, .] matrices);
// the score tile is then block-diagonal.
static void flash_tiling(int L, int max_rows, int& T, int& Rt, int& npad, int& G) {
  T = (L + max_rows - 1) / max_rows;
  G = T == 1 ? max_rows / L : 1;
  Rt = T == 1 ? G * L : (L + T - 1) / T;
  npad = (Rt + 15) & ~15;
}

template <int HD>
static int flash_tmaps(CUtensorMap* tm_main, CUtensorMap* tm_rem, const void* base, int cols, long long rows, int Rt) {
  int rc = encode_tmap_2d_bf16(tm_main, base, (uint64_t)cols, (uint64_t)rows, (uint64_t)cols * 2, 64, (uint32_t)Rt);
  if (rc) return rc;
  if (FlashTile<HD>::kRem != 0)
    return encode_tmap_2d_bf16(tm_rem, base, (uint64_t)cols, (uint64_t)rows, (uint64_t)cols * 2,
                               (uint32_t)FlashTile<HD>::kRem, (uint32_t)Rt, 32);
  *tm_rem = *tm_main;
  return CLIPA_OK;
}

template <int HD>
static int launch_fwd_flash(const void* qkv, void* out, float* lse, int batch, int L, int H, int causal,
                            cudaStream_t stream) {
  const int D = H * HD;
  FlashParams p;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.L = L; p.H = H; p.batch = batch;
  flash_tiling(L, kFlMaxRowsFwd, p.T, p.Rt, p.npad, p.G);
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)HD);
  CUtensorMap tm_main, tm_rem;
  int rc = flash_tmaps<HD>(&tm_main, &tm_rem, qkv, 3 * D, (long long)batch * L, p.Rt);
  if (rc) return rc;
  const long long total = (long long)((batch + p.G - 1) / p.G) * H * p.T;
  CLIPA_REQUIRE(total < (1LL << 31), CLIPA_ERR_UNSUPPORTED, "attention_fwd: too many work items");
  int grid = num_sms();
  if (grid > total) grid = (int)total;
  constexpr int smem_bytes = FlashFwdSmem<HD>::kTotal;
  auto launch = [&](auto kern) -> int {
    CLIPA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    kern<<<grid, kFlThreads, smem_bytes, stream>>>(tm_main, tm_rem, p);
    return CLIPA_OK;
  };
  if (p.npad <= 96)
    rc = causal ? launch(attn_fwd_flash_kernel<HD, true, 3>) : launch(attn_fwd_flash_kernel<HD, false, 3>);
  else
    rc = causal ? launch(attn_fwd_flash_kernel<HD, true, 4>) : launch(attn_fwd_flash_kernel<HD, false, 4>);
  if (rc) return rc;
  CLIPA_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return CLIPA_OK;
}

// ================================================================================================
// BACKWARD, any L, head_dim 64 / 80.  Work item = (sample, head); its T x T tile pairs run on ONE CTA,
// key tile j outer, query tile i inner (96-row operand tiles, M = 128 MMAs over-read 32 rows whose
// accumulator lanes are never read -- same trick as attn_bwd_tc_pipe_kernel):
//   S = Q_i K_j^T, dP = dO_i V_j^T  ->  P = exp(S*scale - lse_i), dS = P o (dP - delta_i) * scale
//   dV_j += P^T dO_i, dK_j += dS^T Q_i      accumulate in TMEM over i, stored after the last i
//   dQ_i += dS K_j                           fresh TMEM tile per pair; summed over j by the worker that
//                                            owns the row, through a CTA-private fp32 scratch (L2-resident,
//                                            no atomics: the same thread reads and writes an element);
//                                            the last j writes bf16 straight to dqkv.  T = 1: no scratch.
// TMEM: S 128 + dP 128 + dQ/dK/dV 3 x HD = 448 / 496 columns.
// smem: P, dS [query][key] bf16 (2 x 2 atoms x 96 rows) | {K_j, V_j} x 2 slots | {Q_i, dO_i, O_i} x 2 slots.
// ================================================================================================
constexpr int kFbRows = 96;

template <int HD>
struct FlashTile96 {
  static constexpr int kRem = HD - 64;
  static constexpr int kMainBytes = kFbRows * 128;
  static constexpr int kRemBytes = kRem ? kFbRows * 32 : 0;
  static constexpr int kOpBytes = kMainBytes + kRemBytes;
  static constexpr uint32_t kRowBytes = 2 * HD;
};

template <int HD, bool PIPE>
struct FlashBwdSmem {
  using Tl = FlashTile96<HD>;
  static constexpr int kPAtom = kFbRows * 128;                 // one 64-key atom of P or dS
  static constexpr int kPdsBytes = 4 * kPAtom;                 // one {P, dS} buffer: 2 x 2 atoms
  static constexpr int kNumPds = PIPE ? 2 : 1;
  static constexpr int kPdsOff = 0;                            // {P, dS} buffers first: the M = 128 over-read of the
  static constexpr int kKvOff = kNumPds * kPdsBytes;           //   last dS atom lands in the operand slots below
  static constexpr int kQdOff = kKvOff + 4 * Tl::kOpBytes;     // 2 slots x {K, V}, then 2 slots x {Q, dO, O}
  static constexpr int kBarOff = kQdOff + 6 * Tl::kOpBytes;
  static constexpr int kDcTiles = 8;                           // delta_i cache: [column half][query tile][row] fp32
  static constexpr int kDcOff = kBarOff + 256;
  static constexpr int kTotal = kDcOff + 2 * kDcTiles * 128 * 4 + 1024;
  static_assert(kTotal <= 227 * 1024, "flash attention backward shared memory budget");
};

struct FlashBwdParams {
  const float* lse;
  __nv_bfloat16* dqkv;
  float* scratch;        // [gridDim.x][T * Rt][HD] fp32 (unused when T == 1)
  int L, H, batch;
  int T, Rt, npad;
  int G;                 // samples packed per tile (see FlashParams)
  int dq_tmem;           // dQ_i of all T query tiles stay in TMEM across the key tiles (no scratch round trip)
  float scale;
};

template <int HD>
__device__ __forceinline__ void flash96_load_op(uint8_t* dst, const CUtensorMap* tm_main, const CUtensorMap* tm_rem,
                                                uint64_t* bar, int col, int row) {
  tma_load_2d(dst, tm_main, bar, col, row);
  if constexpr (HD != 64) tma_load_2d(dst + FlashTile96<HD>::kMainBytes, tm_rem, bar, col + 64, row);
}
template <int HD>
__device__ __forceinline__ void flash96_mma_kmajor(uint32_t d, uint32_t a, uint32_t b, uint32_t idesc, bool issue) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (issue) umma_bf16(d, make_smem_desc_sw128(a + k * 32, 16, 1024), make_smem_desc_sw128(b + k * 32, 16, 1024),
                         idesc, k > 0 ? 1u : 0u);
  if constexpr (HD != 64)
    if (issue) umma_bf16(d, make_smem_desc_sw32(a + FlashTile96<HD>::kMainBytes, 16, 256),
                         make_smem_desc_sw32(b + FlashTile96<HD>::kMainBytes, 16, 256), idesc, 1u);
}
__device__ __forceinline__ uint32_t fb_tile_off(int row, int chunk) {
  return (chunk >> 3) * (kFbRows * 128) + row * 128 + (((chunk & 7) ^ (row & 7)) << 4);
}

// (item, key tile j, query tile i) position of a backward step, i fastest
struct BwdCursor {
  int it, j, i, T;
  int n, h;        // decoded item
  __device__ __forceinline__ void init(int tiles) { it = 0; j = 0; i = 0; T = tiles; n = h = 0; }
  __device__ __forceinline__ bool advance() {   // true when it moved on to the next item
    if (++i == T) {
      i = 0;
      if (++j == T) { j = 0; ++it; return true; }
    }
    return false;
  }
};

// PIPE (head_dim 64: two {P, dS} buffers fit): the worker loop is skewed by one step like
// attn_bwd_tc_pipe_kernel -- scores stage of step k+1 runs while the tensor pipe does the gradient MMAs of
// step k, and S/dP(k+1) are issued right behind pds_full(k):
//     workers:  ... | P/dS(k+1) from S/dP(k+1) | epilogue(k) | P/dS(k+2) | epilogue(k+1) | ...
//     tensor :  ... | S,dP(k+2) | dV,dK,dQ(k+1) | S,dP(k+3) | dV,dK,dQ(k+2) | ...
// (unpipelined, the workers spent 32 % of their samples waiting for the gradient MMAs: ncu source view, r02).
template <int HD, bool CAUSAL, bool PIPE>
__global__ void __launch_bounds__(kFlThreads, 1)
attn_bwd_flash_kernel(const __grid_constant__ CUtensorMap tq_main, const __grid_constant__ CUtensorMap tq_rem,
                      const __grid_constant__ CUtensorMap td_main, const __grid_constant__ CUtensorMap td_rem,
                      const __grid_constant__ CUtensorMap to_main, const __grid_constant__ CUtensorMap to_rem,
                      const FlashBwdParams p) {
  using Tl = FlashTile96<HD>;
  using Sm = FlashBwdSmem<HD, PIPE>;
  constexpr int kPAtom = Sm::kPAtom;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Sm::kBarOff);
  uint64_t* kv_full = bars;           // [2] K_j, V_j landed
  uint64_t* kv_empty = bars + 2;      // [2] last gradient MMA that reads K_j done
  uint64_t* qd_full = bars + 4;       // [2] Q_i, dO_i, O_i landed
  uint64_t* qd_empty = bars + 6;      // [2] gradient MMAs of the step done
  uint64_t* sdp_full = bars + 8;      // S and dP in TMEM
  uint64_t* pds_full = bars + 9;      // [2] P and dS buffer written (8 warp arrivals)
  uint64_t* grad_full = bars + 11;    // dQ (and dK, dV) in TMEM
  uint64_t* t_free = bars + 12;       // gradients read out (8 warp arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L, H = p.H, D = H * HD, T = p.T, Rt = p.Rt;
  const int G = p.G;
  const int total = ((p.batch + G - 1) / G) * H;
  const long long pitch = 3LL * D;

  for (int i = threadIdx.x; i < Sm::kBarOff / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tq_main);
    tma_prefetch_desc(&td_main);
    tma_prefetch_desc(&to_main);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&qd_full[i], 1);
      mbar_init(&qd_empty[i], 1);
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
  // TMEM columns.  Default: S 0, dP 128, one dQ tile 256, dK 256+HD, dV 256+2HD (dQ partial sums over the key tiles
  // go through the fp32 scratch).  dq_tmem (T*HD + 2*npad + 2*HD <= 512, e.g. L = 257 at head_dim 64): S 0, dP npad,
  // dQ_i at 2*npad + i*HD for every query tile of the item -- accumulated by the MMAs over j, read once --, then dK, dV.
  constexpr uint32_t kColS = 0;
  const bool dqt = p.dq_tmem != 0;
  const uint32_t kColDp = dqt ? (uint32_t)p.npad : 128u;
  const uint32_t kColDq = dqt ? 2u * (uint32_t)p.npad : 256u;
  const uint32_t kColDk = dqt ? kColDq + (uint32_t)(T * HD) : 256u + HD;
  const uint32_t kColDv = kColDk + HD;

  const int n_local = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int K = n_local * T * T;   // steps: (item, key tile j, query tile i), i fastest
  auto decode = [&](BwdCursor& c) {          // c.n = FIRST sample of the tile
    const int prob = blockIdx.x + c.it * gridDim.x;
    const int tile = prob / H;
    c.h = prob - tile * H;
    c.n = tile * G;
  };
  // {P, dS} buffer and its barrier for step k
  auto pds_buf = [&](int k) -> int { return PIPE ? (k & 1) : 0; };
  auto pds_par = [&](int k) -> uint32_t { return PIPE ? ((k >> 1) & 1) : (k & 1); };

  if (warp == kFlTmaWarp) {
    if (lane == 0) {
      BwdCursor c;
      c.init(T);
      decode(c);
      int jj = 0;      // running (item, key tile) count
      for (int k = 0; k < K; ++k) {
        if (c.i == 0) {
          const int s = jj & 1;
          mbar_wait(&kv_empty[s], ((jj >> 1) & 1) ^ 1);
          uint8_t* st = smem + Sm::kKvOff + s * 2 * Tl::kOpBytes;
          mbar_expect_tx(&kv_full[s], 2u * (uint32_t)Rt * Tl::kRowBytes);
          flash96_load_op<HD>(st, &tq_main, &tq_rem, &kv_full[s], D + c.h * HD, c.n * L + c.j * Rt);
          flash96_load_op<HD>(st + Tl::kOpBytes, &tq_main, &tq_rem, &kv_full[s], 2 * D + c.h * HD, c.n * L + c.j * Rt);
          ++jj;
        }
        const int s = k & 1;
        mbar_wait(&qd_empty[s], ((k >> 1) & 1) ^ 1);
        uint8_t* st = smem + Sm::kQdOff + s * 3 * Tl::kOpBytes;
        mbar_expect_tx(&qd_full[s], 3u * (uint32_t)Rt * Tl::kRowBytes);
        flash96_load_op<HD>(st, &tq_main, &tq_rem, &qd_full[s], c.h * HD, c.n * L + c.i * Rt);
        flash96_load_op<HD>(st + Tl::kOpBytes, &td_main, &td_rem, &qd_full[s], c.h * HD, c.n * L + c.i * Rt);
        flash96_load_op<HD>(st + 2 * Tl::kOpBytes, &to_main, &to_rem, &qd_full[s], c.h * HD, c.n * L + c.i * Rt);
        if (c.advance() && c.it < n_local) decode(c);
      }
    }
  } else if (warp == kFlMmaWarp) {
    const IssueMode im = issue_mode(lane);
    if (im.in_loop) {
      const uint32_t idesc_s = make_idesc_bf16(128, (uint32_t)p.npad, false, false);   // A K-major, B K-major
      const uint32_t idesc_t64 = make_idesc_bf16(128, 64, true, true);                 // A MN (P^T / dS^T), B MN
      const uint32_t idesc_t16 = make_idesc_bf16(128, 16, true, true);
      const uint32_t idesc_q64 = make_idesc_bf16(128, 64, false, true);                // A K (dS), B MN (K)
      const uint32_t idesc_q16 = make_idesc_bf16(128, 16, false, true);
      const int ksteps = p.npad / 16;
      BwdCursor cs, cg;     // cs: next scores step to issue, cg: gradient step
      cs.init(T); cg.init(T);
      int jj_s = 0, jj_g = 0;   // key-tile counters of the two cursors ((item, j) pairs started so far)
      auto issue_scores = [&](int k) {      // k == cs's step
        if (cs.i == 0) ++jj_s;
        const int sk = (jj_s - 1) & 1, s = k & 1;
        mbar_wait(&qd_full[s], (k >> 1) & 1);
        if (cs.i == 0) mbar_wait(&kv_full[sk], ((jj_s - 1) >> 1) & 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(smem + Sm::kQdOff + s * 3 * Tl::kOpBytes);
        const uint32_t doa = qa + Tl::kOpBytes;
        const uint32_t ka = smem_u32(smem + Sm::kKvOff + sk * 2 * Tl::kOpBytes);
        const uint32_t va = ka + Tl::kOpBytes;
        flash96_mma_kmajor<HD>(tmem_base + kColS, qa, ka, idesc_s, im.issue);
        flash96_mma_kmajor<HD>(tmem_base + kColDp, doa, va, idesc_s, im.issue);
        if (im.issue) umma_commit(sdp_full);
        im.sync();
        cs.advance();
      };
      auto issue_grads = [&](int k) {       // k == cg's step
        if (cg.i == 0) ++jj_g;
        const int sk = (jj_g - 1) & 1, s = k & 1;
        const uint32_t qa = smem_u32(smem + Sm::kQdOff + s * 3 * Tl::kOpBytes);
        const uint32_t doa = qa + Tl::kOpBytes;
        const uint32_t ka = smem_u32(smem + Sm::kKvOff + sk * 2 * Tl::kOpBytes);
        const uint32_t pa = smem_u32(smem + Sm::kPdsOff + pds_buf(k) * Sm::kPdsBytes), dsa = pa + 2 * kPAtom;
        const uint32_t acc0 = cg.i > 0 ? 1u : 0u;   // dK_j, dV_j accumulate over the query tiles
        for (int kk = 0; kk < ksteps; ++kk) {
          // contraction over queries: A = P^T / dS^T (MN-major view of the [query][key] tiles: key atoms
          // one atom apart = LBO, 8-query groups 1 KB apart = SBO), B = dO / Q consumed MN-major
          const uint64_t a_p = make_smem_desc_sw128(pa + kk * 2048, kPAtom, 1024);
          const uint64_t a_ds = make_smem_desc_sw128(dsa + kk * 2048, kPAtom, 1024);
          const uint32_t acc = (acc0 | (kk > 0 ? 1u : 0u));
          if (im.issue) umma_bf16(tmem_base + kColDv, a_p, make_smem_desc_sw128(doa + kk * 2048, 8192, 1024), idesc_t64, acc);
          if (im.issue) umma_bf16(tmem_base + kColDk, a_ds, make_smem_desc_sw128(qa + kk * 2048, 8192, 1024), idesc_t64, acc);
          if constexpr (HD != 64) {
            if (im.issue) umma_bf16(tmem_base + kColDv + 64, a_p,
                                    make_smem_desc_sw32(doa + Tl::kMainBytes + kk * 512, 256, 256), idesc_t16, acc);
            if (im.issue) umma_bf16(tmem_base + kColDk + 64, a_ds,
                                    make_smem_desc_sw32(qa + Tl::kMainBytes + kk * 512, 256, 256), idesc_t16, acc);
          }
        }
        for (int kk = 0; kk < ksteps; ++kk) {
          // contraction over keys: A = dS K-major, B = K_j consumed MN-major; fresh dQ tile
          const uint64_t a_ds = make_smem_desc_sw128(dsa + (kk >> 2) * kPAtom + (kk & 3) * 32, 16, 1024);
          const uint32_t acc = (kk > 0 || (dqt && cg.j > 0)) ? 1u : 0u;
          const uint32_t d_q = tmem_base + kColDq + (dqt ? (uint32_t)(cg.i * HD) : 0u);
          if (im.issue) umma_bf16(d_q, a_ds, make_smem_desc_sw128(ka + kk * 2048, 8192, 1024), idesc_q64, acc);
          if constexpr (HD != 64)
            if (im.issue) umma_bf16(d_q + 64, a_ds,
                                    make_smem_desc_sw32(ka + Tl::kMainBytes + kk * 512, 256, 256), idesc_q16, acc);
        }
        if (im.issue) umma_commit(grad_full);
        if (im.issue) umma_commit(&qd_empty[s]);
        if (cg.i == T - 1 && im.issue) umma_commit(&kv_empty[sk]);
        im.sync();
        cg.advance();
      };
      if (K > 0) issue_scores(0);
      for (int k = 0; k < K; ++k) {
        mbar_wait(&pds_full[pds_buf(k)], pds_par(k));   // P/dS(k) in smem; S/dP(k) consumed
        tc_fence_after();
        if (PIPE && k + 1 < K) issue_scores(k + 1);     // under the workers' epilogue of step k-1
        mbar_wait(t_free, (k & 1) ^ 1);                 // gradients of step k-1 read out of TMEM
        tc_fence_after();
        issue_grads(k);
        if (!PIPE && k + 1 < K) issue_scores(k + 1);
      }
    }
  } else if (fl_is_worker(warp)) {
    const int q = warp & 3;
    const int half = fl_half(warp);
    const int row = q * 32 + lane;          // query index (scores stage, dQ rows) / key index (dK, dV rows)
    const bool warp_writes = q * 32 < p.npad;
    const bool warp_stores = q * 32 < Rt;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float scale_log2 = p.scale * 1.4426950408889634f;
    const int hsplit = ((p.npad + 31) >> 5) << 4;
    const int c_begin = half * hsplit;
    const int c_end = min(p.npad, c_begin + hsplit);
    float* scr = p.scratch + (long long)blockIdx.x * T * Rt * HD;
    float* dcache = reinterpret_cast<float*>(smem + Sm::kDcOff);
    const bool delta_cached = T > 1 && T <= Sm::kDcTiles;
    // packed tiles: this row's sample within the tile and its first row / key column
    const int psample = G > 1 ? min(row, Rt - 1) / L : 0;
    const int plo = psample * L;
    auto row_valid = [&](const BwdCursor& c, int tile_idx) -> bool {   // row of query tile i / key tile j
      return G > 1 ? (row < Rt && c.n + psample < p.batch) : row < min(Rt, L - tile_idx * Rt);
    };
    auto lse_of = [&](const BwdCursor& c) -> float {     // lse is [batch, H, L]
      return row_valid(c, c.i) ? p.lse[((long long)(c.n + psample) * H + c.h) * L + c.i * Rt + row - plo] : 0.f;
    };

    // 32 (or 16) fp32 columns of this thread's TMEM row at column `col` -> bf16 at dst
    auto store32 = [&](__nv_bfloat16* dst, uint32_t col, bool ok) {
      uint32_t v[32];
      tmem_ld_32x32(t_row + col, v);
      tmem_ld_wait();
      if (ok) {
        float f[32];
#pragma unroll
        for (int d = 0; d < 32; ++d) f[d] = __uint_as_float(v[d]);
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) reinterpret_cast<uint4*>(dst)[g8] = pack8_bf16(f + 8 * g8);
      }
    };
    auto store16 = [&](__nv_bfloat16* dst, uint32_t col, bool ok) {
      uint32_t v[16];
      tmem_ld_32x16(t_row + col, v);
      tmem_ld_wait();
      if (ok) {
        float f[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) f[d] = __uint_as_float(v[d]);
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) reinterpret_cast<uint4*>(dst)[g8] = pack8_bf16(f + 8 * g8);
      }
    };

    BwdCursor cs, ce, cl;      // cs: scores stage, ce: epilogue, cl: one step ahead of cs (lse prefetch)
    cs.init(T); ce.init(T); cl.init(T);
    decode(cs);
    ce.n = cl.n = cs.n; ce.h = cl.h = cs.h;
    float lse_raw = K > 0 ? lse_of(cl) : 0.f;
    if (cl.advance() && cl.it < n_local) decode(cl);

    // scores stage of step k (cursor cs): S/dP (TMEM) -> P/dS (bf16, swizzled smem buffer)
    auto scores_stage = [&](int k) {
      const int s = k & 1;
      const bool row_ok = row_valid(cs, cs.i);
      int lo = 0, hi;                                      // valid keys [lo, hi) of this tile for this query row
      if (G > 1) {
        lo = plo;
        hi = CAUSAL ? min(plo + L, row + 1) : plo + L;
      } else {
        hi = min(Rt, L - cs.j * Rt);
        if (CAUSAL) hi = min(hi, cs.i * Rt + row - cs.j * Rt + 1);
      }
      // delta_i = sum_d dO_id * O_id from the swizzled smem tiles (same swizzle in both tiles, so any
      // consistent chunk order gives matching pairs)
      // delta_i does not depend on the key tile: computed in the item's first pass over the query tiles (j == 0) and
      // kept in shared memory, one slot per (column half, query tile, row) so that the reader is the thread that
      // wrote it (recomputed for every j it was 140 of the ~1000 instructions of a step: two LDS.128 + 16 unpacks
      // + 8 FFMA per 8 head dimensions, by both threads of the row)
      float delta = 0.f;
      float* dslot = dcache + (half * Sm::kDcTiles + cs.i) * 128 + row;
      if (delta_cached && cs.j > 0) {
        delta = *dslot;
      } else {
      mbar_wait(&qd_full[s], (k >> 1) & 1);
      if (row_ok && c_begin < c_end) {
        const uint8_t* dot = smem + Sm::kQdOff + s * 3 * Tl::kOpBytes + Tl::kOpBytes;
        const uint8_t* ot = dot + Tl::kOpBytes;
        float d4[4] = {0.f, 0.f, 0.f, 0.f};
        auto acc16 = [&](const uint8_t* pa_, const uint8_t* pb_) {
          const uint4 a = *reinterpret_cast<const uint4*>(pa_);
          const uint4 b = *reinterpret_cast<const uint4*>(pb_);
          d4[0] = fmaf(bf16lo(a.x), bf16lo(b.x), fmaf(bf16hi(a.x), bf16hi(b.x), d4[0]));
          d4[1] = fmaf(bf16lo(a.y), bf16lo(b.y), fmaf(bf16hi(a.y), bf16hi(b.y), d4[1]));
          d4[2] = fmaf(bf16lo(a.z), bf16lo(b.z), fmaf(bf16hi(a.z), bf16hi(b.z), d4[2]));
          d4[3] = fmaf(bf16lo(a.w), bf16lo(b.w), fmaf(bf16hi(a.w), bf16hi(b.w), d4[3]));
        };
        // lanes read the 16-byte chunks of their row in the swizzle's rotated order: conflict-free
        // (the plain order put 8 lanes of a quarter-warp phase on the same 4 banks: 8-way conflicts)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t off = row * 128 + ((c ^ (row & 7)) << 4);
          acc16(ot + off, dot + off);
        }
        if constexpr (HD != 64) {
          const uint32_t o0 = Tl::kMainBytes + row * 32 + (((row >> 2) & 1) << 4);
          acc16(ot + o0, dot + o0);
          acc16(ot + (o0 ^ 16), dot + (o0 ^ 16));
        }
        delta = (d4[0] + d4[1]) + (d4[2] + d4[3]);
      }
      if (delta_cached) *dslot = delta;
      }
      mbar_wait(sdp_full, k & 1);
      tc_fence_after();
      const float lse2 = lse_raw * 1.4426950408889634f;
      const float nds = -delta * p.scale;
      if (k + 1 < K) {                       // next step's lse: consumed one whole stage later
        lse_raw = lse_of(cl);
        if (cl.advance() && cl.it < n_local) decode(cl);
      }
      if (warp_writes) {
        uint8_t* p_buf = smem + Sm::kPdsOff + pds_buf(k) * Sm::kPdsBytes;
        uint8_t* ds_buf = p_buf + 2 * kPAtom;
        for (int c = c_begin; c < c_end; c += 16) {
          float pr[16], ds[16];
          const int kind = attn_chunk_kind(row_ok, c, lo, hi);
          if (kind != 2) {
            uint32_t sv[16], dv[16];
            tmem_ld_32x16(t_row + kColS + c, sv);
            tmem_ld_32x16(t_row + kColDp + c, dv);
            tmem_ld_wait();
            if (kind == 0)
              attn_bwd_chunk16<false>(sv, dv, scale_log2, lse2, p.scale, nds, c, lo, hi, row_ok, pr, ds);
            else
              attn_bwd_chunk16<true>(sv, dv, scale_log2, lse2, p.scale, nds, c, lo, hi, row_ok, pr, ds);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) pr[j] = ds[j] = 0.f;
          }
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            const uint32_t off = fb_tile_off(row, (c >> 3) + g8);
            *reinterpret_cast<uint4*>(p_buf + off) = pack8_bf16(pr + 8 * g8);
            *reinterpret_cast<uint4*>(ds_buf + off) = pack8_bf16(ds + 8 * g8);
          }
        }
        fence_proxy_async_smem();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_full[pds_buf(k)]);
      if (cs.advance() && cs.it < n_local) decode(cs);
    };

    // epilogue of step k (cursor ce).  dQ: this thread owns columns [32*half, 32*half+32) (+ [64, 80) for
    // half 0) of query row i*Rt + row; partial sums over the key tiles go through the CTA-private scratch.
    auto epilogue = [&](int k) {
      const int n = ce.n, h = ce.h, i = ce.i, j = ce.j;
      const bool first_j = j == 0, last_j = j == T - 1;
      const bool row_ok = row_valid(ce, i);
      float* srow = scr + (long long)(i * Rt + row) * HD;
      float acc[32], acc2[16];
      const bool dq_ok = warp_stores && row_ok;
      if (!dqt && !first_j && dq_ok) {             // issued before the wait: L2 latency hides behind the MMAs
#pragma unroll
        for (int d4 = 0; d4 < 8; ++d4)
          *reinterpret_cast<float4*>(acc + 4 * d4) = *reinterpret_cast<const float4*>(srow + half * 32 + 4 * d4);
        if (HD != 64 && half == 0) {
#pragma unroll
          for (int d4 = 0; d4 < 4; ++d4)
            *reinterpret_cast<float4*>(acc2 + 4 * d4) = *reinterpret_cast<const float4*>(srow + 64 + 4 * d4);
        }
      }
      mbar_wait(grad_full, k & 1);
      tc_fence_after();
      if (warp_stores) {
        __nv_bfloat16* qrow = p.dqkv + ((long long)n * L + i * Rt + row) * pitch + h * HD;
        const uint32_t t_dq = t_row + kColDq + (dqt ? (uint32_t)(i * HD) : 0u);
        const bool add_acc = !dqt && !first_j;
        if (!dqt || last_j) {
          uint32_t v[32];
          tmem_ld_32x32(t_dq + half * 32, v);
          tmem_ld_wait();
          if (dq_ok) {
            float f[32];
#pragma unroll
            for (int d = 0; d < 32; ++d) f[d] = __uint_as_float(v[d]) + (add_acc ? acc[d] : 0.f);
            if (last_j) {
#pragma unroll
              for (int g8 = 0; g8 < 4; ++g8) reinterpret_cast<uint4*>(qrow + half * 32)[g8] = pack8_bf16(f + 8 * g8);
            } else {
#pragma unroll
              for (int d4 = 0; d4 < 8; ++d4)
                *reinterpret_cast<float4*>(srow + half * 32 + 4 * d4) = *reinterpret_cast<const float4*>(f + 4 * d4);
            }
          }
        }
        if (HD != 64 && half == 0 && (!dqt || last_j)) {
          uint32_t v[16];
          tmem_ld_32x16(t_dq + 64, v);
          tmem_ld_wait();
          if (dq_ok) {
            float f[16];
#pragma unroll
            for (int d = 0; d < 16; ++d) f[d] = __uint_as_float(v[d]) + (add_acc ? acc2[d] : 0.f);
            if (last_j) {
#pragma unroll
              for (int g8 = 0; g8 < 2; ++g8) reinterpret_cast<uint4*>(qrow + 64)[g8] = pack8_bf16(f + 8 * g8);
            } else {
#pragma unroll
              for (int d4 = 0; d4 < 4; ++d4)
                *reinterpret_cast<float4*>(srow + 64 + 4 * d4) = *reinterpret_cast<const float4*>(f + 4 * d4);
            }
          }
        }
        if (i == T - 1) {                          // dK_j, dV_j complete: rows = keys of tile j
          const bool k_ok = row_valid(ce, j);
          __nv_bfloat16* krow = p.dqkv + ((long long)n * L + j * Rt + row) * pitch + D + h * HD;
          store32(krow + half * 32, kColDk + half * 32, k_ok);
          store32(krow + D + half * 32, kColDv + half * 32, k_ok);
          if (HD != 64 && half == 0) {
            store16(krow + 64, kColDk + 64, k_ok);
            store16(krow + D + 64, kColDv + 64, k_ok);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_free);
      if (ce.advance() && ce.it < n_local) decode(ce);
    };

    if constexpr (PIPE) {
      if (K > 0) scores_stage(0);
      for (int k = 0; k < K; ++k) {
        if (k + 1 < K) scores_stage(k + 1);
        epilogue(k);
      }
    } else {
      for (int k = 0; k < K; ++k) {
        scores_stage(k);
        epilogue(k);
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

// CLIPA_FLASH_DQ_TMEM=0 keeps the scratch path for every shape (A/B)
static bool flash_dq_tmem_enabled() {
  static const bool on = [] { const char* e = getenv("CLIPA_FLASH_DQ_TMEM"); return !(e && e[0] == '0'); }();
  return on;
}

template <int HD>
static int launch_bwd_flash(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                            void* workspace, long long workspace_bytes, int batch, int L, int H, int causal,
                            cudaStream_t stream) {
  const int D = H * HD;
  FlashBwdParams p;
  p.lse = lse;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  p.L = L; p.H = H; p.batch = batch;
  flash_tiling(L, kFlMaxRowsBwd, p.T, p.Rt, p.npad, p.G);
  p.scale = 1.0f / sqrtf((float)HD);
  const long long total = (long long)((batch + p.G - 1) / p.G) * H;
  int grid = num_sms();
  if (grid > total) grid = (int)total;
  p.dq_tmem = (flash_dq_tmem_enabled() && p.T > 1 && p.T * HD + 2 * p.npad + 2 * HD <= 512) ? 1 : 0;
  const long long need = (p.T > 1 && !p.dq_tmem) ? (long long)grid * p.T * p.Rt * HD * 4 : 0;
  CLIPA_REQUIRE(need == 0 || (workspace != nullptr && workspace_bytes >= need), CLIPA_ERR_BAD_ARG,
                "attention_bwd: L=%d needs a %lld-byte workspace (clipa_attention_bwd_workspace), got %lld", L, need,
                workspace_bytes);
  p.scratch = static_cast<float*>(workspace);
  CUtensorMap tq, tqr, td, tdr, to, tor;
  int rc = flash_tmaps<HD>(&tq, &tqr, qkv, 3 * D, (long long)batch * L, p.Rt);
  if (rc) return rc;
  rc = flash_tmaps<HD>(&td, &tdr, dout, D, (long long)batch * L, p.Rt);
  if (rc) return rc;
  rc = flash_tmaps<HD>(&to, &tor, out, D, (long long)batch * L, p.Rt);
  if (rc) return rc;
  constexpr bool kPipe = HD == 64;      // two {P, dS} buffers only fit next to head_dim-64 operand tiles
  constexpr int smem_bytes = FlashBwdSmem<HD, kPipe>::kTotal;
  auto launch = [&](auto kern) -> int {
    CLIPA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    kern<<<grid, kFlThreads, smem_bytes, stream>>>(tq, tqr, td, tdr, to, tor, p);
    return CLIPA_OK;
  };
  rc = causal ? launch(attn_bwd_flash_kernel<HD, true, kPipe>) : launch(attn_bwd_flash_kernel<HD, false, kPipe>);
  if (rc) return rc;
  CLIPA_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return CLIPA_OK;
}

long long attention_bwd_flash_workspace(int batch, int L, int H, int hd) {
  int T, Rt, npad, G;
  flash_tiling(L, kFlMaxRowsBwd, T, Rt, npad, G);
  if (T <= 1) return 0;
  if (flash_dq_tmem_enabled() && T * hd + 2 * npad + 2 * hd <= 512) return 0;   // dQ tiles stay in tensor memory
  long long grid = num_sms();
  if (grid > (long long)batch * H) grid = (long long)batch * H;
  return grid * T * Rt * hd * 4;
}

int attention_bwd_flash(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                        void* workspace, long long workspace_bytes, int batch, int L, int H, int hd, int causal,
                        cudaStream_t stream) {
  if (hd == 64)
    return launch_bwd_flash<64>(qkv, out, dout, lse, dqkv, workspace, workspace_bytes, batch, L, H, causal, stream);
  if (hd == 80)
    return launch_bwd_flash<80>(qkv, out, dout, lse, dqkv, workspace, workspace_bytes, batch, L, H, causal, stream);
  set_error("attention_bwd_flash: head_dim %d not built (64, 80)", hd);
  return CLIPA_ERR_UNSUPPORTED;
}

int attention_fwd_flash(const void* qkv, void* out, float* lse, int batch, int L, int H, int hd, int causal,
                        cudaStream_t stream) {
  if (hd == 64) return launch_fwd_flash<64>(qkv, out, lse, batch, L, H, causal, stream);
  if (hd == 80) return launch_fwd_flash<80>(qkv, out, lse, batch, L, H, causal, stream);
  set_error("attention_fwd_flash: head_dim %d not built (64, 80)", hd);
  return CLIPA_ERR_UNSUPPORTED;
}

}  // namespace clipa
