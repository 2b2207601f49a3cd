// 2-CTA (cta_group::2) variant of the persistent tcgen05 GEMM: a CTA pair (cluster of 2 on one TPC)
// computes a 256 x 256 output tile.  Each CTA stages its own 128 rows of A and HALF of the B tile
// (128 of the 256 N-rows); the leader CTA issues tcgen05.mma.cta_group::2 (M = 256), the hardware
// reads the other half of B from the peer's shared memory.  Per MMA-FLOP this cuts the
// L2 -> shared-memory traffic by a third (32 KB instead of 48 KB per 128x256x64 MACs per CTA) --
// the 1-CTA kernel is L2-feed-bound at ~95 B/clk/SM -- and deepens the ring to 6 stages.
//
//   warp 0 (both CTAs): TMA producer for the CTA's own A / B-half tiles; completion is signalled on
//                       the LEADER's full barrier (cp.async.bulk.tensor ... .cta_group::2)
//   warp 1 (leader)   : MMA issuer; tcgen05.commit.multicast releases smem stages / publishes
//                       accumulators in BOTH CTAs
//   warps 2..9 (both) : epilogue on the CTA's own 128 TMEM lanes (rows r*128.. of the pair tile);
//                       accumulator-drained arrivals go to the leader's tmem_empty barrier
#pragma once
#include "gemm_tc.cuh"

namespace clipa {

constexpr int kBN2 = 256;          // pair tile N
constexpr int kABytes2 = kBM * kBK * 2;        // 128 x 64 bf16
constexpr int kBBytes2 = (kBN2 / 2) * kBK * 2; // this CTA's half of B
constexpr int kStageBytes2 = kABytes2 + kBBytes2;

// The DACT epilogue (dgrad of c_proj fused with gelu') streams a second M x N operand -- the saved
// pre-activation -- through the epilogue.  Read with per-lane LDGs (one row per lane, 64 bytes per
// chunk) every warp-level load touches 32 cache lines and the LSU/L1 path, not the tensor pipe,
// paces the kernel (1130 vs 1475 TFLOP/s for the plain dgrad).  Here each epilogue warp fetches its
// 32 x 32 chunk with one TMA load into a private 64B-swizzled buffer, kDactSideBufs chunks ahead.
#ifndef CLIPA_DACT_STAGES
#define CLIPA_DACT_STAGES 6
#endif
#ifndef CLIPA_DACT_SIDE_BUFS
#define CLIPA_DACT_SIDE_BUFS 1
#endif

// Shared-memory layout of the 2-CTA kernel per epilogue type:
//   [ A/B ring | barriers (512 B) | bias staging (2 x 256 fp32; none for DACT) | per-warp buffers ]
// per-warp buffers: one 2 KB output staging tile (TMA store) + the DACT side-operand tiles, or (BIAS_ACT) a second
// staging tile for the recompute pass's act'(f) output.  The 64B-swizzled tiles need 512-byte alignment; BIAS_ACT
// fills the 227 KB exactly enough that the 1 KB base-alignment slack is dropped (the kernel traps if the dynamic
// shared memory base is not 1024-aligned, which it is whenever the kernel has no static shared memory).
template <int EPI>
struct Tc2Smem {
  static constexpr int kStages = (EPI == EPI_DACT) ? CLIPA_DACT_STAGES : 6;
  static constexpr int kSideBufs = (EPI == EPI_DACT) ? CLIPA_DACT_SIDE_BUFS : 0;
  static constexpr int kBarOffset = kStages * kStageBytes2;
  static constexpr int kBiasOffset = kBarOffset + 512;
  static constexpr int kBiasBytes = (EPI == EPI_DACT) ? 0 : 2048;
  static constexpr int kStoreOffset = ((kBiasOffset + kBiasBytes + 511) / 512) * 512;
  static constexpr int kAuxBufOff = (EPI == EPI_BIAS_ACT) ? 2048 : 0;   // second staging tile (0 = none)
  static constexpr int kWarpBytes = 2048 * (1 + kSideBufs) + kAuxBufOff;
  static constexpr int kSlack = (EPI == EPI_BIAS_ACT) ? 0 : 1024;
  static constexpr int kTotal = kStoreOffset + kNumEpiWarps * kWarpBytes + kSlack;
  static_assert(kTotal <= 227 * 1024, "2-CTA GEMM shared memory budget");
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in the LEADER (even) CTA of the pair: within a CTA
// pair the peer is selected by bit 24 of the shared-window address; clearing it targets rank 0.
__device__ __forceinline__ uint32_t leader_addr(uint32_t local_smem_addr) { return local_smem_addr & 0xFEFFFFFFu; }
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose mbarrier may live in the peer CTA (cluster address)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* tm, uint32_t bar_cluster_addr,
                                                int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once) on the barrier at this offset in BOTH CTAs of the pair when prior MMAs retire
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}

// GemmParams here: m_blocks counts 256-row PAIR tiles, n_blocks counts 256-column tiles.
template <bool A_MN, bool B_MN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_aux,
                const GemmParams p) {
  constexpr uint32_t kTmemCols = 2 * kBN2;
  constexpr uint32_t kIdesc = make_idesc_bf16(2 * kBM, kBN2, A_MN, B_MN);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  using S = Tc2Smem<EPI>;
  if constexpr (S::kSlack == 0) {
    if (raw_addr & 1023u) __trap();   // layout without alignment slack (see Tc2Smem)
  }
  constexpr int kStages2 = S::kStages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);   // used in the leader
  uint64_t* empty_bar = full_bar + kStages2;                              // per CTA
  uint64_t* tmem_full = empty_bar + kStages2;                             // per CTA
  uint64_t* tmem_empty = tmem_full + 2;                                   // used in the leader
  uint64_t* side_bar = tmem_empty + 2;                                    // [8 warps][kSideBufs], DACT only
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(side_bar + kNumEpiWarps * 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < kStages2; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * kNumEpiWarps);  // epilogue warps of both CTAs
    }
    if constexpr (EPI == EPI_DACT) {
      tma_prefetch_desc(&tmap_aux);
      for (int i = 0; i < kNumEpiWarps * 2; ++i) mbar_init(&side_bar[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm<kTmemCols>(tmem_ptr);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int items_mn = p.m_blocks * p.n_blocks;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0) {
    // ======================= TMA producer (both CTAs) =======================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cluster_id; w < p.num_items; w += num_clusters) {
        const int split = w / items_mn;
        const int r = w - split * items_mn;
        const int m_blk = r / p.n_blocks;
        const int n_blk = r - m_blk * p.n_blocks;
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(p.k_blocks, kb_begin + p.kb_per_split);
        const int m0 = m_blk * (2 * kBM) + (int)rank * kBM;        // this CTA's 128 rows of A
        const int n0 = n_blk * kBN2 + (int)rank * (kBN2 / 2);       // this CTA's half of B
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes2;
          uint8_t* sb = sa + kABytes2;
          const uint32_t full_leader = leader_addr(smem_u32(&full_bar[stage]));
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * kStageBytes2);
          if constexpr (!A_MN) {
            tma_load_2d_2sm(sa, &tmap_a, full_leader, kb * kBK, m0);
          } else {
#pragma unroll
            for (int a = 0; a < kBM / 64; ++a)
              tma_load_2d_2sm(sa + a * (kBK * 128), &tmap_a, full_leader, m0 + a * 64, kb * kBK);
          }
          if constexpr (!B_MN) {
            tma_load_2d_2sm(sb, &tmap_b, full_leader, kb * kBK, n0);
          } else {
#pragma unroll
            for (int a = 0; a < (kBN2 / 2) / 64; ++a)
              tma_load_2d_2sm(sb + a * (kBK * 128), &tmap_b, full_leader, n0 + a * 64, kb * kBK);
          }
          if (++stage == kStages2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (leader CTA only) =======================
    const IssueMode im = issue_mode(lane);
    if (leader && im.in_loop) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = cluster_id; w < p.num_items; w += num_clusters) {
        const int split = w / items_mn;
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(p.k_blocks, kb_begin + p.kb_per_split);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + acc * kBN2;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes2);
          const uint32_t sb = sa + kABytes2;
          const uint64_t da = A_MN ? make_smem_desc_sw128(sa, kBK * 128, 1024) : make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t db = B_MN ? make_smem_desc_sw128(sb, kBK * 128, 1024) : make_smem_desc_sw128(sb, 16, 1024);
          constexpr uint32_t a_step = (A_MN ? 2048 : 32) >> 4;
          constexpr uint32_t b_step = (B_MN ? 2048 : 32) >> 4;
          if (im.issue) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_bf16_2sm(d_addr, da + static_cast<uint64_t>(k * a_step), db + static_cast<uint64_t>(k * b_step),
                            kIdesc, (kb > kb_begin || k > 0) ? 1u : 0u);
            umma_commit_2sm(&empty_bar[stage]);
          }
          im.sync();
          if (++stage == kStages2) { stage = 0; phase ^= 1; }
        }
        if (im.issue) umma_commit_2sm(&tmem_full[acc]);
        im.sync();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ======================= epilogue warps (both CTAs) =======================
    const int ew = warp - 2;
    const int q = warp & 3;
    const int half = ew >> 2;
    constexpr int kColsPerWarp = kBN2 / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (EPI == EPI_DACT) {
      // ---- dgrad x gelu'(saved pre-activation): side operand through per-warp TMA loads
      constexpr int NB = S::kSideBufs;
      uint8_t* sbuf = smem + S::kStoreOffset + ew * S::kWarpBytes;   // output staging (TMA store)
      uint8_t* lbuf = sbuf + 2048;                                   // NB side-operand tiles
      uint64_t* sbar = side_bar + ew * 2;
      const int sw = (lane >> 1) & 3;
      // global chunk counter g: tile = g / 4 of this cluster's tile sequence, 32-column chunk = g % 4
      auto issue_side = [&](int g) {
        const int w = cluster_id + (g >> 2) * num_clusters;
        if (w < p.num_items && lane == 0) {
          const int m_blk = w / p.n_blocks;
          const int n_blk = w - m_blk * p.n_blocks;
          uint64_t* b = &sbar[g % NB];
          mbar_expect_tx(b, 2048);
          tma_load_2d(lbuf + (g % NB) * 2048, &tmap_aux, b, n_blk * kBN2 + half * kColsPerWarp + (g & 3) * 32,
                      m_blk * (2 * kBM) + (int)rank * kBM + q * 32);
        }
      };
#pragma unroll
      for (int g0 = 0; g0 < NB; ++g0) issue_side(g0);
      int g = 0;
      for (int w = cluster_id; w < p.num_items; w += num_clusters) {
        const int m_blk = w / p.n_blocks;
        const int n_blk = w - m_blk * p.n_blocks;
        const int row0_warp = m_blk * (2 * kBM) + (int)rank * kBM + q * 32;
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_warp = tmem_base + acc * kBN2 + half * kColsPerWarp + (static_cast<uint32_t>(q * 32) << 16);
        uint32_t vnext[32];
        tmem_ld_32x32(t_warp, vnext);
#pragma unroll 1
        for (int c = 0; c < kColsPerWarp / 32; ++c, ++g) {
          const int col0 = n_blk * kBN2 + half * kColsPerWarp + c * 32;
          float f[32];
          tmem_ld_wait_regs(vnext);
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(vnext[j]);
          if (c + 1 < kColsPerWarp / 32) {
            tmem_ld_32x32(t_warp + (c + 1) * 32, vnext);
          } else {   // accumulator stage drained: hand it back before the math of the last chunk
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(leader_addr(smem_u32(&tmem_empty[acc])));
          }
          mbar_wait(&sbar[g % NB], (g / NB) & 1);
          uint4 side[4];
          const uint8_t* lrow = lbuf + (g % NB) * 2048 + lane * 64;
#pragma unroll
          for (int k = 0; k < 4; ++k) side[k] = *reinterpret_cast<const uint4*>(lrow + ((k ^ sw) << 4));
          __syncwarp();
          issue_side(g + NB);   // the tile just read is free again
          if (col0 >= p.N) continue;  // warp-uniform
          float a[32];
          unpack_bf16x32(side, a);   // rows >= M / columns >= N arrive as zeros and are clipped by the store
          if (p.aux_deriv) {         // the recompute pass stored act'(f): one packed multiply per pair
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float2 r = mul2(make_float2(f[j], f[j + 1]), make_float2(a[j], a[j + 1]));
              f[j] = r.x;
              f[j + 1] = r.y;
            }
          } else if (p.act == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float2 r = mul2(make_float2(f[j], f[j + 1]), gelu_erf_bwd2(make_float2(a[j], a[j + 1])));
              f[j] = r.x;
              f[j + 1] = r.y;
            }
          } else if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= act_bwd(a[j], 1);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= act_bwd(a[j], 2);
          }
          chunk_store_tma(&tmap_c, sbuf, f, col0, row0_warp);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    } else {
    for (int w = cluster_id; w < p.num_items; w += num_clusters) {
      const int split = w / items_mn;
      const int r = w - split * items_mn;
      const int m_blk = r / p.n_blocks;
      const int n_blk = r - m_blk * p.n_blocks;
      const long long row = (long long)m_blk * (2 * kBM) + (long long)rank * kBM + q * 32 + lane;
      const bool row_ok = row < p.M;
      float* sbias_tile = reinterpret_cast<float*>(smem + S::kBiasOffset) + acc * 256;
      epi_stage_bias(p, sbias_tile, n_blk * kBN2, kBN2);   // overlaps the MMAs of this tile
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_warp = tmem_base + acc * kBN2 + half * kColsPerWarp + (static_cast<uint32_t>(q * 32) << 16);
      epi_run_store<EPI>(p, t_warp, row, row_ok, n_blk * kBN2, half * kColsPerWarp, kColsPerWarp, sbias_tile,
                         &tmap_c, &tmap_aux, smem + S::kStoreOffset + ew * S::kWarpBytes, [&]() {
                           tc_fence_before();
                           __syncwarp();
                           if (lane == 0) mbar_arrive_cluster(leader_addr(smem_u32(&tmem_empty[acc])));
                         }, S::kAuxBufOff);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    }
  }

  // teardown: nobody may exit (or free TMEM) while the peer can still touch this CTA's smem/TMEM
  if (warp >= 2 && lane == 0) tma_store_wait<0>();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<kTmemCols>(tmem_base);
  }
}

}  // namespace clipa
