// Host launcher + C ABI for the tcgen05 GEMM family (see gemm_tc.cuh for the kernel).
#include "gemm_tc.cuh"
#include "gemm_tc2.cuh"
#include "host_common.h"

namespace clipa {

template <int BN, bool A_MN, bool B_MN, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                       const CUtensorMap& tx, const GemmParams& p, int grid, cudaStream_t stream) {
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, EPI>;
  static bool attr_set[64] = {false};
  int dev = 0;
  CLIPA_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    CLIPA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          GemmSmem<BN>::kTotal));
    attr_set[dev] = true;
  }
  kern<<<grid, kGemmThreads, GemmSmem<BN>::kTotal, stream>>>(ta, tb, tc, tx, p);
  CLIPA_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return CLIPA_OK;
}

template <bool A_MN, bool B_MN, int EPI>
static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                        const CUtensorMap& tx, const GemmParams& p, int grid, cudaStream_t stream) {
  auto kern = gemm_tc2_kernel<A_MN, B_MN, EPI>;
  static bool attr_set[64] = {false};
  int dev = 0;
  CLIPA_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    CLIPA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Tc2Smem<EPI>::kTotal));
    attr_set[dev] = true;
  }
  kern<<<grid, kGemmThreads, Tc2Smem<EPI>::kTotal, stream>>>(ta, tb, tc, tx, p);
  CLIPA_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return CLIPA_OK;
}

// Output tensor maps for the TMA-store epilogue (bf16 C / aux, 32 x 32 boxes, 64-byte swizzle).
// When the output is not bf16 (fp32 store / atomic accumulate) the maps are unused copies of `ta`.
static int encode_out_maps(GemmParams& p, int epi, const CUtensorMap& ta, CUtensorMap* tc, CUtensorMap* tx) {
  *tc = ta;
  *tx = ta;
  p.tma_store = 0;
  const bool bf16_out = !p.c_f32 && (epi == EPI_STORE || epi == EPI_BIAS_ACT || epi == EPI_DACT);
  if (!bf16_out) return CLIPA_OK;
  int rc = encode_tmap_2d_bf16(tc, p.C, (uint64_t)p.N, (uint64_t)p.M, (uint64_t)p.ldc * 2, 32, 32, 64);
  if (rc) return rc;
  if ((epi == EPI_BIAS_ACT || epi == EPI_DACT) && p.aux) {   // aux: TMA-stored (BIAS_ACT) / TMA-loaded (DACT, 2-CTA)
    rc = encode_tmap_2d_bf16(tx, p.aux, (uint64_t)p.N, (uint64_t)p.M, (uint64_t)p.ldaux * 2, 32, 32, 64);
    if (rc) return rc;
  }
  p.tma_store = 1;
  return CLIPA_OK;
}

// 2-CTA path: 256 x 256 pair tiles (each CTA: 128 rows of A, 128 of the 256 B rows per stage).
static int gemm_dispatch_2cta(GemmParams p, const void* A, long long lda, bool a_mn, const void* B,
                              long long ldb, bool b_mn, int epi, int max_ctas, cudaStream_t stream) {
  p.m_blocks = (p.M + 2 * kBM - 1) / (2 * kBM);
  p.n_blocks = (p.N + kBN2 - 1) / kBN2;
  p.k_blocks = (p.K + kBK - 1) / kBK;
  p.n_per_chunk = 1;
  p.n_chunks = p.n_blocks;
  if (p.split_k < 1) p.split_k = 1;
  if (p.split_k > p.k_blocks) p.split_k = p.k_blocks;
  p.kb_per_split = (p.k_blocks + p.split_k - 1) / p.split_k;
  p.split_k = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;
  p.num_items = p.split_k * p.m_blocks * p.n_blocks;
  CUtensorMap ta, tb;
  int rc;
  if (!a_mn) rc = encode_tmap_2d_bf16(&ta, A, (uint64_t)p.K, (uint64_t)p.M, (uint64_t)lda * 2, kBK, kBM);
  else       rc = encode_tmap_2d_bf16(&ta, A, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)lda * 2, 64, kBK);
  if (rc) return rc;
  if (!b_mn) rc = encode_tmap_2d_bf16(&tb, B, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)ldb * 2, kBK, kBN2 / 2);
  else       rc = encode_tmap_2d_bf16(&tb, B, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)ldb * 2, 64, kBK);
  if (rc) return rc;
  CUtensorMap tc, tx;
  rc = encode_out_maps(p, epi, ta, &tc, &tx);
  if (rc) return rc;
  int clusters = num_sms() / 2;
  if (max_ctas > 1 && max_ctas / 2 < clusters) clusters = max_ctas / 2;
  if (clusters > p.num_items) clusters = p.num_items;
  const int grid = clusters * 2;
#define CLIPA_GEMM2_CASE(AMN_, BMN_, EPI_) \
  if (a_mn == AMN_ && b_mn == BMN_ && epi == EPI_) return launch_gemm2<AMN_, BMN_, EPI_>(ta, tb, tc, tx, p, grid, stream);
  CLIPA_GEMM2_CASE(false, false, EPI_STORE)
  CLIPA_GEMM2_CASE(false, true, EPI_STORE)
  CLIPA_GEMM2_CASE(true, true, EPI_STORE)
  CLIPA_GEMM2_CASE(true, false, EPI_STORE)
  CLIPA_GEMM2_CASE(false, false, EPI_BIAS_ACT)
  CLIPA_GEMM2_CASE(false, true, EPI_DACT)
  CLIPA_GEMM2_CASE(true, true, EPI_ATOMIC_F32)
  CLIPA_GEMM2_CASE(false, false, EPI_ATOMIC_F32)
  CLIPA_GEMM2_CASE(false, true, EPI_ATOMIC_F32)
  CLIPA_GEMM2_CASE(true, false, EPI_ATOMIC_F32)
#undef CLIPA_GEMM2_CASE
  return 1;  // not built for this combination -> caller falls back to the 1-CTA kernel
}

// 0 = auto (2-CTA when the problem is big enough), 1 = force 1-CTA, 2 = force 2-CTA (tests)
static int g_gemm_mode = 0;

// Builds tensor maps and derived tiling, then dispatches on (BN, majors, epilogue).
int gemm_dispatch(GemmParams p, const void* A, long long lda, bool a_mn, const void* B,
                  long long ldb, bool b_mn, int epi, int max_ctas, cudaStream_t stream) {
  CLIPA_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, CLIPA_ERR_BAD_ARG, "gemm: M,N,K must be positive (%d,%d,%d)",
                p.M, p.N, p.K);
  CLIPA_REQUIRE(A && B, CLIPA_ERR_BAD_ARG, "gemm: null operand");
  if (epi <= EPI_ATOMIC_F32 && g_gemm_mode != 1 &&
      (g_gemm_mode == 2 || (p.M >= 1024 && p.N >= 256))) {
    const int rc2 = gemm_dispatch_2cta(p, A, lda, a_mn, B, ldb, b_mn, epi, max_ctas, stream);
    if (rc2 <= 0) return rc2;
  }
  const int BN = (p.N <= 128) ? 128 : 256;
  p.m_blocks = (p.M + kBM - 1) / kBM;
  p.n_blocks = (p.N + BN - 1) / BN;
  p.k_blocks = (p.K + kBK - 1) / kBK;
  if (p.n_per_chunk <= 0) p.n_per_chunk = 1;
  p.n_chunks = (p.n_blocks + p.n_per_chunk - 1) / p.n_per_chunk;
  if (p.split_k < 1) p.split_k = 1;
  if (p.split_k > p.k_blocks) p.split_k = p.k_blocks;
  p.kb_per_split = (p.k_blocks + p.split_k - 1) / p.split_k;
  p.split_k = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;
  p.num_items = p.split_k * p.m_blocks * p.n_chunks;

  CUtensorMap ta, tb;
  int rc;
  if (!a_mn) rc = encode_tmap_2d_bf16(&ta, A, (uint64_t)p.K, (uint64_t)p.M, (uint64_t)lda * 2, kBK, kBM);
  else       rc = encode_tmap_2d_bf16(&ta, A, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)lda * 2, 64, kBK);
  if (rc) return rc;
  if (!b_mn) rc = encode_tmap_2d_bf16(&tb, B, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)ldb * 2, kBK, BN);
  else       rc = encode_tmap_2d_bf16(&tb, B, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)ldb * 2, 64, kBK);
  if (rc) return rc;

  CUtensorMap tc, tx;
  rc = encode_out_maps(p, epi, ta, &tc, &tx);
  if (rc) return rc;
  int grid = num_sms();
  if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
  if (grid > p.num_items) grid = p.num_items;

#define CLIPA_GEMM_CASE(BN_, AMN_, BMN_, EPI_)                                   \
  if (BN == BN_ && a_mn == AMN_ && b_mn == BMN_ && epi == EPI_)                  \
    return launch_gemm<BN_, AMN_, BMN_, EPI_>(ta, tb, tc, tx, p, grid, stream);
#define CLIPA_GEMM_BOTH_BN(AMN_, BMN_, EPI_) \
  CLIPA_GEMM_CASE(256, AMN_, BMN_, EPI_) CLIPA_GEMM_CASE(128, AMN_, BMN_, EPI_)

  CLIPA_GEMM_BOTH_BN(false, false, EPI_STORE)
  CLIPA_GEMM_BOTH_BN(false, true, EPI_STORE)
  CLIPA_GEMM_BOTH_BN(true, false, EPI_STORE)
  CLIPA_GEMM_BOTH_BN(true, true, EPI_STORE)
  CLIPA_GEMM_BOTH_BN(false, false, EPI_BIAS_ACT)
  CLIPA_GEMM_BOTH_BN(false, true, EPI_DACT)
  CLIPA_GEMM_BOTH_BN(false, false, EPI_ATOMIC_F32)
  CLIPA_GEMM_BOTH_BN(false, true, EPI_ATOMIC_F32)
  CLIPA_GEMM_BOTH_BN(true, false, EPI_ATOMIC_F32)
  CLIPA_GEMM_BOTH_BN(true, true, EPI_ATOMIC_F32)
  CLIPA_GEMM_BOTH_BN(false, false, EPI_LSE)
  CLIPA_GEMM_BOTH_BN(false, false, EPI_SOFTMAX_GRAD)
#undef CLIPA_GEMM_BOTH_BN
#undef CLIPA_GEMM_CASE
  set_error("gemm: unsupported combination (epilogue %d, a_major %d, b_major %d)", epi, (int)a_mn,
            (int)b_mn);
  return CLIPA_ERR_UNSUPPORTED;
}

}  // namespace clipa

using namespace clipa;

extern "C" int clipa_set_gemm_mode(int mode) {
  CLIPA_REQUIRE(mode >= 0 && mode <= 2, CLIPA_ERR_BAD_ARG, "set_gemm_mode: mode must be 0 (auto), 1 (1-CTA) or 2 (2-CTA)");
  g_gemm_mode = mode;
  return CLIPA_OK;
}

extern "C" int clipa_gemm(const clipa_gemm_desc* d, void* stream) {
  CLIPA_REQUIRE(d != nullptr, CLIPA_ERR_BAD_ARG, "gemm: null descriptor");
  CLIPA_REQUIRE(d->C != nullptr, CLIPA_ERR_BAD_ARG, "gemm: null output");
  CLIPA_REQUIRE(d->lda % 8 == 0 && d->ldb % 8 == 0, CLIPA_ERR_BAD_ARG,
                "gemm: lda/ldb must be multiples of 8 elements (got %lld, %lld)", (long long)d->lda,
                (long long)d->ldb);
  const bool c_f32 = d->c_dtype == CLIPA_F32;
  CLIPA_REQUIRE(d->ldc % (c_f32 ? 4 : 8) == 0 &&
                    (reinterpret_cast<uintptr_t>(d->C) & 15) == 0,
                CLIPA_ERR_BAD_ARG, "gemm: C must be 16-byte aligned with ldc %% %d == 0", c_f32 ? 4 : 8);
  GemmParams p{};
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.C = d->C; p.ldc = d->ldc; p.c_f32 = c_f32;
  p.alpha = d->alpha;
  p.bias = d->bias; p.bias_f32 = d->bias_dtype == CLIPA_F32;
  p.residual = static_cast<const __nv_bfloat16*>(d->residual); p.ldr = d->ldr;
  p.aux = static_cast<__nv_bfloat16*>(d->aux); p.ldaux = d->ldaux;
  p.act = d->act;
  p.aux_deriv = d->aux_is_derivative;
  p.n_per_chunk = 1;
  p.split_k = 1;
  if (d->bias) CLIPA_REQUIRE((reinterpret_cast<uintptr_t>(d->bias) & 15) == 0, CLIPA_ERR_BAD_ARG,
                             "gemm: bias must be 16-byte aligned");
  int epi;
  switch (d->epilogue) {
    case CLIPA_EPI_STORE:
      epi = EPI_STORE;
      if (d->residual)
        CLIPA_REQUIRE(d->ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(d->residual) & 15) == 0,
                      CLIPA_ERR_BAD_ARG, "gemm: residual must be 16-byte aligned with ldr %% 8 == 0");
      break;
    case CLIPA_EPI_BIAS_ACT:
      epi = EPI_BIAS_ACT;
      CLIPA_REQUIRE(!c_f32 && d->N % 32 == 0, CLIPA_ERR_UNSUPPORTED,
                    "gemm: BIAS_ACT needs bf16 output and N %% 32 == 0 (N=%d)", d->N);
      if (d->aux)
        CLIPA_REQUIRE(d->ldaux % 8 == 0 && (reinterpret_cast<uintptr_t>(d->aux) & 15) == 0,
                      CLIPA_ERR_BAD_ARG, "gemm: aux must be 16-byte aligned with ldaux %% 8 == 0");
      break;
    case CLIPA_EPI_DACT:
      epi = EPI_DACT;
      CLIPA_REQUIRE(!c_f32 && d->N % 32 == 0 && d->aux != nullptr && d->ldaux % 8 == 0 &&
                        (reinterpret_cast<uintptr_t>(d->aux) & 15) == 0,
                    CLIPA_ERR_UNSUPPORTED,
                    "gemm: DACT needs bf16 output, N %% 32 == 0 and an aligned aux (N=%d)", d->N);
      break;
    case CLIPA_EPI_ATOMIC_F32: {
      epi = EPI_ATOMIC_F32;
      CLIPA_REQUIRE(c_f32, CLIPA_ERR_UNSUPPORTED, "gemm: ATOMIC_F32 needs an f32 output");
      int sk = d->split_k;
      if (sk < 0) {
        // auto: the smallest split whose work-item count fills whole waves of persistent CTAs
        // (>= 95 % wave efficiency, >= 2 waves), keeping >= 8 k-blocks per item; every extra split
        // costs one more pass of fp32 atomics over the output.
        const int BN = (d->N <= 128) ? 128 : 256;
        const bool pair = g_gemm_mode != 1 && (g_gemm_mode == 2 || (d->M >= 1024 && d->N >= 256));
        const long long tiles = pair ? (long long)((d->M + 2 * kBM - 1) / (2 * kBM)) * ((d->N + kBN2 - 1) / kBN2)
                                     : (long long)((d->M + kBM - 1) / kBM) * ((d->N + BN - 1) / BN);
        const int kb = (d->K + kBK - 1) / kBK;
        const long long sms = pair ? num_sms() / 2 : num_sms();
        int cap = kb / 8 > 0 ? kb / 8 : 1;
        if (cap > 64) cap = 64;
        int best = 1;
        double best_eff = 0.0;
        for (int s = 1; s <= cap; ++s) {
          const long long items = tiles * s;
          const long long waves = (items + sms - 1) / sms;
          const double eff = (double)items / (double)(waves * sms);
          if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
          if (eff >= 0.95 && waves >= 2) { best = s; break; }
        }
        sk = best;
      }
      p.split_k = sk;
      break;
    }
    default:
      set_error("gemm: unknown epilogue %d", d->epilogue);
      return CLIPA_ERR_BAD_ARG;
  }
  if (epi != EPI_ATOMIC_F32)
    CLIPA_REQUIRE(d->split_k <= 1, CLIPA_ERR_UNSUPPORTED, "gemm: split_k>1 only with ATOMIC_F32");
  return gemm_dispatch(p, d->A, d->lda, d->a_major == CLIPA_MAJOR_MN, d->B, d->ldb,
                       d->b_major == CLIPA_MAJOR_MN, epi, d->max_ctas,
                       static_cast<cudaStream_t>(stream));
}
