// Contrastive head (ClipLoss, open_clip/loss.py:128-157) on top of the tcgen05 GEMM:
// the [B_local x B_global] logits live only in TMEM / registers.
//   clipa_clip_lse          : GEMM with the online log-sum-exp epilogue + a tiny merge kernel
//   clipa_clip_softmax_grad : GEMM whose epilogue emits Pt = softmax - onehot (bf16) and d(scale)
#include "gemm_tc.cuh"
#include "host_common.h"

namespace clipa {

int gemm_dispatch(GemmParams p, const void* A, long long lda, bool a_mn, const void* B,
                  long long ldb, bool b_mn, int epi, int max_ctas, cudaStream_t stream);

struct LseChunking {
  int n_blocks, n_per_chunk, n_chunks;
};
static LseChunking lse_chunking(int b_local, int b_global) {
  const int BN = (b_global <= 128) ? 128 : 256;
  LseChunking c;
  c.n_blocks = (b_global + BN - 1) / BN;
  const int m_blocks = (b_local + kBM - 1) / kBM;
  int want = (2 * num_sms() + m_blocks - 1) / m_blocks;
  if (want < 1) want = 1;
  if (want > c.n_blocks) want = c.n_blocks;
  c.n_per_chunk = (c.n_blocks + want - 1) / want;
  c.n_chunks = (c.n_blocks + c.n_per_chunk - 1) / c.n_per_chunk;
  return c;
}

// merge the per-(chunk, column-half) partials: lse = ln2 * (m + log2(sum_s l_s 2^(m_s - m)))
__global__ void lse_merge_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum,
                                 int slots, int M, float* __restrict__ lse) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  float m = -INFINITY;
  for (int s = 0; s < slots; ++s) m = fmaxf(m, part_max[(size_t)s * M + row]);
  float l = 0.f;
  for (int s = 0; s < slots; ++s)
    l += part_sum[(size_t)s * M + row] * exp2f(part_max[(size_t)s * M + row] - m);
  lse[row] = (m + log2f(l)) * 0.69314718055994531f;
}

}  // namespace clipa

using namespace clipa;

extern "C" int64_t clipa_clip_lse_workspace(int32_t b_local, int32_t b_global) {
  if (b_local <= 0 || b_global <= 0) return 0;
  const LseChunking c = lse_chunking(b_local, b_global);
  return (int64_t)2 * (2 * c.n_chunks) * b_local;
}

extern "C" int clipa_clip_lse(const void* a, const void* b_all, int32_t b_local, int32_t b_global,
                              int32_t E, float scale, const float* scale_dev, int32_t label_offset, float* lse,
                              float* diag, float* workspace, void* stream) {
  CLIPA_REQUIRE(a && b_all && lse && diag && workspace, CLIPA_ERR_BAD_ARG, "clip_lse: null pointer");
  CLIPA_REQUIRE(b_local > 0 && b_global > 0 && E > 0 && E % 8 == 0, CLIPA_ERR_UNSUPPORTED,
                "clip_lse: need E %% 8 == 0 (b_local=%d b_global=%d E=%d)", b_local, b_global, E);
  CLIPA_REQUIRE(label_offset >= 0 && label_offset + b_local <= b_global, CLIPA_ERR_BAD_ARG,
                "clip_lse: labels [%d, %d) fall outside the %d gathered columns", label_offset,
                label_offset + b_local, b_global);
  const LseChunking c = lse_chunking(b_local, b_global);
  const int slots = 2 * c.n_chunks;
  GemmParams p{};
  p.M = b_local; p.N = b_global; p.K = E;
  p.n_per_chunk = c.n_per_chunk;
  p.split_k = 1;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.scale_dev = scale_dev;
  p.label_offset = label_offset;
  p.part_max = workspace;
  p.part_sum = workspace + (size_t)slots * b_local;
  p.diag = diag;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = gemm_dispatch(p, a, E, false, b_all, E, false, EPI_LSE, 0, s);
  if (rc) return rc;
  lse_merge_kernel<<<(b_local + 255) / 256, 256, 0, s>>>(p.part_max, p.part_sum, slots, b_local, lse);
  CLIPA_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return CLIPA_OK;
}

extern "C" int clipa_clip_softmax_grad(const void* a, const void* b_all, int32_t b_local,
                                       int32_t b_global, int32_t E, float scale, const float* scale_dev,
                                       int32_t scale_output, int32_t label_offset, const float* lse,
                                       void* pt, int64_t ldpt, float* dscale_partial, void* stream) {
  CLIPA_REQUIRE(a && b_all && lse && pt && dscale_partial, CLIPA_ERR_BAD_ARG,
                "clip_softmax_grad: null pointer");
  CLIPA_REQUIRE(b_local > 0 && b_global > 0 && E > 0 && E % 8 == 0 && ldpt % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(pt) & 15) == 0,
                CLIPA_ERR_UNSUPPORTED, "clip_softmax_grad: need E %% 8 == 0, ldpt %% 8 == 0, aligned pt");
  GemmParams p{};
  p.M = b_local; p.N = b_global; p.K = E;
  p.n_per_chunk = 1;
  p.split_k = 1;
  p.C = pt; p.ldc = ldpt;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.scale_dev = scale_dev;
  p.scale_out = scale_output;
  p.label_offset = label_offset;
  p.lse = lse;
  p.dscale = dscale_partial;
  return gemm_dispatch(p, a, E, false, b_all, E, false, EPI_SOFTMAX_GRAD, 0,
                       static_cast<cudaStream_t>(stream));
}
