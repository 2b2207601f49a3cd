/*
 * clipa_b200 — C ABI of the B200 (sm_100a) kernels behind the CLIPA training-step hot path.
 *
 * The reference (UCSC-VLAA/CLIPA, clipa_torch/open_clip) has no FFI: its hot path is Python that
 * calls torch ops.  Each entry point below replaces the torch call sites named beside it
 * (paths relative to /root/reference/clipa_torch).  The Python host (clipa_b200/open_clip/*)
 * binds these with ctypes; see INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with `h_`; the library never
 *     allocates, frees or retains caller memory (workspaces are passed in);
 *   - `stream` is a cudaStream_t (passed as void*); all work is enqueued on it, nothing syncs;
 *   - return value: 0 = ok, <0 = error (see clipa_status); clipa_last_error() gives a message
 *     for the calling thread;
 *   - matrices are row-major with an explicit leading dimension in ELEMENTS;
 *   - bf16 = __nv_bfloat16 bit pattern (uint16_t), f32 = IEEE float.
 */
#ifndef CLIPA_B200_H_
#define CLIPA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLIPA_B200_ABI_VERSION 2

typedef enum clipa_status {
  CLIPA_OK = 0,
  CLIPA_ERR_BAD_ARG = -1,     /* null pointer, non-positive dim, misaligned ld/pointer            */
  CLIPA_ERR_UNSUPPORTED = -2, /* shape/dtype outside what the sm_100a kernels implement           */
  CLIPA_ERR_CUDA = -3,        /* a CUDA runtime/driver call failed (message has the CUDA string)  */
  CLIPA_ERR_NO_DEVICE = -4    /* no sm_100 device / driver entry point unavailable                */
} clipa_status;

typedef enum clipa_dtype { CLIPA_BF16 = 0, CLIPA_F32 = 1 } clipa_dtype;

/* Storage order of a GEMM operand.  A is logically [M,K], B is logically [N,K]; C = A * B^T.
 *   CLIPA_MAJOR_K : element (r,k) at ptr[r*ld + k]   (torch `x @ W.T` with W [out,in])
 *   CLIPA_MAJOR_MN: element (r,k) at ptr[k*ld + r]   (the transposed view, read in place) */
typedef enum clipa_major { CLIPA_MAJOR_K = 0, CLIPA_MAJOR_MN = 1 } clipa_major;

typedef enum clipa_act {
  CLIPA_ACT_GELU_ERF = 0,  /* nn.GELU(approximate='none')  open_clip/model.py:128-129 */
  CLIPA_ACT_GELU_TANH = 1, /* nn.GELU(approximate='tanh')  (BigVision configs)         */
  CLIPA_ACT_QUICK_GELU = 2 /* x*sigmoid(1.702x)            open_clip/transformer.py:37-40 */
} clipa_act;

typedef enum clipa_epilogue {
  /* C = alpha*acc (+bias[n]) (+residual[m,n]); C bf16 or f32 */
  CLIPA_EPI_STORE = 0,
  /* f = acc + bias; C = act(f) (bf16); if aux != NULL also writes f (or, with aux_is_derivative, act'(f)) to aux */
  CLIPA_EPI_BIAS_ACT = 1,
  /* C = acc * act'(aux[m,n])  (bf16), aux = saved pre-activation; with aux_is_derivative C = acc * aux[m,n] */
  CLIPA_EPI_DACT = 2,
  /* C (f32) += alpha*acc with red.global.add (split-K capable; caller zero-fills or accumulates) */
  CLIPA_EPI_ATOMIC_F32 = 3
} clipa_epilogue;

typedef struct clipa_gemm_desc {
  int32_t M, N, K;
  const void* A; int64_t lda; int32_t a_major; /* bf16 */
  const void* B; int64_t ldb; int32_t b_major; /* bf16 */
  void* C;       int64_t ldc; int32_t c_dtype; /* clipa_dtype */
  int32_t epilogue;                            /* clipa_epilogue */
  float alpha;
  const void* bias; int32_t bias_dtype;        /* [N] or NULL */
  const void* residual; int64_t ldr;           /* bf16 [M,N] or NULL (STORE only) */
  void* aux; int64_t ldaux;                    /* bf16 [M,N] (see epilogue) */
  int32_t act;                                 /* clipa_act */
  int32_t split_k;                             /* 0/1 = none; >1 only with ATOMIC_F32; -1 = auto */
  int32_t max_ctas;                            /* 0 = one persistent CTA per SM */
  int32_t aux_is_derivative;                   /* BIAS_ACT: aux receives act'(f) instead of f; DACT: aux holds act'(f) */
} clipa_gemm_desc;

/* ---- library ---------------------------------------------------------------------------- */
int clipa_abi_version(void);
const char* clipa_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py: gpu_launches). */
int64_t clipa_launch_count(void);

/* ---- dense GEMM family (tcgen05.mma, TMA-fed, TMEM accumulators) ---------------------------
 * Replaces every F.linear / `@` on the path:
 *   MultiheadAttention in/out projection  open_clip/transformer.py:209,234-236
 *   mlp.c_fc / mlp.c_proj                 open_clip/transformer.py:216-220,249
 *   `pooled @ self.proj`, `x @ self.text_projection`  transformer.py:529, model.py:254-260
 * and their autograd backward (dgrad: B MN-major; wgrad: A and B MN-major, split-K atomic). */
int clipa_gemm(const clipa_gemm_desc* desc, void* stream);
/* Kernel selection for clipa_gemm: 0 = auto (2-CTA cta_group::2 pair tiles for large problems),
 * 1 = always the 1-CTA kernel, 2 = always the 2-CTA kernel where it is built.  Process-wide; meant
 * for tests and A/B measurements. */
int clipa_set_gemm_mode(int mode);

/* ---- LayerNorm ------------------------------------------------------------------------------
 * F.layer_norm via LayerNorm/LayerNormFp32 (open_clip/transformer.py:19-34): fp32 statistics,
 * eps inside the sqrt, affine.  x,y bf16 [rows, D] contiguous; gamma/beta f32 [D];
 * mean/rstd f32 [rows] are saved for backward (may be NULL in fwd). */
int clipa_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean,
                        float* rstd, int64_t rows, int32_t D, float eps, void* stream);
/* dx = (dres ? dres : 0) + LN'(dy); dgamma/dbeta (f32 [D]) are ACCUMULATED (+=).  dxsum (f32 [D], may be NULL):
 * also += the column sums of the output dx -- the bias gradient of the Linear layer that fed the residual
 * stream there (attn.out_proj / mlp.c_proj, open_clip/transformer.py:246-249), fused into this pass. */
int clipa_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean,
                        const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta,
                        float* dxsum, int64_t rows, int32_t D, void* stream);

/* ---- multi-head self-attention core ---------------------------------------------------------
 * The SDPA inside nn.MultiheadAttention(need_weights=False) (open_clip/transformer.py:234-236):
 * O = softmax(Q K^T / sqrt(hd) [+ causal mask]) V, per (sample, head).
 * qkv bf16 [batch*L, 3*D] (the packed in-projection output: columns [0,D)=Q, [D,2D)=K, [2D,3D)=V,
 * head h at columns h*hd..), out bf16 [batch*L, D]; lse f32 [batch, heads, L] (natural-log
 * sum-exp of the scaled scores, saved for backward).  causal != 0 reproduces the text tower's
 * additive -inf upper-triangular mask (open_clip/transformer.py:618-624). */
int clipa_attention_fwd(const void* qkv, void* out, float* lse, int32_t batch, int32_t L,
                        int32_t heads, int32_t head_dim, int32_t causal, void* stream);
/* Backward of the same call site (autograd of F.multi_head_attention_forward's SDPA): dqkv bf16
 * [batch*L, 3*D] receives dQ | dK | dV.  Sequences longer than one tile (L > 128, the 224/336-px
 * fine-tune stages) sum the dQ partials of their key tiles through a caller-provided fp32
 * workspace of clipa_attention_bwd_workspace() bytes (0 for one-tile shapes and for shapes whose
 * dQ tiles all stay in tensor memory, e.g. L = 257 at head_dim 64: pass NULL, 0). */
int64_t clipa_attention_bwd_workspace(int32_t batch, int32_t L, int32_t heads, int32_t head_dim);
int clipa_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                        void* dqkv, void* workspace, int64_t workspace_bytes, int32_t batch, int32_t L,
                        int32_t heads, int32_t head_dim, int32_t causal, void* stream);
/* Kernel selection for the attention core: 0 = auto (tcgen05 one-tile kernels for head_dim 64 and
 * L <= 128, tcgen05 flash kernels for any L with head_dim 64 / 80, mma.sync kernels for head_dim
 * 96 / 128), 1 = mma.sync kernels, 2 = flash kernels wherever head_dim is 64 / 80.  Process-wide;
 * tests and A/B measurements. */
int clipa_set_attention_mode(int mode);

/* ---- column sums (bias gradients): out[n] += sum_m x[m,n]; x bf16 [rows, N], out f32 -------- */
int clipa_colsum_accum(const void* x, int64_t ldx, float* out, int64_t rows, int32_t N,
                       void* stream);

/* ---- contrastive head -------------------------------------------------------------------------
 * ClipLoss.forward (open_clip/loss.py:128-157) with local_loss semantics: the caller passes the
 * LOCAL rows `a` [B_local, E] and the GATHERED other modality `b_all` [B_global, E] (bf16,
 * L2-normalised); logits = scale * a @ b_all^T are never materialised:
 *   clipa_clip_lse        : per-row log-sum-exp and the label logit (label = row + label_offset)
 *                           -> lse[B_local], diag[B_local] (f32, natural log domain)
 *                           workspace: f32 [2 * n_chunks * B_local], n_chunks from
 *                           clipa_clip_lse_workspace().
 *   clipa_clip_softmax_grad: Pt[i,j] = exp(scale*a_i.b_j - lse_i) - [j == i+label_offset] (bf16
 *                           [B_local, B_global]; multiplied by `scale` when scale_output != 0) and
 *                           dscale_partial += sum_ij Pt[i,j]*(a_i.b_j).
 * `scale` is the exponentiated logit scale (CLIP.forward returns logit_scale.exp(), model.py:273).  If
 * `scale_dev` is non-NULL it points to that f32 scalar in DEVICE memory and `scale` is ignored: the
 * training step then never reads the parameter back to the host.
 * The caller (ClipLoss autograd function) turns these into loss, d a, d b_all via clipa_gemm. */
int64_t clipa_clip_lse_workspace(int32_t b_local, int32_t b_global);
int clipa_clip_lse(const void* a, const void* b_all, int32_t b_local, int32_t b_global, int32_t E,
                   float scale, const float* scale_dev, int32_t label_offset, float* lse, float* diag,
                   float* workspace, void* stream);
int clipa_clip_softmax_grad(const void* a, const void* b_all, int32_t b_local, int32_t b_global,
                            int32_t E, float scale, const float* scale_dev, int32_t scale_output,
                            int32_t label_offset, const float* lse, void* pt, int64_t ldpt,
                            float* dscale_partial, void* stream);

/* ---- fused AdamW step ("next" row 8f.1) -----------------------------------------------------------
 * torch.optim.AdamW as built at training/main.py:318-326 (decoupled weight decay, bias correction),
 * over ONE flat fp32 segment of parameters that share a weight-decay value:
 *   g' = grad_scale*g;  p *= 1 - lr*wd;  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;
 *   p -= (lr / (1-b1^step)) * m / (sqrt(v)/sqrt(1-b2^step) + eps)
 * In the same pass it refreshes the bf16 shadow copy the GEMMs read (param_bf16, may be NULL) and,
 * if zero_grad != 0, clears the gradient segment for the next step.  n must be a multiple of 4 and
 * every buffer 16-byte aligned; `step` counts from 1.  grad_scale_dev (may be NULL): f32 scalar in device
 * memory multiplied into grad_scale inside the kernel -- the clip_grad_norm_ factor of train.py:277-283
 * without reading the gradient norm back to the host. */
int clipa_adamw_step(void* param, void* grad, void* exp_avg, void* exp_avg_sq, void* param_bf16, int64_t n,
                     float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                     float grad_scale, const float* grad_scale_dev, int32_t zero_grad, void* stream);

/* ---- tower heads and tails ("next" rows 8f.2 / 8f.4) ----------------------------------------------------
 * The HBM-bound passes around the transformer blocks, one kernel each (forward and backward).
 *
 * clipa_preprocess_u8: images uint8 [n,3,H,W] -> bf16 (x/255 - mean[c]) / std[c]; h_mean3 / h_std3 are HOST
 *   pointers to 3 floats.  Replaces `images.float().div(255)` + transforms.Normalize + cast of the
 *   --to-float-on-device recipe, training/train.py:191-197.
 * clipa_patchify: images (bf16 or f32, clipa_dtype) [n,3,H,W] -> patch rows bf16 [n*(H/ph)*(W/pw), k_padded],
 *   element (c,py,px) of a patch at column c*ph*pw + py*pw + px, columns >= 3*ph*pw zero.  With
 *   conv1.weight.reshape(width, 3*ph*pw) this turns the stride==kernel conv of VisionTransformer.forward
 *   (open_clip/transformer.py:371,491-493) into one clipa_gemm.
 * clipa_assemble_tokens: x[n,0,:] = cls + pos[0]; x[n,1+g,:] = tok[n*G+g,:] + pos[1+g,:]  (transformer.py:495-499);
 *   tok bf16 [n*(L-1), W], cls f32 [W], pos f32 [L, W], x bf16 [n, L, W].  _bwd: dtok = dx[:,1:,:] (the cls / pos
 *   gradients are column sums of dx: clipa_colsum_accum with ldx = L*W).
 * clipa_embed_tokens: x[n,l,:] = table[ids[n,l],:] + pos[l,:] (open_clip/model.py:245-247); ids int64 [n,L],
 *   table f32 [vocab, W], pos f32 [>=L, W].  _bwd: dtable[ids[n,l],:] += dx[n,l,:] (fp32 red.add).
 * clipa_pool_tokens: out[n,:] = x[n,idx,:] with mode 0 first token (CLS, transformer.py:517), 1 last token
 *   (big_vision_last), 2 first position of max(ids[n,:]) (EOT pooling, model.py:254), or the token mean with mode 3
 *   (all tokens, transformer.py:515) / 4 (without the first, big_vision_gap :521-524).  _bwd writes the whole dx.
 * clipa_l2_normalize: y = x / max(||x||_2, 1e-12) per row (F.normalize, model.py:240,263); y has row pitch ldy
 *   so it can be the rank's slice of the all-gather buffer; inv_norm f32 [rows] is kept for _bwd:
 *   dx = inv_norm * (dy - y (y.dy)), dy bf16 or f32. */
int clipa_preprocess_u8(const void* images_u8, void* out_bf16, int64_t n, int32_t height, int32_t width,
                        const float* h_mean3, const float* h_std3, void* stream);
int clipa_patchify(const void* images, int32_t image_dtype, void* patches, int64_t n, int32_t height, int32_t width,
                   int32_t patch_h, int32_t patch_w, int32_t k_padded, void* stream);
int clipa_assemble_tokens(const void* tok, const float* cls, const float* pos, void* x, int64_t n, int32_t L,
                          int32_t W, void* stream);
int clipa_assemble_tokens_bwd(const void* dx, void* dtok, int64_t n, int32_t L, int32_t W, void* stream);
int clipa_embed_tokens(const int64_t* ids, const float* table, const float* pos, void* x, int64_t n, int32_t L,
                       int32_t W, int32_t vocab, void* stream);
int clipa_embed_tokens_bwd(const int64_t* ids, const void* dx, float* dtable, int64_t n, int32_t L, int32_t W,
                           int32_t vocab, void* stream);
int clipa_pool_tokens(const void* x, const int64_t* ids, void* out, int64_t n, int32_t L, int32_t W, int32_t mode,
                      void* stream);
int clipa_pool_tokens_bwd(const void* dout, const int64_t* ids, void* dx, int64_t n, int32_t L, int32_t W,
                          int32_t mode, void* stream);
int clipa_l2_normalize(const void* x, void* y, int64_t ldy, float* inv_norm, int64_t rows, int32_t E, void* stream);
int clipa_l2_normalize_bwd(const void* x, const float* inv_norm, const void* dy, int32_t dy_dtype, void* dx,
                           int64_t rows, int32_t E, void* stream);

/* ---- fp32 parity path (precision='fp32'; north_star "1e-5 (fp32)") ----------------------------------------
 * The same call sites as above on CUDA-core kernels with fp32 storage and fp32 FMA arithmetic (tcgen05 has no
 * fp32 MMA).  Parity only -- not a throughput path.  All pointers f32.
 *   clipa_gemm_f32: C[m,n] (=|+=) alpha * sum_k A[m*rsa + k*csa] * B[n*rsb + k*csb] (+ bias[n]) (+ residual[m,n])
 *                   (element strides: transposed operands are read in place)
 *   clipa_act_f32:  out = act(x) or, with derivative != 0, dy * act'(x)               (clipa_act)
 *   clipa_layernorm_f32_fwd/_bwd: F.layer_norm and its backward (dgamma / dbeta accumulated)
 *   clipa_attention_f32_fwd/_bwd: same contract as clipa_attention_fwd/_bwd with f32 qkv / out / dqkv
 *   clipa_colsum_f32: out[n] += sum_m x[m,n]
 *   clipa_row_lse_f32: lse[m] = logsumexp_n logits[m,n], diag[m] = logits[m, m + label_offset]   (ClipLoss,
 *                   open_clip/loss.py:152-155 over materialised logits)
 *   clipa_softmax_grad_f32: pt = softmax(logits) - onehot; dscale += sum pt * logits / scale (scale in device memory) */
int clipa_gemm_f32(int32_t M, int32_t N, int32_t K, const float* A, int64_t rsa, int64_t csa, const float* B,
                   int64_t rsb, int64_t csb, float* C, int64_t ldc, float alpha, const float* bias,
                   const float* residual, int64_t ldr, int32_t accumulate, void* stream);
int clipa_act_f32(const float* x, const float* dy, float* out, int64_t n, int32_t act, int32_t derivative, void* stream);
int clipa_layernorm_f32_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                            int64_t rows, int32_t D, float eps, void* stream);
int clipa_layernorm_f32_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                            float* dx, float* dgamma, float* dbeta, int64_t rows, int32_t D, void* stream);
int clipa_attention_f32_fwd(const float* qkv, float* out, float* lse, int32_t batch, int32_t L, int32_t heads,
                            int32_t head_dim, int32_t causal, void* stream);
int clipa_attention_f32_bwd(const float* qkv, const float* out, const float* dout, const float* lse, float* dqkv,
                            int32_t batch, int32_t L, int32_t heads, int32_t head_dim, int32_t causal, void* stream);
int clipa_colsum_f32(const float* x, int64_t ldx, float* out, int64_t rows, int32_t N, void* stream);
/* out[n] += sum_k v[k] * W[k*ld + n]: the V third of d(in_proj_bias) = d(out_proj.bias) . W_out (bf16 path too) */
int clipa_gemv_f32_accum(const float* v, const float* W, int64_t ld, float* out, int32_t K, int32_t N, void* stream);
int clipa_row_lse_f32(const float* logits, int64_t ld, int32_t M, int32_t N, int32_t label_offset, float* lse,
                      float* diag, void* stream);
int clipa_softmax_grad_f32(const float* logits, int64_t ld, int32_t M, int32_t N, int32_t label_offset,
                           const float* lse, const float* scale_dev, float* pt, float* dscale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLIPA_B200_H_ */
