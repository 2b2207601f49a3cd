"""precision='fp32': the CLIPA step in fp32 storage and arithmetic on CUDA-core kernels (csrc/fp32_path.cu).

The north_star asks for parity with the reference "within 1e-5 (fp32)"; the tensor cores have no fp32 MMA, so this
mode trades all throughput for reference-exact numerics: every matmul, LayerNorm, activation, the attention core and
the contrastive cross-entropy run in fp32 through the C ABI (`clipa_*_f32`).  Layout glue that moves no arithmetic
(patch reshape, [cls; tokens] concat, positional add, pooling, F.normalize) is plain torch here -- this is a parity
mode, not a hot path.  Entered from CLIP.encode_image / encode_text / ClipLoss when the model was built with
precision='fp32'."""
from __future__ import annotations

import ctypes as C  # noqa: F401
from typing import Optional

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import check
from .ops import _ptr, _stream


def _f(t: torch.Tensor) -> torch.Tensor:
    assert t.is_cuda, "clipa_b200 kernels need CUDA tensors (there is no CPU path)"
    return t if t.dtype == torch.float32 else t.float()


def gemm_f32(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, *, alpha: float = 1.0,
             bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
             accumulate: bool = False) -> torch.Tensor:
    """out[M,N] (=|+=) alpha * a[M,K] @ b[N,K]^T (+ bias) (+ residual); a / b may be arbitrary strided views."""
    M, K = a.shape
    N, Kb = b.shape
    assert K == Kb and a.dtype == torch.float32 and b.dtype == torch.float32
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    assert out.stride(1) == 1 and out.dtype == torch.float32
    if residual is not None:
        assert residual.shape == (M, N) and residual.stride(1) == 1 and residual.dtype == torch.float32
    check(_lib.lib().clipa_gemm_f32(M, N, K, _ptr(a), a.stride(0), a.stride(1), _ptr(b), b.stride(0), b.stride(1),
                                    _ptr(out), out.stride(0), alpha, _ptr(bias), _ptr(residual),
                                    residual.stride(0) if residual is not None else 0, int(accumulate), _stream()),
          "clipa_gemm_f32")
    return out


def colsum_f32(x: torch.Tensor) -> torch.Tensor:
    out = torch.zeros(x.shape[1], dtype=torch.float32, device=x.device)
    check(_lib.lib().clipa_colsum_f32(_ptr(x), x.stride(0), _ptr(out), x.shape[0], x.shape[1], _stream()), "clipa_colsum_f32")
    return out


class LinearF32(torch.autograd.Function):
    """y = x @ W^T (+ b) (+ residual); weight [out, in], or [in, out] with transposed=True."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, transposed):
        x = x.contiguous()
        w = weight.t() if transposed else weight
        y = gemm_f32(x, w, bias=bias, residual=residual)
        ctx.save_for_backward(x, weight, bias)
        ctx.transposed = transposed
        ctx.has_res = residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias = ctx.saved_tensors
        dy = dy.contiguous()
        w = weight.t() if ctx.transposed else weight           # [out, in]
        dx = gemm_f32(dy, w.t()) if ctx.needs_input_grad[0] else None          # dy[M,out] @ W[out,in]
        dw = None
        if ctx.needs_input_grad[1]:
            dw = gemm_f32(dy.t(), x.t())                        # [out, in] = dy^T @ x
            if ctx.transposed:
                dw = dw.t().contiguous()
        db = colsum_f32(dy) if (bias is not None and ctx.needs_input_grad[2]) else None
        return dx, dw, db, (dy if ctx.has_res else None), None


class ActF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        x = x.contiguous()
        out = torch.empty_like(x)
        check(_lib.lib().clipa_act_f32(_ptr(x), None, _ptr(out), x.numel(), act, 0, _stream()), "clipa_act_f32")
        ctx.save_for_backward(x)
        ctx.act = act
        return out

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        out = torch.empty_like(x)
        check(_lib.lib().clipa_act_f32(_ptr(x), _ptr(dy), _ptr(out), x.numel(), ctx.act, 1, _stream()), "clipa_act_f32")
        return out, None


class LayerNormF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1]).contiguous()
        rows, D = x2.shape
        y = torch.empty_like(x2)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        check(_lib.lib().clipa_layernorm_f32_fwd(_ptr(x2), _ptr(weight), _ptr(bias), _ptr(y), _ptr(mean), _ptr(rstd),
                                                 rows, D, eps, _stream()), "clipa_layernorm_f32_fwd")
        ctx.save_for_backward(x2, weight, mean, rstd)
        ctx.shape = shape
        return y.reshape(shape)

    @staticmethod
    def backward(ctx, dy):
        x2, weight, mean, rstd = ctx.saved_tensors
        rows, D = x2.shape
        dy2 = dy.reshape(rows, D).contiguous()
        dx = torch.empty_like(x2)
        dg = torch.zeros(D, dtype=torch.float32, device=x2.device)
        db = torch.zeros(D, dtype=torch.float32, device=x2.device)
        check(_lib.lib().clipa_layernorm_f32_bwd(_ptr(dy2), _ptr(x2), _ptr(weight), _ptr(mean), _ptr(rstd), _ptr(dx),
                                                 _ptr(dg), _ptr(db), rows, D, _stream()), "clipa_layernorm_f32_bwd")
        return dx.reshape(ctx.shape), dg, db, None


class AttentionF32(torch.autograd.Function):
    """qkv [batch*L, 3D] fp32 -> [batch*L, D]; same contract as clipa_attention_fwd/_bwd."""

    @staticmethod
    def forward(ctx, qkv, batch, L, heads, causal):
        qkv = qkv.contiguous()
        D = qkv.shape[1] // 3
        out = torch.empty(batch * L, D, dtype=torch.float32, device=qkv.device)
        lse = torch.empty(batch, heads, L, dtype=torch.float32, device=qkv.device)
        check(_lib.lib().clipa_attention_f32_fwd(_ptr(qkv), _ptr(out), _ptr(lse), batch, L, heads, D // heads, int(causal),
                                                 _stream()), "clipa_attention_f32_fwd")
        ctx.save_for_backward(qkv, out, lse)
        ctx.meta = (batch, L, heads, causal)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse = ctx.saved_tensors
        batch, L, heads, causal = ctx.meta
        dqkv = torch.empty_like(qkv)
        check(_lib.lib().clipa_attention_f32_bwd(_ptr(qkv), _ptr(out), _ptr(dout.contiguous()), _ptr(lse), _ptr(dqkv), batch,
                                                 L, heads, qkv.shape[1] // 3 // heads, int(causal), _stream()),
              "clipa_attention_f32_bwd")
        return dqkv, None, None, None, None


class ClipLossF32(torch.autograd.Function):
    """ClipLoss.forward (open_clip/loss.py:128-157), local-loss form, fp32: logits materialised by clipa_gemm_f32."""

    @staticmethod
    def forward(ctx, img, txt, all_img, all_txt, logit_scale, rank, need_all_grads):
        bl, E = img.shape
        bg = all_img.shape[0]
        off = rank * bl if bg != bl else 0
        dev = img.device
        s = logit_scale.detach().float().reshape(1).contiguous()
        img, txt, all_img, all_txt = (t.contiguous() for t in (img, txt, all_img, all_txt))
        # the reference scales the features first (loss.py:135-136: `logit_scale * image_features @ ...`)
        si, st = img * s, txt * s
        li, lt = gemm_f32(si, all_txt), gemm_f32(st, all_img)                  # [bl, bg]
        lse_i, diag_i, lse_t, diag_t = (torch.empty(bl, dtype=torch.float32, device=dev) for _ in range(4))
        L = _lib.lib()
        check(L.clipa_row_lse_f32(_ptr(li), bg, bl, bg, off, _ptr(lse_i), _ptr(diag_i), _stream()), "clipa_row_lse_f32")
        check(L.clipa_row_lse_f32(_ptr(lt), bg, bl, bg, off, _ptr(lse_t), _ptr(diag_t), _stream()), "clipa_row_lse_f32")
        loss = 0.5 * ((lse_i - diag_i).mean() + (lse_t - diag_t).mean())
        ctx.has_grads = any(ctx.needs_input_grad[:5])
        if not ctx.has_grads:
            return loss
        ds = torch.zeros(1, dtype=torch.float32, device=dev)
        pi, pt = torch.empty_like(li), torch.empty_like(lt)
        check(L.clipa_softmax_grad_f32(_ptr(li), bg, bl, bg, off, _ptr(lse_i), _ptr(s), _ptr(pi), _ptr(ds), _stream()),
              "clipa_softmax_grad_f32")
        check(L.clipa_softmax_grad_f32(_ptr(lt), bg, bl, bg, off, _ptr(lse_t), _ptr(s), _ptr(pt), _ptr(ds), _stream()),
              "clipa_softmax_grad_f32")
        c = 0.5 / bl
        cs = float(c) * s                                                      # device scalar c * s
        d_img = gemm_f32(pi, all_txt.t()) * cs
        d_txt = gemm_f32(pt, all_img.t()) * cs
        d_all_img = d_all_txt = None
        if need_all_grads:
            d_all_txt = gemm_f32(pi.t(), img.t()) * cs
            d_all_img = gemm_f32(pt.t(), txt.t()) * cs
        ctx.save_for_backward(d_img, d_txt, d_all_img, d_all_txt, ds * c)
        return loss

    @staticmethod
    def backward(ctx, g):
        if not ctx.has_grads:
            return (None,) * 7
        d_img, d_txt, d_all_img, d_all_txt, ds = ctx.saved_tensors
        return (d_img * g, d_txt * g, None if d_all_img is None else d_all_img * g,
                None if d_all_txt is None else d_all_txt * g, (ds * g).reshape(()), None, None)


# ------------------------------------------------------------------------------------------------------------------
# towers
# ------------------------------------------------------------------------------------------------------------------
def _ln(mod, x):
    if isinstance(mod, torch.nn.Identity):
        return x
    return LayerNormF32.apply(x, _f(mod.weight), _f(mod.bias), mod.eps)


def _block(blk, x, batch, L, causal):
    """ResidualAttentionBlock.forward (open_clip/transformer.py:238-250), x [batch*L, D] fp32."""
    h = _ln(blk.ln_1, x)
    qkv = LinearF32.apply(h, _f(blk.attn.in_proj_weight), _f(blk.attn.in_proj_bias), None, False)
    o = AttentionF32.apply(qkv, batch, L, blk.n_head, causal)
    x = LinearF32.apply(o, _f(blk.attn.out_proj.weight), _f(blk.attn.out_proj.bias), x, False)
    h = _ln(blk.ln_2, x)
    f = LinearF32.apply(h, _f(blk.mlp.c_fc.weight), _f(blk.mlp.c_fc.bias), None, False)
    g = ActF32.apply(f, blk.act)
    return LinearF32.apply(g, _f(blk.mlp.c_proj.weight), _f(blk.mlp.c_proj.bias), x, False)


def encode_image(visual, images: torch.Tensor) -> torch.Tensor:
    """VisionTransformer.forward (open_clip/transformer.py:480-534) in fp32."""
    N = images.shape[0]
    ph, pw = visual.patch_size
    gh, gw = visual.grid_size
    W = visual.width
    x = _f(images)
    patches = x.reshape(N, 3, gh, ph, gw, pw).permute(0, 2, 4, 1, 3, 5).reshape(N * gh * gw, 3 * ph * pw)
    tok = LinearF32.apply(patches, _f(visual.conv1.weight).reshape(W, -1), None, None, False).reshape(N, gh * gw, W)
    cls = _f(visual.class_embedding).reshape(1, 1, W).expand(N, 1, W)
    x = torch.cat([cls, tok], dim=1) + _f(visual.positional_embedding)
    x = visual.patch_dropout(x)
    L = x.shape[1]
    x = _ln(visual.ln_pre, x).reshape(N * L, W)
    for blk in visual.transformer.resblocks:
        x = _block(blk, x, N, L, False)
    x = x.reshape(N, L, W)
    if visual.pool_style == "open_clip":
        pooled = x.mean(dim=1) if visual.global_average_pool else x[:, 0]
        pooled = _ln(visual.ln_post, pooled)
    elif visual.pool_style == "big_vision_tok":
        pooled = _ln(visual.ln_post, x)[:, 0]
    elif visual.pool_style == "big_vision_gap":
        pooled = _ln(visual.ln_post, x[:, 1:].mean(dim=1))
    else:
        raise ValueError(visual.pool_style)
    if visual.proj is not None:
        pooled = LinearF32.apply(pooled, _f(visual.proj), None, None, True)
    return pooled


def encode_text(mod, text: torch.Tensor, slice_positions: bool) -> torch.Tensor:
    """CLIP.encode_text (open_clip/model.py:245-263) / TextTransformer.forward (transformer.py:638-681) in fp32."""
    N, L = text.shape
    if not slice_positions and L != mod.context_length:
        raise ValueError(f"text length {L} != context_length {mod.context_length}")
    x = F.embedding(text, _f(mod.token_embedding.weight)) + _f(mod.positional_embedding)[:L]
    W = x.shape[-1]
    causal = mod.attn_mask is not None
    x = x.reshape(N * L, W)
    for blk in mod.transformer.resblocks:
        x = _block(blk, x, N, L, causal)
    x = _ln(mod.ln_final, x).reshape(N, L, W)
    if mod.pool_style == "open_clip":
        pooled = x[torch.arange(N, device=x.device), text.argmax(dim=-1)]
    elif mod.pool_style == "big_vision_tok":
        pooled = x[:, 0]
    elif mod.pool_style == "big_vision_last":
        pooled = x[:, -1]
    else:
        raise ValueError(mod.pool_style)
    if mod.text_projection is not None:
        pooled = LinearF32.apply(pooled, _f(mod.text_projection), None, None, True)
    return pooled
