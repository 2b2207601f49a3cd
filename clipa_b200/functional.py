"""torch.autograd.Functions whose forward AND backward are the sm_100a kernels (via clipa_b200.ops).

Activations are bf16, token-major: a tower holds its residual stream as one [batch*L, D] matrix
(sample-major rows), so every projection is a single GEMM over all tokens and the attention
kernel finds a head's Q/K/V as strided 128-byte row segments of the packed QKV matrix.

Parameter handling: "compute" copies of matmul weights are bf16.  In pure-bf16 precision the
parameters already are bf16; with fp32 master weights (amp_bf16 / fp32 flags) a bf16 shadow is
refreshed whenever the parameter's version counter changes.  Weight gradients are produced in fp32
by the split-K wgrad GEMM and returned in the parameter's dtype.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from ._lib import EPI_ATOMIC_F32, EPI_BIAS_ACT, EPI_DACT, EPI_STORE

_BF16 = torch.bfloat16
P_TILDE_BYTES = 512 << 20      # cap of one materialised softmax-gradient block of the contrastive head (per direction)


def compute_copy(p: torch.Tensor) -> torch.Tensor:
    """bf16 view of a parameter for the GEMMs (cached on the tensor, keyed by its version)."""
    if p.dtype == _BF16:
        return p.detach()
    cache = getattr(p, "_clipa_bf16", None)
    if cache is None:
        cache = (p._version, p.detach().to(_BF16))
        p._clipa_bf16 = cache
    elif cache[0] != p._version:
        # refresh IN PLACE: the shadow may be a view into the flat buffer that the fused optimizer
        # kernel keeps up to date (clipa_b200.training.TrainStep)
        cache[1].copy_(p.detach())
        cache = (p._version, cache[1])
        p._clipa_bf16 = cache
    return cache[1]


def _f32(p: torch.Tensor) -> torch.Tensor:
    """LayerNorm parameters are consumed in fp32 (reference keeps them fp32 in every mode)."""
    return p.detach() if p.dtype == torch.float32 else p.detach().float()


def _bias(p: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if p is None:
        return None
    return p.detach() if p.dtype in (_BF16, torch.float32) else p.detach().float()


def grad_sink(p: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """The fp32 `.grad` buffer of a parameter that opted into direct accumulation (TrainStep sets
    `p._clipa_direct_grad` on parameters whose `.grad` is a view of its flat, pre-zeroed gradient
    buffer).  Every gradient kernel of this library ACCUMULATES (split-K `red.add`, atomics), so it can
    write there directly and the autograd node returns None for that input: no zero-fill of a
    temporary and no AccumulateGrad add per parameter (~900 tiny launches and ~5 GB per ViT-L/14 step)."""
    if p is None or not getattr(p, "_clipa_direct_grad", False):
        return None
    g = p.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.shape != p.shape:
        return None
    return g


def _wgrad(dy: torch.Tensor, x: torch.Tensor, like: torch.Tensor, direct: bool = True) -> Optional[torch.Tensor]:
    """dW[out,in] = dy[M,out]^T @ x[M,in]: both operands consumed MN-major, split-K, fp32 atomics.
    Returns None when the result was accumulated straight into `like.grad` (see grad_sink)."""
    sink = grad_sink(like) if direct else None
    dw = sink if sink is not None else torch.zeros(dy.shape[1], x.shape[1], dtype=torch.float32, device=dy.device)
    ops.gemm(dy.t(), x.t(), dw, epilogue=EPI_ATOMIC_F32, split_k=-1)
    if sink is not None:
        return None
    return dw if like.dtype == torch.float32 else dw.to(like.dtype)


def _bgrad(dy: torch.Tensor, like: torch.Tensor) -> Optional[torch.Tensor]:
    sink = grad_sink(like)
    db = sink if sink is not None else torch.zeros(dy.shape[1], dtype=torch.float32, device=dy.device)
    ops.colsum_accum(dy, db)
    if sink is not None:
        return None
    return db if like.dtype == torch.float32 else db.to(like.dtype)


def _in_proj_bias_grad(dqkv: torch.Tensor, b_in: torch.Tensor, db_out: torch.Tensor, w_out: torch.Tensor, D: int):
    """d(in_proj_bias) = column sums of dqkv = [sum dQ | sum dK | sum dV], with two thirds of the pass removed:
      * sum_j dK_j = sum_i (sum_j dS_ij) Q_i = 0 exactly: rows of dS = P o (dP - delta) sum to delta - delta (a constant
        added to every key of a sample does not change its softmax), so the K part is left untouched;
      * sum_j dV_j = sum_i (sum_j P_ij) dO_i = sum_i dO_i, and dO = dx1 @ W_out, so the V part is the tiny product
        d(out_proj.bias) @ W_out (fp32, CUDA-core kernel) of a vector LayerNorm-backward has already produced;
      * only the Q third needs a pass over dqkv."""
    sink = grad_sink(b_in)
    buf = sink if sink is not None else torch.zeros(3 * D, dtype=torch.float32, device=dqkv.device)
    ops.colsum_accum(dqkv[:, :D], buf[:D])
    ops.gemv_f32_accum(db_out, _f32(w_out).contiguous(), buf[2 * D:])          # += d_b_out[k] * W_out[k, :]
    if sink is not None:
        return None
    return buf if b_in.dtype == torch.float32 else buf.to(b_in.dtype)


def _bias_buf(bias: torch.Tensor):
    """(fp32 accumulation target, direct): the parameter's flat .grad buffer when it opted in, else fresh zeros."""
    sink = grad_sink(bias)
    if sink is not None:
        return sink, True
    return torch.zeros(bias.shape, dtype=torch.float32, device=bias.device), False


def _ln_grad_bufs(weight: torch.Tensor, bias: torch.Tensor):
    """(dgamma, dbeta, direct): accumulation targets for layernorm_bwd."""
    sg, sb = grad_sink(weight), grad_sink(bias)
    if sg is not None and sb is not None:
        return sg, sb, True
    D = weight.shape[0]
    return (torch.zeros(D, dtype=torch.float32, device=weight.device),
            torch.zeros(D, dtype=torch.float32, device=weight.device), False)


class LinearFn(torch.autograd.Function):
    """y = x @ W^T (+ b).  `weight` is [out, in] (nn.Linear) or, with transposed=True, [in, out]
    (the `x @ proj` convention of visual.proj / text_projection, open_clip/transformer.py:529)."""

    @staticmethod
    def forward(ctx, x, weight, bias, transposed):
        w = compute_copy(weight)
        wb = w.t() if transposed else w            # logical [out, in]
        y = torch.empty(x.shape[0], wb.shape[0], dtype=_BF16, device=x.device)
        ops.gemm(x, wb, y, bias=_bias(bias))
        ctx.save_for_backward(x, weight, bias)
        ctx.transposed = transposed
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias = ctx.saved_tensors
        dy = dy.contiguous()
        w = compute_copy(weight)
        wb = w.t() if ctx.transposed else w        # [out, in]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            ops.gemm(dy, wb.t(), dx)               # dx[M,in] = dy[M,out] @ W[out,in]
        if ctx.needs_input_grad[1]:
            dw = _wgrad(dy, x, weight, direct=not ctx.transposed)   # [out, in]
            if ctx.transposed:
                dw = dw.t().contiguous()
        if bias is not None and ctx.needs_input_grad[2]:
            db = _bgrad(dy, bias)
        return dx, dw, db, None


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        y, mean, rstd = ops.layernorm_fwd(x, _f32(weight), _f32(bias), eps)
        ctx.save_for_backward(x, weight, bias, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias, mean, rstd = ctx.saved_tensors
        dg, db, direct = _ln_grad_bufs(weight, bias)
        dx = ops.layernorm_bwd(dy.contiguous(), x, _f32(weight), mean, rstd, None, dg, db)
        if direct:
            return dx, None, None, None
        return dx, dg.to(weight.dtype), db.to(bias.dtype), None


class ResidualBlockFn(torch.autograd.Function):
    """ResidualAttentionBlock.forward (open_clip/transformer.py:238-250) as one autograd node.

    Saved for backward: block input x, packed qkv, attention output, x1 (post-attention residual),
    the attention log-sum-exp and the LayerNorm statistics -- 6 activation-sized tensors; with
    `save_ln` also the two LayerNorm outputs (8 tensors; chosen by Transformer when HBM allows, it
    removes two LayerNorm passes per block from backward).  The c_fc GEMM + activation is recomputed in
    backward (+11% FLOPs) instead of keeping the 4x-wide MLP activations resident, unless `keep_mlp`:
    then forward stores act(f) and act'(f) (2 x mlp_ratio more units) and backward starts at the c_proj
    gradients -- Transformer picks it per block when HBM allows (small micro-batches of a GradCache schedule).
    """

    kept_mlp_count = 0      # forwards that kept their MLP activations (tests / bench reporting)

    @staticmethod
    def forward(ctx, x, ln1_w, ln1_b, w_in, b_in, w_out, b_out, ln2_w, ln2_b, w_fc, b_fc, w_proj,
                b_proj, batch, seq, heads, causal, act, save_ln, b_proj_prev=None, skip_b_proj=False, drop_o=False,
                keep_mlp=False):
        M, D = x.shape
        dev = x.device
        h1, mean1, rstd1 = ops.layernorm_fwd(x, _f32(ln1_w), _f32(ln1_b))
        qkv = torch.empty(M, 3 * D, dtype=_BF16, device=dev)
        ops.gemm(h1, compute_copy(w_in), qkv, bias=_bias(b_in))
        if not save_ln:
            h1 = None
        o, lse = ops.attention_fwd(qkv, batch, seq, heads, causal)
        x1 = torch.empty(M, D, dtype=_BF16, device=dev)
        ops.gemm(o, compute_copy(w_out), x1, bias=_bias(b_out), residual=x)
        h2, mean2, rstd2 = ops.layernorm_fwd(x1, _f32(ln2_w), _f32(ln2_b))
        g = torch.empty(M, w_fc.shape[0], dtype=_BF16, device=dev)
        f = None
        if keep_mlp:        # (grad mode is always off inside Function.forward: the caller decides, Transformer._activation_policy)
            ResidualBlockFn.kept_mlp_count += 1
            f = torch.empty_like(g)                        # act'(c_fc(h2)), the same kernel backward would run
            ops.gemm(h2, compute_copy(w_fc), g, epilogue=EPI_BIAS_ACT, bias=_bias(b_fc), aux=f, act=act,
                     aux_is_derivative=True)
        else:
            ops.gemm(h2, compute_copy(w_fc), g, epilogue=EPI_BIAS_ACT, bias=_bias(b_fc), act=act)
        if not save_ln:
            h2 = None
        y = torch.empty(M, D, dtype=_BF16, device=dev)
        ops.gemm(g, compute_copy(w_proj), y, bias=_bias(b_proj), residual=x1)
        if f is None:
            g = None
        if drop_o:
            o = None        # recomputed from qkv in backward (Transformer._activation_policy: HBM is tight)
        ctx.save_for_backward(x, qkv, o, lse, x1, mean1, rstd1, mean2, rstd2, ln1_w, ln1_b, w_in,
                              b_in, w_out, b_out, ln2_w, ln2_b, w_fc, b_fc, w_proj, b_proj, h1, h2, g, f)
        ctx.meta = (batch, seq, heads, causal, act)
        # Bias gradients that are column sums of a LayerNorm-backward OUTPUT are produced by that kernel:
        # d(out_proj.bias) = colsum(dx1) here; d(c_proj.bias) of the PREVIOUS block = colsum(dx), the gradient this
        # block hands down -- the previous block's bias rides along as an extra input (`b_proj_prev`) and that
        # block skips its own column-sum pass (`skip_b_proj`).
        ctx.prev_bias = b_proj_prev
        ctx.skip_b_proj = skip_b_proj
        return y

    @staticmethod
    def backward(ctx, dy):
        (x, qkv, o, lse, x1, mean1, rstd1, mean2, rstd2, ln1_w, ln1_b, w_in, b_in, w_out, b_out,
         ln2_w, ln2_b, w_fc, b_fc, w_proj, b_proj, h1, h2, g, f) = ctx.saved_tensors
        batch, seq, heads, causal, act = ctx.meta
        b_prev = ctx.prev_bias
        M, D = x.shape
        dev = x.device
        dy = dy.contiguous()
        H4 = w_fc.shape[0]
        # ---- MLP: (recompute h2,) f = c_fc(h2), g = act(f)
        if h2 is None:
            h2, _, _ = ops.layernorm_fwd(x1, _f32(ln2_w), _f32(ln2_b), save_stats=False)
        if f is None:                                            # not kept by forward: recompute
            f = torch.empty(M, H4, dtype=_BF16, device=dev)      # receives act'(c_fc(h2)): the dgrad GEMM below only multiplies
            g = torch.empty(M, H4, dtype=_BF16, device=dev)
            ops.gemm(h2, compute_copy(w_fc), g, epilogue=EPI_BIAS_ACT, bias=_bias(b_fc), aux=f, act=act,
                     aux_is_derivative=True)
        d_w_proj = _wgrad(dy, g, w_proj)
        d_b_proj = None if ctx.skip_b_proj else _bgrad(dy, b_proj)     # else: the next block's LayerNorm backward did it
        del g
        df = torch.empty(M, H4, dtype=_BF16, device=dev)
        ops.gemm(dy, compute_copy(w_proj).t(), df, epilogue=EPI_DACT, aux=f, act=act, aux_is_derivative=True)
        del f
        d_w_fc = _wgrad(df, h2, w_fc)
        d_b_fc = _bgrad(df, b_fc)
        del h2
        dh2 = torch.empty(M, D, dtype=_BF16, device=dev)
        ops.gemm(df, compute_copy(w_fc).t(), dh2)
        del df
        d_ln2_w, d_ln2_b, direct2 = _ln_grad_bufs(ln2_w, ln2_b)
        # fresh buffer (not the accumulating .grad sink): this step's value also yields the V-bias gradient below
        db_out = torch.zeros(D, dtype=torch.float32, device=dev)
        dx1 = ops.layernorm_bwd(dh2, x1, _f32(ln2_w), mean2, rstd2, dy, d_ln2_w, d_ln2_b, dxsum=db_out)
        sink_out = grad_sink(b_out)
        if sink_out is not None:
            sink_out.add_(db_out)
            d_b_out = None
        else:
            d_b_out = db_out.to(b_out.dtype)
        del dh2
        # ---- attention
        if o is None:
            o, _ = ops.attention_fwd(qkv, batch, seq, heads, causal)
        d_w_out = _wgrad(dx1, o, w_out)
        do = torch.empty(M, D, dtype=_BF16, device=dev)
        ops.gemm(dx1, compute_copy(w_out).t(), do)
        dqkv = ops.attention_bwd(qkv, o, do, lse, batch, seq, heads, causal)
        del do
        if h1 is None:
            h1, _, _ = ops.layernorm_fwd(x, _f32(ln1_w), _f32(ln1_b), save_stats=False)
        d_w_in = _wgrad(dqkv, h1, w_in)
        d_b_in = _in_proj_bias_grad(dqkv, b_in, db_out, w_out, D)
        del h1
        dh1 = torch.empty(M, D, dtype=_BF16, device=dev)
        ops.gemm(dqkv, compute_copy(w_in).t(), dh1)
        del dqkv
        d_ln1_w, d_ln1_b, direct1 = _ln_grad_bufs(ln1_w, ln1_b)
        db_prev = prev_direct = None
        if b_prev is not None and ctx.needs_input_grad[19]:
            db_prev, prev_direct = _bias_buf(b_prev)
        dx = ops.layernorm_bwd(dh1, x, _f32(ln1_w), mean1, rstd1, dx1, d_ln1_w, d_ln1_b, dxsum=db_prev)
        d_b_prev = None if (db_prev is None or prev_direct) else db_prev.to(b_prev.dtype)
        g_ln1 = (None, None) if direct1 else (d_ln1_w.to(ln1_w.dtype), d_ln1_b.to(ln1_b.dtype))
        g_ln2 = (None, None) if direct2 else (d_ln2_w.to(ln2_w.dtype), d_ln2_b.to(ln2_b.dtype))
        return (dx, g_ln1[0], g_ln1[1], d_w_in, d_b_in, d_w_out, d_b_out,
                g_ln2[0], g_ln2[1], d_w_fc, d_b_fc, d_w_proj, d_b_proj,
                None, None, None, None, None, None, d_b_prev, None, None, None)


class ClipLossFn(torch.autograd.Function):
    """ClipLoss.forward (open_clip/loss.py:128-157), local-loss form: loss and all gradients are
    produced in ONE pass over the (never materialised in forward) [B_local x B_global] logits.

    image_features / text_features: LOCAL rows, bf16 [B_local, E], L2-normalised.
    all_image / all_text: gathered [B_global, E] (the same tensors when world_size == 1).
    Returns the fp32 scalar loss; saves d(local feats), d(gathered feats), d(logit_scale).
    """

    @staticmethod
    def forward(ctx, image_features, text_features, all_image, all_text, logit_scale, rank,
                need_all_grads):
        bl, E = image_features.shape
        bg = all_image.shape[0]
        off = rank * bl if bg != bl else 0
        dev = image_features.device
        # the (exponentiated) logit scale stays on the device: the head kernels read it through a pointer,
        # so the step never drains the launch queue between the towers and the head
        s_dev = logit_scale.detach().to(torch.float32).reshape(1).contiguous()
        img, txt = image_features.contiguous(), text_features.contiguous()
        aimg, atxt = all_image.contiguous(), all_text.contiguous()
        lse_i, diag_i = ops.clip_lse(img, atxt, s_dev, off)
        lse_t, diag_t = ops.clip_lse(txt, aimg, s_dev, off)
        loss = 0.5 * ((lse_i - diag_i).mean() + (lse_t - diag_t).mean())
        ctx.has_grads = any(ctx.needs_input_grad[:5])
        if not ctx.has_grads:          # eval / torch.no_grad(): no backward will follow
            return loss
        # gradients (scaled by grad_output in backward).  P~ = softmax - onehot is materialised in bf16, pre-multiplied
        # by the scale, one COLUMN BLOCK of the gathered batch at a time (<= P_TILDE_BYTES per direction; the whole
        # [B_local x B_global] matrix would be 2 x 1 GiB per GPU at B_global = 65 536): d(local) accumulates over the
        # blocks through the fp32 red.add epilogue, d(gathered) rows of a block are written once.
        ds = torch.zeros(1, dtype=torch.float32, device=dev)
        c = 0.5 / bl
        want_all = need_all_grads and (ctx.needs_input_grad[2] or ctx.needs_input_grad[3])
        cb = bg if bl * bg * 2 <= P_TILDE_BYTES else max(1024, (P_TILDE_BYTES // (2 * bl)) // 256 * 256)
        chunked = cb < bg
        d_img = (torch.zeros if chunked else torch.empty)(bl, E, dtype=torch.float32, device=dev)
        d_txt = (torch.zeros if chunked else torch.empty)(bl, E, dtype=torch.float32, device=dev)
        d_all_img = d_all_txt = None
        if want_all:
            d_all_txt = torch.empty(bg, E, dtype=torch.float32, device=dev)
            d_all_img = torch.empty(bg, E, dtype=torch.float32, device=dev)
        epi = EPI_ATOMIC_F32 if chunked else EPI_STORE
        for c0 in range(0, bg, cb):
            c1 = min(bg, c0 + cb)
            t_blk, i_blk = atxt[c0:c1], aimg[c0:c1]
            p_i = ops.clip_softmax_grad(img, t_blk, s_dev, off - c0, lse_i, ds, scale_output=True)   # s * Pi[:, c0:c1]
            p_t = ops.clip_softmax_grad(txt, i_blk, s_dev, off - c0, lse_t, ds, scale_output=True)
            ops.gemm(p_i, t_blk.t(), d_img, alpha=c, epilogue=epi, split_k=1)  # dI (+)= c*s * Pi @ T_all[block]
            ops.gemm(p_t, i_blk.t(), d_txt, alpha=c, epilogue=epi, split_k=1)  # dT (+)= c*s * Pt @ I_all[block]
            if want_all:
                ops.gemm(p_i.t(), img.t(), d_all_txt[c0:c1], alpha=c)         # dT_all[block] = c*s * Pi^T @ I
                ops.gemm(p_t.t(), txt.t(), d_all_img[c0:c1], alpha=c)         # dI_all[block] = c*s * Pt^T @ T
            del p_i, p_t
        ctx.save_for_backward(d_img, d_txt, d_all_img, d_all_txt, ds * c)
        ctx.dtypes = (image_features.dtype, logit_scale.dtype)
        return loss

    @staticmethod
    def backward(ctx, gout):
        if not ctx.has_grads:
            return (None,) * 7
        d_img, d_txt, d_all_img, d_all_txt, ds = ctx.saved_tensors
        fdt, sdt = ctx.dtypes
        g = gout.float()
        return ((d_img * g).to(fdt), (d_txt * g).to(fdt),
                None if d_all_img is None else (d_all_img * g).to(fdt),
                None if d_all_txt is None else (d_all_txt * g).to(fdt),
                (ds * g).reshape(()).to(sdt), None, None)


# ------------------------------------------------------------------------------------------------
# tower heads and tails (csrc/tower_io.cu): one kernel per pass instead of torch indexing / elementwise chains
# ------------------------------------------------------------------------------------------------
def _sink_or_zeros(p: torch.Tensor, shape=None):
    sink = grad_sink(p)
    if sink is not None:
        return sink, True
    return torch.zeros(shape or p.shape, dtype=torch.float32, device=p.device), False


class AssembleTokensFn(torch.autograd.Function):
    """[cls; patch tokens] + positional table (open_clip/transformer.py:495-499).  tok: [n*(L-1), W] bf16."""

    @staticmethod
    def forward(ctx, tok, cls, pos, n, L):
        x = ops.assemble_tokens(tok.contiguous(), _f32(cls).contiguous(), _f32(pos).contiguous(), n, L)
        ctx.save_for_backward(cls, pos)
        return x

    @staticmethod
    def backward(ctx, dx):
        cls, pos = ctx.saved_tensors
        dx = dx.contiguous()
        n, L, W = dx.shape
        dtok = ops.assemble_tokens_bwd(dx) if ctx.needs_input_grad[0] else None
        flat = dx.view(n, L * W)
        dcls = dpos = None
        if ctx.needs_input_grad[2]:          # learnable table: column sums over the samples; row 0 is also d(cls)
            buf, direct = _sink_or_zeros(pos)
            before = buf.view(-1)[:W].clone() if (direct and ctx.needs_input_grad[1]) else None
            ops.colsum_accum(flat, buf.view(-1))
            dpos = None if direct else buf.to(pos.dtype)
            if ctx.needs_input_grad[1]:
                row0 = buf.view(-1)[:W] - before if before is not None else buf.view(-1)[:W]
                cbuf, cdirect = _sink_or_zeros(cls)
                cbuf.add_(row0)
                dcls = None if cdirect else cbuf.to(cls.dtype)
        elif ctx.needs_input_grad[1]:        # fixed sin-cos table: only d(cls) = sum_n dx[n, 0, :]
            cbuf, cdirect = _sink_or_zeros(cls)
            ops.colsum_accum(flat[:, :W], cbuf)
            dcls = None if cdirect else cbuf.to(cls.dtype)
        return dtok, dcls, dpos, None, None


class EmbedTokensFn(torch.autograd.Function):
    """token_embedding(text) + positional rows (open_clip/model.py:245-247)."""

    @staticmethod
    def forward(ctx, ids, table, pos):
        ids = ids.contiguous()
        x = ops.embed_tokens(ids, _f32(table).contiguous(), _f32(pos).contiguous())
        ctx.save_for_backward(ids, table, pos)
        return x

    @staticmethod
    def backward(ctx, dx):
        ids, table, pos = ctx.saved_tensors
        dx = dx.contiguous()
        n, L, W = dx.shape
        dtable = dpos = None
        if ctx.needs_input_grad[1]:
            buf, direct = _sink_or_zeros(table)
            ops.embed_tokens_bwd(ids, dx, buf)
            dtable = None if direct else buf.to(table.dtype)
        if ctx.needs_input_grad[2]:
            buf, direct = _sink_or_zeros(pos)
            ops.colsum_accum(dx.view(n, L * W), buf.view(-1)[:L * W])
            dpos = None if direct else buf.to(pos.dtype)
        return None, dtable, dpos


class PoolTokensFn(torch.autograd.Function):
    """CLS / first / last / EOT-argmax row or token mean of [n, L, W] (transformer.py:509-529, model.py:251-262)."""

    @staticmethod
    def forward(ctx, x, mode, ids):
        x = x.contiguous()
        ids = ids.contiguous() if ids is not None else None
        ctx.mode, ctx.L = mode, x.shape[1]
        ctx.save_for_backward(ids)
        return ops.pool_tokens(x, mode, ids)

    @staticmethod
    def backward(ctx, dout):
        (ids,) = ctx.saved_tensors
        return ops.pool_tokens_bwd(dout.contiguous(), ctx.mode, ctx.L, ids), None, None


class L2NormalizeFn(torch.autograd.Function):
    """F.normalize(dim=-1) (open_clip/model.py:240,263); `out` may be the rank's slice of an all-gather buffer."""

    @staticmethod
    def forward(ctx, x, out):
        x = x.contiguous()
        y, inv = ops.l2_normalize(x, out)
        ctx.save_for_backward(x, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, inv = ctx.saved_tensors
        return ops.l2_normalize_bwd(x, inv, dy.contiguous()), None
