"""Tensor-level wrappers: torch tensors in, raw device pointers + sizes out to the C ABI.

torch is used here only for memory (allocation, data_ptr) and the current CUDA stream.
Every function fails loudly if its input is not a CUDA tensor or the library is missing.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import (ACT_GELU_ERF, ACT_GELU_TANH, ACT_QUICK_GELU, BF16, EPI_ATOMIC_F32, EPI_BIAS_ACT,
                   EPI_DACT, EPI_STORE, F32, MAJOR_K, MAJOR_MN, GemmDesc, check)

ACT_BY_NAME = {"gelu": ACT_GELU_ERF, "gelu_erf": ACT_GELU_ERF, "none": ACT_GELU_ERF,
               "gelu_tanh": ACT_GELU_TANH, "tanh": ACT_GELU_TANH, "quick_gelu": ACT_QUICK_GELU}


class GemmProfiler:
    """CUDA-event timing of every kernel launch made through this module, on the launching stream
    (bench.py roofline).  Records are (key, flops, bytes, start_event, end_event); key[0] is the
    entry-point name, GEMM keys carry (M, N, K, a_major, b_major, epilogue)."""

    def __init__(self):
        self.records = []

    def summary(self):
        """GEMM launches only -> (achieved TFLOP/s, total ms, number of launches)."""
        torch.cuda.synchronize()
        g = [(f, s.elapsed_time(e)) for k, f, _, s, e in self.records if k[0] == "gemm"]
        ms = sum(t for _, t in g)
        fl = sum(f for f, _ in g)
        return (fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0), ms, len(g)

    def table(self):
        """Per-key totals: {key: dict(calls, ms, tflops, gbs)} sorted by time."""
        torch.cuda.synchronize()
        agg = {}
        for k, f, b, s, e in self.records:
            a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
            a[0] += 1; a[1] += s.elapsed_time(e); a[2] += f; a[3] += b
        rows = []
        for k, (n, ms, f, b) in agg.items():
            rows.append({"key": list(k), "calls": n, "ms": ms, "tflops": f / (ms * 1e-3) / 1e12 if ms else 0.0,
                         "gbs": b / (ms * 1e-3) / 1e9 if ms else 0.0})
        return sorted(rows, key=lambda r: -r["ms"])


PROFILER: Optional[GemmProfiler] = None


class _prof:
    """with _prof(key, flops, bytes): <launch>   -- no-op unless a profiler is installed."""

    def __init__(self, key, flops=0.0, nbytes=0.0):
        self.key, self.flops, self.nbytes = key, flops, nbytes

    def __enter__(self):
        if PROFILER is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if PROFILER is not None and exc[0] is None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILER.records.append((self.key, self.flops, self.nbytes, self.e0, e1))
        return False


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.ClipaError("clipa_b200 kernels need CUDA tensors (there is no CPU path)")
    return t.data_ptr()


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    raise _lib.ClipaError(f"unsupported dtype {t.dtype}")


def _mat(t: torch.Tensor):
    """Returns (ptr, ld, major) for a 2-D bf16 tensor that is row-major or a transposed view."""
    assert t.dim() == 2 and t.dtype == torch.bfloat16, (t.shape, t.dtype)
    if t.stride(1) == 1:
        return _ptr(t), t.stride(0), MAJOR_K
    if t.stride(0) == 1:
        return _ptr(t), t.stride(1), MAJOR_MN
    raise _lib.ClipaError(f"operand strides {t.stride()} are neither row- nor column-major")


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, epilogue: int = EPI_STORE,
         alpha: float = 1.0, bias: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None,
         act: int = ACT_GELU_ERF, split_k: int = 0, max_ctas: int = 0, aux_is_derivative: bool = False) -> torch.Tensor:
    """out[M,N] = epilogue(alpha * a[M,K] @ b[N,K]^T).  `a`/`b` may be transposed views."""
    M, K = a.shape
    N, Kb = b.shape
    assert K == Kb, (a.shape, b.shape)
    assert out.shape == (M, N) and out.stride(1) == 1
    d = GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.A, d.lda, d.a_major = _mat(a)
    d.B, d.ldb, d.b_major = _mat(b)
    d.C, d.ldc, d.c_dtype = _ptr(out), out.stride(0), _dtype_code(out)
    d.epilogue = epilogue
    d.alpha = alpha
    if bias is not None:
        assert bias.numel() == N and bias.is_contiguous()
        d.bias, d.bias_dtype = _ptr(bias), _dtype_code(bias)
    if residual is not None:
        assert residual.shape == (M, N) and residual.stride(1) == 1 and residual.dtype == torch.bfloat16
        d.residual, d.ldr = _ptr(residual), residual.stride(0)
    if aux is not None:
        assert aux.shape == (M, N) and aux.stride(1) == 1 and aux.dtype == torch.bfloat16
        d.aux, d.ldaux = _ptr(aux), aux.stride(0)
    d.act = act
    d.split_k = split_k
    d.max_ctas = max_ctas
    d.aux_is_derivative = int(aux_is_derivative)
    with _prof(("gemm", M, N, K, d.a_major, d.b_major, epilogue), 2.0 * M * N * K,
               2.0 * (M * K + N * K) + out.element_size() * M * N):
        check(_lib.lib().clipa_gemm(C.byref(d), _stream()), "clipa_gemm")
    return out


def layernorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
                  save_stats: bool = True):
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    D = x.shape[-1]
    rows = x.numel() // D
    y = torch.empty_like(x)
    mean = rstd = None
    if save_stats:
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    assert gamma.dtype == torch.float32 and beta.dtype == torch.float32
    with _prof(("ln_fwd", D), 0.0, 4.0 * rows * D):
        check(_lib.lib().clipa_layernorm_fwd(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(y), _ptr(mean),
                                             _ptr(rstd), rows, D, eps, _stream()), "clipa_layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dres, dgamma, dbeta, dxsum=None):
    """Returns dx = dres + LN'(dy); accumulates into dgamma/dbeta (fp32) and, if given, the column sums of dx
    into dxsum (fp32 [D]: a fused bias gradient)."""
    D = x.shape[-1]
    rows = x.numel() // D
    assert dy.is_contiguous() and x.is_contiguous() and (dres is None or dres.is_contiguous())
    dx = torch.empty_like(x)
    with _prof(("ln_bwd", D), 0.0, (8.0 if dres is not None else 6.0) * rows * D):
        check(_lib.lib().clipa_layernorm_bwd(_ptr(dy), _ptr(x), _ptr(gamma), _ptr(mean), _ptr(rstd),
                                             _ptr(dres), _ptr(dx), _ptr(dgamma), _ptr(dbeta), _ptr(dxsum), rows, D,
                                             _stream()), "clipa_layernorm_bwd")
    return dx


def attention_fwd(qkv: torch.Tensor, batch: int, L: int, heads: int, causal: bool):
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous()
    D3 = qkv.shape[-1]
    D = D3 // 3
    hd = D // heads
    out = torch.empty(batch * L, D, dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty(batch, heads, L, dtype=torch.float32, device=qkv.device)
    with _prof(("attn_fwd", L, hd, int(causal)), 4.0 * batch * heads * L * L * hd, 8.0 * batch * L * D):
        check(_lib.lib().clipa_attention_fwd(_ptr(qkv), _ptr(out), _ptr(lse), batch, L, heads, hd,
                                             int(causal), _stream()), "clipa_attention_fwd")
    return out, lse


def attention_bwd(qkv, out, dout, lse, batch: int, L: int, heads: int, causal: bool):
    D = out.shape[-1]
    hd = D // heads
    assert dout.is_contiguous() and out.is_contiguous() and qkv.is_contiguous()
    dqkv = torch.empty_like(qkv)
    nws = int(_lib.lib().clipa_attention_bwd_workspace(batch, L, heads, hd))
    ws = torch.empty(nws, dtype=torch.uint8, device=qkv.device) if nws else None
    with _prof(("attn_bwd", L, hd, int(causal)), 10.0 * batch * heads * L * L * hd, 16.0 * batch * L * D):
        check(_lib.lib().clipa_attention_bwd(_ptr(qkv), _ptr(out), _ptr(dout), _ptr(lse), _ptr(dqkv),
                                             _ptr(ws), nws, batch, L, heads, hd, int(causal), _stream()),
              "clipa_attention_bwd")
    return dqkv


def colsum_accum(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.bfloat16
    assert out.dtype == torch.float32 and out.numel() == x.shape[1]
    with _prof(("colsum", x.shape[1]), 0.0, 2.0 * x.shape[0] * x.shape[1]):
        check(_lib.lib().clipa_colsum_accum(_ptr(x), x.stride(0), _ptr(out), x.shape[0], x.shape[1],
                                            _stream()), "clipa_colsum_accum")
    return out


def _scale_args(scale):
    """(host float, device pointer): a tensor scale stays on the device (no .item() sync)."""
    if isinstance(scale, torch.Tensor):
        assert scale.is_cuda and scale.dtype == torch.float32 and scale.numel() == 1
        return 0.0, _ptr(scale)
    return float(scale), None


def clip_lse(a: torch.Tensor, b_all: torch.Tensor, scale, label_offset: int):
    """Row log-sum-exp and label logit of scale * a @ b_all^T, logits never materialised.
    `scale`: python float, or a 1-element fp32 CUDA tensor (read on the device)."""
    assert a.dtype == torch.bfloat16 and b_all.dtype == torch.bfloat16
    assert a.is_contiguous() and b_all.is_contiguous()
    bl, E = a.shape
    bg = b_all.shape[0]
    nws = _lib.lib().clipa_clip_lse_workspace(bl, bg)
    ws = torch.empty(nws, dtype=torch.float32, device=a.device)
    lse = torch.empty(bl, dtype=torch.float32, device=a.device)
    diag = torch.empty(bl, dtype=torch.float32, device=a.device)
    s_host, s_dev = _scale_args(scale)
    with _prof(("clip_lse", bl, bg, E), 2.0 * bl * bg * E, 2.0 * (bl + bg) * E):
        check(_lib.lib().clipa_clip_lse(_ptr(a), _ptr(b_all), bl, bg, E, s_host, s_dev, label_offset, _ptr(lse),
                                        _ptr(diag), _ptr(ws), _stream()), "clipa_clip_lse")
    return lse, diag


def clip_softmax_grad(a, b_all, scale, label_offset: int, lse: torch.Tensor,
                      dscale_partial: torch.Tensor, scale_output: bool = False):
    """Pt = softmax(scale * a @ b_all^T) - onehot (bf16 [bl, bg]); with scale_output the stored matrix is
    scale * Pt, so the gradient GEMMs that consume it need no logit-scale factor from the host."""
    bl, E = a.shape
    bg = b_all.shape[0]
    ld = (bg + 7) // 8 * 8    # row pitch must stay 16-byte aligned for TMA consumers
    pt = torch.empty(bl, ld, dtype=torch.bfloat16, device=a.device)[:, :bg]
    s_host, s_dev = _scale_args(scale)
    with _prof(("clip_softmax_grad", bl, bg, E), 2.0 * bl * bg * E, 2.0 * (bl + bg) * E + 2.0 * bl * bg):
        check(_lib.lib().clipa_clip_softmax_grad(_ptr(a), _ptr(b_all), bl, bg, E, s_host, s_dev, int(scale_output),
                                                 label_offset, _ptr(lse), _ptr(pt), pt.stride(0),
                                                 _ptr(dscale_partial), _stream()), "clipa_clip_softmax_grad")
    return pt


def adamw_step(param: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor,
               param_bf16: Optional[torch.Tensor], *, lr: float, beta1: float, beta2: float, eps: float,
               weight_decay: float, step: int, grad_scale: float = 1.0, zero_grad: bool = True,
               grad_scale_dev: Optional[torch.Tensor] = None) -> None:
    """Fused AdamW over flat fp32 segments (+ bf16 shadow refresh + gradient clear), in place.
    grad_scale_dev: optional 1-element fp32 CUDA tensor multiplied into grad_scale on the device
    (gradient-clipping factor without a host round trip)."""
    n = param.numel()
    for t in (param, grad, exp_avg, exp_avg_sq):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n
    if param_bf16 is not None:
        assert param_bf16.dtype == torch.bfloat16 and param_bf16.is_contiguous() and param_bf16.numel() == n
    with _prof(("adamw", n), 0.0, (34.0 if param_bf16 is not None else 32.0) * n):
        check(_lib.lib().clipa_adamw_step(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq),
                                          _ptr(param_bf16), n, lr, beta1, beta2, eps, weight_decay, step,
                                          grad_scale, _ptr(grad_scale_dev), int(zero_grad), _stream()),
              "clipa_adamw_step")


# ------------------------------------------------------------------------------------------------
# tower heads and tails (csrc/tower_io.cu)
# ------------------------------------------------------------------------------------------------
def preprocess_u8(images: torch.Tensor, mean, std, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """uint8 [n,3,H,W] (CUDA) -> normalised bf16, training/train.py:191-197."""
    assert images.dtype == torch.uint8 and images.dim() == 4 and images.shape[1] == 3 and images.is_contiguous()
    if out is None:
        out = torch.empty(images.shape, dtype=torch.bfloat16, device=images.device)
    assert out.shape == images.shape and out.dtype == torch.bfloat16 and out.is_contiguous()
    m = (C.c_float * 3)(*[float(v) for v in mean])
    s = (C.c_float * 3)(*[float(v) for v in std])
    with _prof(("preprocess_u8",), 0.0, 3.0 * images.numel()):
        check(_lib.lib().clipa_preprocess_u8(_ptr(images), _ptr(out), images.shape[0], images.shape[2], images.shape[3],
                                             m, s, _stream()), "clipa_preprocess_u8")
    return out


def patchify(images: torch.Tensor, patch_h: int, patch_w: int, k_padded: int) -> torch.Tensor:
    assert images.dim() == 4 and images.shape[1] == 3 and images.is_contiguous()
    n, _, H, W = images.shape
    rows = n * (H // patch_h) * (W // patch_w)
    out = torch.empty(rows, k_padded, dtype=torch.bfloat16, device=images.device)
    with _prof(("patchify",), 0.0, float(images.numel() * images.element_size() + 2 * out.numel())):
        check(_lib.lib().clipa_patchify(_ptr(images), _dtype_code(images), _ptr(out), n, H, W, patch_h, patch_w, k_padded,
                                        _stream()), "clipa_patchify")
    return out


def assemble_tokens(tok: torch.Tensor, cls: torch.Tensor, pos: torch.Tensor, n: int, L: int) -> torch.Tensor:
    W = tok.shape[-1]
    assert tok.dtype == torch.bfloat16 and tok.is_contiguous() and tok.numel() == n * (L - 1) * W
    assert cls.dtype == torch.float32 and pos.dtype == torch.float32 and pos.is_contiguous() and pos.shape == (L, W)
    x = torch.empty(n, L, W, dtype=torch.bfloat16, device=tok.device)
    with _prof(("assemble_tokens",), 0.0, 4.0 * x.numel()):
        check(_lib.lib().clipa_assemble_tokens(_ptr(tok), _ptr(cls), _ptr(pos), _ptr(x), n, L, W, _stream()),
              "clipa_assemble_tokens")
    return x


def assemble_tokens_bwd(dx: torch.Tensor) -> torch.Tensor:
    n, L, W = dx.shape
    assert dx.dtype == torch.bfloat16 and dx.is_contiguous()
    dtok = torch.empty(n * (L - 1), W, dtype=torch.bfloat16, device=dx.device)
    with _prof(("assemble_tokens_bwd",), 0.0, 4.0 * dtok.numel()):
        check(_lib.lib().clipa_assemble_tokens_bwd(_ptr(dx), _ptr(dtok), n, L, W, _stream()), "clipa_assemble_tokens_bwd")
    return dtok


def embed_tokens(ids: torch.Tensor, table: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
    n, L = ids.shape
    V, W = table.shape
    assert ids.dtype == torch.int64 and ids.is_contiguous()
    assert table.dtype == torch.float32 and table.is_contiguous() and pos.dtype == torch.float32 and pos.is_contiguous()
    assert pos.shape[0] >= L and pos.shape[1] == W
    x = torch.empty(n, L, W, dtype=torch.bfloat16, device=table.device)
    with _prof(("embed_tokens",), 0.0, 6.0 * x.numel()):
        check(_lib.lib().clipa_embed_tokens(_ptr(ids), _ptr(table), _ptr(pos), _ptr(x), n, L, W, V, _stream()),
              "clipa_embed_tokens")
    return x


def embed_tokens_bwd(ids: torch.Tensor, dx: torch.Tensor, dtable: torch.Tensor) -> None:
    n, L = ids.shape
    V, W = dtable.shape
    assert dx.dtype == torch.bfloat16 and dx.is_contiguous() and dtable.dtype == torch.float32 and dtable.is_contiguous()
    with _prof(("embed_tokens_bwd",), 0.0, 6.0 * dx.numel()):
        check(_lib.lib().clipa_embed_tokens_bwd(_ptr(ids), _ptr(dx), _ptr(dtable), n, L, W, V, _stream()),
              "clipa_embed_tokens_bwd")


def pool_tokens(x: torch.Tensor, mode: int, ids: Optional[torch.Tensor] = None) -> torch.Tensor:
    n, L, W = x.shape
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    if ids is not None:
        assert ids.dtype == torch.int64 and ids.is_contiguous() and ids.shape == (n, L)
    out = torch.empty(n, W, dtype=torch.bfloat16, device=x.device)
    with _prof(("pool_tokens", mode), 0.0, 2.0 * (x.numel() if mode >= 3 else out.numel()) + 2.0 * out.numel()):
        check(_lib.lib().clipa_pool_tokens(_ptr(x), _ptr(ids), _ptr(out), n, L, W, mode, _stream()), "clipa_pool_tokens")
    return out


def pool_tokens_bwd(dout: torch.Tensor, mode: int, L: int, ids: Optional[torch.Tensor] = None) -> torch.Tensor:
    n, W = dout.shape
    assert dout.dtype == torch.bfloat16 and dout.is_contiguous()
    dx = torch.empty(n, L, W, dtype=torch.bfloat16, device=dout.device)
    with _prof(("pool_tokens_bwd", mode), 0.0, 2.0 * dx.numel() + 2.0 * dout.numel()):
        check(_lib.lib().clipa_pool_tokens_bwd(_ptr(dout), _ptr(ids), _ptr(dx), n, L, W, mode, _stream()),
              "clipa_pool_tokens_bwd")
    return dx


def l2_normalize(x: torch.Tensor, out: Optional[torch.Tensor] = None):
    """Rows of x (bf16 [rows, E]) scaled to unit L2 norm; `out` may be a row-slice of a larger (gather) buffer."""
    rows, E = x.shape
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    assert out.dtype == torch.bfloat16 and out.shape == x.shape and out.stride(1) == 1
    inv = torch.empty(rows, dtype=torch.float32, device=x.device)
    with _prof(("l2_normalize",), 0.0, 4.0 * x.numel()):
        check(_lib.lib().clipa_l2_normalize(_ptr(x), _ptr(out), out.stride(0), _ptr(inv), rows, E, _stream()),
              "clipa_l2_normalize")
    return out, inv


def l2_normalize_bwd(x: torch.Tensor, inv: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    rows, E = x.shape
    assert dy.is_contiguous() and dy.shape == x.shape
    dx = torch.empty_like(x)
    with _prof(("l2_normalize_bwd",), 0.0, 6.0 * x.numel()):
        check(_lib.lib().clipa_l2_normalize_bwd(_ptr(x), _ptr(inv), _ptr(dy), _dtype_code(dy), _ptr(dx), rows, E, _stream()),
              "clipa_l2_normalize_bwd")
    return dx


def gemv_f32_accum(v: torch.Tensor, w: torch.Tensor, out: torch.Tensor) -> None:
    """out[n] += sum_k v[k] * w[k, n]  (fp32, CUDA cores; w row-major [K, N])."""
    K, N = w.shape
    assert v.dtype == w.dtype == out.dtype == torch.float32 and v.numel() == K and out.numel() == N and w.stride(1) == 1
    with _prof(("gemv_f32",), 2.0 * K * N, 4.0 * K * N):
        check(_lib.lib().clipa_gemv_f32_accum(_ptr(v), _ptr(w), w.stride(0), _ptr(out), K, N, _stream()),
              "clipa_gemv_f32_accum")
