"""CPU oracle: a plain-math restatement of the CLIPA training-step hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under clipa_b200/ may import this module; it is used by
tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs as the
checker.  It is written with elementary torch CPU tensor ops (matmul, exp, erf, mean) -- no
nn.MultiheadAttention, F.layer_norm, F.cross_entropy or autograd-free shortcuts -- so that each
line can be read against the reference line it restates.  Parity status: PINNED against outputs of
the reference itself (oracle/make_golden.py imports /root/reference/clipa_torch in the authoring
container and commits the vectors under tests/golden/); the reference ships no tests of its own
(SURVEY.md section 4).

All functions take the reference's own state_dict key schema (open_clip/model.py CLIP):
  visual.conv1.weight, visual.class_embedding, visual.positional_embedding, visual.ln_pre.*,
  visual.transformer.resblocks.{i}.{ln_1,attn.in_proj_*,attn.out_proj,ln_2,mlp.c_fc,mlp.c_proj},
  visual.ln_post.*, visual.proj, token_embedding.weight, positional_embedding,
  transformer.resblocks.{i}.*, ln_final.*, text_projection, logit_scale.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

Tensor = torch.Tensor

# Timing legs (bench.py cpu_baseline / --impl reference) set this to evaluate LayerNorm, GELU and the
# attention core with torch's fused CPU primitives -- the same calls the reference itself makes
# (F.layer_norm, nn.GELU, SDPA inside nn.MultiheadAttention) -- so that the CPU baseline is not
# handicapped by the elementwise restatement.  Parity tests keep it False; tests/test_oracle_golden.py
# checks that both settings agree.
USE_FUSED = False


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """open_clip/transformer.py:19-34 (F.layer_norm): biased variance, eps inside the sqrt."""
    if USE_FUSED:
        return torch.nn.functional.layer_norm(x, (x.shape[-1],), w, b, eps)
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def activation(x: Tensor, kind: str) -> Tensor:
    """nn.GELU(approximate=none|tanh) (open_clip/model.py:128-129) / QuickGELU (transformer.py:37-40)."""
    if USE_FUSED and kind in ("none", "gelu", "gelu_erf", "tanh", "gelu_tanh"):
        return torch.nn.functional.gelu(x, approximate="tanh" if "tanh" in kind else "none")
    if kind in ("none", "gelu", "gelu_erf"):
        return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))
    if kind in ("tanh", "gelu_tanh"):
        return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))
    if kind == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)
    raise ValueError(kind)


def attention(h: Tensor, w_in: Tensor, b_in: Tensor, w_out: Tensor, b_out: Tensor, heads: int,
              mask: Optional[Tensor]) -> Tensor:
    """nn.MultiheadAttention(need_weights=False) as called at open_clip/transformer.py:234-236.

    h: [N, L, D].  Packed in-projection rows 0:D = Q, D:2D = K, 2D:3D = V; head i owns columns
    i*hd:(i+1)*hd; scores scaled by 1/sqrt(hd); additive mask [L, L]; dropout 0."""
    N, L, D = h.shape
    hd = D // heads
    qkv = h @ w_in.t() + b_in
    q, k, v = qkv.split(D, dim=-1)
    q = q.reshape(N, L, heads, hd).transpose(1, 2)
    k = k.reshape(N, L, heads, hd).transpose(1, 2)
    v = v.reshape(N, L, heads, hd).transpose(1, 2)
    if USE_FUSED:
        o = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=mask)
        return o.transpose(1, 2).reshape(N, L, D) @ w_out.t() + b_out
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if mask is not None:
        s = s + mask
    s = s - s.max(-1, keepdim=True).values
    p = torch.exp(s)
    p = p / p.sum(-1, keepdim=True)
    o = (p @ v).transpose(1, 2).reshape(N, L, D)
    return o @ w_out.t() + b_out


def residual_block(x: Tensor, sd: Dict[str, Tensor], pre: str, heads: int, mask: Optional[Tensor],
                   act: str) -> Tensor:
    """ResidualAttentionBlock.forward, open_clip/transformer.py:238-250 (ls_1/ls_2 = Identity)."""
    h = layer_norm(x, sd[pre + "ln_1.weight"], sd[pre + "ln_1.bias"])
    x = x + attention(h, sd[pre + "attn.in_proj_weight"], sd[pre + "attn.in_proj_bias"],
                      sd[pre + "attn.out_proj.weight"], sd[pre + "attn.out_proj.bias"], heads, mask)
    h = layer_norm(x, sd[pre + "ln_2.weight"], sd[pre + "ln_2.bias"])
    f = h @ sd[pre + "mlp.c_fc.weight"].t() + sd[pre + "mlp.c_fc.bias"]
    g = activation(f, act)
    return x + g @ sd[pre + "mlp.c_proj.weight"].t() + sd[pre + "mlp.c_proj.bias"]


def _act_kind(cfg_part: dict, quick_gelu: bool) -> str:
    return "quick_gelu" if quick_gelu else cfg_part.get("gelu_approximate", "none")


def encode_image(images: Tensor, sd: Dict[str, Tensor], cfg: dict) -> Tensor:
    """VisionTransformer.forward, open_clip/transformer.py:480-534 (no patch dropout, no attn pool)."""
    v = cfg["vision_cfg"]
    P = v["patch_size"]
    width = v["width"]
    heads = width // v.get("head_width", 64)
    N, C, Hh, Ww = images.shape
    gh, gw = Hh // P, Ww // P
    # conv1 with stride = kernel = P and no bias == patchify + matmul (transformer.py:371, 491-493)
    patches = images.reshape(N, C, gh, P, gw, P).permute(0, 2, 4, 1, 3, 5).reshape(N, gh * gw, C * P * P)
    x = patches @ sd["visual.conv1.weight"].reshape(width, -1).t()
    cls = sd["visual.class_embedding"].reshape(1, 1, width).expand(N, 1, width)
    x = torch.cat([cls, x], dim=1) + sd["visual.positional_embedding"]
    if v.get("ln_pre", True):
        x = layer_norm(x, sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"])
    act = _act_kind(v, cfg.get("quick_gelu", False))
    for i in range(v["layers"]):
        x = residual_block(x, sd, f"visual.transformer.resblocks.{i}.", heads, None, act)
    style = v.get("pool_style", "open_clip")
    gap = v.get("global_average_pool", False)
    if style == "open_clip":          # transformer.py:514-516, _global_pool :472-478
        pooled = x.mean(1) if gap else x[:, 0]
        pooled = layer_norm(pooled, sd["visual.ln_post.weight"], sd["visual.ln_post.bias"])
    elif style == "big_vision_tok":   # :517-520
        pooled = layer_norm(x, sd["visual.ln_post.weight"], sd["visual.ln_post.bias"])[:, 0]
    elif style == "big_vision_gap":   # :521-524
        pooled = layer_norm(x[:, 1:].mean(1), sd["visual.ln_post.weight"], sd["visual.ln_post.bias"])
    else:
        raise ValueError(style)
    return pooled @ sd["visual.proj"]


def encode_text(text: Tensor, sd: Dict[str, Tensor], cfg: dict) -> Tensor:
    """CLIP.encode_text, open_clip/model.py:242-263."""
    t = cfg["text_cfg"]
    L = t["context_length"]
    assert text.shape[1] == L, "reference adds the un-sliced positional embedding (model.py:247)"
    x = sd["token_embedding.weight"][text] + sd["positional_embedding"]
    mask = None
    if t.get("attention_mask", True):   # transformer.py:618-624
        mask = torch.full((L, L), float("-inf"), dtype=x.dtype).triu(1)
    act = _act_kind(t, cfg.get("quick_gelu", False))
    for i in range(t["layers"]):
        x = residual_block(x, sd, f"transformer.resblocks.{i}.", t["heads"], mask, act)
    x = layer_norm(x, sd["ln_final.weight"], sd["ln_final.bias"])
    style = t.get("pool_style", "open_clip")
    if style == "open_clip":
        pooled = x[torch.arange(x.shape[0]), text.argmax(dim=-1)]
    elif style == "big_vision_tok":
        pooled = x[:, 0]
    elif style == "big_vision_last":
        pooled = x[:, -1]
    else:
        raise ValueError(style)
    return pooled @ sd["text_projection"]


def l2_normalize(x: Tensor, eps: float = 1e-12) -> Tensor:
    """F.normalize(dim=-1), open_clip/model.py:240,263."""
    return x / x.norm(dim=-1, keepdim=True).clamp_min(eps)


def clip_forward(images: Tensor, text: Tensor, sd: Dict[str, Tensor], cfg: dict):
    """CLIP.forward, open_clip/model.py:265-274 -> (image_features, text_features, exp(logit_scale))."""
    return (l2_normalize(encode_image(images, sd, cfg)), l2_normalize(encode_text(text, sd, cfg)),
            sd["logit_scale"].exp())


def cross_entropy_rows(logits: Tensor, labels: Tensor) -> Tensor:
    """F.cross_entropy(reduction='mean'): mean_i (logsumexp_j logits_ij - logits_i,label_i)."""
    m = logits.max(-1, keepdim=True).values
    lse = (logits - m).exp().sum(-1).log() + m.squeeze(-1)
    return (lse - logits[torch.arange(logits.shape[0]), labels]).mean()


def clip_loss(img_local: Tensor, txt_local: Tensor, img_all: Tensor, txt_all: Tensor,
              logit_scale: Tensor, rank: int) -> Tensor:
    """ClipLoss.forward with local_loss=True (open_clip/loss.py:118-120,135-136,150-155).

    `*_all` are the world_size-concatenated features (== the local ones when world_size is 1, where
    this reduces to loss.py:141-142)."""
    bl = img_local.shape[0]
    logits_i = logit_scale * img_local @ txt_all.t()
    logits_t = logit_scale * txt_local @ img_all.t()
    labels = torch.arange(bl) + bl * rank
    return (cross_entropy_rows(logits_i, labels) + cross_entropy_rows(logits_t, labels)) / 2


def train_step_loss(images: Tensor, text: Tensor, sd: Dict[str, Tensor], cfg: dict) -> Tensor:
    """Single-process forward + loss of train_one_epoch (training/train.py:203-212)."""
    fi, ft, s = clip_forward(images, text, sd, cfg)
    return clip_loss(fi, ft, fi, ft, s, 0)
