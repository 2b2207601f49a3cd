"""Deterministic synthetic weights and inputs in the reference's state_dict schema.

TEST INFRASTRUCTURE (see oracle/clip_oracle.py header).  The same (cfg, seed) produces the same
tensors here, in the authoring container where the reference consumes them to make the golden
vectors, and on the GPU box where the CUDA path and the oracle consume them.  Standard deviations
follow TextTransformer.init_parameters (open_clip/transformer.py:599-616); biases and LayerNorm
affine parameters are made non-trivial on purpose so that every term of the path is exercised.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch


def sincos_2d(width: int, grid: int) -> torch.Tensor:
    """Fixed 2-D sin-cos position embedding with a zero CLS row (open_clip/pos_embed.py:20-67).

    First half of the channels encodes the column index ('w goes first', pos_embed.py:27-28),
    second half the row index; each half is [sin(pos*omega), cos(pos*omega)], omega_k = 10000^(-k/(d/2)).
    """
    assert width % 4 == 0
    quarter = width // 4
    omega = 1.0 / (10000.0 ** (np.arange(quarter, dtype=np.float64) / quarter))
    ys, xs = np.meshgrid(np.arange(grid, dtype=np.float64), np.arange(grid, dtype=np.float64), indexing="ij")

    def enc(pos):
        out = pos.reshape(-1, 1) * omega.reshape(1, -1)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)

    emb = np.concatenate([enc(xs), enc(ys)], axis=1)
    emb = np.concatenate([np.zeros((1, width)), emb], axis=0)
    return torch.from_numpy(emb).float()


def make_state_dict(cfg: dict, seed: int, image_size: int = None, pos_embed: str = "learnable") -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    v, t, E = cfg["vision_cfg"], cfg["text_cfg"], cfg["embed_dim"]
    sd: Dict[str, torch.Tensor] = {}

    def blocks(prefix, width, layers, mlp_ratio=4.0):
        proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
        attn_std = width ** -0.5
        fc_std = (2 * width) ** -0.5
        hidden = int(width * mlp_ratio)
        for i in range(layers):
            p = f"{prefix}resblocks.{i}."
            sd[p + "ln_1.weight"] = 1.0 + rn(width, std=0.1)
            sd[p + "ln_1.bias"] = rn(width, std=0.1)
            sd[p + "attn.in_proj_weight"] = rn(3 * width, width, std=attn_std)
            sd[p + "attn.in_proj_bias"] = rn(3 * width, std=0.02)
            sd[p + "attn.out_proj.weight"] = rn(width, width, std=proj_std)
            sd[p + "attn.out_proj.bias"] = rn(width, std=0.02)
            sd[p + "ln_2.weight"] = 1.0 + rn(width, std=0.1)
            sd[p + "ln_2.bias"] = rn(width, std=0.1)
            sd[p + "mlp.c_fc.weight"] = rn(hidden, width, std=fc_std)
            sd[p + "mlp.c_fc.bias"] = rn(hidden, std=0.02)
            sd[p + "mlp.c_proj.weight"] = rn(width, hidden, std=proj_std)
            sd[p + "mlp.c_proj.bias"] = rn(width, std=0.02)

    # ---- vision tower
    W, P = v["width"], v["patch_size"]
    size = image_size or v["image_size"]
    grid = size // P
    sd["visual.class_embedding"] = rn(W, std=W ** -0.5)
    if pos_embed == "sin_cos_2d":
        sd["visual.positional_embedding"] = sincos_2d(W, grid)
    else:
        sd["visual.positional_embedding"] = rn(grid * grid + 1, W, std=W ** -0.5)
    sd["visual.proj"] = rn(W, E, std=W ** -0.5)
    sd["visual.conv1.weight"] = rn(W, 3, P, P, std=(3 * P * P) ** -0.5)
    if v.get("ln_pre", True):
        sd["visual.ln_pre.weight"] = 1.0 + rn(W, std=0.1)
        sd["visual.ln_pre.bias"] = rn(W, std=0.1)
    blocks("visual.transformer.", W, v["layers"], v.get("mlp_ratio", 4.0))
    sd["visual.ln_post.weight"] = 1.0 + rn(W, std=0.1)
    sd["visual.ln_post.bias"] = rn(W, std=0.1)
    # ---- text tower
    TW = t["width"]
    sd["positional_embedding"] = rn(t["context_length"], TW, std=0.01)
    sd["text_projection"] = rn(TW, E, std=TW ** -0.5)
    sd["token_embedding.weight"] = rn(t["vocab_size"], TW, std=0.02)
    blocks("transformer.", TW, t["layers"])
    sd["ln_final.weight"] = 1.0 + rn(TW, std=0.1)
    sd["ln_final.bias"] = rn(TW, std=0.1)
    sd["logit_scale"] = torch.tensor(math.log(1 / 0.07))
    return sd


def make_inputs(cfg: dict, batch: int, seed: int, image_size: int = None):
    """images ~ N(0,1) [B,3,S,S]; text ids uniform in [1, vocab-2] with the LAST position set to
    vocab-1 so argmax pooling picks it like a real EOT (open_clip/model.py:254)."""
    g = torch.Generator().manual_seed(seed)
    size = image_size or cfg["vision_cfg"]["image_size"]
    vocab = cfg["text_cfg"]["vocab_size"]
    L = cfg["text_cfg"]["context_length"]
    images = torch.randn(batch, 3, size, size, generator=g)
    text = torch.randint(1, vocab - 1, (batch, L), generator=g, dtype=torch.int64)
    text[:, -1] = vocab - 1
    return images, text


# Small architectures for CPU-second parity runs (head_dim 64 / 80 like the real models).
TINY_CONFIGS = {
    # CLS pooling, learnable pos-emb, causal text, erf GELU -- the open_clip default path
    "tiny-cls": {"embed_dim": 64,
                 "vision_cfg": {"image_size": 64, "layers": 2, "width": 128, "patch_size": 16},
                 "text_cfg": {"context_length": 16, "vocab_size": 512, "width": 128, "heads": 2, "layers": 2}},
    # GAP pooling (incl. CLS), head_width 80 (ViT-H style), short text
    "tiny-gap-h80": {"embed_dim": 128,
                     "vision_cfg": {"image_size": 48, "layers": 2, "width": 160, "head_width": 80,
                                    "patch_size": 16, "global_average_pool": True},
                     "text_cfg": {"context_length": 8, "vocab_size": 512, "width": 128, "heads": 2, "layers": 2}},
    # BigVision flavour: tanh GELU, no ln_pre, GAP without CLS, bidirectional text, last-token pool
    "tiny-bigvision": {"embed_dim": 64,
                       "vision_cfg": {"image_size": 64, "layers": 2, "width": 128, "patch_size": 16,
                                      "gelu_approximate": "tanh", "ln_pre": False,
                                      "pool_style": "big_vision_gap", "global_average_pool": True},
                       "text_cfg": {"context_length": 16, "vocab_size": 512, "width": 128, "heads": 2,
                                    "layers": 2, "gelu_approximate": "tanh",
                                    "pool_style": "big_vision_last", "attention_mask": False}},
}

# BASELINE.json configs[0]: ViT-B/32 at 192 px (36 patches + CLS), 16 text tokens, batch 8.
CONFIG1 = {"embed_dim": 512,
           "vision_cfg": {"image_size": 224, "layers": 12, "width": 768, "patch_size": 32},
           "text_cfg": {"context_length": 16, "vocab_size": 49408, "width": 512, "heads": 8, "layers": 12}}


# BASELINE.json configs[1..4] with their real widths / head shapes / sequence lengths and REDUCED DEPTH
# (so that the reference finishes on CPU in seconds and the fixtures stay small), batch 32:
#   name -> (cfg, image_size, pos_embed)
def _baseline_dims(vw, vheadw, patch, tw, theads, ctx, embed, vlayers=2, tlayers=2, **vis):
    v = {"image_size": 224, "layers": vlayers, "width": vw, "patch_size": patch}
    if vheadw != 64:
        v["head_width"] = vheadw
    v.update(vis)
    return {"embed_dim": embed, "vision_cfg": v,
            "text_cfg": {"context_length": ctx, "vocab_size": 49408, "width": tw, "heads": theads, "layers": tlayers}}


BASELINE_DIM_CASES = {
    # configs[2]: ViT-L/14, 126 px -> 81 patches + CLS = 82 tokens (16 x 64 heads), 16 text tokens
    "vitl14-i81-t16-d2": (_baseline_dims(1024, 64, 14, 768, 12, 16, 768), 126, "sin_cos_2d"),
    # configs[3]: ViT-L/14 fine-tune, 224 px -> 257 tokens (3 attention tiles), 32 text tokens, learnable pos
    "vitl14-i256-t32-d2": (_baseline_dims(1024, 64, 14, 768, 12, 32, 768), 224, "learnable"),
    # configs[4]: ViT-H/14 (head_dim 80), 84 px -> 37 tokens, 8 text tokens, GAP
    "vith14-i36-t8-d2": (_baseline_dims(1280, 80, 14, 1024, 16, 8, 1024, global_average_pool=True), 84, "sin_cos_2d"),
    # configs[1]: ViT-B/16, 128 px -> 65 tokens, 16 text tokens
    "vitb16-i64-t16-d3": (_baseline_dims(768, 64, 16, 512, 8, 16, 512, vlayers=3, tlayers=3), 128, "sin_cos_2d"),
}
BASELINE_DIM_BATCH = 32
