"""Architecture registry.  The reference keeps one JSON per model under open_clip/model_configs/
(factory.py:26-75); here the ViT + text-transformer configs on the CLIPA path are built from a
size table, and `add_model_config(path)` registers extra JSON files / directories with the same
schema ({"embed_dim", "vision_cfg", "text_cfg"}) exactly like factory.add_model_config.

Names follow the reference: ViT-{S,M,B,L,H,g,bigG}-{patch}[-CL{ctx}][-GAP][-BigVision] ...
"""
from __future__ import annotations

import json
import re
from copy import deepcopy
from pathlib import Path
from typing import Dict, Optional

# vision: (width, layers, head_width, mlp_ratio) ; text: (width, heads, layers) ; embed_dim
_SIZES = {
    "S": ((384, 12, 64, 4.0), (384, 6, 12), 384),
    "M": ((512, 12, 64, 4.0), (512, 8, 12), 512),
    "B": ((768, 12, 64, 4.0), (512, 8, 12), 512),
    "L": ((1024, 24, 64, 4.0), (768, 12, 12), 768),
    "H": ((1280, 32, 80, 4.0), (1024, 16, 24), 1024),
    "g": ((1408, 40, 88, 4.3637), (1024, 16, 24), 1024),
    "bigG": ((1664, 48, 104, 4.9231), (1280, 20, 32), 1280),
}


def _vit(size: str, patch: int, image_size: int = 224, ctx: int = 77, gap: bool = False,
         bigvision: bool = False, text_mask: Optional[str] = None) -> dict:
    (vw, vl, hw, mlp), (tw, th, tl), embed = _SIZES[size]
    vision = {"image_size": image_size, "layers": vl, "width": vw, "patch_size": patch}
    if hw != 64:
        vision["head_width"] = hw
    if mlp != 4.0:
        vision["mlp_ratio"] = mlp
    text = {"context_length": ctx, "vocab_size": 49408, "width": tw, "heads": th, "layers": tl}
    if gap:
        vision["global_average_pool"] = True
    if text_mask:
        text["text_mask"] = text_mask
    if bigvision:  # ViT-*-CL32-GAP-BigVision.json: tanh GELU, no ln_pre, GAP w/o CLS, BERT vocab, no causal mask
        vision.update({"gelu_approximate": "tanh", "ln_pre": False, "pool_style": "big_vision_gap",
                       "global_average_pool": True})
        text.update({"vocab_size": 32000, "bert_tokenizer": True, "gelu_approximate": "tanh",
                     "pool_style": "big_vision_last", "attention_mask": False})
    return {"embed_dim": embed, "vision_cfg": vision, "text_cfg": text}


def _builtin() -> Dict[str, dict]:
    c: Dict[str, dict] = {}
    for size, patches in {"S": (16, 32), "M": (16, 32), "B": (16, 32), "L": (14, 16), "H": (14, 16),
                          "g": (14,), "bigG": (14,)}.items():
        for p in patches:
            c[f"ViT-{size}-{p}"] = _vit(size, p)
            for ctx in (8, 16, 32):
                c[f"ViT-{size}-{p}-CL{ctx}"] = _vit(size, p, ctx=ctx)
                c[f"ViT-{size}-{p}-CL{ctx}-GAP"] = _vit(size, p, ctx=ctx, gap=True)
                c[f"ViT-{size}-{p}-CL{ctx}-Syntax-GAP"] = _vit(size, p, ctx=ctx, gap=True, text_mask="syntax")
                c[f"ViT-{size}-{p}-CL{ctx}-SyntaxMask-GAP"] = _vit(size, p, ctx=ctx, gap=True, text_mask="syntax")
            c[f"ViT-{size}-{p}-CL32-GAP-BigVision"] = _vit(size, p, ctx=32, bigvision=True)
    c["ViT-L-14-280"] = _vit("L", 14, image_size=280)
    c["ViT-L-14-336"] = _vit("L", 14, image_size=336)
    c["ViT-L-16-320"] = _vit("L", 16, image_size=320)
    c["ViT-B-32-quickgelu"] = dict(_vit("B", 32), quick_gelu=True)
    return c


_MODEL_CONFIGS: Dict[str, dict] = _builtin()
_EXTRA_PATHS = []


def _natural_key(s: str):
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", s.lower())]


def add_model_config(path) -> None:
    """Register a JSON file or a directory of JSON files (factory.py:63-68)."""
    path = Path(path)
    _EXTRA_PATHS.append(path)
    files = [path] if path.is_file() else sorted(path.glob("*.json"))
    for cf in files:
        with open(cf, "r") as f:
            cfg = json.load(f)
        if all(k in cfg for k in ("embed_dim", "vision_cfg", "text_cfg")):
            _MODEL_CONFIGS[cf.stem] = cfg


def list_models():
    return sorted(_MODEL_CONFIGS.keys(), key=_natural_key)


def get_model_config(model_name: str):
    cfg = _MODEL_CONFIGS.get(model_name)
    return deepcopy(cfg) if cfg is not None else None
