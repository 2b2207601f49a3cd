"""CLIP model with the reference's public surface (open_clip/model.py:200-274): same constructor
arguments, attributes, parameter names and return conventions; the towers run on the B200
kernels (clipa_b200.functional)."""
from __future__ import annotations

import logging
import math
from dataclasses import dataclass
from typing import Optional, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .. import fp32_path
from .. import functional as Fn
from .transformer import text_tower_forward  # noqa: E402
from .transformer import LayerNorm, TextTransformer, VisionTransformer


@dataclass
class CLIPVisionCfg:
    """Fields of open_clip/model.py:25-50 that the ViT path uses (others are accepted and must stay
    at their defaults: timm / attentional-pool / patchnorm towers are out of scope)."""
    layers: Union[Tuple[int, int, int, int], int] = 12
    width: int = 768
    head_width: int = 64
    mlp_ratio: float = 4.0
    patch_size: int = 16
    image_size: Union[Tuple[int, int], int] = 224
    ls_init_value: Optional[float] = None
    patch_dropout: float = 0.
    input_patchnorm: bool = False
    global_average_pool: bool = False
    attentional_pool: bool = False
    n_queries: int = 256
    attn_pooler_heads: int = 8
    timm_model_name: str = None
    timm_model_pretrained: bool = False
    timm_pool: str = 'avg'
    timm_proj: str = 'linear'
    timm_proj_bias: bool = False
    timm_drop: float = 0.
    timm_drop_path: Optional[float] = None
    output_tokens: bool = False
    pos_embed: str = 'learnable'
    gelu_approximate: str = 'none'
    ln_pre: bool = True
    pool_style: str = 'open_clip'


@dataclass
class CLIPTextCfg:
    """open_clip/model.py:53-75."""
    context_length: int = 77
    vocab_size: int = 49408
    width: int = 512
    heads: int = 8
    layers: int = 12
    ls_init_value: Optional[float] = None
    hf_model_name: str = None
    hf_tokenizer_name: str = None
    hf_model_pretrained: bool = True
    proj: str = 'mlp'
    pooler_type: str = 'mean_pooler'
    embed_cls: bool = False
    pad_id: int = 0
    output_tokens: bool = False
    text_mask: str = 'first'
    gelu_approximate: str = 'none'
    pool_style: str = 'open_clip'
    bert_tokenizer: bool = False
    vocab_path: str = None
    attention_mask: bool = True


def get_cast_dtype(precision: str):
    """open_clip/model.py:78-86."""
    if precision == 'bf16':
        return torch.bfloat16
    if precision == 'fp16':
        raise NotImplementedError("fp16 is not built: the sm_100a kernels are bf16 (tcgen05 kind::f16 with bf16 operands)")
    return None


def _act_name(quick_gelu: bool, gelu_approximate: str) -> str:
    if quick_gelu:
        return "quick_gelu"
    return {"none": "gelu", "tanh": "gelu_tanh"}[gelu_approximate]


def _build_vision_tower(embed_dim, vision_cfg, quick_gelu=False, cast_dtype=None):
    if isinstance(vision_cfg, dict):
        vision_cfg = CLIPVisionCfg(**vision_cfg)
    if vision_cfg.timm_model_name or isinstance(vision_cfg.layers, (tuple, list)) or \
            vision_cfg.attentional_pool or vision_cfg.input_patchnorm:
        raise NotImplementedError("only the ViT tower of the CLIPA hot path is built (no timm / ResNet / "
                                  "attentional pool / patchnorm)")
    return VisionTransformer(
        image_size=vision_cfg.image_size, patch_size=vision_cfg.patch_size, width=vision_cfg.width,
        layers=vision_cfg.layers, heads=vision_cfg.width // vision_cfg.head_width,
        mlp_ratio=vision_cfg.mlp_ratio, ls_init_value=vision_cfg.ls_init_value,
        patch_dropout=vision_cfg.patch_dropout, global_average_pool=vision_cfg.global_average_pool,
        output_tokens=vision_cfg.output_tokens, output_dim=embed_dim,
        act=_act_name(quick_gelu, vision_cfg.gelu_approximate), pos_embed=vision_cfg.pos_embed,
        ln_pre=vision_cfg.ln_pre, pool_style=vision_cfg.pool_style)


def _build_text_tower(embed_dim, text_cfg, quick_gelu=False, cast_dtype=None):
    if isinstance(text_cfg, dict):
        text_cfg = CLIPTextCfg(**text_cfg)
    if text_cfg.hf_model_name:
        raise NotImplementedError("HF text towers are out of scope")
    return TextTransformer(
        context_length=text_cfg.context_length, vocab_size=text_cfg.vocab_size, width=text_cfg.width,
        heads=text_cfg.heads, layers=text_cfg.layers, ls_init_value=text_cfg.ls_init_value,
        output_dim=embed_dim, embed_cls=text_cfg.embed_cls, pad_id=text_cfg.pad_id,
        act=_act_name(quick_gelu, text_cfg.gelu_approximate), pool_style=text_cfg.pool_style,
        attention_mask=text_cfg.attention_mask)


def _reserve_for_later(model, text_tower, image, text) -> None:
    """Tell each tower's activation policy what still has to fit after its own forward: the text tower runs first
    and must leave room for the vision tower's default footprint and the contrastive head; the vision tower for the
    head (`head_reserve_bytes`, set by TrainStep from the global batch; 1 GiB otherwise)."""
    head = int(getattr(model, "head_reserve_bytes", 1 << 30))
    vt = model.visual.transformer
    gh, gw = model.visual.grid_size
    rows = image.shape[0] * (gh * gw + 1)
    text_tower.other_need_bytes = vt.base_need_bytes(rows) + head
    # the (small) text tower keeps its MLP activations only if the vision tower can still keep all of its own
    text_tower.keep_reserve_bytes = vt.full_need_bytes(rows) - vt.base_need_bytes(rows)
    vt.other_need_bytes = head


def _wait_ready(image) -> None:
    """Image batches staged on a copy stream carry their ready events (TrainStep._stage_host_images)."""
    evs = getattr(image, "_clipa_ready", None)
    if evs:
        cur = torch.cuda.current_stream()
        for ev in evs:
            cur.wait_event(ev)


def _l2_normalize(x: torch.Tensor) -> torch.Tensor:
    """F.normalize(dim=-1) (open_clip/model.py:240,263); the norm is taken in fp32, the result is
    rounded once to the feature dtype (bf16) that the contrastive-head GEMM consumes."""
    if x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 2 and x.shape[1] % 8 == 0:
        return Fn.L2NormalizeFn.apply(x, None)
    return F.normalize(x.float(), dim=-1).to(x.dtype)


class CLIP(nn.Module):
    """open_clip/model.py:200-274."""

    def __init__(self, embed_dim: int, vision_cfg: CLIPVisionCfg, text_cfg: CLIPTextCfg,
                 quick_gelu: bool = False, cast_dtype: Optional[torch.dtype] = None,
                 output_dict: bool = False):
        super().__init__()
        self.output_dict = output_dict
        self.visual = _build_vision_tower(embed_dim, vision_cfg, quick_gelu, cast_dtype)
        text = _build_text_tower(embed_dim, text_cfg, quick_gelu, cast_dtype)
        self.transformer = text.transformer
        self.context_length = text.context_length
        self.vocab_size = text.vocab_size
        self.token_embedding = text.token_embedding
        self.positional_embedding = text.positional_embedding
        self.ln_final = text.ln_final
        self.text_projection = text.text_projection
        self.pool_style = text.pool_style
        self.register_buffer('attn_mask', text.attn_mask, persistent=False)
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))

    def lock_image_tower(self, unlocked_groups=0, freeze_bn_stats=False):
        self.visual.lock(unlocked_groups=unlocked_groups, freeze_bn_stats=freeze_bn_stats)

    @torch.jit.ignore
    def set_grad_checkpointing(self, enable=True):
        self.visual.set_grad_checkpointing(enable)
        self.transformer.grad_checkpointing = enable

    def encode_image(self, image, normalize: bool = False):
        _wait_ready(image)
        if check_compute_precision(self) == 'fp32':
            features = fp32_path.encode_image(self.visual, image)
            return F.normalize(features, dim=-1) if normalize else features
        features = self.visual(image)
        return _l2_normalize(features) if normalize else features

    def encode_text(self, text, normalize: bool = False):
        if check_compute_precision(self) == 'fp32':
            x = fp32_path.encode_text(self, text, slice_positions=False)
            return F.normalize(x, dim=-1) if normalize else x
        x = text_tower_forward(self, text, slice_positions=False)
        return _l2_normalize(x) if normalize else x

    def forward(self, image, text):
        # text tower first: it needs no image, so an image batch that is still on its way to the device
        # (TrainStep stages host batches on a copy stream) hides behind it; same results either way
        _reserve_for_later(self, self.transformer, image, text)
        text_features = self.encode_text(text, normalize=True)
        image_features = self.encode_image(image, normalize=True)
        if self.output_dict:
            return {"image_features": image_features, "text_features": text_features,
                    "logit_scale": self.logit_scale.exp()}
        return image_features, text_features, self.logit_scale.exp()


class CustomTextCLIP(nn.Module):
    """open_clip/model.py:277-326: the text tower stays a sub-module (`text.*` state_dict keys) instead of being
    re-homed onto the model; selected by `custom_text` in a model config or --force-custom-text."""

    def __init__(self, embed_dim: int, vision_cfg: CLIPVisionCfg, text_cfg: CLIPTextCfg,
                 quick_gelu: bool = False, cast_dtype: Optional[torch.dtype] = None,
                 output_dict: bool = False):
        super().__init__()
        self.output_dict = output_dict
        self.visual = _build_vision_tower(embed_dim, vision_cfg, quick_gelu, cast_dtype)
        self.text = _build_text_tower(embed_dim, text_cfg, quick_gelu, cast_dtype)
        self.context_length = self.text.context_length
        self.vocab_size = self.text.vocab_size
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))

    def lock_image_tower(self, unlocked_groups=0, freeze_bn_stats=False):
        self.visual.lock(unlocked_groups=unlocked_groups, freeze_bn_stats=freeze_bn_stats)

    def lock_text_tower(self, unlocked_layers: int = 0, freeze_layer_norm: bool = True):
        self.text.lock(unlocked_layers, freeze_layer_norm)

    @torch.jit.ignore
    def set_grad_checkpointing(self, enable=True):
        self.visual.set_grad_checkpointing(enable)
        self.text.set_grad_checkpointing(enable)

    def encode_image(self, image, normalize: bool = False):
        _wait_ready(image)
        if check_compute_precision(self) == 'fp32':
            features = fp32_path.encode_image(self.visual, image)
            return F.normalize(features, dim=-1) if normalize else features
        features = self.visual(image)
        return _l2_normalize(features) if normalize else features

    def encode_text(self, text, normalize: bool = False):
        if check_compute_precision(self) == 'fp32':
            features = fp32_path.encode_text(self.text, text, slice_positions=True)
            return F.normalize(features, dim=-1) if normalize else features
        features = self.text(text)
        return _l2_normalize(features) if normalize else features

    def forward(self, image, text):
        # text tower first: it needs no image, so an image batch that is still on its way to the device
        # (TrainStep stages host batches on a copy stream) hides behind it; same results either way
        _reserve_for_later(self, self.text.transformer, image, text)
        text_features = self.encode_text(text, normalize=True)
        image_features = self.encode_image(image, normalize=True)
        if self.output_dict:
            return {"image_features": image_features, "text_features": text_features,
                    "logit_scale": self.logit_scale.exp()}
        return image_features, text_features, self.logit_scale.exp()


def convert_to_custom_text_state_dict(state_dict: dict):
    """open_clip/model.py:358-373: CLIP-format checkpoint -> CustomTextCLIP keys (text tower under `text.`)."""
    if 'text_projection' in state_dict:
        prefixes = ('text_projection', 'positional_embedding', 'token_embedding', 'transformer', 'ln_final')
        return {('text.' + k if any(k.startswith(p) for p in prefixes) else k): v for k, v in state_dict.items()}
    return state_dict


# Precision flags whose arithmetic this library implements (training/precision.py, open_clip/model.py:78-86):
# fp32 master weights with bf16 tensor-core math ('amp_bf16'), bf16 weights ('bf16'), and 'fp32' = fp32 storage AND
# arithmetic on CUDA-core kernels (clipa_b200/fp32_path.py: the 1e-5 parity mode, not a throughput path).
COMPUTE_PRECISIONS = ('amp_bf16', 'amp_bfloat16', 'bf16', 'pure_bf16', 'fp32')


def check_compute_precision(model) -> str:
    """Returns 'fp32' or 'bf16' (the arithmetic to run).  A model built with any other flag ('amp' = fp16 autocast,
    'fp16') is only a parameter container (state_dict / checkpoint conversion): running it would silently replace the
    requested arithmetic, so it refuses."""
    flag = getattr(model, 'compute_precision', 'amp_bf16')
    if flag not in COMPUTE_PRECISIONS:
        raise NotImplementedError(
            f"precision={flag!r}: no fp16 compute path is built.  Build the model with precision='amp_bf16' (fp32 master "
            "weights + bf16 tensor-core math, the mode of every reference GPU script), 'bf16', or 'fp32' (parity mode).")
    return 'fp32' if flag == 'fp32' else 'bf16'


def resize_pos_embed(state_dict, model, interpolation: str = 'bicubic', antialias: bool = True):
    """Rescale the learnable image position-embedding grid of a checkpoint to the model's grid when they
    differ (fine-tuning at a larger resolution; open_clip/model.py:452-483): the CLS row is kept, the
    patch rows are resampled as a [1, D, g, g] image with F.interpolate(align_corners=False).  In place."""
    old = state_dict.get('visual.positional_embedding', None)
    if old is None or not hasattr(model.visual, 'grid_size'):
        return
    gh, gw = model.visual.grid_size
    extra = 1
    if gh * gw + extra == old.shape[0]:
        return
    tok, img = old[:extra], old[extra:]
    og = int(math.sqrt(len(img)))
    logging.info('Resizing position embedding grid-size from %s to %s', (og, og), (gh, gw))
    img = img.reshape(1, og, og, -1).permute(0, 3, 1, 2)
    img = F.interpolate(img, size=(gh, gw), mode=interpolation, antialias=antialias, align_corners=False)
    img = img.permute(0, 2, 3, 1).reshape(1, gh * gw, -1)[0]
    state_dict['visual.positional_embedding'] = torch.cat([tok, img], dim=0)


def resize_text_pos_embed(state_dict, model, interpolation: str = 'linear', antialias: bool = False):
    """Same for the text position table when the context length differs (open_clip/model.py:486-516)."""
    old = state_dict.get('positional_embedding', None)
    if old is None:
        return
    cur = getattr(model, 'positional_embedding', None)
    if cur is None:
        return
    assert old.shape[1] == cur.shape[1], 'text pos_embed width changed!'
    if old.shape[0] == cur.shape[0]:
        return
    logging.info('Resizing text position embedding num_pos from %s to %s', old.shape[0], cur.shape[0])
    new = F.interpolate(old.reshape(1, old.shape[0], old.shape[1]).permute(0, 2, 1), size=cur.shape[0],
                        mode=interpolation, antialias=antialias, align_corners=False)
    state_dict['positional_embedding'] = new.permute(0, 2, 1)[0]


def convert_weights_to_lp(model: nn.Module, dtype=torch.bfloat16):
    """open_clip/model.py:329-351: Linear / Conv / MultiheadAttention weights+biases and the two
    projection matrices go to low precision; LayerNorm, embeddings, positional/class embeddings
    and logit_scale stay fp32."""

    def _convert(l):
        if isinstance(l, (nn.Conv1d, nn.Conv2d, nn.Linear)):
            l.weight.data = l.weight.data.to(dtype)
            if l.bias is not None:
                l.bias.data = l.bias.data.to(dtype)
        if isinstance(l, nn.MultiheadAttention):
            for attr in ["in_proj_weight", "in_proj_bias", "bias_k", "bias_v"]:
                tensor = getattr(l, attr, None)
                if tensor is not None:
                    tensor.data = tensor.data.to(dtype)
            l.out_proj.weight.data = l.out_proj.weight.data.to(dtype)
            l.out_proj.bias.data = l.out_proj.bias.data.to(dtype)
        for name in ["text_projection", "proj"]:
            if hasattr(l, name):
                attr = getattr(l, name)
                if isinstance(attr, nn.Parameter):
                    attr.data = attr.data.to(dtype)

    model.apply(_convert)


convert_weights_to_fp16 = convert_weights_to_lp
