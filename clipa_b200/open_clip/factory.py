"""Drop-in factory: same call signatures as open_clip/factory.py:121-352, so that
clipa_torch/training/main.py (`create_model_and_transforms`, `create_loss`, `get_model_config`)
runs on top of this package unchanged.

Precision flags (training/precision.py:6-15, open_clip/model.py:78-86):
  'bf16'               weights of matmul-like layers are cast to bf16 (convert_weights_to_lp),
                       LayerNorm/embeddings stay fp32 -- the reference's "pure bf16" mode;
  'amp_bf16'           fp32 master weights; the kernels read bf16 shadow copies refreshed after every
                       optimizer step, gradients are fp32 (the reference's autocast runs the same
                       matmuls in bf16).
  'fp32'               fp32 weights, activations and arithmetic on CUDA-core kernels (clipa_b200/fp32_path.py):
                       the reference-exact parity mode (1e-5), orders of magnitude slower than the tensor-core path.
  'amp' / 'fp16'       NOT computed (no fp16 arithmetic): the model can be built (state_dict schema, checkpoint
                       conversion) but forward() raises instead of silently substituting other math.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, Optional, Tuple, Union

import torch

from .loss import ClipLoss
from .model import (CLIP, COMPUTE_PRECISIONS, CustomTextCLIP, convert_to_custom_text_state_dict, convert_weights_to_lp,
                    get_cast_dtype, resize_pos_embed, resize_text_pos_embed)
from .model_configs import add_model_config, get_model_config, list_models  # noqa: F401

OPENAI_DATASET_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_DATASET_STD = (0.26862954, 0.26130258, 0.27577711)


def load_state_dict(checkpoint_path: str, map_location='cpu'):
    """open_clip/factory.py:99-108."""
    checkpoint = torch.load(checkpoint_path, map_location=map_location)
    state_dict = checkpoint['state_dict'] if isinstance(checkpoint, dict) and 'state_dict' in checkpoint else checkpoint
    if next(iter(state_dict.items()))[0].startswith('module'):
        state_dict = {k[7:]: v for k, v in state_dict.items()}
    return state_dict


def load_checkpoint(model, checkpoint_path, strict=True):
    """open_clip/factory.py:110-118: position tables of the checkpoint are resampled to the model's image
    grid / context length before loading (pre-train at 84-126 px, fine-tune at 224+)."""
    state_dict = load_state_dict(checkpoint_path)
    if 'positional_embedding' in state_dict and not hasattr(model, 'positional_embedding'):
        state_dict = convert_to_custom_text_state_dict(state_dict)       # CLIP checkpoint -> CustomTextCLIP keys
    resize_pos_embed(state_dict, model)
    resize_text_pos_embed(state_dict, model)
    return model.load_state_dict(state_dict, strict=strict)


def create_model(model_name: str, pretrained: Optional[str] = None, precision: str = 'fp32',
                 device: Union[str, torch.device] = 'cpu', jit: bool = False,
                 force_quick_gelu: bool = False, force_custom_text: bool = False,
                 force_patch_dropout: Optional[float] = None,
                 force_image_size: Optional[Union[int, Tuple[int, int]]] = None,
                 pretrained_image: bool = False, pretrained_hf: bool = True,
                 cache_dir: Optional[str] = None, output_dict: Optional[bool] = None,
                 require_pretrained: bool = False, pos_embed: str = None):
    if jit:
        raise NotImplementedError("torch.jit.script is not applicable: blocks are custom CUDA autograd nodes")
    if pretrained_image:
        raise NotImplementedError("timm image towers are outside the CLIPA hot path")
    model_name = model_name.replace('/', '-')
    if isinstance(device, str):
        device = torch.device(device)
    model_cfg = get_model_config(model_name)
    if model_cfg is None:
        raise RuntimeError(f'Model config for {model_name} not found; available models {list_models()}.')
    if force_quick_gelu:
        model_cfg["quick_gelu"] = True
    if force_patch_dropout is not None:
        model_cfg["vision_cfg"]["patch_dropout"] = force_patch_dropout
    if force_image_size is not None:
        model_cfg["vision_cfg"]["image_size"] = force_image_size
    if pos_embed is not None:
        model_cfg["vision_cfg"]["pos_embed"] = pos_embed
    custom_text = model_cfg.pop('custom_text', False) or force_custom_text          # factory.py:203-211
    cast_dtype = get_cast_dtype(precision)
    model = (CustomTextCLIP if custom_text else CLIP)(**model_cfg, cast_dtype=cast_dtype)
    # 'amp' / 'fp16' models can be built, saved and loaded, but refuse to run (model.check_compute_precision);
    # 'fp32' runs the CUDA-core parity path
    model.compute_precision = precision
    if precision not in COMPUTE_PRECISIONS:
        logging.warning("precision=%r: built as a parameter container only; forward() needs 'amp_bf16' or 'bf16'",
                        precision)
    if pretrained:
        import os
        if not os.path.exists(pretrained):
            raise RuntimeError(f'Pretrained weights ({pretrained}) not found: only local checkpoint files '
                               'are supported (no network downloads).')
        logging.info(f'Loading pretrained {model_name} weights ({pretrained}).')
        load_checkpoint(model, pretrained)
    elif require_pretrained:
        raise RuntimeError(f'Pretrained weights were required for {model_name} but not given.')
    model.to(device=device)
    if precision == 'bf16':
        convert_weights_to_lp(model, dtype=torch.bfloat16)
    elif precision == 'fp16':
        raise NotImplementedError("fp16 is not built")
    model.visual.image_mean = OPENAI_DATASET_MEAN
    model.visual.image_std = OPENAI_DATASET_STD
    if output_dict and hasattr(model, "output_dict"):
        model.output_dict = True
    return model


def create_loss(args):
    """open_clip/factory.py:262-290 (ClipLoss only; distill / CoCa losses are out of scope)."""
    if getattr(args, "distill", False) or "coca" in getattr(args, "model", "").lower():
        raise NotImplementedError("only ClipLoss is on the CLIPA hot path")
    return ClipLoss(local_loss=args.local_loss, gather_with_grad=args.gather_with_grad, cache_labels=True,
                    rank=args.rank, world_size=args.world_size, use_horovod=getattr(args, "horovod", False))


def image_transform(image_size, is_train: bool, mean=None, std=None, **_unused):
    """Minimal torchvision pipeline standing in for open_clip/transform.py:91-214 (host-side image
    preprocessing is outside the device hot path)."""
    from torchvision import transforms as T
    mean = mean or OPENAI_DATASET_MEAN
    std = std or OPENAI_DATASET_STD
    size = image_size if isinstance(image_size, int) else tuple(image_size)
    if isinstance(size, tuple) and size[0] == size[1]:
        size = size[0]
    to_rgb = T.Lambda(lambda im: im.convert('RGB'))
    if is_train:
        return T.Compose([T.RandomResizedCrop(size, scale=(0.9, 1.0), interpolation=T.InterpolationMode.BICUBIC),
                          to_rgb, T.ToTensor(), T.Normalize(mean, std)])
    return T.Compose([T.Resize(size, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(size), to_rgb,
                      T.ToTensor(), T.Normalize(mean, std)])


def create_model_and_transforms(model_name: str, pretrained: Optional[str] = None, precision: str = 'fp32',
                                device: Union[str, torch.device] = 'cpu', jit: bool = False,
                                force_quick_gelu: bool = False, force_custom_text: bool = False,
                                force_patch_dropout: Optional[float] = None,
                                force_image_size: Optional[Union[int, Tuple[int, int]]] = None,
                                pretrained_image: bool = False, pretrained_hf: bool = True,
                                image_mean: Optional[Tuple[float, ...]] = None,
                                image_std: Optional[Tuple[float, ...]] = None,
                                aug_cfg: Optional[Dict[str, Any]] = None, cache_dir: Optional[str] = None,
                                output_dict: Optional[bool] = None, to_float_on_device: Optional[bool] = False,
                                pos_embed: str = None, interpolation: str = 'bicubic',
                                square_resize_only: bool = False):
    model = create_model(model_name, pretrained, precision=precision, device=device, jit=jit,
                         force_quick_gelu=force_quick_gelu, force_custom_text=force_custom_text,
                         force_patch_dropout=force_patch_dropout, force_image_size=force_image_size,
                         pretrained_image=pretrained_image, pretrained_hf=pretrained_hf, cache_dir=cache_dir,
                         output_dict=output_dict, pos_embed=pos_embed)
    image_mean = image_mean or getattr(model.visual, 'image_mean', None)
    image_std = image_std or getattr(model.visual, 'image_std', None)
    preprocess_train = image_transform(model.visual.image_size, is_train=True, mean=image_mean, std=image_std)
    preprocess_val = image_transform(model.visual.image_size, is_train=False, mean=image_mean, std=image_std)
    return model, preprocess_train, preprocess_val


def trace_model(model, batch_size=256, device=torch.device('cpu')):
    """open_clip/factory.py (imported by training/main.py:40): torch.jit tracing does not apply to custom CUDA autograd
    nodes."""
    raise NotImplementedError("trace_model: torch.jit tracing is not applicable to the custom-kernel towers")


def get_tokenizer(model_name):
    raise NotImplementedError("tokenizers are host-side string processing outside the hot path; the model "
                              "consumes int64 token ids [batch, context_length] from the reference tokenizer")
