"""One optimizer step of CLIPA training, mirroring `train_one_epoch`'s inner loop
(clipa_torch/training/train.py:180-286) on top of clipa_b200.open_clip:

  H2D copy -> uint8 -> float/255 -> Normalize -> bf16 (`--to-float-on-device`, train.py:191-197)
  -> forward -> ClipLoss (local_loss + gather_with_grad) -> backward -> gradient all-reduce (mean)
  -> AdamW (two weight-decay groups, main.py:311-326) -> logit_scale clamp (train.py:285-286).

When the per-rank batch exceeds `micro_batch`, the GradCache schedule of the reference's
accum_freq > 1 path (train.py:216-256) is used: features of all chunks are first computed without
autograd, the loss and d(features) are evaluated once on the whole (gathered) batch, then every
chunk is re-run with autograd and back-propagated from its slice of d(features).

Data parallelism: one process per GPU.  Parameters' gradients live in ONE flat fp32 buffer, so the
data-parallel reduction is a single NCCL all-reduce (1.7 GB for ViT-L/14, ~4 ms on NVLink 5) issued
after backward; at ~1 s per step it needs no bucketing or overlap.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import open_clip


def exclude_from_wd(name: str, p: torch.Tensor) -> bool:
    """training/main.py:311: no weight decay on gains, biases, LayerNorm, logit_scale."""
    return p.ndim < 2 or "bn" in name or "ln" in name or "bias" in name or "logit_scale" in name


class TrainStep:
    def __init__(self, model: torch.nn.Module, *, rank: int = 0, world_size: int = 1,
                 lr: float = 1.024e-3, betas: Tuple[float, float] = (0.9, 0.95), eps: float = 1e-6,
                 wd: float = 0.2, micro_batch: int = 4096, local_loss: bool = True,
                 gather_with_grad: bool = True, image_mean=None, image_std=None,
                 grad_clip_norm: Optional[float] = None):
        self.model = model
        self.rank, self.world_size = rank, world_size
        self.micro_batch = micro_batch
        self.grad_clip_norm = grad_clip_norm
        self.loss_fn = open_clip.ClipLoss(local_loss=local_loss, gather_with_grad=gather_with_grad,
                                          cache_labels=True, rank=rank, world_size=world_size)
        dev = next(model.parameters()).device
        self.device = dev
        mean = image_mean or getattr(model.visual, "image_mean", open_clip.factory.OPENAI_DATASET_MEAN)
        std = image_std or getattr(model.visual, "image_std", open_clip.factory.OPENAI_DATASET_STD)
        self._mean = torch.tensor(mean, device=dev, dtype=torch.float32).reshape(1, 3, 1, 1)
        self._inv_std = (1.0 / torch.tensor(std, device=dev, dtype=torch.float32)).reshape(1, 3, 1, 1)
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        self.params = [p for _, p in named]
        # flat gradient storage, grouped by dtype (fp32 masters; bf16 parameters in pure-bf16 mode)
        self._flat = {}
        for dt in {p.dtype for p in self.params}:
            ps = [p for p in self.params if p.dtype == dt]
            flat = torch.zeros(sum(p.numel() for p in ps), dtype=dt, device=dev)
            off = 0
            for p in ps:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            self._flat[dt] = flat
        gain = [p for n, p in named if exclude_from_wd(n, p)]
        rest = [p for n, p in named if not exclude_from_wd(n, p)]
        self.optimizer = torch.optim.AdamW(
            [{"params": gain, "weight_decay": 0.}, {"params": rest, "weight_decay": wd}],
            lr=lr, betas=betas, eps=eps, fused=True)

    # -------------------------------------------------------------------------------------
    def preprocess(self, images: torch.Tensor) -> torch.Tensor:
        """uint8 [B,3,H,W] (host or device) -> normalised bf16 on device (train.py:191-197)."""
        images = images.to(self.device, non_blocking=True)
        if images.dtype == torch.uint8:
            images = (images.float().div_(255.0) - self._mean) * self._inv_std
        return images.to(torch.bfloat16)

    def zero_grad(self):
        for flat in self._flat.values():
            flat.zero_()

    def _allreduce_grads(self):
        if self.world_size > 1:
            for flat in self._flat.values():
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
                flat.div_(self.world_size)

    def forward_backward(self, images: torch.Tensor, texts: torch.Tensor) -> torch.Tensor:
        model = self.model
        B = images.shape[0]
        mb = self.micro_batch
        if B <= mb:
            out = model(images, texts)
            feats_i, feats_t, scale = self._unpack(out)
            loss = self.loss_fn(feats_i, feats_t, scale)
            loss.backward()
            return loss.detach()
        chunks = [(s, min(B, s + mb)) for s in range(0, B, mb)]
        with torch.no_grad():
            fi, ft = [], []
            for s, e in chunks:
                a, b, _ = self._unpack(model(images[s:e], texts[s:e]))
                fi.append(a)
                ft.append(b)
        fi = torch.cat(fi).requires_grad_(True)
        ft = torch.cat(ft).requires_grad_(True)
        loss = self.loss_fn(fi, ft, model.logit_scale.exp())
        loss.backward()
        for s, e in chunks:
            a, b, _ = self._unpack(model(images[s:e], texts[s:e]))
            torch.autograd.backward([a, b], [fi.grad[s:e], ft.grad[s:e]])
        return loss.detach()

    @staticmethod
    def _unpack(out):
        if isinstance(out, dict):
            return out["image_features"], out["text_features"], out["logit_scale"]
        return out

    def step(self, images: torch.Tensor, texts: torch.Tensor) -> torch.Tensor:
        """images: uint8 or float [B_local,3,H,W]; texts: int64 [B_local, ctx].  Returns the loss
        (device scalar, no host sync)."""
        images = self.preprocess(images)
        texts = texts.to(self.device, non_blocking=True)
        self.zero_grad()
        loss = self.forward_backward(images, texts)
        self._allreduce_grads()
        if self.grad_clip_norm is not None:
            torch.nn.utils.clip_grad_norm_(self.params, self.grad_clip_norm, norm_type=2.0)
        self.optimizer.step()
        with torch.no_grad():
            self.model.logit_scale.clamp_(0, math.log(100))
        return loss
