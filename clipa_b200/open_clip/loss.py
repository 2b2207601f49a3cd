"""ClipLoss with the reference's constructor and call signature (open_clip/loss.py:92-157) on the
fused sm_100a contrastive-head kernels.

Multi-GPU: the one exchange step of the path.  Per-rank L2-normalised features are all-gathered
over NCCL (NVLink 5 / NVSwitch); each rank then runs the fused log-sum-exp GEMM for its LOCAL rows
against all gathered columns (labels offset by rank*B_local, loss.py:118-120), the backward GEMMs
produce d(local) and d(gathered) features, and d(gathered) is reduce-scattered (SUM) back -- the
autograd transpose of torch.distributed.nn.all_gather that `gather_with_grad` uses (loss.py:75-76).
"""
from __future__ import annotations

import torch
import torch.nn as nn

try:
    import torch.distributed as dist
    has_distributed = True
except ImportError:  # pragma: no cover
    dist = None
    has_distributed = False

from .. import fp32_path
from .. import functional as Fn


class _AllGatherFeatures(torch.autograd.Function):
    """all_gather_into_tensor forward.  Backward, as in open_clip/loss.py:75-88:
      gather_with_grad            -> reduce_scatter(SUM) (the transpose of torch.distributed.nn.all_gather);
      else, not local_loss        -> the local slice only (the reference re-inserts the local features into
                                     the gathered list so that they keep their gradient, loss.py:84-86);
      else (local_loss, no grad)  -> nothing: the gathered tensors carry no gradient at all."""

    @staticmethod
    def forward(ctx, x, world_size, with_grad, local_loss):
        x = x.contiguous()
        out = torch.empty(world_size * x.shape[0], x.shape[1], dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x)
        ctx.with_grad, ctx.local_loss = with_grad, local_loss
        ctx.rank = dist.get_rank()
        ctx.n = x.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        if ctx.with_grad:
            g = g.contiguous()
            out = torch.empty(ctx.n, g.shape[1], dtype=g.dtype, device=g.device)
            dist.reduce_scatter_tensor(out, g, op=dist.ReduceOp.SUM)
            return out, None, None, None
        if not ctx.local_loss:
            return g[ctx.rank * ctx.n:(ctx.rank + 1) * ctx.n].clone(), None, None, None
        return None, None, None, None


def gather_features(image_features, text_features, local_loss=False, gather_with_grad=False,
                    rank=0, world_size=1, use_horovod=False):
    """open_clip/loss.py:31-89 (torch.distributed backend only: no horovod / XLA dispatch)."""
    assert has_distributed, 'torch.distributed did not import correctly'
    if use_horovod:
        raise NotImplementedError("horovod backend is not built (single backend: NCCL via torch.distributed)")
    all_image = _AllGatherFeatures.apply(image_features, world_size, gather_with_grad, local_loss)
    all_text = _AllGatherFeatures.apply(text_features, world_size, gather_with_grad, local_loss)
    return all_image, all_text


class ClipLoss(nn.Module):

    def __init__(self, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0,
                 world_size=1, use_horovod=False):
        super().__init__()
        self.local_loss = local_loss
        self.gather_with_grad = gather_with_grad
        self.cache_labels = cache_labels   # labels are implicit (row + rank*B_local) in the kernel
        self.rank = rank
        self.world_size = world_size
        self.use_horovod = use_horovod

    def forward(self, image_features, text_features, logit_scale, output_dict=False):
        if not image_features.is_cuda:
            raise RuntimeError("clipa_b200.ClipLoss runs on CUDA tensors only (no CPU fallback)")
        if image_features.dtype == torch.float32 and text_features.dtype == torch.float32:
            loss_fn = fp32_path.ClipLossF32        # features of a precision='fp32' model: fp32 head (parity mode)
        else:
            loss_fn = Fn.ClipLossFn
            image_features = image_features.to(torch.bfloat16)
            text_features = text_features.to(torch.bfloat16)
        if self.world_size > 1:
            all_image, all_text = gather_features(
                image_features, text_features, self.local_loss, self.gather_with_grad, self.rank,
                self.world_size, self.use_horovod)
            if self.local_loss:
                # without gather_with_grad the gathered tensors are constants here: skip their gradient GEMMs
                total_loss = loss_fn.apply(image_features, text_features, all_image, all_text,
                                                 logit_scale, self.rank, self.gather_with_grad)
            else:
                # full-matrix form (loss.py:138-139): every rank evaluates all rows
                total_loss = loss_fn.apply(all_image, all_text, all_image, all_text,
                                                 logit_scale, 0, True)
        else:
            total_loss = loss_fn.apply(image_features, text_features, image_features,
                                             text_features, logit_scale, 0, True)
        return {"contrastive_loss": total_loss} if output_dict else total_loss
