"""clipa_b200: the CLIPA training-step hot path on B200 (sm_100a) behind the open_clip API.

    from clipa_b200 import open_clip
    model, _, _ = open_clip.create_model_and_transforms('ViT-L-14-CL16', precision='amp_bf16', device='cuda')
    loss_fn = open_clip.ClipLoss(local_loss=True, gather_with_grad=True, rank=rank, world_size=world)
"""
from . import _lib, ops  # noqa: F401
from . import open_clip  # noqa: F401

__version__ = "0.1.0"
