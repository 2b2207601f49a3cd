"""Fixed 2-D sin-cos position embedding (the `--pos-embed sin_cos_2d` of every CLIPA pre-training
script).  Restates open_clip/pos_embed.py:20-67: the first half of the channels encodes the
column index, the second half the row index; each half is [sin(p * w_k), cos(p * w_k)] with
w_k = 10000^(-k / (D/4)); an all-zero row is prepended for the CLS token."""
from __future__ import annotations

import math

import torch


def get_2d_sincos_pos_embed(embed_dim: int, grid_size: int, cls_token: bool = False) -> torch.Tensor:
    if embed_dim % 4 != 0:
        raise ValueError("embed_dim must be a multiple of 4")
    quarter = embed_dim // 4
    omega = torch.exp(-math.log(10000.0) * torch.arange(quarter, dtype=torch.float64) / quarter)
    idx = torch.arange(grid_size, dtype=torch.float64)
    rows = idx.repeat_interleave(grid_size)   # row index of each patch (row-major patch order)
    cols = idx.repeat(grid_size)              # column index

    def encode(pos):
        ang = pos[:, None] * omega[None, :]
        return torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)

    emb = torch.cat([encode(cols), encode(rows)], dim=1)
    if cls_token:
        emb = torch.cat([torch.zeros(1, embed_dim, dtype=torch.float64), emb], dim=0)
    return emb.float()
