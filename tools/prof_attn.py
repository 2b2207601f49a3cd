"""Tiny driver for ncu: a few attention fwd/bwd launches at the config-3 per-layer shape."""
import sys
import torch
from clipa_b200 import ops
B, L, H, hd = 1024, 82, 16, 64
dev = torch.device("cuda:0")
qkv = (torch.randn(B * L, 3 * H * hd, device=dev)).bfloat16()
dout = torch.randn(B * L, H * hd, device=dev).bfloat16()
for _ in range(3):
    out, lse = ops.attention_fwd(qkv, B, L, H, False)
    ops.attention_bwd(qkv, out, dout, lse, B, L, H, False)
torch.cuda.synchronize()
