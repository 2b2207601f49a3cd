"""Tiny ncu driver: flash attention fwd/bwd at the config-4 per-layer shape (reduced batch)."""
import sys
import torch
from clipa_b200 import ops
B, L, H, hd = (int(a) for a in (sys.argv[1:5] if len(sys.argv) >= 5 else (256, 257, 16, 64)))
dev = torch.device("cuda:0")
qkv = (torch.randn(B * L, 3 * H * hd, device=dev)).bfloat16()
dout = torch.randn(B * L, H * hd, device=dev).bfloat16()
for _ in range(2):
    out, lse = ops.attention_fwd(qkv, B, L, H, False)
    ops.attention_bwd(qkv, out, dout, lse, B, L, H, False)
torch.cuda.synchronize()
