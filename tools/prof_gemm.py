"""Tiny driver for ncu: the four forward GEMM shapes of a ViT-L/14 block at 82 tokens x SAMPLES samples
(default 1024; `python tools/prof_gemm.py 4096` = the bench's per-GPU shard), the c_fc recompute (GELU + GELU' outputs), one dgrad, the dgrad x GELU' (DACT) GEMM and one wgrad."""
import sys
import torch
from clipa_b200 import ops
from clipa_b200._lib import EPI_ATOMIC_F32, EPI_BIAS_ACT, EPI_DACT
dev = torch.device("cuda:0")
M, D = 82 * (int(sys.argv[1]) if len(sys.argv) > 1 else 1024), 1024
x = torch.randn(M, D, device=dev).bfloat16()
x4 = torch.randn(M, 4 * D, device=dev).bfloat16()
w_in = torch.randn(3 * D, D, device=dev).bfloat16()
w_fc = torch.randn(4 * D, D, device=dev).bfloat16()
w_pr = torch.randn(D, 4 * D, device=dev).bfloat16()
b4 = torch.randn(4 * D, device=dev)
for _ in range(2):
    ops.gemm(x, w_in, torch.empty(M, 3 * D, dtype=torch.bfloat16, device=dev))
    ops.gemm(x, w_fc, torch.empty(M, 4 * D, dtype=torch.bfloat16, device=dev), epilogue=EPI_BIAS_ACT, bias=b4)
    ops.gemm(x, w_fc, torch.empty(M, 4 * D, dtype=torch.bfloat16, device=dev), epilogue=EPI_BIAS_ACT, bias=b4,
             aux=x4, aux_is_derivative=True)
    ops.gemm(x4, w_pr, torch.empty(M, D, dtype=torch.bfloat16, device=dev), residual=x)
    ops.gemm(x4, w_fc.t(), torch.empty(M, D, dtype=torch.bfloat16, device=dev))
    ops.gemm(x, w_pr.t(), torch.empty(M, 4 * D, dtype=torch.bfloat16, device=dev), epilogue=EPI_DACT, aux=x4, aux_is_derivative=True)
    ops.gemm(x4.t(), x.t(), torch.zeros(4 * D, D, dtype=torch.float32, device=dev), epilogue=EPI_ATOMIC_F32, split_k=-1)
torch.cuda.synchronize()
