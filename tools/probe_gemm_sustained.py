"""Sustained (power-capped clocks) throughput of the block GEMMs with their epilogues at the bench's shard shape
(M = 82 x SAMPLES tokens, default 4096 samples): each case runs back to back for ~1 s before it is timed."""
import sys
import torch
from clipa_b200 import ops
from clipa_b200._lib import EPI_ATOMIC_F32, EPI_BIAS_ACT, EPI_DACT
dev = torch.device("cuda:0")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
D = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
M = 82 * S
mk = lambda shape, s=1.0: (torch.randn(*shape, device=dev) * s).bfloat16()
x = mk((M, D)); dy = mk((M, D)); res = mk((M, D))
w_fc = mk((4 * D, D), 0.03); w_pr = mk((D, 4 * D), 0.03); w_in = mk((3 * D, D), 0.03)
b4 = torch.randn(4 * D, device=dev)
g = torch.empty(M, 4 * D, dtype=torch.bfloat16, device=dev)
f = torch.empty(M, 4 * D, dtype=torch.bfloat16, device=dev)
outD = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
qkv = torch.empty(M, 3 * D, dtype=torch.bfloat16, device=dev)
dw = torch.zeros(4 * D, D, dtype=torch.float32, device=dev)
F4 = 2.0 * M * 4 * D * D
cases = [
    ("qkv STORE(bias)", F4 * 0.75, lambda: ops.gemm(x, w_in, qkv, bias=b4[:3 * D].contiguous())),
    ("c_fc STORE(bias)", F4, lambda: ops.gemm(x, w_fc, g, bias=b4)),
    ("c_fc BIAS_ACT erf (forward)", F4, lambda: ops.gemm(x, w_fc, g, epilogue=EPI_BIAS_ACT, bias=b4)),
    ("c_fc BIAS_ACT erf + GELU' aux (recompute)", F4,
     lambda: ops.gemm(x, w_fc, g, epilogue=EPI_BIAS_ACT, bias=b4, aux=f, aux_is_derivative=True)),
    ("c_fc BIAS_ACT erf + pre-act aux", F4, lambda: ops.gemm(x, w_fc, g, epilogue=EPI_BIAS_ACT, bias=b4, aux=f)),
    ("c_proj dgrad STORE", F4, lambda: ops.gemm(dy, w_pr.t(), g)),
    ("c_proj dgrad x GELU' aux (DACT)", F4,
     lambda: ops.gemm(dy, w_pr.t(), g, epilogue=EPI_DACT, aux=f, aux_is_derivative=True)),
    ("c_proj dgrad DACT from pre-act", F4, lambda: ops.gemm(dy, w_pr.t(), g, epilogue=EPI_DACT, aux=f)),
    ("c_proj STORE(bias,res)", F4, lambda: ops.gemm(g, w_pr, outD, bias=b4[:D].contiguous(), residual=res)),
    ("c_fc wgrad split-K", F4, lambda: ops.gemm(g.t(), x.t(), dw, epilogue=EPI_ATOMIC_F32, split_k=-1)),
]
REP = int(sys.argv[3]) if len(sys.argv) > 3 else 400
for name, fl, fn in cases:
    for _ in range(REP):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(f"PERF sustained M={M} D={D} {name:44s} {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s", flush=True)
