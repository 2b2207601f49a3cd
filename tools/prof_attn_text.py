"""Times the tcgen05 attention kernels at the text-tower shapes (packed tiles) and the image shape."""
import torch
from clipa_b200 import ops

dev = torch.device("cuda:0")


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for B, L, H, causal in [(4096, 16, 12, True), (4096, 8, 16, True), (4096, 32, 12, True), (1024, 82, 16, False),
                        (4096, 50, 12, False)]:
    D = H * 64
    qkv = torch.randn(B * L, 3 * D, device=dev).bfloat16()
    dout = torch.randn(B * L, D, device=dev).bfloat16()
    out, lse = ops.attention_fwd(qkv, B, L, H, causal)
    tf = timeit(lambda: ops.attention_fwd(qkv, B, L, H, causal))
    tb = timeit(lambda: ops.attention_bwd(qkv, out, dout, lse, B, L, H, causal))
    gb_f = (B * L * 4 * D * 2) / 1e9
    gb_b = (B * L * 8 * D * 2) / 1e9
    print(f"PERF attn B={B} L={L} H={H} causal={causal}: fwd {tf:.1f} us ({gb_f / tf * 1e6:.0f} GB/s)  "
          f"bwd {tb:.1f} us ({gb_b / tb * 1e6:.0f} GB/s)")
