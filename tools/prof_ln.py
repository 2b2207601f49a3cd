"""Times LayerNorm forward/backward and the bias-gradient column sum at the ViT-L/14 and text-tower
shapes; reports algorithmic GB/s (bf16 rows in + out, as DESIGN.md counts them)."""
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


for rows, D in [(82 * 4096, 1024), (16 * 4096, 768), (37 * 2048, 1280)]:
    x = torch.randn(rows, D, device=dev).bfloat16()
    dy = torch.randn(rows, D, device=dev).bfloat16()
    dres = torch.randn(rows, D, device=dev).bfloat16()
    g = torch.randn(D, device=dev)
    b = torch.randn(D, device=dev)
    dg = torch.zeros(D, device=dev)
    db = torch.zeros(D, device=dev)
    y, mean, rstd = ops.layernorm_fwd(x, g, b)
    tf = timeit(lambda: ops.layernorm_fwd(x, g, b))
    tb = timeit(lambda: ops.layernorm_bwd(dy, x, g, mean, rstd, dres, dg, db))
    tb2 = timeit(lambda: ops.layernorm_bwd(dy, x, g, mean, rstd, None, dg, db))
    dxs = torch.zeros(D, device=dev)
    tb3 = timeit(lambda: ops.layernorm_bwd(dy, x, g, mean, rstd, dres, dg, db, dxsum=dxs))
    tc = timeit(lambda: ops.colsum_accum(dy, dxs))
    n = rows * D
    print(f"PERF ln rows={rows} D={D}: fwd {tf:.1f} us ({4 * n / tf / 1e3:.0f} GB/s)  "
          f"bwd+dres {tb:.1f} us ({8 * n / tb / 1e3:.0f} GB/s)  bwd {tb2:.1f} us ({6 * n / tb2 / 1e3:.0f} GB/s)  "
          f"bwd+dres+dxsum {tb3:.1f} us ({8 * n / tb3 / 1e3:.0f} GB/s; separate column-sum pass: {tc:.1f} us)")
