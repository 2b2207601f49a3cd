"""GPU bring-up probe: runs one named group of kernel checks against torch fp32 math and prints a
compact PASS/FAIL table with error maps for failures.  Run each group in its own process
(tools/gpu_probe.sh does, under `timeout`) so that one hung or faulting kernel cannot take the
other groups down with it.  Not part of the product; tests/ holds the real parity suite.
"""
import sys
import time

import torch

from clipa_b200 import ops
from clipa_b200._lib import (ACT_GELU_ERF, ACT_GELU_TANH, ACT_QUICK_GELU, EPI_ATOMIC_F32,
                             EPI_BIAS_ACT, EPI_DACT, EPI_STORE)

dev = torch.device("cuda:0")
torch.manual_seed(0)
results = []


def errmap(got, ref, rb=128, cb=64):
    e = (got.float() - ref.float()).abs()
    M, N = e.shape
    lines = []
    for r0 in range(0, min(M, rb * 6), rb):
        row = []
        for c0 in range(0, min(N, cb * 10), cb):
            row.append(f"{e[r0:r0 + rb, c0:c0 + cb].max().item():8.2e}")
        lines.append(" ".join(row))
    return "\n".join(lines)


def report(name, got, ref, tol):
    got = got.float()
    ref = ref.float()
    scale = ref.abs().max().item() + 1e-30
    err = (got - ref).abs().max().item() / scale
    bad = not (err <= tol) or not torch.isfinite(got).all().item()
    results.append((name, err, tol, not bad))
    print(f"{'PASS' if not bad else 'FAIL'} {name:60s} rel_max_err={err:.3e} tol={tol:.1e}", flush=True)
    if bad:
        print(errmap(got, ref), flush=True)
    return not bad


def mk(shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)


def gemm_case(M, N, K, a_mn=False, b_mn=False, **kw):
    A = mk((K, M)).t() if a_mn else mk((M, K))
    B = mk((K, N)).t() if b_mn else mk((N, K))
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    ops.gemm(A, B, out, **kw)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    return report(f"gemm M{M} N{N} K{K} a_mn={int(a_mn)} b_mn={int(b_mn)}", out, ref, 1e-2)


def group_gemm_kk():
    gemm_case(128, 256, 64)
    gemm_case(128, 256, 256)
    gemm_case(256, 512, 1024)
    gemm_case(128, 128, 128)
    gemm_case(4096, 1024, 1024)
    gemm_case(8192, 3072, 1024)


def group_gemm_tails():
    gemm_case(100, 256, 64)
    gemm_case(333, 776, 200)
    gemm_case(1000, 88, 72)
    gemm_case(129, 264, 1032)
    gemm_case(82 * 7, 768, 768)


def group_gemm_kmn():
    gemm_case(128, 256, 64, b_mn=True)
    gemm_case(256, 512, 512, b_mn=True)
    gemm_case(4096, 1024, 4096, b_mn=True)
    gemm_case(333, 776, 200, b_mn=True)
    gemm_case(128, 128, 128, b_mn=True)


def group_gemm_mnmn():
    gemm_case(128, 256, 64, a_mn=True, b_mn=True)
    gemm_case(256, 512, 512, a_mn=True, b_mn=True)
    gemm_case(1024, 1024, 8192, a_mn=True, b_mn=True)
    gemm_case(776, 336, 1000, a_mn=True, b_mn=True)
    gemm_case(128, 256, 64, a_mn=True, b_mn=False)
    gemm_case(512, 384, 640, a_mn=True, b_mn=False)


def group_gemm_epi():
    M, N, K = 1024, 1024, 512
    A, B = mk((M, K)), mk((N, K), 0.05)
    bias = torch.randn(N, device=dev)
    res = mk((M, N))
    ref = A.float() @ B.float().t()
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    ops.gemm(A, B, out, bias=bias, residual=res)
    report("epi STORE bias(f32)+residual", out, ref + bias + res.float(), 1e-2)
    ops.gemm(A, B, out, bias=bias.bfloat16(), alpha=0.5)
    report("epi STORE bias(bf16) alpha", out, 0.5 * ref + bias.bfloat16().float(), 1e-2)
    out32 = torch.empty(M, N, dtype=torch.float32, device=dev)
    ops.gemm(A, B, out32)
    report("epi STORE f32 out", out32, ref, 1e-5)
    for act, name, fn in [(ACT_GELU_ERF, "erf", lambda x: torch.nn.functional.gelu(x)),
                          (ACT_GELU_TANH, "tanh", lambda x: torch.nn.functional.gelu(x, approximate="tanh")),
                          (ACT_QUICK_GELU, "quick", lambda x: x * torch.sigmoid(1.702 * x))]:
        aux = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        ops.gemm(A, B, out, epilogue=EPI_BIAS_ACT, bias=bias, aux=aux, act=act)
        f = ref + bias
        report(f"epi BIAS_ACT {name} (aux)", aux, f, 1e-2)
        report(f"epi BIAS_ACT {name} (out)", out, fn(f.bfloat16().float()), 1e-2)
        # DACT: C = (A@B^T) * act'(aux)
        x = aux.float().requires_grad_(True)
        fn(x).sum().backward()
        W = mk((K, N), 0.05)  # dgrad style: b MN-major [N,K] view
        Bt = W.t()
        out2 = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        ops.gemm(A, Bt, out2, epilogue=EPI_DACT, aux=aux, act=act)
        report(f"epi DACT {name}", out2, (A.float() @ W.float()) * x.grad, 1e-2)
    # atomic split-K wgrad style
    Mt = 20000
    dY, X = mk((Mt, 768), 0.1), mk((Mt, 512), 0.1)
    acc = torch.zeros(768, 512, dtype=torch.float32, device=dev)
    ops.gemm(dY.t(), X.t(), acc, epilogue=EPI_ATOMIC_F32, split_k=-1)
    ops.gemm(dY.t(), X.t(), acc, epilogue=EPI_ATOMIC_F32, split_k=3, alpha=2.0)
    refw = 3.0 * (dY.float().t() @ X.float())
    report("epi ATOMIC_F32 split-K accumulate (MN,MN)", acc, refw, 1e-4)


def group_ln():
    for rows, D in [(1000, 768), (4099, 1024), (513, 1280), (64, 512), (300, 256)]:
        x = mk((rows, D), 2.0) + 0.5
        g = torch.randn(D, device=dev)
        b = torch.randn(D, device=dev)
        y, mean, rstd = ops.layernorm_fwd(x, g, b)
        xr = x.float().requires_grad_(True)
        gr = g.clone().requires_grad_(True)
        br = b.clone().requires_grad_(True)
        yr = torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-5)
        report(f"ln_fwd rows{rows} D{D}", y, yr, 1e-2)
        report(f"ln_fwd mean rows{rows} D{D}", mean, xr.mean(-1), 1e-5)
        dy = mk((rows, D))
        dres = mk((rows, D))
        yr.backward(dy.float())
        dg = torch.zeros(D, device=dev)
        db = torch.zeros(D, device=dev)
        dx = ops.layernorm_bwd(dy, x, g, mean, rstd, dres, dg, db)
        report(f"ln_bwd dx rows{rows} D{D}", dx, xr.grad + dres.float(), 1e-2)
        report(f"ln_bwd dgamma rows{rows} D{D}", dg, gr.grad, 1e-3)
        report(f"ln_bwd dbeta rows{rows} D{D}", db, br.grad, 1e-3)
    x = mk((5000, 776))
    out = torch.zeros(776, device=dev)
    ops.colsum_accum(x, out)
    report("colsum 5000x776", out, x.float().sum(0), 1e-4)


def group_loss():
    for bl, bg, E, off in [(256, 256, 512, 0), (300, 1000, 768, 300), (1024, 4096, 768, 2048)]:
        a = torch.nn.functional.normalize(torch.randn(bl, E, device=dev), dim=-1).bfloat16()
        b = torch.nn.functional.normalize(torch.randn(bg, E, device=dev), dim=-1).bfloat16()
        scale = 14.285714
        logits = scale * (a.float() @ b.float().t())
        lse_ref = torch.logsumexp(logits, dim=-1)
        labels = torch.arange(bl, device=dev) + off
        diag_ref = logits[torch.arange(bl, device=dev), labels]
        lse, diag = ops.clip_lse(a, b, scale, off)
        report(f"clip_lse lse bl{bl} bg{bg}", lse, lse_ref, 1e-5)
        report(f"clip_lse diag bl{bl} bg{bg}", diag, diag_ref, 1e-5)
        ds = torch.zeros(1, device=dev)
        pt = ops.clip_softmax_grad(a, b, scale, off, lse, ds)
        p_ref = torch.softmax(logits, -1)
        p_ref[torch.arange(bl, device=dev), labels] -= 1.0
        report(f"clip_softmax_grad pt bl{bl} bg{bg}", pt, p_ref, 1e-2)
        report(f"clip_softmax_grad dscale bl{bl} bg{bg}", ds, (p_ref * logits / scale).sum().reshape(1), 2e-3)


def group_attn():
    import math
    for (B, L, H, hd, causal) in [(3, 82, 16, 64, False), (5, 16, 12, 64, True), (2, 37, 16, 80, False),
                                  (2, 257, 4, 64, False), (4, 8, 16, 64, True), (2, 65, 12, 64, False),
                                  (3, 32, 12, 64, True), (2, 100, 2, 64, True)]:
        D = H * hd
        qkv = mk((B * L, 3 * D))
        out, lse = ops.attention_fwd(qkv, B, L, H, causal)
        x = qkv.float().reshape(B, L, 3, H, hd).requires_grad_(True)
        q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
        s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
        if causal:
            s = s + torch.full((L, L), float("-inf"), device=dev).triu(1)
        ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * L, D)
        tag = f"B{B} L{L} H{H} hd{hd} causal{int(causal)}"
        report(f"attn_fwd out {tag}", out, ref, 2e-2)
        report(f"attn_fwd lse {tag}", lse, torch.logsumexp(s, -1), 1e-2)
        dout = mk((B * L, D))
        ref.backward(dout.float())
        dqkv = ops.attention_bwd(qkv, out, dout, lse, B, L, H, causal)
        gref = x.grad.reshape(B * L, 3 * D)
        report(f"attn_bwd dq {tag}", dqkv[:, :D], gref[:, :D], 3e-2)
        report(f"attn_bwd dk {tag}", dqkv[:, D:2 * D], gref[:, D:2 * D], 3e-2)
        report(f"attn_bwd dv {tag}", dqkv[:, 2 * D:], gref[:, 2 * D:], 3e-2)
    # timing at the config-3 shape (per layer): B=1024 x 16 heads, L=82
    B, L, H, hd = 1024, 82, 16, 64
    qkv = mk((B * L, 3 * H * hd))
    for _ in range(3):
        out, lse = ops.attention_fwd(qkv, B, L, H, False)
    dout = mk((B * L, H * hd))
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for _ in range(10):
        out, lse = ops.attention_fwd(qkv, B, L, H, False)
    e1.record()
    for _ in range(10):
        ops.attention_bwd(qkv, out, dout, lse, B, L, H, False)
    e2.record()
    torch.cuda.synchronize()
    fb = B * L * H * hd * 2 * 4
    print(f"PERF attn B{B} L{L}: fwd {e0.elapsed_time(e1) / 10:.3f} ms ({fb / (e0.elapsed_time(e1) / 10) / 1e6:.0f} GB/s), "
          f"bwd {e1.elapsed_time(e2) / 10:.3f} ms ({fb * 2 / (e1.elapsed_time(e2) / 10) / 1e6:.0f} GB/s)", flush=True)


def group_gemm_perf():
    for (M, N, K) in [(8192, 8192, 8192), (82 * 1024, 3072, 1024), (82 * 1024, 4096, 1024),
                      (82 * 1024, 1024, 4096), (82 * 1024, 1024, 1024)]:
        A, B = mk((M, K)), mk((N, K))
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        for _ in range(3):
            ops.gemm(A, B, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gemm(A, B, out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        tf = 2.0 * M * N * K / ms / 1e9
        for _ in range(3):
            torch.matmul(A, B.t(), out=out)
        e0.record()
        for _ in range(10):
            torch.matmul(A, B.t(), out=out)
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / 10
        print(f"PERF gemm M{M} N{N} K{K}: ours {ms:.3f} ms {tf:.0f} TFLOP/s | cublas {ms2:.3f} ms "
              f"{2.0 * M * N * K / ms2 / 1e9:.0f} TFLOP/s", flush=True)


def group_gemm_epi_perf():
    """Epilogue cost at the c_fc / c_proj-dgrad shapes (M = 82*1024 tokens)."""
    M, D = 82 * 1024, 1024
    x = mk((M, D)); w_fc = mk((4 * D, D), 0.03); b4 = torch.randn(4 * D, device=dev)
    g = torch.empty(M, 4 * D, dtype=torch.bfloat16, device=dev)
    f = torch.empty(M, 4 * D, dtype=torch.bfloat16, device=dev)
    dy = mk((M, D)); w_pr = mk((D, 4 * D), 0.03); res = mk((M, D))
    outD = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
    cases = {
        "c_fc STORE(bias)": lambda: ops.gemm(x, w_fc, g, bias=b4),
        "c_fc BIAS_ACT erf": lambda: ops.gemm(x, w_fc, g, epilogue=EPI_BIAS_ACT, bias=b4),
        "c_fc BIAS_ACT erf +aux": lambda: ops.gemm(x, w_fc, g, epilogue=EPI_BIAS_ACT, bias=b4, aux=f),
        "c_proj-dgrad STORE": lambda: ops.gemm(dy, w_pr.t(), g),
        "c_proj-dgrad DACT erf": lambda: ops.gemm(dy, w_pr.t(), g, epilogue=EPI_DACT, aux=f),
        "c_proj STORE(bias,res)": lambda: ops.gemm(g, w_pr, outD, bias=b4[:D].contiguous(), residual=res),
        "out_proj STORE(bias,res)": lambda: ops.gemm(x, w_fc[:D], outD, bias=b4[:D].contiguous(), residual=res),
        "out_proj STORE(bias)": lambda: ops.gemm(x, w_fc[:D], outD, bias=b4[:D].contiguous()),
    }
    flops = {"c_fc": 2.0 * M * 4 * D * D, "c_pr": 2.0 * M * 4 * D * D, "out_": 2.0 * M * D * D}
    for name, fn in cases.items():
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"PERF {name:28s} {ms:.3f} ms  {flops[name[:4]] / ms / 1e9:.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    import os
    from clipa_b200 import _lib
    mode = int(os.environ.get("CLIPA_GEMM_MODE", "0"))
    _lib.check(_lib.lib().clipa_set_gemm_mode(mode), "set_gemm_mode")
    print(f"gemm mode {mode}", flush=True)
    g = sys.argv[1]
    t0 = time.time()
    globals()["group_" + g]()
    torch.cuda.synchronize()
    nfail = sum(1 for r in results if not r[3])
    print(f"GROUP {g}: {len(results) - nfail} passed, {nfail} failed, {time.time() - t0:.1f}s", flush=True)
    sys.exit(1 if nfail else 0)
