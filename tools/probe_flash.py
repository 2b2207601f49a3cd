"""GPU probe of the flash tcgen05 attention kernels (attention_flash.cu): per-output error breakdown
(by tensor, by query/key tile, main 64 columns vs head_dim-80 remainder) so that ONE run localises a bug,
then timing against the mma.sync kernels at the BASELINE per-layer shapes.
  python tools/probe_flash.py [check] [perf]"""
import math
import sys

import torch

from clipa_b200 import _lib, ops


def set_mode(m):
    _lib.check(_lib.lib().clipa_set_attention_mode(m), "set_attention_mode")


def rel(a, b):
    a, b = a.float(), b.float()
    if not torch.isfinite(a).all():
        return float("nan")
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def check_case(B, L, H, hd, causal, dev):
    torch.manual_seed(B * 100 + L)
    D = H * hd
    qkv = torch.randn(B * L, 3 * D, device=dev).bfloat16()
    x = qkv.float().reshape(B, L, 3, H, hd).requires_grad_(True)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if causal:
        s = s + torch.full((L, L), float("-inf"), device=dev).triu(1)
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * L, D)
    dout = torch.randn(B * L, D, device=dev).bfloat16()
    ref.backward(dout.float())
    gref = x.grad.reshape(B * L, 3 * D)
    set_mode(2)
    tag = f"B={B} L={L} H={H} hd={hd} causal={causal}"
    try:
        out, lse = ops.attention_fwd(qkv, B, L, H, causal)
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        print(f"FAIL fwd {tag}: {e}")
        set_mode(0)
        return False
    e_out, e_lse = rel(out, ref), rel(lse, torch.logsumexp(s, -1))
    ok = e_out < 2e-2 and e_lse < 1e-4
    print(f"{'PASS' if ok else 'FAIL'} fwd {tag}: out {e_out:.2e} lse {e_lse:.2e}")
    T = (L + 95) // 96
    Rt = (L + T - 1) // T
    if not ok:
        o3, r3 = out.float().reshape(B, L, H, hd), ref.reshape(B, L, H, hd)
        for i in range(T):
            sl = slice(i * Rt, min(L, (i + 1) * Rt))
            print(f"    q-tile {i}: main {rel(o3[:, sl, :, :64], r3[:, sl, :, :64]):.2e}"
                  + (f" rem {rel(o3[:, sl, :, 64:], r3[:, sl, :, 64:]):.2e}" if hd > 64 else "")
                  + f" lse {rel(lse[:, :, sl], torch.logsumexp(s, -1)[:, :, sl]):.2e}")
        for b in range(min(B, 3)):
            print(f"    sample {b}: {rel(o3[b], r3[b]):.2e}  heads " +
                  " ".join(f"{rel(o3[b, :, h], r3[b, :, h]):.1e}" for h in range(min(H, 4))))
    # backward always from the REFERENCE forward outputs, so a forward bug does not mask the backward's
    out_ref = ref.detach().bfloat16()
    lse_ref = torch.logsumexp(s, -1).detach().float().contiguous()
    try:
        dqkv = ops.attention_bwd(qkv, out_ref, dout, lse_ref, B, L, H, causal)
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        print(f"FAIL bwd {tag}: {e}")
        set_mode(0)
        return False
    names = ("dq", "dk", "dv")
    floor = 1e-5 * dout.float().abs().max().item()      # L == 1: dQ = dK = 0 exactly in the reference

    def relf(a, b):
        a, b = a.float(), b.float()
        return float("nan") if not torch.isfinite(a).all() else ((a - b).abs().max() / (b.abs().max() + floor)).item()
    errs = [relf(dqkv[:, i * D:(i + 1) * D], gref[:, i * D:(i + 1) * D]) for i in range(3)]
    okb = all(e < 3e-2 for e in errs)
    print(f"{'PASS' if okb else 'FAIL'} bwd {tag}: " + " ".join(f"{n} {e:.2e}" for n, e in zip(names, errs)))
    if not okb:
        g3, r3 = dqkv.float().reshape(B, L, 3, H, hd), gref.reshape(B, L, 3, H, hd)
        for i in range(T):
            sl = slice(i * Rt, min(L, (i + 1) * Rt))
            print(f"    tile {i}: " + " ".join(
                f"{n} main {rel(g3[:, sl, c, :, :64], r3[:, sl, c, :, :64]):.1e}"
                + (f" rem {rel(g3[:, sl, c, :, 64:], r3[:, sl, c, :, 64:]):.1e}" if hd > 64 else "")
                for c, n in enumerate(names)))
    set_mode(0)
    return ok and okb


def perf(dev):
    def t(fn, n=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    for (B, L, H, hd) in ((512, 257, 16, 64), (2048, 37, 16, 80), (256, 257, 16, 80), (128, 577, 16, 64),
                          (1024, 82, 16, 64), (1024, 65, 16, 80)):
        D = H * hd
        qkv = torch.randn(B * L, 3 * D, device=dev).bfloat16()
        dout = torch.randn(B * L, D, device=dev).bfloat16()
        line = f"PERF attn B={B} L={L} H={H} hd={hd}:"
        for mode, name in ((2, "flash"), (1, "mma.sync"), (0, "auto")):
            set_mode(mode)
            try:
                out, lse = ops.attention_fwd(qkv, B, L, H, False)
                tf = t(lambda: ops.attention_fwd(qkv, B, L, H, False))
                tb = t(lambda: ops.attention_bwd(qkv, out, dout, lse, B, L, H, False))
                fl = 4.0 * B * H * L * L * hd
                line += (f"  {name} fwd {tf:.0f} us ({fl / tf / 1e6:.0f} TF/s, {8.0 * B * L * D / tf / 1e3:.0f} GB/s)"
                         f" bwd {tb:.0f} us ({2.5 * fl / tb / 1e6:.0f} TF/s, {16.0 * B * L * D / tb / 1e3:.0f} GB/s)")
            except Exception as e:  # noqa
                line += f"  {name} unsupported ({str(e)[:60]})"
        set_mode(0)
        print(line)


def main():
    dev = torch.device("cuda:0")
    what = sys.argv[1:] or ["check", "perf"]
    if "check" in what:
        cases = [(2, 37, 16, 80, False), (2, 82, 4, 64, False), (2, 257, 4, 64, False), (2, 257, 4, 80, False),
                 (3, 129, 4, 64, False), (2, 192, 2, 64, True), (2, 193, 2, 80, True), (1, 577, 2, 64, False),
                 (40, 257, 16, 64, False), (700, 37, 16, 80, False), (2, 16, 12, 64, True), (1, 1, 2, 80, False),
                 (5, 37, 16, 80, False), (7, 33, 4, 64, True), (300, 16, 12, 64, True), (11, 48, 2, 80, True)]
        n_ok = sum(check_case(*c, dev) for c in cases)
        print(f"GROUP flash check: {n_ok}/{len(cases)} ok")
    if "perf" in what:
        perf(dev)


if __name__ == "__main__":
    main()
