"""GPU parity tests of every C-ABI entry point against plain fp32 torch math of the same op
(floating-point kernels: the fp32 reference is the oracle at this granularity; tolerances are
bf16 round-off of the OUTPUT, stated per test)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

BF16_OUT = 1e-2     # max-abs error / max-abs value for a bf16-rounded output of O(1)-conditioned math
F32_OUT = 1e-5


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def ops():
    from clipa_b200 import ops as o
    return o


@pytest.fixture(params=[1, 2], ids=["1cta", "2cta"])
def gemm_mode(request):
    """Force the 1-CTA kernel / the cta_group::2 pair kernel for every GEMM in the test."""
    from clipa_b200 import _lib
    _lib.check(_lib.lib().clipa_set_gemm_mode(request.param), "set_gemm_mode")
    yield request.param
    _lib.check(_lib.lib().clipa_set_gemm_mode(0), "set_gemm_mode")


def relmax(got, ref):
    got, ref = got.float(), ref.float()
    assert torch.isfinite(got).all()
    return ((got - ref).abs().max() / (ref.abs().max() + 1e-30)).item()


def mk(shape, dev, scale=1.0, seed=None):
    return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 1024), (128, 128, 128), (100, 256, 64),
                                   (333, 776, 200), (1000, 88, 72), (129, 264, 1032), (82 * 7, 768, 768),
                                   (4096, 3072, 1024)])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
def test_gemm_majors(dev, gemm_mode, M, N, K, a_mn, b_mn):
    if (a_mn and M % 8) or (b_mn and N % 8) or (not a_mn and K % 8) or (not b_mn and K % 8):
        pytest.skip("row pitch of an operand would not be 16-byte aligned (rejected by the ABI; see test_gemm_rejects_bad_arguments)")
    torch.manual_seed(M * 7 + N * 3 + K)
    A = mk((K, M), dev).t() if a_mn else mk((M, K), dev)
    B = mk((K, N), dev).t() if b_mn else mk((N, K), dev)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    ops().gemm(A, B, out)
    assert relmax(out, A.float() @ B.float().t()) < BF16_OUT


def test_gemm_epilogues(dev, gemm_mode):
    from clipa_b200._lib import (ACT_GELU_ERF, ACT_GELU_TANH, ACT_QUICK_GELU, EPI_ATOMIC_F32,
                                 EPI_BIAS_ACT, EPI_DACT)
    o = ops()
    torch.manual_seed(1)
    M, N, K = 1024, 1024, 512
    A, B = mk((M, K), dev), mk((N, K), dev, 0.05)
    bias = torch.randn(N, device=dev)
    res = mk((M, N), dev)
    ref = A.float() @ B.float().t()
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    o.gemm(A, B, out, bias=bias, residual=res)
    assert relmax(out, ref + bias + res.float()) < BF16_OUT
    o.gemm(A, B, out, bias=bias.bfloat16(), alpha=0.5)
    assert relmax(out, 0.5 * ref + bias.bfloat16().float()) < BF16_OUT
    out32 = torch.empty(M, N, dtype=torch.float32, device=dev)
    o.gemm(A, B, out32)
    assert relmax(out32, ref) < F32_OUT
    F = torch.nn.functional
    for act, fn in [(ACT_GELU_ERF, lambda x: F.gelu(x)), (ACT_GELU_TANH, lambda x: F.gelu(x, approximate="tanh")),
                    (ACT_QUICK_GELU, lambda x: x * torch.sigmoid(1.702 * x))]:
        aux = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        o.gemm(A, B, out, epilogue=EPI_BIAS_ACT, bias=bias, aux=aux, act=act)
        f = ref + bias
        assert relmax(aux, f) < BF16_OUT
        assert relmax(out, fn(f.bfloat16().float())) < BF16_OUT
        out_noaux = torch.empty_like(out)
        o.gemm(A, B, out_noaux, epilogue=EPI_BIAS_ACT, bias=bias, act=act)
        assert relmax(out_noaux, fn(f)) < BF16_OUT
        x = aux.float().requires_grad_(True)
        fn(x).sum().backward()
        W = mk((K, N), dev, 0.05)
        out2 = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        o.gemm(A, W.t(), out2, epilogue=EPI_DACT, aux=aux, act=act)
        assert relmax(out2, (A.float() @ W.float()) * x.grad) < BF16_OUT
        # recompute-pass variant: aux receives act'(f) (of the UNROUNDED f), the dgrad GEMM only multiplies by it
        daux = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        o.gemm(A, B, out, epilogue=EPI_BIAS_ACT, bias=bias, aux=daux, act=act, aux_is_derivative=True)
        xf = f.clone().requires_grad_(True)
        fn(xf).sum().backward()
        assert relmax(out, fn(f)) < BF16_OUT and relmax(daux, xf.grad) < BF16_OUT
        o.gemm(A, W.t(), out2, epilogue=EPI_DACT, aux=daux, act=act, aux_is_derivative=True)
        assert relmax(out2, (A.float() @ W.float()) * daux.float()) < BF16_OUT
    Mt = 20000
    dY, X = mk((Mt, 768), dev, 0.1), mk((Mt, 512), dev, 0.1)
    acc = torch.zeros(768, 512, dtype=torch.float32, device=dev)
    o.gemm(dY.t(), X.t(), acc, epilogue=EPI_ATOMIC_F32, split_k=-1)
    o.gemm(dY.t(), X.t(), acc, epilogue=EPI_ATOMIC_F32, split_k=3, alpha=2.0)
    assert relmax(acc, 3.0 * (dY.float().t() @ X.float())) < 1e-4


def test_gemm_rejects_bad_arguments(dev):
    from clipa_b200._lib import ClipaError
    o = ops()
    A, B = mk((64, 36), dev), mk((64, 36), dev)      # ld = 36 is not a multiple of 8
    with pytest.raises(ClipaError):
        o.gemm(A, B, torch.empty(64, 64, dtype=torch.bfloat16, device=dev))
    with pytest.raises(ClipaError):                  # CPU tensors are refused, not silently handled
        o.gemm(torch.zeros(64, 64, dtype=torch.bfloat16), torch.zeros(64, 64, dtype=torch.bfloat16),
               torch.zeros(64, 64, dtype=torch.bfloat16))


@pytest.mark.parametrize("rows,D", [(1000, 768), (4099, 1024), (513, 1280), (64, 512), (300, 256), (7, 1664)])
def test_layernorm(dev, rows, D):
    o = ops()
    torch.manual_seed(rows + D)
    x = mk((rows, D), dev, 2.0) + 0.5
    g, b = torch.randn(D, device=dev), torch.randn(D, device=dev)
    y, mean, rstd = o.layernorm_fwd(x, g, b)
    xr, gr, br = x.float().requires_grad_(True), g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-5)
    assert relmax(y, yr) < BF16_OUT
    assert relmax(mean, xr.mean(-1)) < F32_OUT
    dy, dres = mk((rows, D), dev), mk((rows, D), dev)
    yr.backward(dy.float())
    dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    dxs = torch.full((D,), 0.25, device=dev)
    dx = o.layernorm_bwd(dy, x, g, mean, rstd, dres, dg, db, dxsum=dxs)
    assert relmax(dx, xr.grad + dres.float()) < BF16_OUT
    assert relmax(dg, gr.grad) < 1e-3 and relmax(db, br.grad) < 1e-3
    # fused column sums of the output (a bias gradient): accumulated (+=) onto the buffer, unrounded dx
    assert relmax(dxs, 0.25 + (xr.grad + dres.float()).sum(0)) < 2e-3
    dx2 = o.layernorm_bwd(dy, x, g, mean, rstd, dres, torch.zeros_like(dg), torch.zeros_like(db))
    assert torch.equal(dx2, dx)


def test_layernorm_constant_rows(dev):
    """Size-independent property: a constant row normalises to exactly beta."""
    o = ops()
    D = 1024
    x = torch.full((33, D), 3.25, device=dev).bfloat16()
    g, b = torch.randn(D, device=dev), torch.randn(D, device=dev)
    y, _, _ = o.layernorm_fwd(x, g, b)
    assert torch.equal(y, b.bfloat16().expand(33, D))


def test_colsum(dev):
    o = ops()
    x = mk((5000, 776), dev)
    out = torch.ones(776, device=dev)
    o.colsum_accum(x, out)
    assert relmax(out, 1.0 + x.float().sum(0)) < 1e-4


@pytest.mark.parametrize("B,L,H,hd,causal", [(3, 82, 16, 64, False), (5, 16, 12, 64, True), (2, 37, 16, 80, False),
                                             (2, 257, 4, 64, False), (4, 8, 16, 64, True), (2, 65, 12, 64, False),
                                             (3, 32, 12, 64, True), (2, 100, 2, 64, True), (1, 1, 2, 64, True),
                                             # packed tiles (several samples per 128-row tile): partial last
                                             # tile, blocks straddling 32-column chunks, G*L < 128, one sample
                                             (19, 16, 12, 64, True), (21, 16, 4, 64, False), (9, 8, 16, 64, True),
                                             (7, 33, 4, 64, True), (5, 48, 2, 64, False), (11, 63, 2, 64, True),
                                             (300, 16, 12, 64, True), (3, 64, 2, 64, False), (1, 5, 2, 64, True),
                                             # pipelined backward (64 < L <= 96), several problems per CTA
                                             (40, 82, 16, 64, False), (37, 96, 8, 64, True), (3, 70, 4, 64, True),
                                             (75, 65, 4, 64, False)])
def test_attention(dev, B, L, H, hd, causal):
    _attention_case(dev, B, L, H, hd, causal)


@pytest.fixture
def flash_mode():
    """Force the flash tcgen05 kernels (attention_flash.cu) for every head_dim 64 / 80 shape."""
    from clipa_b200 import _lib
    _lib.check(_lib.lib().clipa_set_attention_mode(2), "set_attention_mode")
    yield
    _lib.check(_lib.lib().clipa_set_attention_mode(0), "set_attention_mode")


def _attention_case(dev, B, L, H, hd, causal, tol_out=2e-2, tol_grad=3e-2):
    o = ops()
    torch.manual_seed(B * 100 + L)
    D = H * hd
    qkv = mk((B * L, 3 * D), dev)
    out, lse = o.attention_fwd(qkv, B, L, H, causal)
    x = qkv.float().reshape(B, L, 3, H, hd).requires_grad_(True)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if causal:
        s = s + torch.full((L, L), float("-inf"), device=dev).triu(1)
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * L, D)
    assert relmax(out, ref) < tol_out
    assert relmax(lse, torch.logsumexp(s, -1)) < 1e-4
    dout = mk((B * L, D), dev)
    ref.backward(dout.float())
    dqkv = o.attention_bwd(qkv, out, dout, lse, B, L, H, causal)
    gref = x.grad.reshape(B * L, 3 * D)
    floor = 1e-5 * dout.float().abs().max().item()      # L == 1: dQ = dK = 0 exactly in the reference; fp32 round-off here
    for i in range(3):
        got, ref = dqkv[:, i * D:(i + 1) * D].float(), gref[:, i * D:(i + 1) * D]
        assert torch.isfinite(got).all()
        assert ((got - ref).abs().max() / (ref.abs().max() + floor)).item() < tol_grad


# Flash tcgen05 kernels: every BASELINE shape outside the one-tile kernels -- L = 257 (config 4: 3 tiles of
# 86), head_dim 80 (config 5 at L = 37; ViT-H fine-tune at 257) -- plus 336-px fine-tune lengths (577 = 7
# tiles of 83, 401), tile-boundary lengths, causal multi-tile masks, more work items than CTAs, and
# (forced) the one-tile shapes themselves.
@pytest.mark.parametrize("B,L,H,hd,causal", [
    (2, 257, 4, 64, False), (2, 37, 16, 80, False), (2, 257, 4, 80, False), (1, 577, 2, 64, False),
    (1, 577, 2, 80, False), (2, 401, 2, 64, False), (3, 129, 4, 64, False), (2, 192, 2, 64, True),
    (2, 193, 2, 80, True), (2, 97, 4, 80, False), (3, 96, 4, 80, True), (5, 82, 16, 64, False),
    (2, 16, 12, 64, True), (3, 8, 16, 80, True), (1, 1, 2, 80, False), (40, 257, 16, 64, False),
    (700, 37, 16, 80, False), (30, 65, 16, 80, False), (9, 128, 8, 64, True),
    # packed short sequences (block-diagonal tiles): partial last tile, causal, G*L < tile rows
    (5, 37, 16, 80, False), (7, 33, 4, 64, True), (300, 16, 12, 64, True), (11, 48, 2, 80, True), (4, 64, 2, 64, False),
    (1, 5, 2, 64, True), (1, 577, 2, 80, True)])
def test_attention_flash(dev, flash_mode, B, L, H, hd, causal):
    _attention_case(dev, B, L, H, hd, causal)


def test_attention_flash_is_the_default_for_long_and_wide(dev):
    """Without any override the BASELINE shapes with L > 128 or head_dim 80 take the tcgen05 flash kernels:
    same results as the forced path, bit for bit."""
    from clipa_b200 import _lib
    o = ops()
    for (B, L, H, hd) in ((2, 257, 4, 64), (3, 37, 16, 80)):
        qkv = mk((B * L, 3 * H * hd), dev, seed=L)
        out0, lse0 = o.attention_fwd(qkv, B, L, H, False)
        _lib.check(_lib.lib().clipa_set_attention_mode(2), "set_attention_mode")
        out2, lse2 = o.attention_fwd(qkv, B, L, H, False)
        _lib.check(_lib.lib().clipa_set_attention_mode(1), "set_attention_mode")
        out1, _ = o.attention_fwd(qkv, B, L, H, False)          # mma.sync kernels: close, not identical
        _lib.check(_lib.lib().clipa_set_attention_mode(0), "set_attention_mode")
        assert torch.equal(out0, out2) and torch.equal(lse0, lse2)
        assert relmax(out1, out0) < 2e-2


def test_attention_uniform_values_long(dev):
    """Size-independent property at the config-4 per-layer shape (L = 257, 16 heads; 256 samples): identical
    value rows -> attention returns that row; and d(qkv) of a constant upstream gradient has zero dQ, dK
    (softmax rows sum to one, so constant values carry no score gradient)."""
    o = ops()
    B, L, H, hd = 256, 257, 16, 64
    D = H * hd
    qkv = mk((B * L, 3 * D), dev)
    vrow = torch.randn(D, device=dev).bfloat16()
    qkv[:, 2 * D:] = vrow
    out, lse = o.attention_fwd(qkv, B, L, H, False)
    assert relmax(out, vrow.float().expand(B * L, D)) < 1e-2
    dout = mk((B * L, D), dev)
    dqkv = o.attention_bwd(qkv, out, dout, lse, B, L, H, False)
    assert dqkv[:, :2 * D].float().abs().max().item() < 2e-2 * dout.float().abs().max().item()


def test_attention_uniform_values(dev):
    """Property at full per-GPU size of config 3 (B=4096 x 16 heads would be 10 GB of qkv; one
    eighth here): if every key carries the same value row, attention returns exactly that row."""
    o = ops()
    B, L, H, hd = 512, 82, 16, 64
    D = H * hd
    qkv = mk((B * L, 3 * D), dev)
    vrow = torch.randn(D, device=dev).bfloat16()
    qkv[:, 2 * D:] = vrow
    out, _ = o.attention_fwd(qkv, B, L, H, False)
    assert relmax(out, vrow.float().expand(B * L, D)) < 1e-2


@pytest.mark.parametrize("bl,bg,E,off", [(256, 256, 512, 0), (300, 1000, 768, 300), (1024, 4096, 768, 2048),
                                         (8, 8, 64, 0), (130, 260, 1024, 130)])
def test_clip_head_kernels(dev, bl, bg, E, off):
    o = ops()
    torch.manual_seed(bl + bg)
    a = torch.nn.functional.normalize(torch.randn(bl, E, device=dev), dim=-1).bfloat16()
    b = torch.nn.functional.normalize(torch.randn(bg, E, device=dev), dim=-1).bfloat16()
    scale = 14.285714
    logits = scale * (a.float() @ b.float().t())
    idx = torch.arange(bl, device=dev)
    lse, diag = o.clip_lse(a, b, scale, off)
    assert relmax(lse, torch.logsumexp(logits, -1)) < F32_OUT
    assert relmax(diag, logits[idx, idx + off]) < F32_OUT
    ds = torch.zeros(1, device=dev)
    pt = o.clip_softmax_grad(a, b, scale, off, lse, ds)
    p_ref = torch.softmax(logits, -1)
    p_ref[idx, idx + off] -= 1.0
    assert relmax(pt, p_ref) < BF16_OUT
    assert relmax(ds, (p_ref * logits / scale).sum().reshape(1)) < 2e-3


def test_clip_loss_identical_pairs_full_size(dev):
    """BASELINE size (B_local 4096 x B_global 32768, E 768): with identical samples every logit is
    equal, so both cross-entropies are ln(B_global) -- the reference's SyntheticDataset property
    (training/data.py:469-486)."""
    from clipa_b200.open_clip import ClipLoss
    bl, bg, E = 4096, 32768, 768
    v = torch.nn.functional.normalize(torch.randn(1, E, device=dev), dim=-1).bfloat16()
    a, b = v.expand(bl, E).contiguous(), v.expand(bg, E).contiguous()
    lse, diag = ops().clip_lse(a, b, 14.285714, 4096)
    loss = (lse - diag).mean().item()
    assert abs(loss - math.log(bg)) < 1e-4
    # and through the public module at world_size 1 (8192 pairs)
    f = v.expand(8192, E).contiguous().requires_grad_(True)
    out = ClipLoss()(f, f, torch.tensor(14.285714, device=dev))
    assert abs(out.item() - math.log(8192)) < 1e-4


@pytest.mark.parametrize("wd", [0.0, 0.2])
def test_fused_adamw_matches_torch(dev, wd):
    """clipa_adamw_step == torch.optim.AdamW (training/main.py:318-326 recipe) over several steps,
    incl. the bf16 shadow copy, the gradient clear and the folded gradient scale."""
    o = ops()
    torch.manual_seed(3)
    n = 4096 * 33
    p0 = torch.randn(n, device=dev)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=1.024e-3, betas=(0.9, 0.95), eps=1e-6, weight_decay=wd)
    p, m, v = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    shadow = torch.empty(n, dtype=torch.bfloat16, device=dev)
    for step in range(1, 6):
        g = torch.randn(n, device=dev) * 0.1
        ref.grad = (g * 0.5).clone()
        opt.step()
        gbuf = g.clone()
        o.adamw_step(p, gbuf, m, v, shadow, lr=1.024e-3, beta1=0.9, beta2=0.95, eps=1e-6, weight_decay=wd,
                     step=step, grad_scale=0.5, zero_grad=True)
        assert torch.count_nonzero(gbuf).item() == 0
    assert relmax(p, ref.detach()) < 1e-6
    assert torch.equal(shadow, p.bfloat16())


# ---- tower heads and tails (csrc/tower_io.cu) against the torch expressions the reference uses ------------------
def test_preprocess_u8_matches_reference_recipe(dev):
    """training/train.py:191-197: images.float().div(255) -> Normalize(mean, std) -> cast."""
    o = ops()
    mean, std = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)
    for (n, h, w) in ((3, 126, 126), (2, 7, 9), (1, 224, 224)):
        img = torch.randint(0, 256, (n, 3, h, w), dtype=torch.uint8, device=dev)
        got = o.preprocess_u8(img, mean, std)
        m = torch.tensor(mean, device=dev).view(1, 3, 1, 1)
        s = torch.tensor(std, device=dev).view(1, 3, 1, 1)
        ref = ((img.float().div(255) - m) / s)
        assert got.dtype == torch.bfloat16 and relmax(got, ref) < 4e-3


@pytest.mark.parametrize("n,size,patch,width", [(3, 126, 14, 64), (2, 64, 16, 128), (2, 96, 32, 32)])
def test_patchify_plus_gemm_equals_conv(dev, n, size, patch, width):
    """conv1 with stride == kernel (open_clip/transformer.py:371,491-493) == patchify + GEMM."""
    o = ops()
    torch.manual_seed(n + size)
    img = torch.randn(n, 3, size, size, device=dev)
    wgt = torch.randn(width, 3, patch, patch, device=dev) * 0.05
    K = 3 * patch * patch
    Kp = (K + 7) // 8 * 8
    for x in (img, img.bfloat16()):
        p = o.patchify(x.contiguous(), patch, patch, Kp)
        assert p.shape == (n * (size // patch) ** 2, Kp) and (p[:, K:] == 0).all()
        ref = torch.nn.functional.conv2d(x.float().bfloat16().float(), wgt.bfloat16().float(), stride=patch)
        ref = ref.reshape(n, width, -1).permute(0, 2, 1).reshape(-1, width)
        got = p[:, :K].float() @ wgt.reshape(width, K).bfloat16().float().t()
        assert relmax(got, ref) < 1e-3


def test_assemble_embed_pool_normalize_match_torch(dev):
    from clipa_b200 import functional as Fn
    from clipa_b200._lib import POOL_ARGMAX_ID, POOL_FIRST, POOL_LAST, POOL_MEAN_ALL, POOL_MEAN_SKIP_FIRST
    torch.manual_seed(0)
    n, L, W, V = 5, 10, 64, 97
    # [cls; tok] + pos, forward and backward (learnable table and fixed table)
    tok = mk((n * (L - 1), W), dev).requires_grad_(True)
    cls = torch.randn(W, device=dev, requires_grad=True)
    for learn in (True, False):
        pos = torch.randn(L, W, device=dev, requires_grad=learn)
        x = Fn.AssembleTokensFn.apply(tok, cls, pos, n, L)
        ref = torch.cat([cls.expand(n, 1, W), tok.float().reshape(n, L - 1, W)], 1) + pos
        assert relmax(x, ref) < BF16_OUT
        g = mk((n, L, W), dev)
        gt, gc, gp = torch.autograd.grad(x, [tok, cls] + ([pos] if learn else []), g) + ((None,) if not learn else ())
        rt, rc, rp = torch.autograd.grad(ref, [tok, cls] + ([pos] if learn else []), g.float()) + ((None,) if not learn else ())
        assert relmax(gt, rt) < 1e-6 and relmax(gc, rc) < 1e-5
        if learn:
            assert relmax(gp, rp) < 1e-5
    # token embedding gather + positions
    ids = torch.randint(0, V, (n, L), device=dev)
    table = torch.randn(V, W, device=dev, requires_grad=True)
    pos = torch.randn(L + 3, W, device=dev, requires_grad=True)
    x = Fn.EmbedTokensFn.apply(ids, table, pos)
    ref = table[ids] + pos[:L]
    assert relmax(x, ref) < BF16_OUT
    g = mk((n, L, W), dev)
    gt, gp = torch.autograd.grad(x, [table, pos], g)
    rt, rp = torch.autograd.grad(ref, [table, pos], g.float())
    assert relmax(gt, rt) < 1e-5 and relmax(gp, rp) < 1e-5
    # pooling modes
    xx = mk((n, L, W), dev).requires_grad_(True)
    refs = {POOL_FIRST: lambda t: t[:, 0], POOL_LAST: lambda t: t[:, -1],
            POOL_ARGMAX_ID: lambda t: t[torch.arange(n, device=dev), ids.argmax(-1)],
            POOL_MEAN_ALL: lambda t: t.mean(1), POOL_MEAN_SKIP_FIRST: lambda t: t[:, 1:].mean(1)}
    for mode, fn in refs.items():
        y = Fn.PoolTokensFn.apply(xx, mode, ids if mode == POOL_ARGMAX_ID else None)
        r = fn(xx.float())
        assert relmax(y, r) < BF16_OUT
        g = mk((n, W), dev)
        assert relmax(torch.autograd.grad(y, xx, g)[0], torch.autograd.grad(r, xx, g.float())[0]) < BF16_OUT
    # ties in the ids: torch.argmax returns the FIRST maximal position
    tie = torch.zeros(n, L, dtype=torch.int64, device=dev)
    tie[:, 3] = 7; tie[:, 6] = 7
    assert torch.equal(Fn.PoolTokensFn.apply(xx, POOL_ARGMAX_ID, tie), xx[:, 3].detach())
    # F.normalize and its backward; output written into a slice of a wider buffer
    f = mk((33, 128), dev).requires_grad_(True)
    y = Fn.L2NormalizeFn.apply(f, None)
    r = torch.nn.functional.normalize(f.float(), dim=-1)
    assert relmax(y, r) < BF16_OUT
    g = torch.randn(33, 128, device=dev)
    assert relmax(torch.autograd.grad(y, f, g)[0], torch.autograd.grad(r, f, g)[0]) < 2e-2
    buf = torch.zeros(99, 128, dtype=torch.bfloat16, device=dev)
    y2 = Fn.L2NormalizeFn.apply(f, buf[33:66])
    assert torch.equal(buf[33:66], y.detach()) and (buf[:33] == 0).all() and (buf[66:] == 0).all()
    assert y2.data_ptr() == buf[33:66].data_ptr()


def test_clip_loss_column_chunking_matches_unchunked(dev, monkeypatch):
    """The head materialises its softmax-gradient matrix one column block at a time once it would exceed
    functional.P_TILDE_BYTES (config 5: 2 x 1 GiB per GPU otherwise); same loss and gradients as the one-block path."""
    from clipa_b200 import functional as Fn
    from clipa_b200.open_clip import ClipLoss
    torch.manual_seed(4)
    bl, bg, E = 512, 4096, 256
    a = torch.nn.functional.normalize(torch.randn(bl, E, device=dev), dim=-1).bfloat16()
    b = torch.nn.functional.normalize(torch.randn(bl, E, device=dev), dim=-1).bfloat16()
    a_all = torch.nn.functional.normalize(torch.randn(bg, E, device=dev), dim=-1).bfloat16()
    b_all = torch.nn.functional.normalize(torch.randn(bg, E, device=dev), dim=-1).bfloat16()
    a_all[1024:1024 + bl], b_all[1024:1024 + bl] = a, b           # this rank's rows sit at offset 1024 (rank 2 of 8)
    scale = torch.tensor(14.285714, device=dev, requires_grad=True)
    res = []
    for cap in (1 << 40, 2 * bl * 1024):                           # one block / four column blocks of 1024
        monkeypatch.setattr(Fn, "P_TILDE_BYTES", cap)
        ins = [t.clone().requires_grad_(True) for t in (a, b, a_all, b_all)]
        loss = Fn.ClipLossFn.apply(ins[0], ins[1], ins[2], ins[3], scale, 2, True)
        grads = torch.autograd.grad(loss, ins + [scale])
        res.append((loss.item(), grads))
    (l0, g0), (l1, g1) = res
    assert abs(l0 - l1) < 1e-6
    for x, y in zip(g0, g1):
        assert relmax(y, x) < 1e-2        # outputs are bf16-rounded (1 ulp = 3.9e-3); fp32 accumulation order differs
