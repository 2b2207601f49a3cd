"""GPU parity of the whole hot path (towers -> features -> ClipLoss -> backward) through the public
open_clip-style API, against
  (a) the committed golden vectors the REFERENCE produced (tests/golden, fp32 and pure-bf16 runs),
  (b) the CPU oracle on the same seeded weights/inputs.

Tolerance model (bf16): the reference's own bf16 run deviates from its fp32 run by e_ref (stored
in the goldens).  A correct bf16 implementation lands at a comparable distance from the fp32 truth,
so each quantity must satisfy  err(ours, ref_fp32) <= 2 * e_ref + floor.  The loss must be within
2.5e-3 relative of the fp32 reference loss (or 2*e_ref if larger): north_star asks 1e-3, which
the reference's OWN pure-bf16 run misses in 3 of these 4 cases (5.5e-4 .. 5.9e-3, see the goldens);
the measured values (2e-4 .. 1.6e-3 here) are recorded in the parity report.
Measured errors are appended to gpurun_out/parity_report.jsonl for the round report.
"""
import json
import os
from pathlib import Path

import pytest
import torch

from tests.helpers import load_golden, rel_err

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
CASES = ["tiny-cls", "tiny-gap-h80", "tiny-bigvision", "config1-vitb32"]


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _report(rec):
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    with open(out / "parity_report.jsonl", "a") as f:
        f.write(json.dumps(rec) + "\n")


def build_model(meta, precision, dev):
    import tempfile
    from clipa_b200 import open_clip
    from oracle.weights import make_state_dict
    tmp = Path(tempfile.mkdtemp())
    name = f"golden-{meta['name']}"
    (tmp / f"{name}.json").write_text(json.dumps(meta["cfg"]))
    open_clip.add_model_config(tmp)
    model, _, _ = open_clip.create_model_and_transforms(
        name, precision=precision, device=dev, force_image_size=meta["image_size"],
        pos_embed=meta["pos_embed"], output_dict=True)
    sd = make_state_dict(meta["cfg"], meta["seed"], image_size=meta["image_size"], pos_embed=meta["pos_embed"])
    model.load_state_dict(sd, strict=True)
    model.train()
    return model


def run_ours(meta, precision, dev):
    from clipa_b200 import open_clip
    from oracle.weights import make_inputs
    model = build_model(meta, precision, dev)
    images, text = make_inputs(meta["cfg"], meta["batch"], meta["seed"] + 1000, image_size=meta["image_size"])
    out = model(images.to(dev), text.to(dev))
    out["image_features"].retain_grad()
    out["text_features"].retain_grad()
    loss = open_clip.ClipLoss()(out["image_features"], out["text_features"], out["logit_scale"])
    loss.backward()
    return model, text, out, loss


@pytest.mark.parametrize("precision", ["amp_bf16", "bf16"])
@pytest.mark.parametrize("name", CASES)
def test_step_matches_reference(dev, name, precision):
    meta, g32 = load_golden(name, "fp32")
    _, g16 = load_golden(name, "bf16")
    model, text, out, loss = run_ours(meta, precision, dev)
    params = dict(model.named_parameters())
    last = meta["cfg"]["vision_cfg"]["layers"] - 1
    rows = torch.as_tensor(g32["token_rows_idx"])
    ours = {
        "image_features": out["image_features"].detach().float().cpu(),
        "text_features": out["text_features"].detach().float().cpu(),
        "d_image_features": out["image_features"].grad.float().cpu(),
        "d_text_features": out["text_features"].grad.float().cpu(),
        "g_visual_proj": params["visual.proj"].grad.float().cpu(),
        "g_text_projection": params["text_projection"].grad.float().cpu(),
        "g_v0_in_proj_weight": params["visual.transformer.resblocks.0.attn.in_proj_weight"].grad.float().cpu()[:64, :64],
        "g_v0_in_proj_bias": params["visual.transformer.resblocks.0.attn.in_proj_bias"].grad.float().cpu(),
        "g_vlast_c_fc_bias": params[f"visual.transformer.resblocks.{last}.mlp.c_fc.bias"].grad.float().cpu(),
        "g_t0_ln_1_weight": params["transformer.resblocks.0.ln_1.weight"].grad.float().cpu(),
        "g_class_embedding": params["visual.class_embedding"].grad.float().cpu(),
        "g_token_rows": params["token_embedding.weight"].grad.float().cpu()[rows],
    }
    rec = {"case": name, "precision": precision, "errors": {}}
    failures = []
    for k, v in ours.items():
        e_ours = rel_err(v, g32[k])
        e_ref = rel_err(g16[k], g32[k])
        rec["errors"][k] = {"ours_vs_ref_fp32": e_ours, "ref_bf16_vs_ref_fp32": e_ref}
        if not e_ours <= 2.0 * e_ref + 2e-3:
            failures.append((k, e_ours, e_ref))
    l32, l16 = float(g32["loss"]), float(g16["loss"])
    e_loss = abs(loss.item() - l32) / l32
    e_loss_ref = abs(l16 - l32) / l32
    rec["loss"] = {"ours": loss.item(), "ref_fp32": l32, "ref_bf16": l16, "rel_err": e_loss, "ref_rel_err": e_loss_ref}
    gs, gs32, gs16 = params["logit_scale"].grad.item(), float(g32["g_logit_scale"]), float(g16["g_logit_scale"])
    rec["g_logit_scale"] = {"ours": gs, "ref_fp32": gs32, "ref_bf16": gs16}
    _report(rec)
    assert not failures, failures
    assert e_loss <= max(2.5e-3, 2 * e_loss_ref), rec["loss"]
    assert abs(gs - gs32) <= 2 * abs(gs16 - gs32) + 2e-2 * abs(gs32) + 1e-4, rec["g_logit_scale"]
    assert loss.dtype == torch.float32


BASELINE_DIM = ["vitl14-i81-t16-d2", "vitl14-i256-t32-d2", "vith14-i36-t8-d2", "vitb16-i64-t16-d3"]


@pytest.mark.parametrize("name", BASELINE_DIM)
def test_step_matches_reference_amp_bf16_baseline_dims(dev, name):
    """north_star parity bar, same mode, BASELINE dimensions: the CUDA step in amp_bf16 against the reference's
    own amp_bf16 run (fp32 master weights + bf16 autocast, tests/golden/*_amp_bf16.npz) at ViT-L/14 (82 and
    257 tokens), ViT-H/14 (head_dim 80) and ViT-B/16 shapes, reduced depth, batch 32:
      loss within 1e-3 relative of the same-mode reference loss (and of the fp32 reference loss);
      features / gradients within 2x the reference's own amp-vs-fp32 deviation (+2e-3) of the fp32 run."""
    meta, g32 = load_golden(name, "fp32")
    _, ga = load_golden(name, "amp_bf16")
    model, text, out, loss = run_ours(meta, "amp_bf16", dev)
    params = dict(model.named_parameters())
    last = meta["cfg"]["vision_cfg"]["layers"] - 1
    rows = torch.as_tensor(g32["token_rows_idx"])
    ours = {
        "image_features": out["image_features"].detach().float().cpu(),
        "text_features": out["text_features"].detach().float().cpu(),
        "d_image_features": out["image_features"].grad.float().cpu(),
        "d_text_features": out["text_features"].grad.float().cpu(),
        "g_visual_proj": params["visual.proj"].grad.float().cpu()[:64, :64],
        "g_text_projection": params["text_projection"].grad.float().cpu()[:64, :64],
        "g_v0_in_proj_weight": params["visual.transformer.resblocks.0.attn.in_proj_weight"].grad.float().cpu()[:64, :64],
        "g_v0_in_proj_bias": params["visual.transformer.resblocks.0.attn.in_proj_bias"].grad.float().cpu(),
        "g_vlast_c_fc_bias": params[f"visual.transformer.resblocks.{last}.mlp.c_fc.bias"].grad.float().cpu(),
        "g_t0_ln_1_weight": params["transformer.resblocks.0.ln_1.weight"].grad.float().cpu(),
        "g_class_embedding": params["visual.class_embedding"].grad.float().cpu(),
        "g_token_rows": params["token_embedding.weight"].grad.float().cpu()[rows],
    }
    rec = {"case": name, "precision": "amp_bf16", "same_mode_reference": True, "batch": meta["batch"], "errors": {}}
    failures = []
    for k, v in ours.items():
        e_ours, e_ref = rel_err(v, g32[k]), rel_err(ga[k], g32[k])
        rec["errors"][k] = {"ours_vs_ref_fp32": e_ours, "ref_amp_vs_ref_fp32": e_ref, "ours_vs_ref_amp": rel_err(v, ga[k])}
        if not e_ours <= 2.0 * e_ref + 2e-3:
            failures.append((k, e_ours, e_ref))
    l32, la = float(g32["loss"]), float(ga["loss"])
    rec["loss"] = {"ours": loss.item(), "ref_fp32": l32, "ref_amp_bf16": la, "rel_err_vs_amp": abs(loss.item() - la) / la,
                   "rel_err_vs_fp32": abs(loss.item() - l32) / l32, "ref_amp_vs_fp32": abs(la - l32) / l32}
    _report(rec)
    assert rec["loss"]["rel_err_vs_amp"] <= 1e-3, rec["loss"]
    assert rec["loss"]["rel_err_vs_fp32"] <= 1e-3, rec["loss"]
    assert not failures, failures


@pytest.mark.parametrize("name", CASES[:3])
def test_step_matches_oracle_features(dev, name):
    """Same weights and inputs through the CPU oracle (fp32) and the CUDA path (bf16)."""
    from oracle import clip_oracle as O
    from oracle.weights import make_inputs, make_state_dict
    meta, _ = load_golden(name, "fp32")
    sd = make_state_dict(meta["cfg"], meta["seed"], image_size=meta["image_size"], pos_embed=meta["pos_embed"])
    images, text = make_inputs(meta["cfg"], meta["batch"], meta["seed"] + 1000, image_size=meta["image_size"])
    fi, ft, s = O.clip_forward(images, text, sd, meta["cfg"])
    ref_loss = O.clip_loss(fi, ft, fi, ft, s, 0).item()
    model, _, out, loss = run_ours(meta, "amp_bf16", dev)
    assert rel_err(out["image_features"].detach().float().cpu(), fi) < 2e-2
    assert rel_err(out["text_features"].detach().float().cpu(), ft) < 2e-2
    # batch 4-6: the loss is maximally sensitive to bf16 noise -- the reference's OWN amp_bf16 run sits e_ref from its
    # fp32 run on these cases (2.9e-3 for tiny-bigvision); the 1e-3 bar is asserted at batch 32 on the BASELINE shapes
    _, g32 = load_golden(name, "fp32")
    _, ga = load_golden(name, "amp_bf16")
    e_ref = abs(float(ga["loss"]) - float(g32["loss"])) / float(g32["loss"])
    assert abs(loss.item() - ref_loss) / ref_loss < max(2e-3, 2 * e_ref), (loss.item(), ref_loss, e_ref)
    # un-normalised encode_* API (zero-shot path, training/zero_shot.py:36,75)
    with torch.no_grad():
        raw = model.encode_image(images.to(dev))
    assert rel_err(raw.float().cpu(), O.encode_image(images, sd, meta["cfg"])) < 2e-2


def test_state_dict_schema_and_checkpoint_roundtrip(dev, tmp_path):
    from clipa_b200 import open_clip
    meta, _ = load_golden("tiny-cls", "fp32")
    model = build_model(meta, "amp_bf16", dev)
    path = tmp_path / "ckpt.pt"
    torch.save({"epoch": 1, "name": "t", "state_dict": model.state_dict()}, path)   # training/main.py:436-448
    m2 = build_model(meta, "amp_bf16", dev)
    open_clip.load_checkpoint(m2, str(path))
    for (k1, v1), (k2, v2) in zip(model.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)


def test_optimizer_step_reduces_loss(dev):
    """A few AdamW steps of the reference training recipe (training/main.py:311-326,
    train.py:200-215,285-286) on a fixed batch drive the loss down: forward, loss, backward and
    the bf16 shadow-weight refresh after optimizer.step() all work together."""
    import math
    from clipa_b200 import open_clip
    from oracle.weights import make_inputs
    meta, _ = load_golden("tiny-cls", "fp32")
    model = build_model(meta, "amp_bf16", dev)
    images, text = make_inputs(meta["cfg"], 16, 5, image_size=meta["image_size"])
    images, text = images.to(dev), text.to(dev)
    exclude = lambda n, p: p.ndim < 2 or "bn" in n or "ln" in n or "bias" in n or "logit_scale" in n
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    opt = torch.optim.AdamW([{"params": [p for n, p in named if exclude(n, p)], "weight_decay": 0.},
                             {"params": [p for n, p in named if not exclude(n, p)], "weight_decay": 0.2}],
                            lr=1e-3, betas=(0.9, 0.98), eps=1e-6)
    loss_fn = open_clip.ClipLoss()
    losses = []
    for _ in range(8):
        opt.zero_grad()
        out = model(images, text)
        loss = loss_fn(**out)
        loss.backward()
        opt.step()
        with torch.no_grad():
            model.logit_scale.clamp_(0, math.log(100))
        losses.append(loss.item())
    assert losses[-1] < 0.7 * losses[0], losses


def test_fused_train_step_matches_torch_optimizer_step(dev):
    """TrainStep with the fused flat-buffer AdamW kernel == TrainStep with torch.optim.AdamW."""
    from clipa_b200.training import TrainStep
    from oracle.weights import make_inputs
    meta, _ = load_golden("tiny-cls", "fp32")
    images, text = make_inputs(meta["cfg"], 16, 5, image_size=meta["image_size"])
    results = []
    for fused in (True, False):
        model = build_model(meta, "amp_bf16", dev)
        ts = TrainStep(model, micro_batch=16, fused_optimizer=fused, lr=1e-3)
        assert ts.fused == fused
        losses = [ts.step(images, text).item() for _ in range(4)]
        results.append((losses, {k: v.detach().float().clone() for k, v in model.state_dict().items()}))
    (l_f, sd_f), (l_t, sd_t) = results
    assert l_f[-1] < l_f[0], l_f
    for a, b in zip(l_f, l_t):
        assert abs(a - b) / abs(b) < 2e-2, (l_f, l_t)      # bf16 forward, atomics in wgrad: not bit-identical
    # Adam's update is ~ +-lr per step whatever the gradient magnitude, so elements whose (noisy, bf16)
    # gradient is near zero may legitimately differ by up to 2*lr*steps; nothing may differ by more.
    bound = 2 * 1e-3 * 4 + 1e-4
    worst = max(((sd_f[k] - sd_t[k]).abs().max().item(), k) for k in sd_f)
    assert worst[0] <= bound, worst
    moved = max((sd_f[k] - sd_t[k]).abs().mean().item() for k in sd_f)
    assert moved < 1e-3, moved


@pytest.mark.parametrize("micro", [4, 5, 8])
def test_gradcache_chunks_match_single_pass(dev, micro):
    """TrainStep's GradCache schedule (train.py:216-256; ragged last chunk, last chunk's graph kept) gives the
    same loss and parameter gradients as one pass over the whole batch -- except logit_scale, whose gradient the
    reference's accumulation loop counts once per chunk (reproduced by default, tests/test_host_logic_cpu.py)."""
    from clipa_b200.training import TrainStep
    from oracle.weights import make_inputs
    meta, _ = load_golden("tiny-cls", "fp32")
    images, text = make_inputs(meta["cfg"], 16, 5, image_size=meta["image_size"])
    out = []
    for mb, quirk in ((16, True), (micro, False), (micro, True)):
        model = build_model(meta, "amp_bf16", dev)
        ts = TrainStep(model, micro_batch=mb, lr=1e-3, reference_accum_logit_scale=quirk)
        loss = ts.forward_backward(ts.preprocess(images), text.to(dev))
        out.append((loss.item(), {n: p.grad.float().clone() for n, p in model.named_parameters() if p.grad is not None}))
    (l0, g0), (l1, g1), (l2, g2) = out
    assert abs(l0 - l1) / abs(l0) < 2e-3, (l0, l1)       # features are bf16 either way; chunking changes GEMM tiling only
    for n in g0:
        assert rel_err(g1[n].cpu(), g0[n].cpu()) < 2e-2, n
    n_chunks = -(-16 // micro)
    assert rel_err(g2["logit_scale"].cpu(), n_chunks * g0["logit_scale"].cpu()) < 2e-2
    assert rel_err(g2["visual.proj"].cpu(), g0["visual.proj"].cpu()) < 2e-2


def test_patch_dropout_runs_on_device(dev):
    """PatchDropout (open_clip/transformer.py:53-90, pinned on CPU in tests/test_host_logic_cpu.py) inside the
    CUDA tower: the shortened sequence (CLS + kept patches) goes through the packed-tile attention path;
    deterministic for fixed scores, identity in eval mode, gradients finite."""
    from clipa_b200 import open_clip
    from clipa_b200.open_clip import PatchDropout
    from oracle.weights import make_inputs
    meta, _ = load_golden("tiny-cls", "fp32")
    model = build_model(meta, "amp_bf16", dev)
    images, text = make_inputs(meta["cfg"], 6, 3, image_size=meta["image_size"])
    images, text = images.to(dev), text.to(dev)
    model.eval()
    with torch.no_grad():
        base = model(images, text)["image_features"].float()
    n_patch = model.visual.positional_embedding.shape[0] - 1
    pd = PatchDropout(0.5)
    scores = torch.randn(6, n_patch, generator=torch.Generator().manual_seed(5)).to(dev)
    pd.score_fn = lambda b, n, d: scores
    model.visual.patch_dropout = pd
    model.eval()                       # the freshly attached module must follow the model's mode
    with torch.no_grad():
        assert torch.equal(model(images, text)["image_features"].float(), base)     # eval: identity
    model.train()
    out1 = model(images, text)
    with torch.no_grad():
        out2 = model(images, text)
    assert torch.equal(out1["image_features"], out2["image_features"])              # same scores -> same tokens
    assert (out1["image_features"].float() - base).abs().max() > 1e-3                # half the patches are gone
    loss = open_clip.ClipLoss()(out1["image_features"], out1["text_features"], out1["logit_scale"])
    loss.backward()
    g = model.visual.conv1.weight.grad
    assert g is not None and torch.isfinite(g.float()).all() and g.float().abs().sum() > 0


def test_activation_policy_save_ln_equals_recompute(dev):
    """The activation-memory levels (keep LayerNorm outputs / recompute them / also recompute the attention output /
    keep the MLP activations of all or of the last block) are the same math."""
    from clipa_b200.open_clip.transformer import Transformer
    meta, _ = load_golden("tiny-gap-h80", "fp32")
    grads = []
    old = (Transformer.save_ln_outputs, Transformer.recompute_attn_out, Transformer.keep_mlp_blocks)
    try:
        for save_ln, drop_o, keep in ((True, False, 0), (False, False, 0), (False, True, 0), (True, False, 99), (False, False, 1)):
            Transformer.save_ln_outputs, Transformer.recompute_attn_out, Transformer.keep_mlp_blocks = save_ln, drop_o, keep
            from clipa_b200.functional import ResidualBlockFn
            kept0 = ResidualBlockFn.kept_mlp_count
            model, _, _, loss = run_ours(meta, "amp_bf16", dev)
            n_blocks = len(model.visual.transformer.resblocks) + len(model.transformer.resblocks)
            assert ResidualBlockFn.kept_mlp_count - kept0 == (0 if keep == 0 else (n_blocks if keep == 99 else 2))
            grads.append((loss.item(), {n: p.grad.float().clone() for n, p in model.named_parameters() if p.grad is not None}))
    finally:
        Transformer.save_ln_outputs, Transformer.recompute_attn_out, Transformer.keep_mlp_blocks = old
    for l1, g1 in grads[1:]:
        assert grads[0][0] == l1
        for n in grads[0][1]:
            assert rel_err(grads[0][1][n].cpu(), g1[n].cpu()) < 1e-3, n     # only fp32-atomic summation order differs


def test_custom_text_clip_matches_clip(dev):
    """CustomTextCLIP (open_clip/model.py:277-326; text tower under `text.`, TextTransformer.forward
    transformer.py:638-681) computes the same features, loss and gradients as CLIP on converted weights."""
    import tempfile
    from clipa_b200 import open_clip
    from oracle.weights import make_inputs, make_state_dict
    meta, _ = load_golden("tiny-cls", "fp32")
    clip = build_model(meta, "amp_bf16", dev)
    ct, _, _ = open_clip.create_model_and_transforms(f"golden-{meta['name']}", precision="amp_bf16", device=dev,
                                                     force_image_size=meta["image_size"], pos_embed=meta["pos_embed"],
                                                     output_dict=True, force_custom_text=True)
    assert isinstance(ct, open_clip.CustomTextCLIP)
    ct.load_state_dict(open_clip.convert_to_custom_text_state_dict(clip.state_dict()), strict=True)
    ct.train()
    images, text = make_inputs(meta["cfg"], meta["batch"], meta["seed"] + 1000, image_size=meta["image_size"])
    images, text = images.to(dev), text.to(dev)
    outs = []
    for m in (clip, ct):
        o = m(images, text)
        loss = open_clip.ClipLoss()(o["image_features"], o["text_features"], o["logit_scale"])
        loss.backward()
        outs.append((o, loss))
    assert torch.equal(outs[0][0]["text_features"], outs[1][0]["text_features"])
    assert torch.equal(outs[0][0]["image_features"], outs[1][0]["image_features"])
    assert outs[0][1].item() == outs[1][1].item()
    g0 = clip.transformer.resblocks[0].attn.in_proj_weight.grad
    g1 = ct.text.transformer.resblocks[0].attn.in_proj_weight.grad
    assert torch.allclose(g0, g1, rtol=1e-4, atol=1e-7)       # fp32 atomics: same sum, different order
    # TextTransformer.forward slices the position table: a shorter text runs (CLIP.encode_text refuses it)
    short = text[:, :8].clone()
    short[:, -1] = meta["cfg"]["text_cfg"]["vocab_size"] - 1
    with torch.no_grad():
        assert ct.encode_text(short).shape == (meta["batch"], meta["cfg"]["embed_dim"])
        with pytest.raises(ValueError):
            clip.encode_text(short)


FP32_CASES = CASES + ["vitl14-i81-t16-d2", "vith14-i36-t8-d2", "vitb16-i64-t16-d3", "vitl14-i256-t32-d2"]


@pytest.mark.parametrize("name", FP32_CASES)
def test_fp32_mode_matches_reference_fp32(dev, name):
    """north_star "within 1e-5 (fp32)": precision='fp32' (CUDA-core fp32 kernels, clipa_b200/fp32_path.py) against the
    reference's fp32 run -- features and loss to 1e-5 relative, gradients to 5e-5 -- on the small cases and at the
    BASELINE widths / head shapes / sequence lengths (batch 32)."""
    meta, g = load_golden(name, "fp32")
    model, text, out, loss = run_ours(meta, "fp32", dev)
    assert out["image_features"].dtype == torch.float32 and loss.dtype == torch.float32
    params = dict(model.named_parameters())
    compact = meta.get("compact", False)
    cut = (lambda t: t[:64, :64]) if compact else (lambda t: t)
    last = meta["cfg"]["vision_cfg"]["layers"] - 1
    rows = torch.as_tensor(g["token_rows_idx"])
    rec = {"case": name, "precision": "fp32", "errors": {}}
    checks = {
        "image_features": (out["image_features"].detach(), 1e-5), "text_features": (out["text_features"].detach(), 1e-5),
        "d_image_features": (out["image_features"].grad, 5e-5), "d_text_features": (out["text_features"].grad, 5e-5),
        "g_visual_proj": (cut(params["visual.proj"].grad), 5e-5),
        "g_text_projection": (cut(params["text_projection"].grad), 5e-5),
        "g_v0_in_proj_weight": (params["visual.transformer.resblocks.0.attn.in_proj_weight"].grad[:64, :64], 1e-4),
        "g_v0_in_proj_bias": (params["visual.transformer.resblocks.0.attn.in_proj_bias"].grad, 1e-4),
        "g_vlast_c_fc_bias": (params[f"visual.transformer.resblocks.{last}.mlp.c_fc.bias"].grad, 1e-4),
        "g_t0_ln_1_weight": (params["transformer.resblocks.0.ln_1.weight"].grad, 1e-4),
        "g_class_embedding": (params["visual.class_embedding"].grad, 1e-4),
        "g_token_rows": (params["token_embedding.weight"].grad[rows.to(dev)], 1e-4),
    }
    bad = []
    for k, (v, tol) in checks.items():
        e = rel_err(v.float().cpu(), g[k])
        rec["errors"][k] = e
        if not e < tol:
            bad.append((k, e, tol))
    l32 = float(g["loss"])
    rec["loss"] = {"ours": loss.item(), "ref_fp32": l32, "rel_err": abs(loss.item() - l32) / l32}
    gs = params["logit_scale"].grad.item()
    rec["g_logit_scale"] = {"ours": gs, "ref_fp32": float(g["g_logit_scale"])}
    _report(rec)
    assert rec["loss"]["rel_err"] < 1e-5, rec["loss"]
    assert abs(gs - float(g["g_logit_scale"])) <= 1e-4 * abs(float(g["g_logit_scale"])) + 1e-7, rec["g_logit_scale"]
    assert not bad, bad


def test_fp16_precision_refuses_to_run_on_gpu(dev):
    from clipa_b200 import open_clip
    m = open_clip.create_model("ViT-B-32-CL16", precision="amp", device=dev, force_image_size=64)
    with pytest.raises(NotImplementedError, match="amp_bf16"):
        m(torch.zeros(2, 3, 64, 64, device=dev), torch.zeros(2, 16, dtype=torch.long, device=dev))


def test_device_side_grad_clip_and_resume(dev):
    """TrainStep with grad_clip_norm: the clip factor is computed and applied on the device (no .item()); the step
    equals torch's clip_grad_norm_ + AdamW.  state_dict()/load_state_dict() resume: a restored trainer continues
    on the same trajectory as the uninterrupted one."""
    from clipa_b200.training import TrainStep
    from oracle.weights import make_inputs
    meta, _ = load_golden("tiny-cls", "fp32")
    images, text = make_inputs(meta["cfg"], 6, 5, image_size=meta["image_size"])
    images, text = images.to(dev), text.to(dev)

    def run(fused, steps, resume_at=None):
        model = build_model(meta, "amp_bf16", dev)
        ts = TrainStep(model, micro_batch=16, lr=1e-3, grad_clip_norm=0.5, fused_optimizer=fused)
        losses = []
        for i in range(steps):
            if resume_at is not None and i == resume_at:
                sd_opt, sd_model = ts.state_dict(), {k: v.clone() for k, v in model.state_dict().items()}
                model = build_model(meta, "amp_bf16", dev)
                model.load_state_dict(sd_model)
                ts = TrainStep(model, micro_batch=16, lr=1e-3, grad_clip_norm=0.5, fused_optimizer=fused)
                ts.load_state_dict(sd_opt)
            losses.append(ts.step(images, text).item())
        return losses, model
    l_fused, m_fused = run(True, 4)
    l_torch, m_torch = run(False, 4)
    for a, b in zip(l_fused, l_torch):
        assert abs(a - b) < 2e-2 * abs(b) + 2e-3, (l_fused, l_torch)      # the loss falls 15x in 4 steps: bf16 noise compounds
    w_f = m_fused.visual.transformer.resblocks[0].mlp.c_fc.weight.float()
    w_t = m_torch.visual.transformer.resblocks[0].mlp.c_fc.weight.float()
    assert ((w_f - w_t).norm() / w_t.norm()).item() < 2e-3
    l_res, _ = run(True, 4, resume_at=2)
    for a, b in zip(l_res, l_fused):
        assert abs(a - b) < 2e-2 * abs(b) + 2e-3, (l_res, l_fused)
