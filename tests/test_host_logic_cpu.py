"""CPU: host-side mirror of the reference interface -- registry, module tree / state_dict schema,
precision policy, weight-decay grouping, position embedding -- against facts dumped from the
reference (tests/golden/reference_schema.json, sincos_ref.npz)."""
import json

import numpy as np
import pytest
import torch

from clipa_b200 import open_clip
from oracle.weights import CONFIG1, TINY_CONFIGS
from tests.helpers import GOLD


@pytest.fixture(scope="module")
def schema():
    return json.loads((GOLD / "reference_schema.json").read_text())


def _register(tmp_path, name, cfg):
    p = tmp_path / f"{name}.json"
    p.write_text(json.dumps(cfg))
    open_clip.add_model_config(p)


@pytest.mark.parametrize("name,cfg,size", [("tiny-cls", TINY_CONFIGS["tiny-cls"], 64), ("ViT-B-32-ctx16", CONFIG1, 192)])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_state_dict_schema_and_dtypes_match_reference(tmp_path, schema, name, cfg, size, precision):
    _register(tmp_path, f"schema-{name}", cfg)
    m = open_clip.create_model(f"schema-{name}", precision=precision, device="cpu", force_image_size=size,
                               pos_embed="sin_cos_2d")
    ref = schema[f"{name}/{precision}"]
    ours = {k: [list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()}
    assert list(ours.keys()) == list(ref["state_dict"].keys())          # same keys, same order
    assert ours == ref["state_dict"]                                    # same shapes and dtypes
    assert {k: bool(p.requires_grad) for k, p in m.named_parameters()} == ref["requires_grad"]
    assert [k for k, _ in m.named_buffers() if k not in m.state_dict()] == ref["buffers_non_persistent"]


def test_builtin_registry_covers_baseline_configs():
    want = {
        "ViT-B-16-CL16": (768, 12, 512, 12, 16, 512),
        "ViT-L-14-CL16": (1024, 24, 768, 12, 16, 768),
        "ViT-L-14-CL32": (1024, 24, 768, 12, 32, 768),
        "ViT-H-14-CL8-SyntaxMask-GAP": (1280, 32, 1024, 24, 8, 1024),
        "ViT-L-14": (1024, 24, 768, 12, 77, 768),
    }
    for name, (vw, vl, tw, tl, ctx, e) in want.items():
        c = open_clip.get_model_config(name)
        assert c is not None, name
        assert (c["vision_cfg"]["width"], c["vision_cfg"]["layers"], c["text_cfg"]["width"],
                c["text_cfg"]["layers"], c["text_cfg"]["context_length"], c["embed_dim"]) == (vw, vl, tw, tl, ctx, e)
    assert open_clip.get_model_config("ViT-H-14")["vision_cfg"]["head_width"] == 80
    bv = open_clip.get_model_config("ViT-L-14-CL32-GAP-BigVision")
    assert bv["vision_cfg"]["pool_style"] == "big_vision_gap" and bv["text_cfg"]["attention_mask"] is False
    assert open_clip.get_model_config("no-such-model") is None
    with pytest.raises(RuntimeError):
        open_clip.create_model("no-such-model")


def test_sincos_matches_reference():
    from clipa_b200.open_clip.pos_embed import get_2d_sincos_pos_embed
    z = np.load(GOLD / "sincos_ref.npz")
    assert np.abs(get_2d_sincos_pos_embed(128, 3, True).numpy() - z["w128_g3"]).max() < 1e-6
    assert np.abs(get_2d_sincos_pos_embed(768, 6, True).numpy() - z["w768_g6"]).max() < 1e-6


def test_weight_decay_split_follows_reference_rule(tmp_path):
    from clipa_b200.training import exclude_from_wd
    _register(tmp_path, "wd-tiny", TINY_CONFIGS["tiny-cls"])
    m = open_clip.create_model("wd-tiny")
    no_decay = {n for n, p in m.named_parameters() if exclude_from_wd(n, p)}
    assert "logit_scale" in no_decay and "visual.class_embedding" in no_decay
    assert "visual.transformer.resblocks.0.ln_1.weight" in no_decay
    assert "transformer.resblocks.1.mlp.c_fc.bias" in no_decay
    assert "visual.transformer.resblocks.0.attn.in_proj_weight" not in no_decay
    assert "visual.proj" not in no_decay and "token_embedding.weight" not in no_decay


def test_unsupported_features_raise(tmp_path):
    cfg = json.loads(json.dumps(TINY_CONFIGS["tiny-cls"]))
    cfg["vision_cfg"]["timm_model_name"] = "resnet50"
    _register(tmp_path, "bad-timm", cfg)
    with pytest.raises(NotImplementedError):
        open_clip.create_model("bad-timm")
    with pytest.raises(NotImplementedError):
        open_clip.create_model("ViT-B-32", precision="fp16")
    # PatchDropout is built: the factory override reaches the tower
    from clipa_b200.open_clip import PatchDropout
    m = open_clip.create_model("ViT-B-32", force_patch_dropout=0.5)
    assert isinstance(m.visual.patch_dropout, PatchDropout) and m.visual.patch_dropout.prob == 0.5


def test_create_loss_signature():
    from types import SimpleNamespace
    args = SimpleNamespace(distill=False, model="ViT-L-14", local_loss=True, gather_with_grad=True, rank=3,
                           world_size=8, horovod=False)
    loss = open_clip.create_loss(args)
    assert isinstance(loss, open_clip.ClipLoss) and loss.rank == 3 and loss.world_size == 8 and loss.local_loss


def test_gradcache_schedule_matches_single_pass_cpu():
    """TrainStep.forward_backward's chunked (GradCache, train.py:216-256) schedule -- including the kept graph
    of the last chunk and a ragged tail -- accumulates the same gradients as one pass.  Pure host logic: a small
    torch stand-in model and a torch contrastive loss replace the CUDA path."""
    import torch
    import torch.nn.functional as F
    from clipa_b200.training import TrainStep

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(0)
            self.vis = torch.nn.Linear(3 * 4 * 4, 8)
            self.txt = torch.nn.Embedding(11, 8)
            self.logit_scale = torch.nn.Parameter(torch.tensor(2.0))
            self.visual = self.vis
            self.calls = 0

        def forward(self, images, texts):
            self.calls += 1
            i = F.normalize(self.vis(images.float().flatten(1)), dim=-1)
            t = F.normalize(self.txt(texts).mean(1), dim=-1)
            return i, t, self.logit_scale.exp()

    def loss_fn(i, t, scale):
        logits = scale * i @ t.t()
        labels = torch.arange(i.shape[0])
        return (F.cross_entropy(logits, labels) + F.cross_entropy(logits.t(), labels)) / 2

    torch.manual_seed(1)
    images = torch.randn(10, 3, 4, 4)
    texts = torch.randint(0, 11, (10, 5))
    grads = []
    for mb in (10, 4, 3):
        model = Toy()
        ts = TrainStep(model, micro_batch=mb, image_mean=(0., 0., 0.), image_std=(1., 1., 1.),
                       reference_accum_logit_scale=False)
        ts.loss_fn = loss_fn
        loss = ts.forward_backward(images, texts)
        n_chunks = -(-10 // mb)
        assert model.calls == (1 if n_chunks == 1 else 2 * n_chunks - 1)
        grads.append((loss.item(), [p.grad.clone() for p in ts.params]))
    for l, g in grads[1:]:
        assert abs(l - grads[0][0]) < 1e-6
        for a, b in zip(g, grads[0][1]):
            assert torch.allclose(a, b, atol=1e-6, rtol=1e-5)
    # default: the reference's accumulation loop (train.py:243-256) back-propagates the FULL loss once per chunk,
    # so its logit_scale gradient is n_chunks x the single-pass one; every other parameter is unaffected
    model = Toy()
    ts = TrainStep(model, micro_batch=4, image_mean=(0., 0., 0.), image_std=(1., 1., 1.))
    ts.loss_fn = loss_fn
    ts.forward_backward(images, texts)
    single = dict(zip([id(p) for p in grads[0][1]], grads[0][1]))
    for (name, p), ref in zip(model.named_parameters(), grads[0][1]):
        if name == "logit_scale":
            assert torch.allclose(p.grad, 3 * ref, atol=1e-6, rtol=1e-5)
        else:
            assert torch.allclose(p.grad, ref, atol=1e-6, rtol=1e-5)


def test_reference_accumulation_multiplies_the_logit_scale_gradient():
    """The quirk reproduced above, checked against the reference's own loop body (train.py:216-256) run on the
    same toy model: cached no-grad features, then per chunk a re-forward + full loss + backward."""
    import torch
    import torch.nn.functional as F
    torch.manual_seed(0)
    vis, txt = torch.nn.Linear(12, 8), torch.nn.Linear(5, 8)
    logit_scale = torch.nn.Parameter(torch.tensor(2.0))
    x, y = torch.randn(9, 12), torch.randn(9, 5)

    def fwd(a, b):
        return F.normalize(vis(a), dim=-1), F.normalize(txt(b), dim=-1), logit_scale.exp()

    def loss_fn(i, t, s):
        lg = s * i @ t.t()
        lab = torch.arange(i.shape[0])
        return (F.cross_entropy(lg, lab) + F.cross_entropy(lg.t(), lab)) / 2
    loss_fn(*fwd(x, y)).backward()
    g_single = logit_scale.grad.clone()
    g_w = vis.weight.grad.clone()
    for p in (vis.weight, vis.bias, txt.weight, txt.bias, logit_scale):
        p.grad = None
    chunks = [(0, 3), (3, 6), (6, 9)]
    with torch.no_grad():
        cache = [fwd(x[s:e], y[s:e])[:2] for s, e in chunks]
    for j, (s, e) in enumerate(chunks):                 # train.py:243-256
        i_j, t_j, sc = fwd(x[s:e], y[s:e])
        ii = torch.cat([c[0] for c in cache[:j]] + [i_j] + [c[0] for c in cache[j + 1:]])
        tt = torch.cat([c[1] for c in cache[:j]] + [t_j] + [c[1] for c in cache[j + 1:]])
        loss_fn(ii, tt, sc).backward()
    assert torch.allclose(logit_scale.grad, 3 * g_single, rtol=1e-5)
    assert torch.allclose(vis.weight.grad, g_w, rtol=1e-4, atol=1e-6)


def _host_gold():
    import numpy as np
    from pathlib import Path
    return np.load(Path(__file__).parent / "golden" / "host_ops_ref.npz")


def test_patch_dropout_matches_reference_golden():
    """PatchDropout (open_clip/transformer.py:53-90) with the reference's scores injected: same kept tokens in
    the same (top-k) order, CLS in front; identity in eval mode."""
    import torch
    from clipa_b200.open_clip import PatchDropout
    g = _host_gold()
    for tag in ("pd_a", "pd_b", "pd_c"):
        x = torch.from_numpy(g[f"{tag}_x"])
        scores = torch.from_numpy(g[f"{tag}_scores"])
        pd = PatchDropout(float(g[f"{tag}_prob"]))
        pd.score_fn = lambda b, n, dev, s=scores: s
        pd.train()
        y = pd(x)
        assert y.shape == tuple(g[f"{tag}_y"].shape)
        assert torch.equal(y, torch.from_numpy(g[f"{tag}_y"])), tag
        pd.eval()
        assert pd(x) is x
    # default score_fn draws on the activations' device and keeps max(1, int(n * (1 - p))) patch tokens
    pd = PatchDropout(0.5)
    pd.train()
    y = pd(torch.randn(2, 82, 4))
    assert y.shape == (2, 1 + int(81 * 0.5), 4)


def test_pos_embed_resizers_match_reference_golden():
    """resize_pos_embed / resize_text_pos_embed (open_clip/model.py:452-516) against the reference's outputs."""
    import torch
    from types import SimpleNamespace
    from clipa_b200.open_clip import resize_pos_embed, resize_text_pos_embed
    g = _host_gold()
    for tag in ("pe_up", "pe_down"):
        new = g[f"{tag}_new"]
        grid = int(round((new.shape[0] - 1) ** 0.5))
        sd = {"visual.positional_embedding": torch.from_numpy(g[f"{tag}_old"].copy())}
        resize_pos_embed(sd, SimpleNamespace(visual=SimpleNamespace(grid_size=(grid, grid))))
        assert torch.equal(sd["visual.positional_embedding"], torch.from_numpy(new)), tag
    for tag in ("te_up", "te_down"):
        new = g[f"{tag}_new"]
        sd = {"positional_embedding": torch.from_numpy(g[f"{tag}_old"].copy())}
        resize_text_pos_embed(sd, SimpleNamespace(positional_embedding=torch.zeros(new.shape)))
        assert torch.equal(sd["positional_embedding"], torch.from_numpy(new)), tag
    # same grid / length: untouched
    sd = {"visual.positional_embedding": torch.zeros(1 + 9, 4), "positional_embedding": torch.zeros(16, 4)}
    keep = sd["visual.positional_embedding"]
    resize_pos_embed(sd, SimpleNamespace(visual=SimpleNamespace(grid_size=(3, 3))))
    resize_text_pos_embed(sd, SimpleNamespace(positional_embedding=torch.zeros(16, 4)))
    assert sd["visual.positional_embedding"] is keep


def test_load_checkpoint_resamples_position_tables(tmp_path):
    """factory.load_checkpoint resamples a checkpoint saved at another image grid / context length
    (open_clip/factory.py:110-118) instead of failing on the shape mismatch."""
    import torch
    from clipa_b200 import open_clip
    from oracle.weights import TINY_CONFIGS
    import copy, json
    cfg = copy.deepcopy(TINY_CONFIGS["tiny-cls"])
    d = tmp_path / "cfgs"
    d.mkdir()
    (d / "ckpt-tiny.json").write_text(json.dumps(cfg))
    open_clip.add_model_config(d)
    small = open_clip.create_model("ckpt-tiny", precision="fp32", device="cpu", force_image_size=32)
    torch.save({"state_dict": {"module." + k: v for k, v in small.state_dict().items()}}, tmp_path / "small.pt")
    big = open_clip.create_model("ckpt-tiny", precision="fp32", device="cpu", force_image_size=64,
                                 pretrained=str(tmp_path / "small.pt"))
    gs, gb = small.visual.grid_size[0], big.visual.grid_size[0]
    assert gb == 2 * gs and big.visual.positional_embedding.shape[0] == 1 + gb * gb
    assert torch.equal(big.visual.positional_embedding[0], small.visual.positional_embedding[0])      # CLS row kept
    assert torch.equal(big.visual.conv1.weight, small.visual.conv1.weight)


def test_grad_sink_is_opt_in():
    """functional.grad_sink: gradient kernels write straight into `.grad` only for parameters that opted in
    (TrainStep's flat-buffer views) and whose `.grad` is a contiguous fp32 tensor of the parameter's shape."""
    import torch
    from clipa_b200.functional import grad_sink
    p = torch.nn.Parameter(torch.zeros(4, 8))
    assert grad_sink(None) is None and grad_sink(p) is None            # no flag
    p._clipa_direct_grad = True
    assert grad_sink(p) is None                                          # no .grad yet
    p.grad = torch.zeros(4, 8)
    assert grad_sink(p) is p.grad
    pb = torch.nn.Parameter(torch.zeros(4, 8, dtype=torch.bfloat16))
    pb._clipa_direct_grad = True
    pb.grad = torch.zeros(4, 8, dtype=torch.bfloat16)
    assert grad_sink(pb) is None                                         # only fp32 gradient buffers qualify
    flat = torch.zeros(64)
    p.grad = None
    p.grad = flat[:32].view(4, 8)
    assert grad_sink(p) is p.grad and grad_sink(p).data_ptr() == flat.data_ptr()
    q = torch.nn.Parameter(torch.zeros(8, 4))
    q._clipa_direct_grad = True
    q.grad = torch.zeros(8, 4)
    q.grad = q.grad.t().contiguous().t()                                 # non-contiguous view
    assert grad_sink(q) is None


def test_unbuilt_precisions_refuse_to_run():
    """precision='amp' (fp16 autocast) and 'fp16' have no arithmetic on this path: the model is a parameter container
    (schema / checkpoints) and forward() raises instead of silently running other math.  'fp32' (the factory default)
    is the CUDA-core parity mode: it passes the precision check and then, like every mode, refuses CPU tensors."""
    import torch
    from clipa_b200 import open_clip
    m = open_clip.create_model("ViT-B-32-CL16", precision="amp", device="cpu", force_image_size=64)
    assert m.compute_precision == "amp" and next(m.parameters()).dtype == torch.float32
    with pytest.raises(NotImplementedError, match="amp_bf16"):
        m(torch.zeros(1, 3, 64, 64), torch.zeros(1, 16, dtype=torch.long))
    with pytest.raises(NotImplementedError, match="amp_bf16"):
        m.encode_text(torch.zeros(1, 16, dtype=torch.long))
    with pytest.raises(NotImplementedError):
        open_clip.create_model("ViT-B-32-CL16", precision="fp16", device="cpu")
    for prec in ("amp_bf16", "fp32"):
        ok = open_clip.create_model("ViT-B-32-CL16", precision=prec, device="cpu", force_image_size=64)
        with pytest.raises(Exception) as e:     # the precision check passes; the kernels then refuse CPU tensors
            ok.encode_text(torch.zeros(1, 16, dtype=torch.long))
        assert not isinstance(e.value, NotImplementedError)


def test_custom_text_clip_schema_and_checkpoint_conversion():
    """CustomTextCLIP (open_clip/model.py:277-326, --force-custom-text): text tower under `text.*`, same tensors
    as CLIP; a CLIP-format checkpoint loads through convert_to_custom_text_state_dict (factory.py:110-118)."""
    import torch
    from clipa_b200 import open_clip
    clip = open_clip.create_model("ViT-B-32-CL16", precision="amp_bf16", device="cpu", force_image_size=64)
    ct = open_clip.create_model("ViT-B-32-CL16", precision="amp_bf16", device="cpu", force_image_size=64,
                                force_custom_text=True)
    assert isinstance(ct, open_clip.CustomTextCLIP) and isinstance(ct.text, open_clip.TextTransformer)
    sd = clip.state_dict()
    conv = open_clip.convert_to_custom_text_state_dict(sd)
    assert set(conv) == set(ct.state_dict())
    assert {k for k in conv if k.startswith("text.")} == {"text." + k for k in sd if k.split(".")[0] in
                                                          ("text_projection", "positional_embedding", "token_embedding",
                                                           "transformer", "ln_final")}
    ct.load_state_dict(conv, strict=True)
    assert torch.equal(ct.text.text_projection, clip.text_projection)
    assert ct.context_length == 16 and ct.vocab_size == clip.vocab_size
    ct.lock_text_tower(unlocked_layers=1)
    assert not ct.text.token_embedding.weight.requires_grad
    assert ct.text.transformer.resblocks[-1].mlp.c_fc.weight.requires_grad
    assert not ct.text.transformer.resblocks[0].mlp.c_fc.weight.requires_grad


def test_train_step_optimizer_state_dict_is_adamw_compatible():
    """TrainStep.state_dict()/load_state_dict(): torch.optim.AdamW layout (state[idx] = step / exp_avg /
    exp_avg_sq, two param groups in main.py:318-326 order), so --resume checkpoints interchange with the
    reference optimizer; flat segments are padded to 8 elements (16-byte aligned bf16 shadows)."""
    import torch
    from clipa_b200.training import TrainStep, exclude_from_wd
    torch.manual_seed(0)
    model = torch.nn.Sequential()
    model.visual = torch.nn.Linear(5, 3)            # 15 + 3 elements: not multiples of 8
    model.ln = torch.nn.LayerNorm(3)
    model.logit_scale = torch.nn.Parameter(torch.tensor(1.0))
    ts = TrainStep(model, micro_batch=4, image_mean=(0., 0., 0.), image_std=(1., 1., 1.))
    assert ts.fused
    for grp in ts._groups:
        assert all(sz % 8 == 0 for sz in grp["sizes"])
        off = 0
        for p, sz in zip(grp["params"], grp["sizes"]):
            assert p.data_ptr() == grp["p"][off:].data_ptr() and (p.data_ptr() - grp["p"].data_ptr()) % 32 == 0
            off += sz
    named = list(model.named_parameters())
    gain = [p for n, p in named if exclude_from_wd(n, p)]
    rest = [p for n, p in named if not exclude_from_wd(n, p)]
    ref = torch.optim.AdamW([{"params": gain, "weight_decay": 0.}, {"params": rest, "weight_decay": 0.2}],
                            lr=1e-3, betas=(0.9, 0.95), eps=1e-6)
    for p in gain + rest:
        p.grad = torch.randn_like(p)
    ref.step()
    ref_sd = ref.state_dict()
    ts.load_state_dict(ref_sd)
    assert ts.step_count == 1
    mine = ts.state_dict()
    assert [g["params"] for g in mine["param_groups"]] == [g["params"] for g in ref_sd["param_groups"]]
    assert [g["weight_decay"] for g in mine["param_groups"]] == [0.0, 0.2]
    for k, st in ref_sd["state"].items():
        assert torch.equal(mine["state"][k]["exp_avg"], st["exp_avg"])
        assert torch.equal(mine["state"][k]["exp_avg_sq"], st["exp_avg_sq"])
        assert float(mine["state"][k]["step"]) == float(st["step"])
    ref.load_state_dict(mine)                       # and back into the reference optimizer


def test_activation_levels_follow_the_free_hbm():
    """Transformer's activation-memory policy as plain arithmetic: ViT-L/14 vision tower,
    82 tokens; one unit = one [pairs*82, 1024] bf16 tensor."""
    from clipa_b200.open_clip.transformer import activation_levels
    GiB = 1 << 30
    L, r2 = 24, 8.0
    unit4096 = 4096 * 82 * 1024 * 2
    unit2048 = unit4096 // 2
    # 4096 pairs, ~165 GB free for the tower: LayerNorm outputs kept, a handful of blocks keep their MLP activations
    save_ln, drop_o, keep = activation_levels(165 * GiB, unit4096, L, r2)
    assert save_ln and not drop_o and 0 < keep < 8
    # 2048 pairs with the same room: every block keeps them
    assert activation_levels(165 * GiB, unit2048, L, r2) == (True, False, L)
    # tight: below the 8-unit level -> recompute the LayerNorm outputs, keep nothing; tighter still -> also drop O
    need6 = (6 * L + 10) * unit4096
    assert activation_levels(need6 + 4 * GiB, unit4096, L, r2) == (False, False, 0)
    assert activation_levels(need6 + 1 * GiB, unit4096, L, r2) == (False, True, 0)
    # the reservation for the tower that runs afterwards only limits the MLP-keeping decision
    full = activation_levels(165 * GiB, unit2048, L, r2, keep_reserve=0)[2]
    held = activation_levels(165 * GiB, unit2048, L, r2, keep_reserve=100 * GiB)[2]
    assert full == L and held < L
    # forced levels override the arithmetic and are clamped to the depth of the tower
    assert activation_levels(0, unit4096, L, r2, save_ln=True, drop_o=False, keep_mlp=99) == (True, False, L)
    assert activation_levels(1 << 50, unit4096, L, r2, save_ln=False, drop_o=True, keep_mlp=0) == (False, True, 0)


def test_bench_micro_batch_choice():
    """bench.py: one plain step when the per-GPU batch fits, else the workload's GradCache chunk size."""
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("clipa_bench", Path(__file__).resolve().parent.parent / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    wl = bench.WORKLOADS["vitl14_i81_t16_gb32k"]
    assert [bench.pick_micro_batch(wl, wl["global_batch"] // n) for n in (1, 2, 4, 8)] == [2048, 2048, 2048, 4096]
    for name, w in bench.WORKLOADS.items():
        plain, chunk = w["micro"]
        assert chunk <= plain and w["global_batch"] % chunk == 0, name
        assert (w["global_batch"] // 8) % bench.pick_micro_batch(w, w["global_batch"] // 8) == 0, name
