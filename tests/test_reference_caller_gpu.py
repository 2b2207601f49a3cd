"""Drop-in proof: the REFERENCE's own training loop (clipa_torch/training/train.py:158-314 `train_one_epoch`,
unmodified, from baseline/_ref) and its zero-shot evaluator (training/zero_shot.py) run over this package aliased as
`open_clip`: bf16 autocast, the reference's optimizer construction and loss call
(`loss(**model_out, output_dict=True)`), both the plain and the accum_freq > 1 (GradCache) branches.
Skipped where baseline/_ref is not installed (tools/install_reference.sh)."""
import importlib
import importlib.machinery
import math
import sys
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.fixture()
def reference_training(dev):
    """Imports the reference's `training` package with `open_clip` resolving to clipa_b200.open_clip."""
    from baseline import ref_loader
    root = ref_loader.reference_root()
    if root is None:
        pytest.skip("baseline/_ref not installed")
    import clipa_b200.open_clip as ours
    saved = {k: v for k, v in sys.modules.items() if k == "open_clip" or k.startswith("open_clip.") or
             k == "training" or k.startswith("training.")}
    for k in saved:
        del sys.modules[k]
    # stubs for the tokenizer / data-loader dependencies that are not installed (see baseline/ref_loader.py)
    for n in ref_loader._STUBS:
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                from unittest.mock import MagicMock
                m = MagicMock()
                m.__spec__ = importlib.machinery.ModuleSpec(n, None)
                m.__path__ = []
                sys.modules[n] = m
    sys.modules["open_clip"] = ours
    for sub in ("factory", "model", "loss", "transformer", "pos_embed", "model_configs"):
        sys.modules[f"open_clip.{sub}"] = importlib.import_module(f"clipa_b200.open_clip.{sub}")
    sys.path.insert(0, str(root))
    try:
        train = importlib.import_module("training.train")
        zero_shot = importlib.import_module("training.zero_shot")
        assert train.CLIP is ours.CLIP                      # train.py:28 imported OUR classes
        yield SimpleNamespace(train=train, zero_shot=zero_shot, open_clip=ours)
    finally:
        sys.path.remove(str(root))
        for k in [k for k in sys.modules if k == "open_clip" or k.startswith("open_clip.") or k == "training" or
                  k.startswith("training.")]:
            del sys.modules[k]
        sys.modules.update(saved)


class _Loader(list):
    """The two attributes train_one_epoch reads from a dataloader (train.py:171-172)."""

    def __init__(self, batches, batch_size):
        super().__init__(batches)
        self.num_batches = len(batches)
        self.num_samples = len(batches) * batch_size


def _args(dev, accum_freq, batch_size):
    return SimpleNamespace(device=str(dev), precision="amp_bf16", accum_freq=accum_freq, distill=False, skip_scheduler=True,
                           to_float_on_device=True, image_mean=None, image_std=None, grad_clip_norm=1.0, horovod=False,
                           distributed=False, log_every_n_steps=1, batch_size=batch_size, world_size=1, rank=0, local_rank=0,
                           wandb=False, zeroshot_steps=0, val_steps=0, local_loss=True, gather_with_grad=True,
                           model="ViT-B-32-CL16", lr=1e-3, beta1=0.9, beta2=0.95, eps=1e-6, wd=0.2)


def _build(ref, dev, args):
    torch.manual_seed(0)
    model, _, _ = ref.open_clip.create_model_and_transforms(args.model, precision=args.precision, device=dev,
                                                            force_image_size=64, pos_embed="sin_cos_2d", output_dict=True)
    model.set_grad_checkpointing()
    exclude = lambda n, p: p.ndim < 2 or "bn" in n or "ln" in n or "bias" in n or "logit_scale" in n   # main.py:311-326
    named = list(model.named_parameters())
    optimizer = torch.optim.AdamW(
        [{"params": [p for n, p in named if exclude(n, p) and p.requires_grad], "weight_decay": 0.},
         {"params": [p for n, p in named if not exclude(n, p) and p.requires_grad], "weight_decay": args.wd}],
        lr=args.lr, betas=(args.beta1, args.beta2), eps=args.eps)
    return model, optimizer, ref.open_clip.create_loss(args)


def _batches(n, bs, vocab):
    g = torch.Generator().manual_seed(3)
    out = []
    for _ in range(n):
        t = torch.randint(1, vocab - 1, (bs, 16), generator=g)
        t[:, -1] = vocab - 1
        out.append((torch.randint(0, 256, (bs, 3, 64, 64), generator=g, dtype=torch.uint8), t))
    return out


@pytest.mark.parametrize("accum_freq", [1, 2])
def test_reference_train_one_epoch_runs_over_the_alias(dev, reference_training, accum_freq, caplog):
    ref = reference_training
    bs = 16
    args = _args(dev, accum_freq, bs)
    model, optimizer, loss = _build(ref, dev, args)
    from training.data import DataInfo
    batches = _batches(6, bs, model.vocab_size)
    data = {"train": DataInfo(dataloader=_Loader(batches, bs))}
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    import logging
    with caplog.at_level(logging.INFO):
        for epoch in range(3):
            ref.train.train_one_epoch(model, data, loss, epoch, optimizer, None, None, None, args)
    torch.cuda.synchronize()
    logged = [float(r.getMessage().split("Contrastive_loss: ")[1].split()[0]) for r in caplog.records
              if "Contrastive_loss" in r.getMessage()]
    assert len(logged) >= 6 and all(math.isfinite(v) for v in logged)
    assert logged[-1] < logged[0] - 0.05, logged             # it trains: the same 6 batches, three epochs
    moved = sum(float((p.detach() - before[n]).abs().sum()) for n, p in model.named_parameters())
    assert moved > 0 and all(torch.isfinite(p).all() for p in model.parameters())
    assert 0 <= model.logit_scale.item() <= math.log(100) + 1e-6
    # the bf16 shadow weights the GEMMs read follow the torch optimizer's in-place updates
    from clipa_b200.functional import compute_copy
    w = model.visual.transformer.resblocks[0].mlp.c_fc.weight
    assert torch.equal(compute_copy(w), w.detach().to(torch.bfloat16))


def test_first_step_loss_equals_trainstep(dev, reference_training):
    """Same init, same batch: the loss the reference loop logs for its first step == TrainStep's first loss."""
    ref = reference_training
    bs = 16
    args = _args(dev, 1, bs)
    model, optimizer, loss = _build(ref, dev, args)
    batch = _batches(1, bs, model.vocab_size)[0]
    from clipa_b200.training import TrainStep
    m2, _, _ = _build(ref, dev, args)
    m2.load_state_dict(model.state_dict())
    ts_loss = TrainStep(m2, micro_batch=bs, lr=args.lr).step(*batch).item()
    images = batch[0].to(dev).float().div(255)
    import torchvision.transforms as T
    images = T.Normalize(mean=model.visual.image_mean, std=model.visual.image_std)(images)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(images, batch[1].to(dev))
        ref_loss = sum(loss(**out, output_dict=True).values()).item()
    assert abs(ref_loss - ts_loss) / ref_loss < 1e-3, (ref_loss, ts_loss)


def test_reference_zero_shot_run_over_the_alias(dev, reference_training):
    """training/zero_shot.py `run` (unmodified) with our model and a classifier from clipa_b200.zero_shot; the
    reference's accuracy numbers equal ours."""
    ref = reference_training
    from clipa_b200 import zero_shot as zs
    args = _args(dev, 1, 8)
    model, _, _ = _build(ref, dev, args)
    model.eval()
    g = torch.Generator().manual_seed(5)
    n_cls, n_prompts = 16, 3
    ids = []
    for _ in range(n_cls):
        t = torch.randint(1, model.vocab_size - 1, (n_prompts, 16), generator=g)
        t[:, -1] = model.vocab_size - 1
        ids.append(t)
    classifier = zs.zero_shot_classifier(model, class_token_ids=ids)
    assert classifier.shape == (model.visual.output_dim, n_cls)
    assert torch.allclose(classifier.float().norm(dim=0), torch.ones(n_cls, device=dev), atol=1e-2)
    loader = _Loader([(torch.randint(0, 256, (8, 3, 64, 64), generator=g, dtype=torch.uint8),
                       torch.randint(0, n_cls, (8,), generator=g)) for _ in range(3)], 8)
    ref_top1, ref_top5 = ref.zero_shot.run(model, classifier.float(), loader, args)
    from clipa_b200.training import TrainStep
    pre = TrainStep(model, micro_batch=8).preprocess
    top1, top5 = zs.run(model, classifier, loader, preprocess=pre)
    assert abs(top1 - ref_top1) <= 1 / 24 + 1e-9 and abs(top5 - ref_top5) <= 2 / 24 + 1e-9   # bf16 near-ties may flip one sample
