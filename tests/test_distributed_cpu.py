"""CPU, world_size 2 over gloo: the host-side multi-rank logic of the path -- the differentiable
feature all-gather (forward all-gather, backward reduce-scatter SUM) and the flat-buffer gradient
all-reduce of TrainStep -- checked against the reference-produced 2-rank golden."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import GOLD


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from clipa_b200.open_clip.loss import gather_features
        from oracle import clip_oracle as O
        z = np.load(GOLD / "cliploss_world2_gloo.npz")
        img_all, txt_all = torch.tensor(z["img"]), torch.tensor(z["txt"])
        bl = img_all.shape[0] // world
        img = img_all[rank * bl:(rank + 1) * bl].clone().requires_grad_(True)
        txt = txt_all[rank * bl:(rank + 1) * bl].clone().requires_grad_(True)
        scale = torch.tensor(z["scale"]).requires_grad_(True)
        # host path under test: differentiable gather; the loss math itself is the oracle's here
        # (the CUDA loss kernels are covered by the -m gpu tests)
        gi, gt = gather_features(img, txt, local_loss=True, gather_with_grad=True, rank=rank, world_size=world)
        assert torch.equal(gi.detach(), img_all) and torch.equal(gt.detach(), txt_all)
        loss = O.clip_loss(img, txt, gi, gt, scale, rank)
        loss.backward()
        res = {"loss": loss.item(), "d_img": img.grad.numpy(), "d_txt": txt.grad.numpy(), "d_scale": scale.grad.item()}
        ok = (abs(res["loss"] - float(z[f"loss_r{rank}"])) < 1e-5
              and np.abs(res["d_img"] - z[f"d_img_r{rank}"]).max() < 1e-6
              and np.abs(res["d_txt"] - z[f"d_txt_r{rank}"]).max() < 1e-6
              and abs(res["d_scale"] - float(z[f"d_scale_r{rank}"])) < 1e-5)
        # without gather_with_grad only the local slice of d(gathered) comes back
        img2 = img.detach().clone().requires_grad_(True)
        gi2, _ = gather_features(img2, txt.detach(), gather_with_grad=False, rank=rank, world_size=world)
        (gi2 * torch.arange(gi2.shape[0]).float()[:, None]).sum().backward()
        ok = ok and torch.allclose(img2.grad, torch.arange(rank * bl, (rank + 1) * bl).float()[:, None].expand(bl, gi2.shape[1]))
        # flat-buffer gradient averaging of TrainStep
        from clipa_b200.training import TrainStep
        lin = torch.nn.Linear(4, 3)
        with torch.no_grad():
            for p in lin.parameters():
                p.fill_(1.0)
        ts = TrainStep.__new__(TrainStep)
        ts.world_size, ts.params = world, list(lin.parameters())
        flat = torch.zeros(sum(p.numel() for p in ts.params))
        off = 0
        for p in ts.params:
            p.grad = flat[off:off + p.numel()].view_as(p); off += p.numel()
        ts._flat = {torch.float32: flat}
        flat.fill_(float(rank + 1))
        ts._allreduce_grads()
        ok = ok and torch.allclose(lin.weight.grad, torch.full((3, 4), 1.5)) and torch.allclose(lin.bias.grad, torch.full((3,), 1.5))
        # bucketed (overlapped) all-reduce bookkeeping of TrainStep: ranges reduced by the block hooks + the complement
        # reduced after backward cover every element exactly once
        from clipa_b200.open_clip.transformer import Transformer
        net = torch.nn.Module()
        net.visual = torch.nn.Linear(8, 8)              # not in any bucket: reduced by the complement pass
        net.transformer = Transformer(width=64, layers=5, heads=1)
        net.logit_scale = torch.nn.Parameter(torch.tensor(1.0))
        ts2 = TrainStep(net, rank=rank, world_size=world, micro_batch=4, image_mean=(0., 0., 0.), image_std=(1., 1., 1.),
                        allreduce_bucket_blocks=2, overlap_grad_allreduce=True)
        assert ts2.fused and ts2.overlap
        ts2._setup_overlap()
        assert sorted(b0 for (_, b0) in ts2._buckets) == [0, 2, 4]
        for grp in ts2._groups:
            grp["g"].fill_(float(rank + 1))
        ts2._armed = True
        for b0 in (4, 2):                                # backward order: last blocks first; bucket 0 never fires here
            ts2._on_blocks_ready(net.transformer, b0)
        covered = sum(hi - lo for _, lo, hi in ts2._done_ranges)
        assert 0 < covered < sum(g["g"].numel() for g in ts2._groups)
        ts2._allreduce_grads(average=True)
        ok = ok and all(torch.allclose(g["g"], torch.full_like(g["g"], 1.5)) for g in ts2._groups) and not ts2._armed
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


def test_world2_gather_and_grad_sync(tmp_path):
    mp.spawn(_worker, args=(2, 29617, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert (tmp_path / f"ok{r}").read_text() == "1"
