"""Multi-GPU check of the path's one exchange step (run under torchrun, NCCL):
  N ranks x (B/N pairs each)  ==  1 rank x B pairs
for the loss and for every parameter gradient after the data-parallel reduction.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/ddp_check.py
"""
import json
import os
import sys
import tempfile
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from clipa_b200 import open_clip  # noqa: E402
from clipa_b200.training import TrainStep  # noqa: E402
from oracle.weights import TINY_CONFIGS, make_inputs, make_state_dict  # noqa: E402


def build(dev):
    cfg = TINY_CONFIGS["tiny-cls"]
    tmp = Path(tempfile.mkdtemp())
    (tmp / "ddp-tiny.json").write_text(json.dumps(cfg))
    open_clip.add_model_config(tmp)
    model, _, _ = open_clip.create_model_and_transforms("ddp-tiny", precision="amp_bf16", device=dev, output_dict=True)
    model.load_state_dict(make_state_dict(cfg, 11), strict=True)
    model.train()
    return cfg, model


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg, model = build(dev)
    B = 16 * world
    images, text = make_inputs(cfg, B, 99)
    bl = B // world
    # --- N-rank step (micro-batch smaller than the local batch -> also exercises the GradCache path)
    for mb, overlap in ((bl, False), (bl // 2, False), (bl, True), (bl // 2, True)):
        # single-pass logit_scale gradient on both sides (the reference's accumulation loop counts it once per chunk)
        ts = TrainStep(model, rank=rank, world_size=world, micro_batch=mb, reference_accum_logit_scale=False,
                       overlap_grad_allreduce=overlap, allreduce_bucket_blocks=1)
        ts.zero_grad()
        loss = ts.forward_backward(ts.preprocess(images[rank * bl:(rank + 1) * bl]), text[rank * bl:(rank + 1) * bl].to(dev))
        ts._allreduce_grads()
        lsum = loss.clone()
        dist.all_reduce(lsum)
        grads = {n: p.grad.clone() for n, p in model.named_parameters()}
        # --- single-rank reference on the whole batch (same weights)
        _, ref_model = build(dev)
        ts1 = TrainStep(ref_model, rank=0, world_size=1, micro_batch=B)
        ts1.zero_grad()
        loss1 = ts1.forward_backward(ts1.preprocess(images), text.to(dev))
        worst = 0.0
        for n, p in ref_model.named_parameters():
            g1, gN = p.grad.float(), grads[n].float()
            err = ((g1 - gN).norm() / g1.norm().clamp_min(1e-12)).item()
            worst = max(worst, err)
        if rank == 0:
            print(f"[ddp_check] world={world} micro_batch={mb} overlap={overlap}: mean-over-ranks loss {lsum.item() / world:.6f} vs "
                  f"single-rank {loss1.item():.6f}; worst parameter-gradient rel err {worst:.3e}", flush=True)
        assert abs(lsum.item() / world - loss1.item()) / loss1.item() < 2e-3
        assert worst < 5e-2, worst   # bf16 activations; different micro-batching changes rounding only
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
