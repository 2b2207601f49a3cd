"""The "library Blackwell path to beat" (SURVEY 8d): the oracle restatement of the reference step evaluated
with torch's own CUDA kernels (cuBLASLt GEMMs, F.layer_norm, nn.GELU, SDPA, autograd) under bf16 autocast --
what the unmodified reference does on a B200.  A measurement tool, not part of the product path and not a
bench.py line.

  python tools/torch_path_baseline.py [--workload vitl14_i81_t16_gb32k] [--batch 1024] [--steps 5]

Prints one JSON line: pairs/s of forward + loss + backward at the given per-GPU batch (torch autograd keeps
every activation, so the reference's own per-GPU batch of 4096 does not fit without checkpointing; the
per-pair rate is flat in the batch once the GEMMs are large)."""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402  (workload table)
from clipa_b200.open_clip import get_model_config  # noqa: E402
from oracle import clip_oracle as O  # noqa: E402
from oracle.weights import make_inputs, make_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="vitl14_i81_t16_gb32k", choices=list(bench.WORKLOADS))
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--device", default="cuda")
    args = ap.parse_args()
    wl = bench.WORKLOADS[args.workload]
    dev = torch.device(args.device)
    O.USE_FUSED = True
    cfg = get_model_config(wl["model"])
    sd = {k: v.to(dev).requires_grad_(True) for k, v in
          make_state_dict(cfg, 0, image_size=wl["image"], pos_embed=wl["pos"]).items()}
    images, text = make_inputs(cfg, args.batch, 1, image_size=wl["image"])
    images, text = images.to(dev), text.to(dev)
    if dev.type == "cuda":
        torch.backends.cuda.matmul.allow_tf32 = True      # training/main.py:85-91
        torch.backends.cudnn.allow_tf32 = True

    def step():
        for v in sd.values():
            v.grad = None
        with torch.autocast(device_type=dev.type, dtype=torch.bfloat16):
            loss = O.train_step_loss(images, text, sd, cfg)
        loss.backward()
        return loss

    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step()
    sync()
    dt = (time.perf_counter() - t0) / args.steps
    out = {"path": "torch library kernels (oracle restatement, bf16 autocast)", "workload": args.workload,
           "batch": args.batch, "ms_per_step": dt * 1e3, "pairs_per_s": args.batch / dt, "loss": float(loss),
           "optimizer": "not included"}
    if dev.type == "cuda":
        out["peak_hbm_gb"] = round(torch.cuda.max_memory_allocated(dev) / 2**30, 1)
        out["mfu_vs_1466"] = out["pairs_per_s"] * wl["gflop_per_pair"] / 1466.2e3
    print(json.dumps(out))


if __name__ == "__main__":
    main()
