"""CLI for baseline/library_step.py: the unmodified reference on torch's CUDA kernels at a bench.py workload.

  python tools/library_baseline.py --workload vitl14_i81_t16_gb32k --batch 2048 [--no-ckpt]
Prints one JSON line (pairs/s of a full optimizer step at the given per-GPU batch)."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402  (workload table)
from baseline.library_step import library_baseline  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="vitl14_i81_t16_gb32k", choices=list(bench.WORKLOADS))
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--no-ckpt", action="store_true")
    a = ap.parse_args()
    wl = bench.WORKLOADS[a.workload]
    out = library_baseline(wl, a.batch, a.steps, a.warmup, grad_checkpointing=not a.no_ckpt)
    out["workload"] = a.workload
    pk, _ = bench.peaks()
    out["mfu_vs_sustained_peak"] = out["pairs_per_s"] * wl["gflop_per_pair"] / (pk["bf16_tflops_sustained"] * 1e3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
