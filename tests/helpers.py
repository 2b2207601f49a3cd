"""Shared helpers for the parity tests (oracle = test infrastructure, see oracle/clip_oracle.py)."""
import json
from pathlib import Path

import numpy as np
import torch

GOLD = Path(__file__).resolve().parent / "golden"


def load_golden(name, precision):
    z = np.load(GOLD / f"{name}_{precision}.npz", allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    return meta, {k: z[k] for k in z.files if k != "meta"}


def rel_err(got, ref):
    got = torch.as_tensor(np.asarray(got), dtype=torch.float64)
    ref = torch.as_tensor(np.asarray(ref), dtype=torch.float64)
    return ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item()


def max_rel_err(got, ref):
    got = torch.as_tensor(np.asarray(got), dtype=torch.float64)
    ref = torch.as_tensor(np.asarray(ref), dtype=torch.float64)
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()
