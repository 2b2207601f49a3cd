"""Imports the UNMODIFIED reference (`open_clip` + `training` of UCSC-VLAA/CLIPA clipa_torch) for
measurement and drop-in tests.  Never imported by the product path (clipa_b200/).

Where it comes from: `baseline/_ref/` -- `python -m pip install --no-index --no-build-isolation --no-deps
--target baseline/_ref <copy of /root/reference/clipa_torch>` (tools/install_reference.sh; git-ignored,
travels to the GPU box) -- or, in the authoring container only, /root/reference/clipa_torch itself.

The reference's tokenizer / data modules import ftfy, tensorflow(_text), webdataset, braceexpand
unconditionally (open_clip/tokenizer.py:11-18, training/data.py:9,17-22); none of them is on the
hot path and none is installed, so they are stubbed.  The wheel does not carry the JSON model
configs (setup.py has no package_data): they are registered through the reference's own
`add_model_config` from clipa_b200's size table, which tests pin against the reference's JSON files.
"""
from __future__ import annotations

import importlib.machinery
import json
import sys
import tempfile
from pathlib import Path
from unittest.mock import MagicMock

ROOT = Path(__file__).resolve().parent.parent
CANDIDATES = (ROOT / "baseline" / "_ref", Path("/root/reference/clipa_torch"))
_STUBS = ("ftfy", "tensorflow", "tensorflow_text", "webdataset", "webdataset.filters",
          "webdataset.tariterators", "braceexpand", "fsspec", "timm", "horovod", "horovod.torch")


def reference_root():
    for c in CANDIDATES:
        if (c / "open_clip" / "factory.py").exists():
            return c
    return None


def available() -> bool:
    return reference_root() is not None


def import_reference(register_configs: bool = True):
    """Returns the reference's `open_clip` module (its `training` package becomes importable too)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not installed: run tools/install_reference.sh in the authoring container")
    for n in _STUBS:
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                m = MagicMock()
                m.__spec__ = importlib.machinery.ModuleSpec(n, None)
                m.__path__ = []
                sys.modules[n] = m
    if str(root) not in sys.path:
        sys.path.insert(0, str(root))
    import open_clip  # noqa: the reference's, top-level
    assert Path(open_clip.__file__).resolve().is_relative_to(root.resolve()), open_clip.__file__
    if register_configs:
        from clipa_b200.open_clip.model_configs import _MODEL_CONFIGS
        tmp = Path(tempfile.mkdtemp(prefix="clipa_ref_cfg_"))
        for name, cfg in _MODEL_CONFIGS.items():
            if open_clip.get_model_config(name) is None:
                (tmp / f"{name}.json").write_text(json.dumps(cfg))
        open_clip.add_model_config(tmp)
    return open_clip
