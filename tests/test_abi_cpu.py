"""CPU: the C-ABI library loads, exports every symbol include/clipa_b200.h declares, and rejects
bad arguments before touching a device (no compute calls here)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from clipa_b200 import _lib
    if not _lib.LIB_PATH.exists():
        from clipa_b200 import build
        build.build()
    return _lib.lib()


def declared_symbols():
    text = (ROOT / "include" / "clipa_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(clipa_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    from clipa_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), "ctypes binding and header disagree"
    assert lib.clipa_abi_version() == _lib.ABI_VERSION


def test_struct_layout_matches_header():
    """Field order of the ctypes mirror == declaration order in clipa_gemm_desc."""
    from clipa_b200._lib import GemmDesc
    text = (ROOT / "include" / "clipa_b200.h").read_text()
    body = re.search(r"typedef struct clipa_gemm_desc \{(.*?)\} clipa_gemm_desc;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        head, *rest = decl.split(",")
        fields.append(head.split()[-1].lstrip("*"))
        fields += [r.strip().lstrip("*") for r in rest]
    assert fields == [f[0] for f in GemmDesc._fields_]


def test_null_and_bad_arguments_are_rejected_with_messages(lib):
    assert lib.clipa_gemm(None, None) == -1
    assert b"null descriptor" in lib.clipa_last_error()
    from clipa_b200._lib import GemmDesc
    d = GemmDesc()
    d.M, d.N, d.K = 128, 128, 64
    d.lda = d.ldb = d.ldc = 36          # not a multiple of 8
    d.C = 16
    assert lib.clipa_gemm(C.byref(d), None) == -1
    assert b"multiples of 8" in lib.clipa_last_error()
    assert lib.clipa_layernorm_fwd(None, None, None, None, None, None, 4, 64, 1e-5, None) == -1
    assert lib.clipa_attention_fwd(None, None, None, 1, 1, 1, 64, 0, None) == -1
    assert lib.clipa_clip_lse(None, None, 8, 8, 64, 1.0, None, 0, None, None, None, None) == -1
    assert lib.clipa_attention_bwd(None, None, None, None, None, None, 0, 1, 1, 1, 64, 0, None) == -1
    assert lib.clipa_attention_bwd_workspace(2, 577, 4, 64) >= 2 * 4 * 577 * 64 * 4   # 7 tiles of 83 rows, fp32 dQ partials
    assert lib.clipa_attention_bwd_workspace(2, 257, 4, 80) >= 2 * 4 * 257 * 80 * 4
    assert lib.clipa_attention_bwd_workspace(2, 257, 4, 64) == 0                      # 3 dQ tiles stay in tensor memory
    assert lib.clipa_attention_bwd_workspace(2, 82, 4, 64) == 0
    assert lib.clipa_clip_lse_workspace(4096, 32768) > 0
    assert lib.clipa_launch_count() == 0   # nothing was launched by the rejected calls


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from clipa_b200 import _lib
    monkeypatch.setenv("CLIPA_B200_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(_lib.ClipaError):
        _lib.lib()


def test_cpu_tensors_are_refused():
    import torch
    from clipa_b200 import ops
    from clipa_b200._lib import ClipaError
    x = torch.zeros(64, 64, dtype=torch.bfloat16)
    with pytest.raises(ClipaError):
        ops.gemm(x, x, torch.zeros(64, 64, dtype=torch.bfloat16))
    with pytest.raises(ClipaError):
        ops.layernorm_fwd(x, torch.ones(64), torch.zeros(64))
    from clipa_b200.open_clip import ClipLoss
    with pytest.raises(RuntimeError):
        ClipLoss()(x.float(), x.float(), torch.tensor(10.0))


def test_product_code_never_imports_the_oracle():
    for f in (ROOT / "clipa_b200").rglob("*.py"):
        src = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
