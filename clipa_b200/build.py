"""In-tree build of the C-ABI library (nvcc -> clipa_b200/lib/libclipa_b200.so) for sm_100a.

No torch involvement: the library is plain CUDA C++ with an `extern "C"` surface
(include/clipa_b200.h).  nvcc cross-compiles without a GPU, so this runs in the authoring
container; the built .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
OBJDIR = PKG / "lib" / "obj"
LIB = LIBDIR / "libclipa_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
]


def _extra_flags() -> list:
    """Extra nvcc flags for experiment builds, e.g. CLIPA_B200_NVCC_FLAGS="-DCLIPA_UNIFORM_ISSUE=1"
    together with CLIPA_B200_LIB_NAME=libclipa_b200_uni.so (loaded through the CLIPA_B200_LIB variable)."""
    return os.environ.get("CLIPA_B200_NVCC_FLAGS", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _digest(src: Path) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS + _extra_flags()).encode())
    for f in sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [src, PKG.parent / "include" / "clipa_b200.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every csrc/*.cu and link the shared library. Incremental via content hashes.
    Experiment builds (CLIPA_B200_LIB_NAME set) get their own object directory and library file."""
    nvcc = _nvcc()
    name = os.environ.get("CLIPA_B200_LIB_NAME", "")
    OBJDIR = PKG / "lib" / ("obj_" + Path(name).stem if name else "obj")
    LIB = LIBDIR / name if name else LIBDIR / "libclipa_b200.so"
    OBJDIR.mkdir(parents=True, exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    jobs = []
    for src in sources:
        obj = OBJDIR / (src.stem + ".o")
        stamp = OBJDIR / (src.stem + ".sha")
        dig = _digest(src)
        if not force and obj.exists() and stamp.exists() and stamp.read_text() == dig:
            continue
        jobs.append((src, obj, stamp, dig))

    def compile_one(job):
        src, obj, stamp, dig = job
        cmd = [nvcc, *NVCC_FLAGS, *_extra_flags(), "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, file=sys.stderr)
        stamp.write_text(dig)
        return src.name

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for name in ex.map(compile_one, jobs):
                if verbose:
                    print(f"[clipa_b200.build] compiled {name}", file=sys.stderr)
    objs = [str(OBJDIR / (s.stem + ".o")) for s in sources]
    if jobs or not LIB.exists():
        cmd = [nvcc, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-cudart=static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
