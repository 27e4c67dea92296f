"""Builds libpylbl_b200.so (sm_100a) in-tree with nvcc.

    python -m pylbl_b200.build            # build if sources are newer than the library
    python -m pylbl_b200.build --force

The library is a plain C-ABI shared object (include/pylbl_b200.h); it has no Python or
torch dependency.  nvcc cross-compiles for sm_100a without a GPU present.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libpylbl_b200.so"
SOURCES = ["lbl_api.cu", "lbl_db.cpp", "lbl_pack.cpp"]
HEADERS = ["lbl_core.cuh", "lbl_threads.cuh", "lbl_kernels.cuh", "lbl_continuum.cuh", "lbl_db.h", "lbl_cheb.h", "lbl_bands.h",
           "../../include/pylbl_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
    "-Xptxas", "-v",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any((CSRC / f).stat().st_mtime > t for f in SOURCES + HEADERS) or \
        Path(__file__).stat().st_mtime > t


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", str(LIB)] + [str(CSRC / s) for s in SOURCES] + ["-ldl"]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    (PKG / "build.log").write_text(" ".join(cmd) + "\n" + proc.stdout)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libpylbl_b200.so (see pylbl_b200/build.log)")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
