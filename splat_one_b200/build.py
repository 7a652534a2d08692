"""Build libb200splat.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m splat_one_b200.build [--force] [--verbose]

One object per translation unit (compiled in parallel, cached on mtime), linked into
splat_one_b200/libb200splat.so with a static cudart, so the library can be loaded with
ctypes next to any PyTorch build and travels to the GPU box with the source snapshot.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "_build"
LIB = PKG / "libb200splat.so"

NVCC_FLAGS = [
    "-O3",
    "--use_fast_math",          # same numerics class as the reference build (G/cuda/_backend.py:93-100)
    "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "--expt-relaxed-constexpr",
    "--expt-extended-lambda",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libb200splat.so cannot be built")
    return nvcc


def _deps_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "b200splat.h", Path(__file__)]
    return max(p.stat().st_mtime for p in hdrs)


def _flags():
    # B200SPLAT_TUNING=1: also compile the A/B kernel variants selected by B200SPLAT_TUNING_VARIANT
    return NVCC_FLAGS + (["-DB2S_TUNING"] if os.environ.get("B200SPLAT_TUNING", "0") == "1" else [])


def _compile(src: Path, force: bool, verbose: bool) -> Path:
    obj = OBJ / (src.stem + ".o")
    if not force and obj.exists() and obj.stat().st_mtime >= max(src.stat().st_mtime, _deps_mtime()):
        return obj
    cmd = [_nvcc(), *_flags(), "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    (OBJ / (src.stem + ".ptxas.log")).write_text(r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"[build] {src.name} ok")
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    if force or not LIB.exists() or LIB.stat().st_mtime < max(o.stat().st_mtime for o in objs):
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
               "-o", str(LIB), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[build] linked {LIB}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
