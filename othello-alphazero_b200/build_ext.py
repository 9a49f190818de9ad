"""Build libothello_b200.so (CUDA kernels + C ABI) for sm_100a, in-tree.

nvcc cross-compiles without a GPU.  The library lands next to the Python package
(othello_mcts/libothello_b200.so) so it travels with the repo snapshot and is the file the
GPU tests / bench load.  Usage: ``python build_ext.py [--force] [--verbose]``.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "othello_mcts" / "libothello_b200.so"
OBJ = HERE / "build"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-strict-aliasing"]

# Per-file extra flags.  The tree kernels reproduce the reference's float decisions bit for
# bit, so FMA contraction is off there (the reference build has none: SURVEY.md section 7).
SOURCES = {
    "oz_rules_api.cu": [],
    "oz_tree.cu": ["-fmad=false"],
    "oz_resnet.cu": [],
    "oz_engine_api.cu": ["-fmad=false"],
    "oz_train.cu": [],
}


def _deps() -> float:
    return max(p.stat().st_mtime for p in list(CSRC.glob("*")) + [HERE.parent / "include" / "othello_b200.h"])


def _compile(name: str, flags: list[str], verbose: bool) -> Path:
    obj = OBJ / (name + ".o")
    cmd = [NVCC, *ARCH, *COMMON, *flags, *os.environ.get("OZ_NVCC_FLAGS", "").split(), "-c", str(CSRC / name), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = {n: f for n, f in SOURCES.items() if (CSRC / n).exists()}
    if OUT.exists() and not force and OUT.stat().st_mtime >= _deps():
        return OUT
    OBJ.mkdir(exist_ok=True)
    with ThreadPoolExecutor(max_workers=4) as pool:
        objs = list(pool.map(lambda kv: _compile(kv[0], kv[1], verbose), srcs.items()))
    cmd = [NVCC, *ARCH, "-shared", "-o", str(OUT), *map(str, objs), "-lcudart", "-ldl"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"link failed:\n{proc.stdout}\n{proc.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
