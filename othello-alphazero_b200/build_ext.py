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
    "oz_wgrad.cu": [],
}


BINDING_SRC = CSRC / "othello_mcts_pybind.cpp"
BINDING_PKG = HERE / "cpp_binding" / "othello_mcts"


def _deps() -> float:
    return max(p.stat().st_mtime for p in [q for q in CSRC.glob("*") if q != BINDING_SRC] + [HERE.parent / "include" / "othello_b200.h"])


def binding_path() -> Path:
    import sysconfig
    return BINDING_PKG / f"_othello_mcts_impl{sysconfig.get_config_var('EXT_SUFFIX')}"


def build_binding(force: bool = False) -> Path:
    """The reference's pybind11 binding over the C ABI (csrc/othello_mcts_pybind.cpp): g++ against this image's
    pybind11 + LibTorch headers, linked to libothello_b200.so (rpath = the ctypes package's directory)."""
    import sysconfig
    out = binding_path()
    hdr = HERE.parent / "include" / "othello_b200.h"
    if out.exists() and not force and out.stat().st_mtime >= max(BINDING_SRC.stat().st_mtime, hdr.stat().st_mtime):
        return out
    import pybind11
    import torch
    from torch.utils import cpp_extension
    inc = [f"-I{HERE.parent / 'include'}", f"-I{pybind11.get_include()}", f"-I{sysconfig.get_paths()['include']}"]
    inc += [f"-I{p}" for p in cpp_extension.include_paths()]
    torch_lib = Path(torch.__file__).parent / "lib"
    cmd = (["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-fvisibility=hidden",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", "-DTORCH_EXTENSION_NAME=_othello_mcts_impl"] + inc +
           [str(BINDING_SRC), "-o", str(out), f"-L{torch_lib}", f"-Wl,-rpath,{torch_lib}", f"-L{OUT.parent}",
            "-Wl,-rpath,$ORIGIN/../../othello_mcts", "-lothello_b200", "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_python"])
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"g++ failed for the pybind11 binding:\n{proc.stdout}\n{proc.stderr}")
    return out


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
    lib = build_library(force, verbose)
    build_binding(force)
    return lib


def build_library(force: bool = False, verbose: bool = False) -> Path:
    srcs = {n: f for n, f in SOURCES.items() if (CSRC / n).exists()}
    if OUT.exists() and not force and OUT.stat().st_mtime >= _deps():
        return OUT
    OBJ.mkdir(exist_ok=True)
    with ThreadPoolExecutor(max_workers=4) as pool:
        objs = list(pool.map(lambda kv: _compile(kv[0], kv[1], verbose), srcs.items()))
    cmd = [NVCC, *ARCH, "-shared", "-o", str(OUT), *map(str, objs), "-lcudart"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"link failed:\n{proc.stdout}\n{proc.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
