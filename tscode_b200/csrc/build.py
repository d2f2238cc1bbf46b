"""Build libtscode_b200.so in-tree with nvcc for sm_100a (no torch glue: plain C-ABI)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
SOURCES = ["capi.cu", "pack.cu", "rmsd_sim.cu", "rmsd_screen.cu", "rmsd_verify.cu", "eliminate.cu", "clash.cu", "rotcorr.cu", "tfd_moi.cu"]
HEADERS = ["tsc_common.cuh", "tsc_math.cuh", "screen_common.cuh"]
LIB = os.path.join(PKG, "libtscode_b200.so")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS + ["build.py"])


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    flags = [f for f in FLAGS if not f.startswith("--use_fast_math")]
    cmd = [nvcc()] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB] + \
          [os.path.join(HERE, s) for s in SOURCES] + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        print(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
