"""Builds libcattus_b200.so (the C-ABI library of include/cattus_b200.h) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this also runs in the build container (`__graft_entry__.build()`).
The built .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libcattus_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread",
    # host arithmetic of the MCTS driver must stay un-fused so its f32 selection scores match the reference's
    "-Xcompiler", "-ffp-contract=off",
    # lets g++ vectorise the selection loop (a masked IEEE division); results are unchanged, only FP exception flags differ
    "-Xcompiler", "-fno-trapping-math",
    # the self-play driver keeps typed headers and rows in one pool of 32-bit words
    "-Xcompiler", "-fno-strict-aliasing",
]


def sources():
    return (sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.hpp")) + sorted(CSRC.glob("*.cpp"))
            + sorted((ROOT / "include").glob("*.h")))


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(s.stat().st_mtime > t for s in sources())


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: cannot build libcattus_b200.so (there is no CPU fallback)")


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", str(LIB), str(CSRC / "engine.cu"), str(CSRC / "selfplay.cpp"), str(CSRC / "chess_api.cpp")]
    for knob in ("DS_SELECT_MIN_BLOCKS", "DS_EXPAND_MIN_BLOCKS"):  # occupancy experiments on the device-search kernels
        if os.environ.get(knob):
            cmd.insert(1, f"-D{knob}={int(os.environ[knob])}")
    if os.environ.get("CATTUS_B200_TRUNK_TRACE"):  # diagnostic build with clock64 trace points (tools/trace_trunk.py)
        cmd.insert(1, "-DCB2_TRUNK_TRACE")
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
