"""Builds libgcp_b200.so in-tree with nvcc for sm_100a (seconds; no torch headers involved)."""
from __future__ import annotations

import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.path.join(_PKG, "libgcp_b200.so")
SOURCES = ["gcp_abi.cu"]
HEADERS = ["gcp_device.cuh", "gcp_fwd.cuh", "gcp_bwd.cuh", "gcp_blk.cuh"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(_ROOT, "include", "gcp_abi.h")]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library if missing or older than its sources.  Needs nvcc, not a GPU."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        if os.path.exists(LIB_PATH):
            return LIB_PATH  # GPU box without toolkit on PATH: use the prebuilt library
        raise RuntimeError("nvcc not found and libgcp_b200.so is not built")
    cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(_ROOT, "include"), "-o", LIB_PATH,
           *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB_PATH


if __name__ == "__main__":
    print(build_lib(force=True, verbose=True))
