"""Builds libgcp_b200.so in-tree with nvcc for sm_100a (no torch headers involved).

Each .cu is compiled to an object only when it (or a header) changed; the objects are then linked."""
from __future__ import annotations

import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
CSRC = os.path.join(_PKG, "csrc")
OBJ = os.path.join(_PKG, "build")
LIB_PATH = os.path.join(_PKG, "libgcp_b200.so")
# translation unit -> headers it depends on
SOURCES = {
    "gcp_abi.cu": ["gcp_device.cuh", "gcp_fwd.cuh", "gcp_bwd.cuh", "gcp_blk.cuh"],
    "gcp_splat.cu": [],
    "gcp_tile.cu": [],
    "gcp_host.cu": [],
}
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fopenmp"]   # OpenMP: the host-side key packing of gcp_host.cu


def _mtime(p):
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def _obj_stale(src, hdrs, obj):
    t = _mtime(obj)
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in hdrs] + \
           [os.path.join(_ROOT, "include", "gcp_abi.h")]
    return t == 0.0 or any(_mtime(d) > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library if missing or older than its sources.  Needs nvcc, not a GPU."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        if os.path.exists(LIB_PATH):
            return LIB_PATH  # GPU box without the toolkit on PATH: use the prebuilt library
        raise RuntimeError("nvcc not found and libgcp_b200.so is not built")
    os.makedirs(OBJ, exist_ok=True)
    relink = force or not os.path.exists(LIB_PATH)
    objs, procs = [], []
    for src, hdrs in SOURCES.items():
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _obj_stale(src, hdrs, obj):
            cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(_ROOT, "include"), "-c", os.path.join(CSRC, src), "-o", obj]
            if verbose:
                cmd[1:1] = ["-Xptxas", "-v"]
            procs.append((src, subprocess.Popen(cmd, cwd=CSRC)))
            relink = True
    for src, p in procs:
        if p.wait() != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    if relink or any(_mtime(o) > _mtime(LIB_PATH) for o in objs):
        subprocess.check_call([nvcc, "-shared", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fopenmp", "-o", LIB_PATH,
                               *objs])
    return LIB_PATH


if __name__ == "__main__":
    print(build_lib(force=True, verbose=True))
