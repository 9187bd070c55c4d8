"""TEST INFRASTRUCTURE — not product code.

numpy/ctypes front end of the C oracle (oracle/gcp_oracle.c).  Importable only from
tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference``
legs.  The shipped package never imports this module.

Every function names the reference lines it restates; see the header of
gcp_oracle.c for how the oracle is pinned (KAT1 cuda_test.py:19-34, KAT2
uitility.py:383-393, and oracle/_ref on the GPU box).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgcp_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle with the distro gcc (seconds)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "gcp_oracle.c"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libgcp_oracle.so"])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        fp, ip, dp, lp = (ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32),
                          ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64))
        i64 = ctypes.c_int64
        L.gcp_oracle_cumprod_fwd_f64.argtypes = [fp, ip, dp, i64]
        L.gcp_oracle_cumprod_fwd_f32.argtypes = [fp, ip, fp, i64]
        L.gcp_oracle_cumsum_fwd_f64.argtypes = [fp, ip, dp, i64]
        L.gcp_oracle_cumsum_fwd_f32.argtypes = [fp, ip, fp, i64]
        L.gcp_oracle_cumprod_bwd_refloop_f32.argtypes = [fp, fp, fp, ip, ip, fp, i64]
        L.gcp_oracle_cumprod_bwd_ref_f64.argtypes = [fp, fp, fp, ip, ip, dp, i64]
        L.gcp_oracle_cumprod_bwd_exact_f64.argtypes = [fp, fp, ip, dp, i64]
        L.gcp_oracle_segment_starts.argtypes = [ip, i64, lp]
        L.gcp_oracle_segment_starts.restype = i64
        L.gcp_oracle_max_threads.restype = ctypes.c_int
        L.gcp_oracle_fwd_bwd_f32_omp.argtypes = [fp, fp, lp, i64, fp, fp, ctypes.c_int]
        for f in ("cumprod_fwd_f64", "cumprod_fwd_f32", "cumsum_fwd_f64", "cumsum_fwd_f32",
                  "cumprod_bwd_refloop_f32", "cumprod_bwd_ref_f64", "cumprod_bwd_exact_f64",
                  "fwd_bwd_f32_omp"):
            getattr(L, "gcp_oracle_" + f).restype = None
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def cumprod_fwd(x, key, dtype=np.float64):
    """grouped_cumprod_forward (cuda_kernel/grouped_cumprod_forward.cu:17-23)."""
    x, key = _f32(x), _i32(key)
    y = np.empty(x.shape[0], dtype=dtype)
    if dtype == np.float64:
        lib().gcp_oracle_cumprod_fwd_f64(_p(x, ctypes.c_float), _p(key, ctypes.c_int32),
                                         _p(y, ctypes.c_double), x.shape[0])
    else:
        lib().gcp_oracle_cumprod_fwd_f32(_p(x, ctypes.c_float), _p(key, ctypes.c_int32),
                                         _p(y, ctypes.c_float), x.shape[0])
    return y


def cumsum_fwd(x, key, dtype=np.float64):
    """grouped_cumsum_forward (cuda_kernel/grouped_cumsum_forward.cu:17-23)."""
    x, key = _f32(x), _i32(key)
    y = np.empty(x.shape[0], dtype=dtype)
    if dtype == np.float64:
        lib().gcp_oracle_cumsum_fwd_f64(_p(x, ctypes.c_float), _p(key, ctypes.c_int32),
                                        _p(y, ctypes.c_double), x.shape[0])
    else:
        lib().gcp_oracle_cumsum_fwd_f32(_p(x, ctypes.c_float), _p(key, ctypes.c_int32),
                                        _p(y, ctypes.c_float), x.shape[0])
    return y


def cumprod_bwd_refloop_f32(x, y, g, inv, seg_end):
    """The reference kernel's loop, literally (grouped_cumprod_backward.cu:18-29); O(sum L^2)."""
    x, y, g, inv, seg_end = _f32(x), _f32(y), _f32(g), _i32(inv), _i32(seg_end)
    out = np.empty_like(x)
    lib().gcp_oracle_cumprod_bwd_refloop_f32(
        _p(x, ctypes.c_float), _p(y, ctypes.c_float), _p(g, ctypes.c_float),
        _p(inv, ctypes.c_int32), _p(seg_end, ctypes.c_int32), _p(out, ctypes.c_float), x.shape[0])
    return out


def cumprod_bwd_ref(x, y, g, inv, seg_end):
    """Reference formula (grouped_cumprod_backward.cu:18-29) in fp64, O(n)."""
    x, y, g, inv, seg_end = _f32(x), _f32(y), _f32(g), _i32(inv), _i32(seg_end)
    out = np.empty(x.shape[0], dtype=np.float64)
    lib().gcp_oracle_cumprod_bwd_ref_f64(
        _p(x, ctypes.c_float), _p(y, ctypes.c_float), _p(g, ctypes.c_float),
        _p(inv, ctypes.c_int32), _p(seg_end, ctypes.c_int32), _p(out, ctypes.c_double), x.shape[0])
    return out


def cumprod_bwd_exact(x, g, inv):
    """Exact division-free gradient E_i*S_i in fp64 (equals the reference where no x==0)."""
    x, g, inv = _f32(x), _f32(g), _i32(inv)
    out = np.empty(x.shape[0], dtype=np.float64)
    lib().gcp_oracle_cumprod_bwd_exact_f64(_p(x, ctypes.c_float), _p(g, ctypes.c_float),
                                           _p(inv, ctypes.c_int32), _p(out, ctypes.c_double),
                                           x.shape[0])
    return out


def segment_starts(key):
    key = _i32(key)
    starts = np.empty(key.shape[0] + 1, dtype=np.int64)
    k = lib().gcp_oracle_segment_starts(_p(key, ctypes.c_int32), key.shape[0],
                                        _p(starts, ctypes.c_int64))
    return starts[: k + 1].copy()


def max_threads() -> int:
    return int(lib().gcp_oracle_max_threads())


def fwd_bwd_f32_omp(x, g, starts, nthreads=0, y=None, gin=None):
    """CPU baseline ("port"): fp32 fwd + division-free bwd, OpenMP over segments."""
    x, g = _f32(x), _f32(g)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    if y is None:
        y = np.empty_like(x)
    if gin is None:
        gin = np.empty_like(x)
    lib().gcp_oracle_fwd_bwd_f32_omp(_p(x, ctypes.c_float), _p(g, ctypes.c_float),
                                     _p(starts, ctypes.c_int64), starts.shape[0] - 1,
                                     _p(y, ctypes.c_float), _p(gin, ctypes.c_float), int(nthreads))
    return y, gin


# --------------------------------------------------------------------------
# Pure-Python loops (tiny inputs only): an independent second statement of the
# same semantics, used by tests/test_oracle.py to cross-check the C code.
# --------------------------------------------------------------------------
def py_scan(x, key, op):
    out, acc = [], None
    for i, (v, k) in enumerate(zip(x, key)):
        if i == 0 or key[i - 1] != k:
            acc = float(v)
        else:
            acc = acc * float(v) if op == "mul" else acc + float(v)
        out.append(acc)
    return np.array(out, dtype=np.float64)


def py_bwd_bruteforce(x, g, inv):
    """dL/dx_i by the product rule, term by term (no recurrences)."""
    n = len(x)
    out = np.zeros(n, dtype=np.float64)
    for i in range(n):
        b = i
        while b > 0 and inv[b - 1] == inv[i]:
            b -= 1
        e = i
        while e + 1 < n and inv[e + 1] == inv[i]:
            e += 1
        tot = 0.0
        for k in range(i, e + 1):
            p = 1.0
            for j in range(b, k + 1):
                if j != i:
                    p *= float(x[j])
            tot += float(g[k]) * p
        out[i] = tot
    return out


def allclose(a, b, rtol=1e-5, atol=1e-6):
    """The north-star criterion: |a-b| <= atol + rtol*|b| (b = fp64 oracle)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return bool(np.all(np.abs(a - b) <= atol + rtol * np.abs(b)))


def max_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0, 0.0
    d = np.abs(a - b)
    return float(d.max()), float((d / (1e-6 / 1e-5 + np.abs(b))).max())
