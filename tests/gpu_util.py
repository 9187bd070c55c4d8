"""Helpers shared by the -m gpu parity tests (test infrastructure)."""
import importlib.util
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL, ATOL = 1e-5, 1e-6  # BASELINE.json north_star: rel 1e-5 / abs 1e-6 vs the fp64 oracle


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def seg_arrays(lengths):
    lengths = np.asarray(lengths, dtype=np.int64)
    inv = np.repeat(np.arange(len(lengths), dtype=np.int32), lengths)
    seg_end = np.cumsum(lengths).astype(np.int32)
    return inv, seg_end


def run_fwd(op, x, key, variant=None, offset=(0, 0, 0)):
    """Run our op on the GPU.  `offset` = element offsets of (x, key, y) inside over-allocated buffers,
    to exercise unaligned base pointers (sliced tensors)."""
    import grouped_cumprod as gc
    from simplegaussiansplat_tk71_b200 import ops

    n = len(x)
    ox, ok, oy = offset
    xb = torch.zeros(n + ox, dtype=torch.float32, device="cuda")
    kb = torch.zeros(n + ok, dtype=torch.int32, device="cuda")
    yb = torch.full((n + oy,), float("nan"), dtype=torch.float32, device="cuda")
    xb[ox:] = dev(x, torch.float32)
    kb[ok:] = dev(key, torch.int32)
    if variant is not None:
        ops.set_variant("fwd", variant)
    try:
        fn = gc.grouped_cumprod_forward if op == "mul" else gc.grouped_cumsum_forward
        fn(xb[ox:], kb[ok:], yb[oy:])
        torch.cuda.synchronize()
    finally:
        ops.set_variant("fwd", -1)
    assert ops.workspace_status() == 0, "watchdog fired"
    assert bool(torch.isnan(yb[:oy]).all()), "the op wrote in front of its output"
    return yb[oy:].cpu().numpy()


def run_bwd(x, y, g, inv, seg_end, variant=None, offset=0):
    import grouped_cumprod as gc
    from simplegaussiansplat_tk71_b200 import ops

    n = len(x)

    def buf(a, dt):
        b = torch.zeros(n + offset, dtype=dt, device="cuda")
        b[offset:] = dev(a, dt)
        return b[offset:]

    xb, yb, gb = buf(x, torch.float32), buf(y, torch.float32), buf(g, torch.float32)
    ib = buf(inv, torch.int32)
    out_full = torch.full((n + offset,), float("nan"), dtype=torch.float32, device="cuda")
    out = out_full[offset:]
    if variant is not None:
        ops.set_variant("bwd", variant)
    try:
        gc.grouped_cumprod_backward(xb, yb, gb, ib, out, dev(seg_end, torch.int32))
        torch.cuda.synchronize()
    finally:
        ops.set_variant("bwd", -1)
    assert ops.workspace_status() == 0, "watchdog fired"
    assert bool(torch.isnan(out_full[:offset]).all()), "the op wrote in front of its output"
    return out.cpu().numpy()


def assert_close(got, ref, what, scale=None, rtol=RTOL, atol=ATOL):
    """|got-ref| <= atol + rtol*|ref|  (or rtol*scale when a condition-aware scale is given)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    if got.size == 0:
        return
    nf = np.flatnonzero(np.isfinite(got) != np.isfinite(ref))
    assert nf.size == 0, (f"{what}: non-finite mismatch at {nf.size} positions, first {nf[:8].tolist()} last "
                          f"{nf[-3:].tolist()} of {got.size}: got {got[nf[:4]].tolist()} ref {ref[nf[:4]].tolist()}")
    m = np.isfinite(ref)
    bound = atol + rtol * (np.abs(ref[m]) if scale is None else np.asarray(scale, dtype=np.float64)[m])
    err = np.abs(got[m] - ref[m])
    bad = err > bound
    if bad.any():
        i = int(np.argmax(err / bound))
        idx = np.flatnonzero(m)[i]
        raise AssertionError(
            f"{what}: {int(bad.sum())}/{got.size} outside tolerance; worst at {idx}: got {got[idx]!r} "
            f"ref {ref[idx]!r} err {err[i]:.3e} bound {bound[i]:.3e}")


_ref_mod = None


def reference_ops():
    """The reference's own CUDA ops built unchanged by oracle/build_ref.sh, or None if not built."""
    global _ref_mod
    if _ref_mod is None:
        path = os.path.join(ROOT, "oracle", "_ref", "grouped_cumprod_ref.so")
        if not os.path.exists(path):
            _ref_mod = False
        else:
            try:
                spec = importlib.util.spec_from_file_location("grouped_cumprod_ref", path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                _ref_mod = mod
            except Exception as e:  # noqa: BLE001
                print("reference ops failed to load:", e)
                _ref_mod = False
    return _ref_mod or None
