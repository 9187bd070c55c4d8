"""A/B of GCP_OPT_CHAIN (blocked backward: contiguous tile range per CTA, carries chained in registers) on C3, C4
and a single giant segment: time per launch and agreement of the two modes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import grouped_cumprod as gc  # noqa: E402
from simplegaussiansplat_tk71_b200 import ops, workloads as wl  # noqa: E402


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


cases = [("c3", lambda: wl.c3("cuda")), ("c4", lambda: wl.c4("cuda"))]
for name, mk in cases:
    e = mk()
    y = torch.empty_like(e.x)
    gc.grouped_cumprod_forward(e.x, e.key, y)
    outs = []
    for chain in (0, 1, 2, 2):
        ops.set_option(1, chain)
        gin = torch.full_like(e.x, float("nan"))
        def fb():
            gc.grouped_cumprod_forward(e.x, e.key, y)
            gc.grouped_cumprod_backward(e.x, y, e.grad_out, e.inv, gin, e.seg_end)
        fwd_ms = timeit(lambda: gc.grouped_cumprod_forward(e.x, e.key, y))
        ms = timeit(fb) - fwd_ms      # alternating, as in a training step (the hint comes from the forward)
        st = ops.workspace_status()
        ab = wl.algorithmic_bytes(e.n, e.k)["bwd"]
        abf = wl.algorithmic_bytes(e.n, e.k)["fwd"]
        print(f"{name} chain={chain}: bwd {ms:.4f} ms  {ab / ms / 1e6:7.1f} GB/s  frac {ab / ms / 1e6 / 6553.9:.3f} | "
              f"fwd {fwd_ms:.4f} ms  frac {abf / fwd_ms / 1e6 / 6553.9:.3f}  status {st}", flush=True)
        outs.append(gin)
    ys = []
    for chain in (0, 1):
        ops.set_option(2, chain)
        yy = torch.full_like(e.x, float("nan"))
        gc.grouped_cumprod_forward(e.x, e.key, yy)
        ys.append(yy)
    print(f"   forward: max |diff| {float((ys[0] - ys[1]).abs().max()):.3e}, finite {bool(torch.isfinite(ys[1]).all())}")
    d = (outs[0] - outs[1]).abs()
    ref = outs[0].abs()
    print(f"   max |diff| {float(d.max()):.3e}, max rel (|ref|>1e-6) {float((d / ref.clamp_min(1e-6)).max()):.3e}, "
          f"finite {bool(torch.isfinite(outs[1]).all())}", flush=True)
    del e, y, outs
ops.set_option(1, 1)
ops.set_option(2, 0)
