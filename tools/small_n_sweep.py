"""Where the persistent blocked kernels (variant 1: one cooperative launch, TMA ring, grid barrier) and the plain-load
kernel pair (variant 0: K1 + fix-up launch) cross over as the element list shrinks: per-op device time of both for
n = 16 Ki ... 64 Mi (C1's segment-length law), with the L2 flushed before every op and with a warm L2."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import grouped_cumprod as gc  # noqa: E402
from simplegaussiansplat_tk71_b200 import ops, workloads as wl  # noqa: E402


def median_ms(fn, reps=15, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
sizes = [int(s) for s in sys.argv[1:]] or [1 << 14, 1 << 16, 1 << 18, 1 << 20, 1 << 22, 1 << 24, 1 << 26]
print(f"{'n':>10} | {'flushed L2: fwd v0 / v1':>24} | {'bwd v0 / v1':>17} | {'warm L2: fwd v0 / v1':>22} | {'bwd v0 / v1':>17}  (us)")
for n in sizes:
    e = wl.build(f"n{n}", wl.lengths_c1(n, max(1, n // 16)), 1 << 14, 1, 0, "cuda")
    y = torch.empty_like(e.x)
    gin = torch.empty_like(e.x)
    gc.grouped_cumprod_forward(e.x, e.key, y)
    row = []
    for fl in (flush, None):
        for op in ("fwd", "bwd"):
            for v in (0, 1):
                ops.set_variant(op, v)
                if op == "fwd":
                    t = median_ms(lambda: gc.grouped_cumprod_forward(e.x, e.key, y), flush=fl)
                else:
                    t = median_ms(lambda: gc.grouped_cumprod_backward(e.x, y, e.grad_out, e.inv, gin, e.seg_end), flush=fl)
                row.append(1e3 * t)
                ops.set_variant(op, -1)
    assert ops.workspace_status() == 0
    print(f"{n:>10} | {row[0]:>11.1f} / {row[1]:<10.1f} | {row[2]:>7.1f} / {row[3]:<7.1f} | {row[4]:>10.1f} / {row[5]:<9.1f} | "
          f"{row[6]:>7.1f} / {row[7]:<7.1f}", flush=True)
