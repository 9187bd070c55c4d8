"""Runs one op a few times on the C3 workload for a given kernel variant (ncu / timing driver)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import grouped_cumprod as gc  # noqa: E402
from simplegaussiansplat_tk71_b200 import ops, workloads as wl  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--op", default="fwd")
ap.add_argument("--variant", type=int, default=-1)
ap.add_argument("--halo", type=int, default=1)
ap.add_argument("--workload", default="c3")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--chain", type=int, default=1, help="GCP_OPT_CHAIN of the blocked backward: 0 tickets, 1 chained, 2 auto")
a = ap.parse_args()
e = wl.c4("cuda", scale=a.scale) if a.workload == "c4" else wl.c3("cuda", scale=a.scale)
y = torch.empty_like(e.x)
gin = torch.empty_like(e.x)
gc.grouped_cumprod_forward(e.x, e.key, y)
ops.set_option(0, a.halo)
ops.set_option(1, a.chain)
ops.set_variant(a.op, a.variant)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(a.reps):
    if i == a.reps - 1:
        ev0.record()
    if a.op == "fwd":
        gc.grouped_cumprod_forward(e.x, e.key, y)
    else:
        gc.grouped_cumprod_backward(e.x, y, e.grad_out, e.inv, gin, e.seg_end)
ev1.record()
torch.cuda.synchronize()
print(f"{a.op} v{a.variant} halo={a.halo} n={e.n}: {ev0.elapsed_time(ev1):.4f} ms, status {ops.workspace_status()}")
