"""Times one splat step (compositor forward + backward of one view) on the splat-route workload."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simplegaussiansplat_tk71_b200 import workloads as wl  # noqa: E402
from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--c2", type=int, default=-1, help="use view K of the bundled-scene workload (C2) instead")
ap.add_argument("--route", default=None, help="compositor route: tiles | lists (default: compositor.ROUTE)")
a = ap.parse_args()
if a.route:
    from simplegaussiansplat_tk71_b200 import compositor  # noqa: E402
    compositor.ROUTE = a.route
v = wl.bundled_views("cuda")[a.c2] if a.c2 >= 0 else wl.splat_view(a.width, a.height, a.n, device="cuda")
print(v.name, "elements", v.elements)
mean = v.mean.float().requires_grad_(True)
lam = v.lam.clone().requires_grad_(True)
opac = v.opacity.clone().requires_grad_(True)
l_d = v.l_d.clone().requires_grad_(True)
gI = torch.rand(v.height + 1, v.width + 1, 3, device="cuda") + 0.1
W, H = torch.tensor(v.width), torch.tensor(v.height)
batch = torch.tensor([v.n])


def step():
    for t in (mean, lam, opac, l_d):
        t.grad = None
    img = F.apply(v.boxsize, batch, v.startpoint, v.endpoint, mean, lam, opac, l_d, W, H)
    img.backward(gI)
    return img


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
tf = tb = 0.0
per_step = []
for _ in range(a.steps):
    for t in (mean, lam, opac, l_d):
        t.grad = None
    e0.record()
    img = F.apply(v.boxsize, batch, v.startpoint, v.endpoint, mean, lam, opac, l_d, W, H)
    e1.record()
    img.backward(gI)
    e2.record()
    torch.cuda.synchronize()
    tf += e0.elapsed_time(e1)
    tb += e1.elapsed_time(e2)
    per_step.append((round(e0.elapsed_time(e1), 3), round(e1.elapsed_time(e2), 3)))
print("per step (fwd, bwd) ms:", per_step)
print(f"splat step: fwd {tf / a.steps:.3f} ms  bwd {tb / a.steps:.3f} ms  total {(tf + tb) / a.steps:.3f} ms "
      f"({v.elements / ((tf + tb) / a.steps) / 1e6:.2f} Gelem/s), peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")

if os.environ.get("SPLAT_PROFILE"):
    # per-kernel device times of one step under the torch (kineto) profiler — for a breakdown, not a bench value
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
    if os.environ.get("SPLAT_TRACE"):
        prof.export_chrome_trace(os.environ["SPLAT_TRACE"])
