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
a = ap.parse_args()
v = wl.splat_view(a.width, a.height, a.n, device="cuda")
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
print(f"splat step: fwd {tf / a.steps:.3f} ms  bwd {tb / a.steps:.3f} ms  total {(tf + tb) / a.steps:.3f} ms "
      f"({v.elements / ((tf + tb) / a.steps) / 1e6:.2f} Gelem/s), peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
