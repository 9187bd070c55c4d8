"""A/B of the multi-view loop with and without compositor.plan_view (same process, same box)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simplegaussiansplat_tk71_b200 import compositor, workloads as wl  # noqa: E402
from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F  # noqa: E402

dev = torch.device("cuda", 0)
scenes = [wl.splat_view(1920, 1080, 1_000_000, seed=1080 + i, device=dev) for i in range(2)]
leaves = [[sc.mean.float().requires_grad_(True), sc.lam.clone().requires_grad_(True),
           sc.opacity.clone().requires_grad_(True), sc.l_d.clone().requires_grad_(True)] for sc in scenes]
gI = torch.rand(1081, 1921, 3, device=dev) + 0.1
W, H = torch.tensor(1920), torch.tensor(1080)


n_param = 1_000_000
bucket = torch.zeros(n_param * 38, dtype=torch.float32, device=dev)
BUCKET = os.environ.get("MV_BUCKET", "1") == "1"


def one_view(i, plan_next):
    sc, (m, lam, o, l) = scenes[i % 2], leaves[i % 2]
    if plan_next:
        nx = scenes[(i + 1) % 2]
        compositor.plan_view(nx.boxsize, nx.startpoint, nx.endpoint, 1920, 1080)
    for t_ in (m, lam, o, l):
        t_.grad = None
    img = F.apply(sc.boxsize, torch.tensor([sc.n]), sc.startpoint, sc.endpoint, m, lam, o, l, W, H)
    img.backward(gI)
    if BUCKET:
        k = sc.n
        bucket[0:2 * k] += m.grad.reshape(-1)
        bucket[2 * n_param:2 * n_param + 4 * k] += lam.grad.reshape(-1)
        bucket[6 * n_param:6 * n_param + k] += o.grad.reshape(-1)
        bucket[7 * n_param:7 * n_param + 3 * k] += l.grad.reshape(-1)


if os.environ.get("MV_LIKE_BENCH"):
    # the state bench.py is in when it reaches its multi-view leg: C3 arrays resident, pinned host copies alive,
    # a host-buffer streamer created and dropped
    from simplegaussiansplat_tk71_b200.host import HostStreamer
    import grouped_cumprod as gcp_mod
    e = wl.c3(dev)
    y = torch.empty_like(e.x); gin = torch.empty_like(e.x)
    for _ in range(5):
        gcp_mod.grouped_cumprod_forward(e.x, e.key, y)
        gcp_mod.grouped_cumprod_backward(e.x, y, e.grad_out, e.inv, gin, e.seg_end)
    hx, hk, hg = (t_.cpu().pin_memory() for t_ in (e.x, e.key, e.grad_out))
    hy = torch.empty(e.n, dtype=torch.float32).pin_memory(); hgin = torch.empty(e.n, dtype=torch.float32).pin_memory()
    st = HostStreamer(dev, chunk_elems=8 << 20, depth=3)
    st.fwd_bwd(hx, hk, hg, hy, hgin)
    torch.cuda.synchronize()
    del st
held = None
if os.environ.get("MV_HOLD"):
    sc, (m, lam, o, l) = scenes[0], leaves[0]
    for _ in range(9):
        for t_ in (m, lam, o, l):
            t_.grad = None
        held = F.apply(sc.boxsize, torch.tensor([sc.n]), sc.startpoint, sc.endpoint, m, lam, o, l, W, H)
        held.backward(gI)
        torch.cuda.synchronize()
V = int(os.environ.get("MV_V", "32"))
for rep in range(3):
    for plan in (False, True):
        compositor._plans.clear()
        for i in range(2):
            one_view(i, plan)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        import time
        h0 = time.perf_counter()
        for i in range(V):
            one_view(i, plan)
        host_ms = (time.perf_counter() - h0) * 1e3
        b.record()
        torch.cuda.synchronize()
        print(f"rep {rep} plan_next={plan}: {a.elapsed_time(b) / V:.3f} ms per view (host enqueue {host_ms / V:.3f} ms per view)")
