"""The batch of views (views.NativeViewBatch, gcp_views_step) over 1..4 stream lanes: step time of V distinct 1080p
views.  GCP_WALK_PER_SM / GCP_BATCH_PDL in the environment override the walk-grid cap and the dependent launches."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simplegaussiansplat_tk71_b200 import views as vw, workloads as wl  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 32
lanes_list = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 3, 4]
dev = torch.device("cuda", 0)
W, H, n = 1920, 1080, 1_000_000
views = [wl.splat_view_device(W, H, n, seed=1080 + v, device=dev) for v in range(V)]
target = torch.rand(H + 1, W + 1, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(64))
g = [torch.zeros(n, k, device=dev) for k in (2, 4)] + [torch.zeros(n, device=dev), torch.zeros(n, 3, device=dev)]
for lanes in lanes_list:
    b = vw.NativeViewBatch(views, W, H, targets=[target] * V, lanes=lanes)
    for _ in range(2):
        b.step(*g)
        torch.cuda.synchronize()
        b.finish()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(3):
        e0.record()
        b.step(*g)
        e1.record()
        torch.cuda.synchronize()
        assert b.finish()
        ts.append(e0.elapsed_time(e1))
    print(f"lanes {lanes} cap {os.environ.get('GCP_WALK_PER_SM', 'default')} pdl {os.environ.get('GCP_BATCH_PDL', 'default')}: "
          f"{min(ts):.2f} ms for {V} views = {min(ts) / V:.4f} ms per view", flush=True)
    del b
