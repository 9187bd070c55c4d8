"""Host-side phase times of the tile-route forward of one 1080p view (perf_counter between the calls), and the
device time of the whole forward: where the wall time beyond the kernels goes."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simplegaussiansplat_tk71_b200 import _lib, compositor as C, workloads as wl  # noqa: E402

v = wl.splat_view(1920, 1080, 1_000_000, device="cuda")
L = _lib.lib()
dev = v.startpoint.device
W, H, n = v.width, v.height, v.n
p = lambda t: t.data_ptr()  # noqa: E731
mean, lam, opac, l_d = v.mean.float(), v.lam, v.opacity, v.l_d


def forward(marks):
    t = time.perf_counter
    marks.append(("start", t()))
    stream = torch.cuda.current_stream(dev).cuda_stream
    sp, ep, toff, totals = C._prologue_tiles(L, dev, v.startpoint, v.endpoint, n, W, H)
    marks.append(("prologue queued", t()))
    host, event = C._totals_to_host(totals, dev)
    marks.append(("totals copy queued", t()))
    l_d_ = C._aligned(l_d.detach().to(torch.float32))
    mean_ = C._aligned(mean.detach().to(torch.float32))
    lam_ = C._aligned(lam.detach().to(torch.float32).reshape(n, 4))
    opac_ = opac.detach().to(torch.float32).reshape(n).contiguous()
    rec = torch.empty((n, 16), dtype=torch.int32, device=dev)
    _lib.check(L.gcp_tile_pack(p(mean_), p(lam_), p(opac_), p(l_d_), p(sp), p(ep), p(toff), n, W, H, p(rec), stream), "pack")
    image = torch.empty((H + 1, W + 1, 3), dtype=torch.float32, device=dev)
    tstart = torch.empty(int(L.gcp_tile_num_tiles(W, H)) + 1, dtype=torch.int32, device=dev)
    marks.append(("pack queued", t()))
    event.synchronize()
    (P,) = host.tolist()
    marks.append(("pair count on host", t()))
    plan = torch.empty(int(L.gcp_tile_plan_ints(P, W, H)), dtype=torch.int32, device=dev)
    pstate = torch.empty(int(L.gcp_tile_state_floats(P, W, H)), dtype=torch.float32, device=dev)
    pgid = torch.empty(max(P, 1), dtype=torch.int32, device=dev)
    tkeep = torch.empty(max(P, 1) * 32, dtype=torch.float32, device=dev)
    temp = C._scratch_bytes(dev, "bin", int(L.gcp_tile_bin_bytes(P, W, H)))
    marks.append(("allocations", t()))
    _lib.check(L.gcp_tile_bin(p(sp), p(ep), p(toff), n, P, W, H, p(tstart), p(plan), p(pgid), p(temp), temp.numel(), stream), "bin")
    marks.append(("bin queued", t()))
    _lib.check(L.gcp_tile_render(p(tstart), p(plan), p(pgid), p(rec), P, W, H, p(image), p(tkeep), p(pstate), stream), "render")
    marks.append(("render queued", t()))
    torch.cuda.synchronize()
    marks.append(("device idle", t()))
    return image


for _ in range(3):
    forward([])
res = []
for _ in range(7):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m = []
    torch.cuda.synchronize()
    a.record()
    forward(m)
    b.record()
    torch.cuda.synchronize()
    res.append((a.elapsed_time(b), m))
res.sort(key=lambda r: r[0])
ms, m = res[len(res) // 2]
print(f"forward (events): {ms:.3f} ms")
t0 = m[0][1]
prev = t0
for name, tt in m[1:]:
    print(f"  {name:22s} +{(tt - prev) * 1e6:7.1f} us   (at {(tt - t0) * 1e6:7.1f} us)")
    prev = tt
