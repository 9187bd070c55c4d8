"""Host-buffer entry point (host.py) on the C3 arrays for several chunk sizes / pipeline depths."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simplegaussiansplat_tk71_b200 import workloads as wl  # noqa: E402
from simplegaussiansplat_tk71_b200.host import HostStreamer  # noqa: E402

e = wl.c3("cpu")
n = e.x.numel()
hx, hk, hg = (t.pin_memory() for t in (e.x, e.key, e.grad_out))
hy = torch.empty(n).pin_memory()
hgin = torch.empty(n).pin_memory()
for chunk, depth in ((8 << 20, 3), (4 << 20, 3), (4 << 20, 4), (8 << 20, 4), (16 << 20, 3), (2 << 20, 6)):
    st = HostStreamer("cuda", chunk_elems=chunk, depth=depth)
    st.fwd_bwd(hx, hk, hg, hy, hgin)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(4):
        up, down = st.fwd_bwd(hx, hk, hg, hy, hgin)
    b.record()
    host_ms = (time.perf_counter() - t0) * 1e3 / 4
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 4
    print(f"chunk {chunk >> 20:3d} Mi depth {depth}: {ms:7.2f} ms/step (host loop {host_ms:6.2f} ms)  {n / ms / 1e6:5.2f} Gelem/s  "
          f"up {up / ms / 1e6:5.1f} GB/s down {down / ms / 1e6:5.1f} GB/s", flush=True)
    del st
# one direction at a time, for the link's own limits
d = torch.empty(n, device="cuda")
for name, fn in (("H2D only", lambda: d.copy_(hx, non_blocking=True)), ("D2H only", lambda: hy.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); fn(); b.record(); torch.cuda.synchronize()
    print(f"{name}: {2 * n * 4 / a.elapsed_time(b) / 1e6:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty(n, device="cuda")
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s1):
    d.copy_(hx, non_blocking=True); d.copy_(hg, non_blocking=True)
with torch.cuda.stream(s2):
    hy.copy_(d2, non_blocking=True); hgin.copy_(d2, non_blocking=True)
torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
b.record(); torch.cuda.synchronize()
print(f"both directions at once: {2 * n * 4 / a.elapsed_time(b) / 1e6:.1f} GB/s each way")
