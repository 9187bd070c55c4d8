"""Runs the reference's CUDA ops (oracle/_ref) and ours once each on C1 and C3, for an ncu launch list
(`ncu --metrics gpu__time_duration.sum`): kernel-only device times of thrust's DeviceScanByKey kernels and of the
reference's backward kernel beside ours — without the cudaMalloc/cudaFree/sync that the reference op pays per call
and that bench.py's `ref_cuda_ops` leg (op-level, CUDA events) includes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import grouped_cumprod as ours  # noqa: E402
from oracle import ref_function as rf  # noqa: E402
from simplegaussiansplat_tk71_b200 import workloads as wl  # noqa: E402

ref = rf.reference_ops()
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
for name, e in (("c1", wl.c1("cuda")), ("c3", wl.c3("cuda"))):
    y, s, gin = (torch.empty_like(e.x) for _ in range(3))
    for mod in (ours, ref):
        for _ in range(2):   # the second round is the one to read (warm instruction caches)
            flush.fill_(0.0)
            mod.grouped_cumprod_forward(e.x, e.key, y)
            flush.fill_(0.0)
            mod.grouped_cumsum_forward(e.grad_out, e.key, s)
            flush.fill_(0.0)
            mod.grouped_cumprod_backward(e.x, y, e.grad_out, e.inv, gin, e.seg_end)
    torch.cuda.synchronize()
    print(name, "done", e.n)
