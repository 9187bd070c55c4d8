"""Generates tests/golden/ref_ops_fixture.npz by running the REFERENCE's own CUDA ops (built unchanged
from /root/reference/cuda_kernel by oracle/build_ref.sh -> oracle/_ref/grouped_cumprod_ref.so) on a
seeded input on a B200.  Run on the GPU box:

    gpurun -- python tests/golden/make_ref_fixture.py      # writes gpurun_out/ref_ops_fixture.npz

then copy the file to tests/golden/.  tests/test_oracle.py::test_oracle_matches_reference_ops_fixture
checks the oracle against it on CPU; tests/test_gpu_parity.py checks the CUDA path against it.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_util import reference_ops  # noqa: E402

ref = reference_ops()
assert ref is not None, "build oracle/_ref first (bash oracle/build_ref.sh)"
rng = np.random.default_rng(20261018)
L = np.maximum(1, np.rint(rng.lognormal(np.log(12), 1.0, 700))).astype(np.int64)
L[50] = 5000           # one long segment
inv = np.repeat(np.arange(len(L), dtype=np.int32), L)
key = (inv.astype(np.int64) * 37 % 10007).astype(np.int32)     # non-monotone keys, runs preserved
key = np.where(np.r_[True, key[1:] != key[:-1]] | True, key, key)
n = inv.size
a = 1.0 / (1.0 + np.exp(-rng.normal(1.735, 1.432, n)))
x = (1.0 - a * np.exp(-4.5 * rng.uniform(size=n))).astype(np.float32)
g = rng.uniform(0.0, 1.0, n).astype(np.float32)
seg_end = np.cumsum(L).astype(np.int32)
dx, dg, dk, di, ds = (torch.from_numpy(v).cuda() for v in (x, g, key, inv, seg_end))
y = torch.zeros_like(dx)
s = torch.zeros_like(dx)
b = torch.zeros_like(dx)
ref.grouped_cumprod_forward(dx, dk, y)
ref.grouped_cumsum_forward(dg, dk, s)
ref.grouped_cumprod_backward(dx, y, dg, di, b, ds)
torch.cuda.synchronize()
out = os.path.join(ROOT, "gpurun_out", "ref_ops_fixture.npz")
os.makedirs(os.path.dirname(out), exist_ok=True)
np.savez_compressed(out, x=x, g=g, key=key, inv=inv, seg_end=seg_end, y=y.cpu().numpy(),
                    cumsum=s.cpu().numpy(), grad_in=b.cpu().numpy(),
                    meta=np.array(["reference ops @ sm_100a, CUDA 12.9 / CCCL 2.8.2, torch 2.11; "
                                   + torch.cuda.get_device_name(0)]))
print("wrote", out, "n =", n)
