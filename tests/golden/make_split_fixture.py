"""Writes tests/golden/split_by_cumsum.json from the REFERENCE's own Utilities.split_by_cumsum_parallel
(/root/reference/uitility.py:478-488), imported unmodified (stub modules for pycolmap / kornia, absent here)."""
import json
import os
import sys
import types

import torch

for name in ("kornia", "kornia.metrics", "pycolmap"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["kornia"].metrics = sys.modules["kornia.metrics"]
sys.path.insert(0, "/root/reference")
from uitility import Utilities  # noqa: E402

g = torch.Generator().manual_seed(0)
cases = []
for n, limit in ((7, 3.0), (40, 10.0), (100, 25.5), (64, 1000.0), (30, 0.75)):
    x = (torch.rand(n, generator=g) * 4).round() / 2 + 0.5          # multiples of 0.5: exact in fp32
    counts = Utilities.split_by_cumsum_parallel(x, limit)
    cases.append({"x": x.tolist(), "limit": limit, "counts": counts.tolist()})
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "split_by_cumsum.json")
json.dump(cases, open(out, "w"))
print("wrote", out, [c["counts"] for c in cases])
