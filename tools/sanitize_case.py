"""Small end-to-end case for compute-sanitizer: every op, every default kernel path (blocked TMA kernels with
fix-up phase, LDG + K2 for unaligned pointers), the compositor forward + backward."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import grouped_cumprod as gc  # noqa: E402
from simplegaussiansplat_tk71_b200 import ops  # noqa: E402
from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F  # noqa: E402

rng = np.random.default_rng(0)
L = np.maximum(1, np.rint(rng.lognormal(np.log(20), 1.0, 2500))).astype(np.int64)
L[100] = 12000          # unresolved tiles -> fix-up phase, long trailing run -> CTA re-run
inv = torch.from_numpy(np.repeat(np.arange(len(L), dtype=np.int32), L)).cuda()
n = inv.numel()
x = (1 - 0.5 * torch.rand(n, device="cuda") ** 3).contiguous()
g = torch.rand(n, device="cuda")
se = torch.from_numpy(np.cumsum(L).astype(np.int32)).cuda()
for halo in (1, 0):
    ops.set_option(0, halo)
    for off in (0, 1):          # aligned -> blocked TMA kernels; off by one element -> LDG kernels + K2
        xs, gs, ks = (torch.cat([t[:1], t])[1 - off + off:] if False else t for t in (x, g, inv))
        if off:
            xb = torch.zeros(n + 1, device="cuda"); xb[1:] = x; xs = xb[1:]
            gb = torch.zeros(n + 1, device="cuda"); gb[1:] = g; gs = gb[1:]
            kb = torch.zeros(n + 1, device="cuda", dtype=torch.int32); kb[1:] = inv; ks = kb[1:]
        y = torch.empty(n, device="cuda"); s = torch.empty(n, device="cuda"); gin = torch.empty(n, device="cuda")
        gc.grouped_cumprod_forward(xs, ks, y)
        gc.grouped_cumsum_forward(gs, ks, s)
        gc.grouped_cumprod_backward(xs, y, gs, ks, gin, se)
        torch.cuda.synchronize()
        assert ops.workspace_status() == 0
ops.set_option(0, 1)
f = np.load(os.path.join(ROOT, "tests", "golden", "compositor_fixture.npz"))
t = lambda k: torch.from_numpy(np.ascontiguousarray(f["dense/" + k])).cuda()  # noqa: E731
W, H = (int(v) for v in f["dense/WH"])
o = t("opacity").requires_grad_(True)
img = F.apply(t("boxsize"), torch.tensor([0]), t("startpoint"), t("endpoint"), t("mean").float(), t("lam"), o, t("l_d"), W, H)
(img * t("grad_image")).sum().backward()
torch.cuda.synchronize()
assert np.allclose(img.detach().cpu().numpy(), f["dense/image"], rtol=2e-4, atol=2e-5)
print("sanitize case ok, n =", n)
