"""-m gpu: parity at BASELINE.json's full sizes.

C3 (1080p, 68 M elements) is small enough for the sequential C oracle (fractions of a second per pass), so the
CUDA path is compared element by element against fp64 at full size; C4 (4K + deep segments) is checked at a
quarter of its pixel rows against the oracle, and at full size through size-independent properties:
  * cumsum of ones   = position inside the pixel list (exact small integers)
  * cumprod of ones  = 1 everywhere; cumprod of exact powers of two is bit-exact in any association order
  * backward with x = 1 and g = 1 = number of elements from i to the end of its list (exact)
  * the last element of every list in the forward equals the product over the list (segment_reduce)
"""
import numpy as np
import pytest
import torch

from gpu_util import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[1, 0, 2], ids=["chain", "tickets", "chain-by-hint"], autouse=True)
def chain_mode(request):
    """GCP_OPT_CHAIN of the blocked backward: default (follows the workspace hint), tickets, contiguous ranges."""
    from simplegaussiansplat_tk71_b200 import ops

    ops.set_option(1, request.param)
    yield request.param
    ops.set_option(1, 1)


def _ops():
    import grouped_cumprod as gc
    from simplegaussiansplat_tk71_b200 import ops

    return gc, ops


def test_c3_full_size_against_fp64_oracle(oracle):
    from simplegaussiansplat_tk71_b200 import workloads as wl

    gc, ops = _ops()
    e = wl.c3("cuda")
    y = torch.empty_like(e.x)
    s = torch.empty_like(e.x)
    gin = torch.empty_like(e.x)
    gpos = e.grad_out.abs()
    gc.grouped_cumprod_forward(e.x, e.key, y)
    gc.grouped_cumsum_forward(gpos, e.key, s)
    gc.grouped_cumprod_backward(e.x, y, gpos, e.inv, gin, e.seg_end)
    torch.cuda.synchronize()
    assert ops.workspace_status() == 0
    assert ops.validate_segments(e.inv, e.seg_end) == 0          # integer side of the workload: bit-exact contract
    x, key, inv, g = (t.cpu().numpy() for t in (e.x, e.key, e.inv, gpos))
    assert_close(y.cpu().numpy(), oracle.cumprod_fwd(x, key), "C3 fwd")
    assert_close(s.cpu().numpy(), oracle.cumsum_fwd(g, key), "C3 cumsum")
    assert_close(gin.cpu().numpy(), oracle.cumprod_bwd_exact(x, g, inv), "C3 bwd")


def test_c4_quarter_size_with_deep_segments_against_oracle(oracle):
    from simplegaussiansplat_tk71_b200 import workloads as wl

    gc, ops = _ops()
    e = wl.c4("cuda", scale=0.25)            # 3840 x 540 pixel lists + 128 deep lists of 8 Ki..256 Ki elements
    y = torch.empty_like(e.x)
    gin = torch.empty_like(e.x)
    gpos = e.grad_out.abs()
    gc.grouped_cumprod_forward(e.x, e.key, y)
    gc.grouped_cumprod_backward(e.x, y, gpos, e.inv, gin, e.seg_end)
    torch.cuda.synchronize()
    assert ops.workspace_status() == 0
    x, key, inv, g = (t.cpu().numpy() for t in (e.x, e.key, e.inv, gpos))
    ref = oracle.cumprod_fwd(x, key)
    seq = np.abs(oracle.cumprod_fwd(x, key, np.float32) - ref)
    got = y.cpu().numpy()
    # lists of 1e5 elements: the bound adds the error the sequential fp32 evaluation itself makes
    assert np.all(np.abs(got - ref) <= 1e-6 + 1e-5 * np.abs(ref) + 8 * seq.max()), np.abs(got - ref).max()
    assert_close(gin.cpu().numpy(), oracle.cumprod_bwd_exact(x, g, inv), "C4/4 bwd", rtol=2e-4, atol=1e-5)


def test_c4_full_size_properties():
    from simplegaussiansplat_tk71_b200 import workloads as wl

    gc, ops = _ops()
    e = wl.c4("cuda")                        # 475 M elements, 8.29 M lists, 512 of them 8 Ki..256 Ki long
    n = e.n
    start = torch.zeros(e.k, dtype=torch.int64, device="cuda")
    start[1:] = e.seg_end[:-1].long()
    pos = (torch.arange(n, device="cuda", dtype=torch.int64) - start[e.inv.long()]).float()   # < 2^24: exact
    length = (e.seg_end.long() - start)[e.inv.long()].float()
    del start
    one = torch.ones(n, device="cuda")
    out = torch.empty(n, device="cuda")
    gc.grouped_cumsum_forward(one, e.key, out)
    assert torch.equal(out, pos + 1)                                   # position inside the list
    gc.grouped_cumprod_forward(one, e.key, out)
    assert torch.equal(out, one)
    gc.grouped_cumprod_backward(one, one, one, e.inv, out, e.seg_end)
    assert torch.equal(out, length - pos)                              # elements from i to the end of its list
    # exact powers of two: x = 2 at even positions, 1/2 at odd -> inclusive product is 2 or 1, in any order
    xp = torch.where((pos.long() & 1) == 0, 2.0 * one, 0.5 * one)
    gc.grouped_cumprod_forward(xp, e.key, out)
    assert torch.equal(out, torch.where((pos.long() & 1) == 0, 2.0 * one, one))
    del xp, pos, length, one
    # real values: the last element of each list = product over the list (fp64 log-sum, loose) and is in (0, 1]
    gc.grouped_cumprod_forward(e.x, e.key, out)
    torch.cuda.synchronize()
    assert ops.workspace_status() == 0
    last = out[(e.seg_end.long() - 1)]
    assert bool(((last >= 0) & (last <= 1)).all())
    logsum = torch.zeros(e.k, dtype=torch.float64, device="cuda").index_add_(0, e.inv.long(), torch.log(e.x.double()))
    ok = torch.isclose(last.double(), torch.exp(logsum), rtol=2e-3, atol=1e-30)
    assert bool(ok.all()), int((~ok).sum())


def test_splat_step_full_size_routes_agree_and_exact_properties(monkeypatch):
    """BASELINE.json's splat-step workload at full size (1920x1080, 1 M Gaussians, 37 M elements): the fused tile
    route against the element-list route (scan ops a1/a3 inside), plus properties that hold exactly:
      * opacity 0 everywhere -> a black image and zero gradients (T = 1 all along every list);
      * the tile route is bitwise reproducible (no float atomics)."""
    from simplegaussiansplat_tk71_b200 import compositor, workloads as wl
    from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F

    v = wl.splat_view(1920, 1080, 1_000_000, device="cuda")
    gI = torch.rand(v.height + 1, v.width + 1, 3, device="cuda") + 0.1

    def step(route, opacity):
        monkeypatch.setattr(compositor, "ROUTE", route)
        m, lam, o, l = (v.mean.float().requires_grad_(True), v.lam.clone().requires_grad_(True),
                        opacity.clone().requires_grad_(True), v.l_d.clone().requires_grad_(True))
        img = F.apply(v.boxsize, torch.tensor([v.n]), v.startpoint, v.endpoint, m, lam, o, l, v.width, v.height)
        img.backward(gI)
        return [img.detach()] + [t.grad for t in (m, lam, o, l)]

    tiles, tiles2, lists = step("tiles", v.opacity), step("tiles", v.opacity), step("lists", v.opacity)
    for a, b in zip(tiles, tiles2):
        assert torch.equal(a, b)
    for name, a, b in zip(("image", "mean", "lambda", "opacity", "l"), tiles, lists):
        assert torch.isfinite(a).all(), name
        scale = float(b.abs().max())
        assert torch.allclose(a, b, rtol=2e-4, atol=2e-6 * scale), (name, float((a - b).abs().max()), scale)
    black = step("tiles", torch.zeros_like(v.opacity))
    assert float(black[0].abs().max()) == 0.0
    for g in (black[1], black[2], black[4]):      # d_mean, d_Lambda, d_l carry a factor alpha = 0
        assert float(g.abs().max()) == 0.0
    # d_opacity = sum over the box of g * <dL/dI, l> with T = 1 and U = 0: positive, finite
    assert torch.isfinite(black[3]).all() and float(black[3].min()) >= 0.0
