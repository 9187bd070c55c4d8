"""Compositor (rows a5-a9 of SURVEY.md §8): the native `custom_autograd_grouped_cumprod` against fixtures the
reference's own Function produced (tests/golden/compositor_fixture.npz) and against the fp64 oracle.

CPU variant: the host logic only, with the two scan ops swapped for the oracle (test-only monkeypatch).
GPU variant (-m gpu): the real thing, through the C ABI."""
import os
import types

import numpy as np
import pytest
import torch

from test_compositor_oracle import CASES, FIX, load_case


ROUTES = ["tiles", "lists"]   # compositor.ROUTE: the fused tile walk / the element lists + scan ops


@pytest.fixture(params=ROUTES)
def route(request, monkeypatch):
    from simplegaussiansplat_tk71_b200 import compositor

    monkeypatch.setattr(compositor, "ROUTE", request.param)
    return request.param


@pytest.fixture
def lists_route(monkeypatch):
    from simplegaussiansplat_tk71_b200 import compositor

    monkeypatch.setattr(compositor, "ROUTE", "lists")


def _run(case, device, F=None):
    if F is None:
        from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F

    t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(device if dt is None else device, dtype=dt)  # noqa: E731
    boxsize = t(case["boxsize"])
    sp, ep = t(case["sp"]), t(case["ep"])
    mean = t(case["mean"]).float().requires_grad_(True)
    lam = t(case["lam"]).requires_grad_(True)
    opac = t(case["opac"]).requires_grad_(True)
    l_d = t(case["l_d"]).requires_grad_(True)
    n = boxsize.numel()
    img = F.apply(boxsize, torch.tensor([n]), sp, ep, mean, lam, opac, l_d, torch.tensor(case["W"]), torch.tensor(case["H"]))
    (img * t(case["grad_image"])).sum().backward()
    return [v.detach().cpu().numpy() for v in (img, mean.grad, lam.grad, opac.grad, l_d.grad)]


RTOL, ATOL = 1e-5, 1e-6   # BASELINE.json north_star: rel 1e-5 / abs 1e-6 in fp32, against the fp64 oracle
NAMES = ("image", "grad_mean", "grad_lambda", "grad_opacity", "grad_l")


def _oracle(case):
    """fp64 oracle results and condition scales (sum of |terms|, oracle.compositor_oracle.backward(scales=True))."""
    from oracle import compositor_oracle as co

    img, cache = co.forward(case["boxsize"], case["sp"], case["ep"], case["mean"], case["lam"], case["opac"],
                            case["l_d"], case["W"], case["H"])
    grads, sc = co.backward(cache, case["grad_image"], scales=True)
    return (img,) + tuple(grads), (co.image_scale(cache, case["W"], case["H"]),) + tuple(sc)


def _check(got, case, fixture=True, oracle=None):
    """The north star's tolerance: |got - fp64| <= 1e-6 + 1e-5 * scale for all five outputs, scale = the sum of the
    absolute values of the terms an output is made of (= |fp64| where nothing cancels; measured worst ratio over
    all scenes of tools/compositor_errors.py: 0.18 image, 0.39 gradients — profiles/r02_compositor_error_table.txt).
    Against the fixture the reference's own fp32 Function produced the bound is doubled: both sides are within one
    bound of the fp64 value (the reference's worst ratio to the oracle is 0.16, tests/test_compositor_oracle.py)."""
    ref, scales = oracle if oracle is not None else _oracle(case)
    assert got[0].shape == (case["H"] + 1, case["W"] + 1, 3)
    for name, a, b, sc in zip(NAMES, got, ref, scales):
        a = np.asarray(a, np.float64).reshape(b.shape)
        bound = ATOL + RTOL * np.asarray(sc).reshape(b.shape)
        err = np.abs(a - b)
        assert np.all(err <= bound), (name, "vs fp64 oracle", float((err / bound).max()), float(err.max()))
        if fixture and name in case:
            f = np.asarray(case[name], np.float64).reshape(b.shape)
            errf = np.abs(a - f)
            assert np.all(errf <= 2 * bound), (name, "vs reference fixture", float((errf / bound).max()))


@pytest.fixture
def oracle_backed_ops(monkeypatch):
    """The torch fp32 formulation (oracle/compositor_torch.py) with the scan ops backed by the C oracle."""
    from oracle import oracle as orc
    from oracle import compositor_torch as compositor

    def fwd(x, key, y):
        y.copy_(torch.from_numpy(orc.cumprod_fwd(x.numpy(), key.numpy(), np.float32)))

    def bwd(param, cp, gout, inv, gin, inv_len):
        gin.copy_(torch.from_numpy(orc.cumprod_bwd_exact(param.numpy(), gout.numpy(), inv.numpy()).astype(np.float32)))

    monkeypatch.setattr(compositor, "ops", types.SimpleNamespace(grouped_cumprod_forward=fwd,
                                                                 grouped_cumprod_backward=bwd))


@pytest.mark.parametrize("name", CASES)
def test_torch_formulation_matches_reference_fixture_cpu(name, oracle_backed_ops):
    """Pins the division-free algorithm itself (T_i U_i through the backward op) to the reference's Function."""
    from oracle.compositor_torch import custom_autograd_grouped_cumprod as F

    case = load_case(np.load(FIX), name)
    _check(_run(case, "cpu", F), case)


def test_native_compositor_refuses_cpu_tensors():
    from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F

    case = load_case(np.load(FIX), "small")
    with pytest.raises(RuntimeError, match="no CPU path"):
        _run(case, "cpu", F)


def test_element_plan_integer_side_is_bit_exact():
    """Sort permutation, segment ids and offsets against torch.sort(stable=True) / unique_consecutive."""
    from oracle.compositor_torch import ElementPlan
    from oracle import compositor_oracle as co

    case = load_case(np.load(FIX), "wide")
    plan = ElementPlan(torch.from_numpy(case["boxsize"]), torch.from_numpy(case["sp"]), torch.from_numpy(case["ep"]))
    gid, px, py = co.expand(case["boxsize"], case["sp"], case["ep"])
    key = (py * 10000 + px).astype(np.int32)
    order = np.argsort(key, kind="stable")
    assert np.array_equal(plan.key_s.numpy(), key[order])
    assert np.array_equal(plan.gid_s.numpy(), gid[order])
    _, counts = torch.unique_consecutive(plan.key_s, return_counts=True)
    assert np.array_equal(plan.seg_end.numpy(), np.cumsum(counts.numpy()).astype(np.int32))
    assert np.array_equal(plan.inv.numpy(), np.repeat(np.arange(len(counts)), counts.numpy()).astype(np.int32))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_native_compositor_matches_reference_fixture_gpu(name, route):
    case = load_case(np.load(FIX), name)
    _check(_run(case, "cuda"), case)


@pytest.mark.gpu
@pytest.mark.parametrize("placement", [True, False], ids=["counting-placement", "expand+radix-sort"])
@pytest.mark.parametrize("name", ["wide", "dense", "opaque"])
def test_native_element_list_is_bit_exact_gpu(name, placement, monkeypatch, lists_route):
    """Integer side of the native path: keys, Gaussian ids and their stable order vs numpy argsort(stable),
    for both ways of building the list."""
    from oracle import compositor_oracle as co
    from simplegaussiansplat_tk71_b200 import compositor

    monkeypatch.setattr(compositor, "USE_PLACEMENT", placement)
    case = load_case(np.load(FIX), name)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    _, view = compositor._render_forward(t(case["boxsize"]), t(case["sp"]), t(case["ep"]), t(case["mean"]).float(),
                                         t(case["lam"]), t(case["opac"]), t(case["l_d"]), case["W"], case["H"])
    gid, px, py = co.expand(case["boxsize"], case["sp"], case["ep"])
    key = (py * 10000 + px).astype(np.int32)
    order = np.argsort(key, kind="stable")
    assert np.array_equal(view.key_s.cpu().numpy(), key[order])
    assert np.array_equal(view.gid_s.cpu().numpy(), gid[order].astype(np.int32))


@pytest.mark.gpu
def test_placement_equals_sort_on_a_1080p_scene_gpu(monkeypatch, lists_route):
    """The two list builders agree bit for bit at full scale (0.2 M Gaussians, 1920x1080, ~8 M elements)."""
    from simplegaussiansplat_tk71_b200 import compositor, workloads as wl

    v = wl.splat_view(1920, 1080, 200_000, seed=7, device="cuda")
    outs = []
    for placement in (True, False):
        monkeypatch.setattr(compositor, "USE_PLACEMENT", placement)
        img, view = compositor._render_forward(v.boxsize, v.startpoint, v.endpoint, v.mean.float(), v.lam, v.opacity,
                                               v.l_d, v.width, v.height)
        outs.append((view.key_s.clone(), view.gid_s.clone(), img))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.allclose(outs[0][2], outs[1][2], rtol=1e-5, atol=1e-6)   # colour sums use float atomics


@pytest.mark.gpu
def test_placement_equals_sort_on_long_pixel_lists_gpu(monkeypatch, lists_route):
    """Bundled-scene view (C2): per-pixel lists of hundreds of elements take the transposed (ballot-compacting)
    fill and the warp-per-list key kernel; the result must still be the stable sort, bit for bit."""
    from simplegaussiansplat_tk71_b200 import compositor, workloads as wl

    v = wl.bundled_views("cuda", n_views=3)[2]
    outs = []
    for placement in (True, False):
        monkeypatch.setattr(compositor, "USE_PLACEMENT", placement)
        img, view = compositor._render_forward(v.boxsize, v.startpoint, v.endpoint, v.mean.float(), v.lam, v.opacity,
                                               v.l_d, v.width, v.height)
        outs.append((view.key_s.clone(), view.gid_s.clone(), img))
        del view
    assert outs[0][0].numel() == v.elements
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.allclose(outs[0][2], outs[1][2], rtol=1e-4, atol=1e-5)   # colour sums use float atomics


@pytest.mark.gpu
def test_long_list_backward_equals_per_element_backward_gpu(lists_route):
    """C2 view: the cell-walking, shared-memory-transposing backward un-sort (long pixel lists) must give exactly
    the gradients of the per-element scatter version (the reduction that follows is deterministic)."""
    from simplegaussiansplat_tk71_b200 import _lib, workloads as wl
    from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F

    v = wl.bundled_views("cuda", n_views=3)[2]
    gI = torch.rand(v.height + 1, v.width + 1, 3, device="cuda") + 0.1
    L = _lib.lib()
    grads = []
    try:
        for thr in (8, 1 << 20):  # default (long-list kernels) / short-list kernels forced
            L.gcp_splat_set_long_list_threshold(thr)
            m, lam, o, l = (v.mean.float().requires_grad_(True), v.lam.clone().requires_grad_(True),
                            v.opacity.clone().requires_grad_(True), v.l_d.clone().requires_grad_(True))
            img = F.apply(v.boxsize, torch.tensor([v.n]), v.startpoint, v.endpoint, m, lam, o, l, v.width, v.height)
            img.backward(gI)
            grads.append([t.grad.clone() for t in (m, lam, o, l)])
            del img
    finally:
        L.gcp_splat_set_long_list_threshold(8)
    for a, b in zip(*grads):
        assert torch.isfinite(a).all()
        assert torch.equal(a, b)


@pytest.mark.gpu
def test_native_compositor_matches_oracle_on_a_larger_scene_gpu(route):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from oracle import compositor_oracle as co
    import make_compositor_fixture as mk

    W, H, n = 160, 120, 4000
    boxsize, sp, ep, mean, lam, opac, l_d = mk.make_scene(11, W, H, n, 9, opaque=20)
    rng = np.random.default_rng(5)
    gI = (rng.uniform(0.1, 1.0, (H + 1, W + 1, 3))).astype(np.float32)
    case = dict(boxsize=boxsize.numpy(), sp=sp.numpy(), ep=ep.numpy(), mean=mean.numpy(), lam=lam.numpy(),
                opac=opac.numpy(), l_d=l_d.numpy(), W=W, H=H, grad_image=gI)
    _check(_run(case, "cuda"), case, fixture=False)


@pytest.mark.gpu
def test_native_compositor_edge_cases_gpu(route):
    from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F

    dev = "cuda"
    W, H = 20, 10
    # no Gaussians at all: a black image and empty gradients
    z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=dev)  # noqa: E731
    mean = z(0, 2).requires_grad_(True)
    img = F.apply(z(0, dt=torch.int64), torch.tensor([0]), z(0, 2, dt=torch.int32), z(0, 2, dt=torch.int32), mean,
                  z(0, 2, 2), z(0, 1), z(0, 3), W, H)
    assert img.shape == (H + 1, W + 1, 3) and float(img.abs().sum()) == 0.0
    # one single-pixel Gaussian with opacity 0.5 sitting exactly on its pixel: image = 0.5 * l there, T = 1
    sp = torch.tensor([[7, 3]], dtype=torch.int32, device=dev)
    o = torch.tensor([[0.5]], device=dev, requires_grad=True)
    l = torch.tensor([[0.2, 0.4, 0.8]], device=dev, requires_grad=True)
    lam = torch.eye(2, device=dev)[None].clone().requires_grad_(True)
    m = torch.tensor([[7.0, 3.0]], device=dev, requires_grad=True)
    img = F.apply(torch.tensor([1], device=dev), torch.tensor([1]), sp, sp.clone(), m, lam, o, l, torch.tensor(W),
                  torch.tensor(H))
    assert torch.allclose(img[3, 7], torch.tensor([0.1, 0.2, 0.4], device=dev))
    assert float(img.sum()) == pytest.approx(0.7)
    img.sum().backward()
    assert torch.allclose(o.grad, torch.tensor([[1.4]], device=dev))        # g * <1, l> = 0.2+0.4+0.8
    # the reference's d/l for the colour gradient (gs_model.py:763-766): d = T alpha <1,l> = 0.7
    assert torch.allclose(l.grad, torch.tensor([[3.5, 1.75, 0.875]], device=dev))
    assert float(m.grad.abs().sum()) == 0.0 and float(lam.grad.abs().sum()) == 0.0   # r - m = 0


@pytest.mark.gpu
def test_bundled_scene_front_slice_against_oracle_gpu(route):
    """C2 (BASELINE.json configs[1]): the nearest 4000 Gaussians of a bundled-scene view — boxes of hundreds of
    pixels, per-pixel lists of hundreds of elements (most tiles go through the fix-up phase)."""
    from oracle import compositor_oracle as co
    from simplegaussiansplat_tk71_b200 import workloads as wl

    sc = wl.bundled_views("cpu", n_views=2)[1]
    k = 4000
    case = dict(boxsize=sc.boxsize[:k].numpy(), sp=sc.startpoint[:k].numpy(), ep=sc.endpoint[:k].numpy(),
                mean=sc.mean[:k].numpy(), lam=sc.lam[:k].numpy(), opac=sc.opacity[:k].numpy(), l_d=sc.l_d[:k].numpy(),
                W=sc.width, H=sc.height)
    rng = np.random.default_rng(9)
    gI = rng.uniform(0.1, 1.0, (sc.height + 1, sc.width + 1, 3)).astype(np.float32)
    case.update(grad_image=gI)
    # sums over boxes of thousands of elements, per-pixel lists of hundreds: the condition-aware bound of _check
    _check(_run(case, "cuda"), case, fixture=False)


@pytest.mark.gpu
def test_planned_view_equals_unplanned_gpu(route):
    """compositor.plan_view only moves the prologue to a side stream: same element list, same image, same grads."""
    from simplegaussiansplat_tk71_b200 import compositor, workloads as wl
    from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F

    v = wl.splat_view(640, 360, 60_000, seed=3, device="cuda")
    gI = torch.rand(v.height + 1, v.width + 1, 3, device="cuda") + 0.1
    res = []
    for planned in (False, True, True):
        if planned:
            compositor.plan_view(v.boxsize, v.startpoint, v.endpoint, v.width, v.height)
            assert len(compositor._plans) == 1
        m, lam, o, l = (v.mean.float().requires_grad_(True), v.lam.clone().requires_grad_(True),
                        v.opacity.clone().requires_grad_(True), v.l_d.clone().requires_grad_(True))
        img = F.apply(v.boxsize, torch.tensor([v.n]), v.startpoint, v.endpoint, m, lam, o, l, v.width, v.height)
        assert len(compositor._plans) == 0          # consumed
        img.backward(gI)
        res.append([img.detach().clone()] + [t.grad.clone() for t in (m, lam, o, l)])
    for other in res[1:]:
        assert torch.allclose(res[0][0], other[0], rtol=1e-5, atol=1e-6)   # colour sums use float atomics
        for a, b in zip(res[0][1:], other[1:]):
            assert torch.equal(a, b)


def _arena_arrays(L, view, W, H):
    """The integer arrays of a rendered tile-route view, read out of its two arenas (gcp_view_layout)."""
    import ctypes

    out = (ctypes.c_int64 * 16)()
    assert L.gcp_view_layout(view.n, W, H, view.pairs.cap, out) == 0
    ntiles = L.gcp_tile_num_tiles(W, H)
    plan = view.plan.buf.cpu().numpy()
    pairs = view.pairs.buf.cpu().numpy()
    i32 = lambda buf, off, cnt: buf[off:off + 4 * cnt].view(np.int32)  # noqa: E731
    return dict(toff=i32(plan, out[0], view.n + 1), tcount=i32(plan, out[1], ntiles), tstart=i32(plan, out[2], ntiles + 1),
                pextra=i32(plan, out[3], ntiles), hdr=i32(plan, out[4], 64), pgid=i32(pairs, out[6], view.pairs.cap),
                ptile_x=i32(pairs, out[7], out[8]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["wide", "dense", "opaque"])
def test_tile_pair_list_is_bit_exact_gpu(name, monkeypatch):
    """Integer side of the tile route: pair offsets per Gaussian, pairs per tile, every tile's list in the order
    of a stable sort by tile, the tile offsets and the pieces of long lists, against the numpy restatement
    (oracle/tile_oracle.py, itself pinned on CPU to the reference's sorted element list by
    tests/test_compositor_oracle.py).  The native binning is a stable radix sort of the Gaussian-major pair list by
    tile id: the RESULT must be the stable sort, bit for bit."""
    from oracle import tile_oracle as to
    from simplegaussiansplat_tk71_b200 import _lib, compositor

    monkeypatch.setattr(compositor, "ROUTE", "tiles")
    case = load_case(np.load(FIX), name)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    L = _lib.lib()
    default = L.gcp_tile_piece_pairs()
    W, H = case["W"], case["H"]
    assert (L.gcp_tile_width(), L.gcp_tile_height()) == (to.TW, to.TH)
    ntx, nty = to.num_tiles(W, H)
    ntiles = ntx * nty
    assert ntiles == L.gcp_tile_num_tiles(W, H)
    tiles, gids, toff = to.tile_pairs(case["sp"], case["ep"], W, H)
    gid_s, start, _ = to.sort_by_tile(tiles, gids, ntiles)
    counts = np.diff(start)
    try:
        for piece in (32, default):
            assert L.gcp_tile_set_piece_pairs(piece) == 0
            _, view = compositor._render_forward(t(case["boxsize"]), t(case["sp"]), t(case["ep"]),
                                                 t(case["mean"]).float(), t(case["lam"]), t(case["opac"]),
                                                 t(case["l_d"]), W, H)
            torch.cuda.synchronize()
            a = _arena_arrays(L, view, W, H)
            assert view.P == len(tiles)
            assert np.array_equal(a["toff"], toff.astype(np.int32))
            assert np.array_equal(a["tcount"], counts)
            assert np.array_equal(a["tstart"], start)
            for tl in range(ntiles):
                seg = a["pgid"][a["tstart"][tl]:a["tstart"][tl] + counts[tl]]
                assert np.array_equal(seg, gid_s[start[tl]:start[tl + 1]]), f"tile {tl}"
            # pieces: a tile with more than `piece` pairs owns ceil(len / piece) consecutive slots of the extra table
            npieces = np.where(counts > piece, -(-counts // piece), 0)
            assert a["hdr"][3] == int(npieces.sum())
            seen = np.zeros(int(npieces.sum()), bool)
            for tl in range(ntiles):
                x = a["pextra"][tl]
                if npieces[tl] == 0:
                    assert x == -1
                    continue
                assert a["ptile_x"][x] == -1 and np.all(a["ptile_x"][x + 1:x + npieces[tl]] == tl)
                assert not seen[x:x + npieces[tl]].any()
                seen[x:x + npieces[tl]] = True
            assert seen.all()
    finally:
        L.gcp_tile_set_piece_pairs(default)


@pytest.mark.gpu
def test_tile_binning_of_long_lists_is_bit_exact_gpu(monkeypatch):
    """Tile lists of thousands of pairs (bundled scene): still the stable order."""
    from simplegaussiansplat_tk71_b200 import _lib, compositor, workloads as wl

    monkeypatch.setattr(compositor, "ROUTE", "tiles")
    L = _lib.lib()
    classes = set()
    for v in wl.bundled_views("cuda", n_views=3):
        _, view = compositor._render_forward(v.boxsize, v.startpoint, v.endpoint, v.mean.float(), v.lam, v.opacity,
                                             v.l_d, v.width, v.height)
        torch.cuda.synchronize()
        a = _arena_arrays(L, view, v.width, v.height)
        assert a["tcount"].max() > 2048, "this scene is expected to have long tile lists"
        assert int(a["tcount"].sum()) == view.P
        # every segment of every tile strictly increasing (distinct Gaussian ids in depth order)
        d = np.diff(a["pgid"][:view.P].astype(np.int64))
        ends = (a["tstart"][:-1] + a["tcount"])[a["tcount"] > 0] - 1          # last pair of every non-empty tile
        inner = np.ones(max(view.P - 1, 0), bool)
        inner[ends[ends < view.P - 1]] = False
        assert np.all(d[inner] > 0), "a tile list is not in Gaussian order"
        classes |= {"huge"} if (a["tcount"] > 4096).any() else set()
        classes |= {"long"} if ((a["tcount"] > 512) & (a["tcount"] <= 4096)).any() else set()
        classes |= {"warp"} if ((a["tcount"] > 1) & (a["tcount"] <= 512)).any() else set()
    assert classes == {"huge", "long", "warp"}, classes
    tw, th = L.gcp_tile_width(), L.gcp_tile_height()
    ntx = (v.width + tw) // tw
    sp, ep = v.startpoint.cpu().numpy(), v.endpoint.cpu().numpy()
    rng = np.random.default_rng(0)
    for tl in np.concatenate([np.argsort(a["tcount"])[-4:], rng.integers(0, len(a["tcount"]), 60)]):
        seg = a["pgid"][a["tstart"][tl]:a["tstart"][tl] + a["tcount"][tl]]
        assert np.all(np.diff(seg) > 0)
        # ... and exactly the Gaussians whose clipped box touches the tile
        ty, tx = divmod(int(tl), ntx)
        hit = (np.maximum(sp[:, 0], 0) // tw <= tx) & (np.minimum(ep[:, 0], v.width) // tw >= tx) & \
              (np.maximum(sp[:, 1], 0) // th <= ty) & (np.minimum(ep[:, 1], v.height) // th >= ty) & \
              (ep[:, 0] >= sp[:, 0]) & (ep[:, 1] >= sp[:, 1])
        assert np.array_equal(seg, np.flatnonzero(hit).astype(np.int32)), f"tile {tl}"


@pytest.mark.gpu
@pytest.mark.parametrize("W,H,n,max_half", [(1920, 1080, 200_000, None), (3840, 2160, 300_000, None),
                                            (16383, 4095, 20_000, 40), (5, 3, 40, 3), (2047, 2047, 3_000, 600),
                                            (1279, 719, 30_000, 12)],
                         ids=["1080p-2x8bit", "4k-2x9bit", "2M-tiles-3x7bit", "one-tile-1bit", "big-boxes", "holes"])
def test_radix_binning_is_the_stable_sort_by_tile_gpu(W, H, n, max_half, monkeypatch):
    """The binning of the tile route — a stable radix sort of the Gaussian-major pair list on the tile id — against
    the numpy restatement (oracle/tile_oracle.py: numpy.argsort(kind="stable")), bit for bit: pair offsets, pairs
    per tile, tile offsets and every tile's list, whatever the number of digit passes the tile count asks for
    (1080p: 2 x 8 bits, 4K: 2 x 9, 2 Mi tiles: 3 x 7, one tile: 1 x 1), for boxes of thousands of tiles, and for
    long runs of Gaussians without any pair (the global-memory path of k_view_pairs)."""
    from oracle import tile_oracle as to
    from simplegaussiansplat_tk71_b200 import _lib, compositor, workloads as wl

    monkeypatch.setattr(compositor, "ROUTE", "tiles")
    L = _lib.lib()
    if max_half is None:
        v = wl.splat_view(W, H, n, seed=7, device="cuda")
    else:
        rng = np.random.default_rng(W + H)
        half = rng.integers(1, max_half + 1, (n, 2))
        c = np.stack((rng.integers(-2, W + 3, n), rng.integers(-2, H + 3, n)), 1)
        sp = np.clip(c - half, 0, [W, H]).astype(np.int32)
        ep = np.clip(c + half, 0, [W, H]).astype(np.int32)
        if (W, H) == (1279, 719):
            # 15 000 consecutive Gaussians without any pair (empty boxes): the 4096 pairs of a block of k_view_pairs
            # then span more Gaussians than it stages in shared memory — its global-memory path
            ep[5_000:20_000, 0] = sp[5_000:20_000, 0] - 1
        sig = half.astype(np.float64) / 3.0 + 0.5
        lam = np.zeros((n, 2, 2), np.float32)
        lam[:, 0, 0], lam[:, 1, 1] = 1.0 / sig[:, 0] ** 2, 1.0 / sig[:, 1] ** 2
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
        v = wl.SplatView(name=f"boxes {W}x{H}", boxsize=t(np.maximum(ep - sp + 1, 0).astype(np.int64).prod(1)), startpoint=t(sp),
                         endpoint=t(ep), mean=t(c.astype(np.float32)), lam=t(lam),
                         opacity=t(rng.uniform(0.01, 0.3, (n, 1)).astype(np.float32)),
                         l_d=t(rng.uniform(0.2, 1.0, (n, 3)).astype(np.float32)), width=W, height=H)
    img, view = compositor._render_forward(v.boxsize, v.startpoint, v.endpoint, v.mean.float(), v.lam, v.opacity, v.l_d,
                                           v.width, v.height)
    torch.cuda.synchronize()
    assert torch.isfinite(img).all()
    a = _arena_arrays(L, view, v.width, v.height)
    ntx, nty = to.num_tiles(W, H)
    tiles, gids, toff = to.tile_pairs_np(v.startpoint.cpu().numpy(), v.endpoint.cpu().numpy(), W, H)
    gid_s, start, _ = to.sort_by_tile(tiles, gids, ntx * nty)
    assert view.P == len(tiles) > 0
    assert np.array_equal(a["toff"], toff.astype(np.int32))
    assert np.array_equal(a["tstart"], start)
    assert np.array_equal(a["tcount"], np.diff(start))
    assert np.array_equal(a["pgid"][:view.P], gid_s)


def _both_routes(v, gI, monkeypatch, repeats=1):
    from simplegaussiansplat_tk71_b200 import compositor
    from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F

    res = {}
    for r in ["tiles"] * repeats + ["lists"]:
        monkeypatch.setattr(compositor, "ROUTE", r)
        m, lam, o, l = (v.mean.float().requires_grad_(True), v.lam.clone().requires_grad_(True),
                        v.opacity.clone().requires_grad_(True), v.l_d.clone().requires_grad_(True))
        img = F.apply(v.boxsize, torch.tensor([v.n]), v.startpoint, v.endpoint, m, lam, o, l, v.width, v.height)
        img.backward(gI)
        res.setdefault(r, []).append([img.detach().clone()] + [t.grad.clone() for t in (m, lam, o, l)])
        del img
    return res


def _assert_routes_agree(res, rtol, atol_scale):
    a, b = res["tiles"][0], res["lists"][0]
    for name, x, y in zip(("image", "mean", "lambda", "opacity", "l"), a, b):
        assert torch.isfinite(x).all(), name
        scale = float(y.abs().max())
        assert torch.allclose(x, y, rtol=rtol, atol=atol_scale * scale), (name, float((x - y).abs().max()), scale)


@pytest.mark.gpu
def test_tile_route_equals_list_route_on_a_1080p_scene_gpu(monkeypatch):
    """Both routes compute the same sums in different fp32 orders; the tile route has no float atomics and must
    be bitwise reproducible."""
    from simplegaussiansplat_tk71_b200 import workloads as wl

    v = wl.splat_view(1920, 1080, 200_000, seed=7, device="cuda")
    gI = torch.rand(v.height + 1, v.width + 1, 3, device="cuda") + 0.1
    res = _both_routes(v, gI, monkeypatch, repeats=2)
    _assert_routes_agree(res, rtol=2e-4, atol_scale=2e-6)
    for x, y in zip(*res["tiles"]):
        assert torch.equal(x, y)


@pytest.mark.gpu
def test_tile_route_equals_list_route_on_long_pixel_lists_gpu(monkeypatch):
    """Bundled-scene view (C2): lists of hundreds of elements per pixel, boxes of 10^5 pixels."""
    from simplegaussiansplat_tk71_b200 import workloads as wl

    v = wl.bundled_views("cuda", n_views=3)[2]
    gI = torch.rand(v.height + 1, v.width + 1, 3, device="cuda") + 0.1
    res = _both_routes(v, gI, monkeypatch)
    _assert_routes_agree(res, rtol=2e-3, atol_scale=2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_tile_route_with_32_pair_pieces_gpu(name, monkeypatch):
    """Every tile list longer than 32 pairs cut into pieces walked independently: the carries between the pieces
    (T forward, U backward) must reproduce the reference fixtures, and a larger scene must equal the result with
    the default piece length up to fp32 rounding."""
    from simplegaussiansplat_tk71_b200 import _lib, compositor, workloads as wl

    monkeypatch.setattr(compositor, "ROUTE", "tiles")
    L = _lib.lib()
    default = L.gcp_tile_piece_pairs()
    assert L.gcp_tile_set_piece_pairs(33) != 0          # must be a multiple of 32
    try:
        assert L.gcp_tile_set_piece_pairs(32) == 0
        case = load_case(np.load(FIX), name)
        _check(_run(case, "cuda"), case)
        if name == "dense":
            # a larger scene with long lists (the nearest 20 000 Gaussians of a bundled-scene view): both piece
            # lengths within the tolerance of the fp64 oracle
            v = wl.bundled_views("cpu", n_views=2)[1]
            k = 20000
            big = dict(boxsize=v.boxsize[:k].numpy(), sp=v.startpoint[:k].numpy(), ep=v.endpoint[:k].numpy(),
                       mean=v.mean[:k].numpy(), lam=v.lam[:k].numpy(), opac=v.opacity[:k].numpy(),
                       l_d=v.l_d[:k].numpy(), W=v.width, H=v.height,
                       grad_image=np.random.default_rng(3).uniform(0.1, 1.0, (v.height + 1, v.width + 1, 3)).astype(np.float32))
            ref = _oracle(big)
            for pairs in (32, default):
                assert L.gcp_tile_set_piece_pairs(pairs) == 0
                _check(_run(big, "cuda"), big, fixture=False, oracle=ref)
    finally:
        L.gcp_tile_set_piece_pairs(default)


@pytest.mark.gpu
def test_render_without_gradients_skips_the_kept_T_gpu(route):
    """No input requires a gradient: same image; the tile route then does not store T for a backward."""
    from simplegaussiansplat_tk71_b200 import compositor
    from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F

    case = load_case(np.load(FIX), "dense")
    want = _run(case, "cuda")[0]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    args = (t(case["boxsize"]), torch.tensor([0]), t(case["sp"]), t(case["ep"]), t(case["mean"]).float(),
            t(case["lam"]), t(case["opac"]), t(case["l_d"]), case["W"], case["H"])
    with torch.no_grad():
        img = F.apply(*args)
    assert np.array_equal(img.cpu().numpy(), want) or np.allclose(img.cpu().numpy(), want, rtol=1e-5, atol=1e-6)
    _, view = compositor._render_forward(*args[:1], *args[2:], keep=False)
    if route == "tiles":
        assert view.keep is False
        with pytest.raises(RuntimeError, match="without keeping T"):
            compositor._render_backward(view, torch.zeros(case["H"] + 1, case["W"] + 1, 3, device="cuda"))


@pytest.mark.gpu
@pytest.mark.parametrize("W,H,n,max_half,piece", [
    (17, 9, 300, 4, 32),        # image sizes that are not multiples of the 8x4 tile; every list in several pieces
    (31, 33, 900, 12, 32),
    (64, 3, 500, 40, 64),       # a single tile row, boxes wider than the image
    (7, 7, 200, 2, 128),        # image smaller than two tiles
    (100, 50, 3000, 9, 128),
    (8, 4, 64, 1, 32),          # W+1 = 9, H+1 = 5: the last tile column / row hold one pixel
])
def test_native_compositor_ragged_image_sizes_against_oracle_gpu(W, H, n, max_half, piece, route):
    """Random scenes on awkward image sizes (partial tiles, single rows, degenerate boxes of one pixel column)
    against the fp64 oracle, for both routes and several piece lengths of the tile route."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from oracle import compositor_oracle as co
    from simplegaussiansplat_tk71_b200 import _lib
    import make_compositor_fixture as mk

    boxsize, sp, ep, mean, lam, opac, l_d = mk.make_scene(100 + W, W, H, n, max_half, opaque=n // 20)
    # a few degenerate boxes: one pixel wide / one pixel high, and one covering the whole image
    sp[0], ep[0] = torch.tensor([0, 0]), torch.tensor([W, H])
    ep[1, 0] = sp[1, 0]
    ep[2, 1] = sp[2, 1]
    boxsize = torch.prod((ep - sp + 1).to(torch.int64), dim=1)
    rng = np.random.default_rng(W * 1000 + H)
    gI = rng.uniform(0.1, 1.0, (H + 1, W + 1, 3)).astype(np.float32)
    case = dict(boxsize=boxsize.numpy(), sp=sp.numpy(), ep=ep.numpy(), mean=mean.numpy(), lam=lam.numpy(),
                opac=opac.numpy(), l_d=l_d.numpy(), W=W, H=H, grad_image=gI)
    L = _lib.lib()
    default = L.gcp_tile_piece_pairs()
    try:
        assert L.gcp_tile_set_piece_pairs(piece) == 0
        got = _run(case, "cuda")
    finally:
        L.gcp_tile_set_piece_pairs(default)
    _check(got, case, fixture=False)


@pytest.mark.gpu
def test_one_call_forward_on_a_known_capacity_gpu(monkeypatch):
    """gcp_view_forward (plan + render in one call, no host wait in between) on a pair arena the caller already
    owns: same bits as the exact two-call pass; a capacity that is too small is reported through the totals and
    nothing is rendered."""
    from simplegaussiansplat_tk71_b200 import _lib, compositor, workloads as wl

    monkeypatch.setattr(compositor, "ROUTE", "tiles")
    L = _lib.lib()
    v = wl.splat_view(640, 360, 60_000, seed=5, device="cuda")
    W, H, n = v.width, v.height, v.n
    img_exact, view = compositor._render_forward(v.boxsize, v.startpoint, v.endpoint, v.mean.float(), v.lam, v.opacity,
                                                 v.l_d, W, H)
    torch.cuda.synchronize()
    assert view.P > 10_000
    p = lambda t: t.data_ptr()  # noqa: E731
    mean, lam, opac, l_d = v.mean.float().contiguous(), v.lam.reshape(n, 4).contiguous(), v.opacity.reshape(n).contiguous(), v.l_d
    stream = torch.cuda.current_stream().cuda_stream
    for cap, fits in ((view.P, True), (view.P * 3, True), (view.P - 1, False), (1024, False)):
        plan = torch.empty(int(L.gcp_view_plan_bytes(n, W, H)), dtype=torch.uint8, device="cuda")
        pairs = torch.empty(int(L.gcp_view_pair_bytes(cap, W, H)), dtype=torch.uint8, device="cuda")
        totals = torch.zeros(2, dtype=torch.int64).pin_memory()
        img = torch.full((H + 1, W + 1, 3), float("nan"), device="cuda")
        _lib.check(L.gcp_view_forward(p(v.startpoint), p(v.endpoint), p(mean), p(lam), p(opac), p(l_d), n, W, H, p(plan),
                                      plan.numel(), p(pairs), pairs.numel(), cap, 1, p(img), p(totals), stream), "fwd")
        torch.cuda.synchronize()
        assert int(totals[0]) == view.P
        if fits:
            assert torch.equal(img, img_exact)
        else:
            assert bool(torch.isnan(img).all()), "an overflowing view must not be rendered at all"


@pytest.mark.gpu
def test_speculative_forward_equals_the_synchronous_one_gpu(monkeypatch):
    """compositor.SPECULATE: plan and render queued in one call on a pooled arena's capacity, the pair count read from
    pinned memory afterwards.  Same bits as the plan / wait / render sequence — also when the pooled arena is too
    small and the view has to be rendered again (the first attempt must leave nothing behind)."""
    from simplegaussiansplat_tk71_b200 import compositor, workloads as wl
    from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F

    monkeypatch.setattr(compositor, "ROUTE", "tiles")
    small = wl.splat_view(640, 360, 5_000, seed=11, device="cuda")
    big = wl.splat_view(640, 360, 80_000, seed=12, device="cuda")
    gI = torch.rand(361, 641, 3, device="cuda") + 0.1

    def run(v):
        m, lam, o, l = (v.mean.float().requires_grad_(True), v.lam.clone().requires_grad_(True),
                        v.opacity.clone().requires_grad_(True), v.l_d.clone().requires_grad_(True))
        img = F.apply(v.boxsize, torch.tensor([v.n]), v.startpoint, v.endpoint, m, lam, o, l, v.width, v.height)
        img.backward(gI)
        torch.cuda.synchronize()
        return [img.detach().clone()] + [t.grad.clone() for t in (m, lam, o, l)]

    monkeypatch.setattr(compositor, "SPECULATE", False)
    want_small, want_big = run(small), run(big)
    monkeypatch.setattr(compositor, "SPECULATE", True)
    compositor._free_pair.clear()
    compositor._free_plan.clear()
    run(small)                                   # leaves a small pair arena in the pool
    key = compositor._pool_key(torch.device("cuda", torch.cuda.current_device()))
    assert compositor._free_pair.get(key), "the small view's arena should be back in the pool"
    cap_small = max(a.cap for a in compositor._free_pair[key])
    got_big = run(big)                           # speculates on it, overflows, is rendered again
    assert max(a.cap for a in compositor._free_pair[key]) > cap_small
    got_big2 = run(big)                          # now fits: the one-call path
    got_small = run(small)                       # a much larger arena than needed
    for got, want in ((got_big, want_big), (got_big2, want_big), (got_small, want_small)):
        for a, b in zip(got, want):
            assert torch.equal(a, b)


@pytest.mark.gpu
def test_native_compositor_1080p_200k_view_against_oracle_gpu(route):
    """A 1920x1080 view of 200 000 Gaussians (7.5 M elements, the largest the fp64 oracle finishes in seconds) at
    the north star's tolerance, both routes."""
    from simplegaussiansplat_tk71_b200 import workloads as wl

    sc = wl.splat_view(1920, 1080, 200_000, seed=1080, device="cpu")
    rng = np.random.default_rng(9)
    case = dict(boxsize=sc.boxsize.numpy(), sp=sc.startpoint.numpy(), ep=sc.endpoint.numpy(), mean=sc.mean.numpy(),
                lam=sc.lam.numpy(), opac=sc.opacity.numpy(), l_d=sc.l_d.numpy(), W=sc.width, H=sc.height,
                grad_image=rng.uniform(0.1, 1.0, (sc.height + 1, sc.width + 1, 3)).astype(np.float32))
    _check(_run(case, "cuda"), case, fixture=False)


@pytest.mark.gpu
def test_backward_is_linear_in_grad_image_at_full_size_gpu(monkeypatch):
    """Size-independent property at BASELINE.json's full splat size (1080p, 1 M Gaussians, 37 M elements): the
    backward is linear in dL/dimage — grads(a gI1 + b gI2) = a grads(gI1) + b grads(gI2) — and two runs of the same
    step are bitwise equal (no float atomics on the tile route)."""
    from simplegaussiansplat_tk71_b200 import compositor, workloads as wl
    from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F

    monkeypatch.setattr(compositor, "ROUTE", "tiles")
    v = wl.splat_view_device(1920, 1080, 1_000_000, seed=1080, device="cuda")
    g1 = torch.rand(v.height + 1, v.width + 1, 3, device="cuda")
    g2 = torch.rand(v.height + 1, v.width + 1, 3, device="cuda") - 0.5

    def grads(gI):
        m, lam, o, l = (v.mean.float().requires_grad_(True), v.lam.clone().requires_grad_(True),
                        v.opacity.clone().requires_grad_(True), v.l_d.clone().requires_grad_(True))
        img = F.apply(v.boxsize, torch.tensor([v.n]), v.startpoint, v.endpoint, m, lam, o, l, v.width, v.height)
        img.backward(gI)
        return [img.detach()] + [t.grad for t in (m, lam, o, l)]

    a, b = 0.75, -1.5
    r1, r2, r12, r12b = grads(g1), grads(g2), grads(a * g1 + b * g2), grads(a * g1 + b * g2)
    assert torch.isfinite(r12[0]).all() and float(r12[0].max()) > 0
    for x, y in zip(r12, r12b):
        assert torch.equal(x, y)
    # |g1| and |g2| bound the terms: the scale of the comparison is grads(|a| |g1| + |b| |g2|) >= sum of |terms|
    rs = grads(abs(a) * g1.abs() + abs(b) * g2.abs())
    for name, x1, x2, x12, s in zip(("mean", "lambda", "opacity", "l"), r1[1:], r2[1:], r12[1:], rs[1:]):
        want = a * x1.double() + b * x2.double()
        # the scale run has cancellation of its own (mean / lambda terms change sign with d): use the array's norm
        # per Gaussian as a floor so that the bound is never tighter than fp32 rounding of the terms
        bound = 1e-6 + 1e-4 * torch.maximum(s.double().abs(), want.abs())
        err = (x12.double() - want).abs()
        frac_bad = float((err > bound).double().mean())
        assert frac_bad < 1e-3, (name, frac_bad, float((err / bound).max()))


@pytest.mark.gpu
def test_speculative_forward_in_an_unsynchronised_render_loop_gpu(monkeypatch):
    """A no_grad render loop never synchronises and drops every view's arenas at once, so the host runs views ahead
    of the device and the pooled arenas (with their pinned pair-count word) are reused while earlier plan kernels
    are still queued: every image must still be the one of its own view (sizes alternate, so a count read from the
    wrong view would overflow or mis-size the arena)."""
    from simplegaussiansplat_tk71_b200 import compositor, workloads as wl
    from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F

    monkeypatch.setattr(compositor, "ROUTE", "tiles")
    views = [wl.splat_view_device(640, 360, n, seed=200 + i, device="cuda")
             for i, n in enumerate((3_000, 90_000, 6_000, 120_000, 2_000, 60_000))]

    def render(v):
        with torch.no_grad():
            return F.apply(v.boxsize, torch.tensor([v.n]), v.startpoint, v.endpoint, v.mean.float(), v.lam, v.opacity,
                           v.l_d, v.width, v.height)

    monkeypatch.setattr(compositor, "SPECULATE", False)
    want = []
    for v in views:
        want.append(render(v).clone())
        torch.cuda.synchronize()
    monkeypatch.setattr(compositor, "SPECULATE", True)
    compositor._free_pair.clear()
    compositor._free_plan.clear()
    for rep in range(3):
        got = [render(v) for v in views]            # no synchronisation in between
        torch.cuda.synchronize()
        for a, b in zip(got, want):
            assert torch.equal(a, b), rep


def _deep_view(W=2047, H=2047, n=535_000, half=32, seed=231):
    """A view of MORE THAN 2**31 elements: n boxes of (2 half + 1)^2 pixels, all inside the image, ~520 Gaussians
    over every pixel, opacities small enough that T is still ~0.2 at the end of a list."""
    rng = np.random.default_rng(seed)
    cx = rng.integers(half, W - half + 1, n)
    cy = rng.integers(half, H - half + 1, n)
    sp = np.stack((cx - half, cy - half), 1).astype(np.int32)
    ep = np.stack((cx + half, cy + half), 1).astype(np.int32)
    mean = (np.stack((cx, cy), 1) + rng.uniform(-0.5, 0.5, (n, 2))).astype(np.float32)
    s0, s1 = rng.uniform(8.0, 16.0, n), rng.uniform(8.0, 16.0, n)
    rho = rng.uniform(-0.4, 0.4, n)
    det = (s0 * s1) ** 2 * (1 - rho ** 2)
    lam = np.stack((s1 ** 2 / det, -rho * s0 * s1 / det, -rho * s0 * s1 / det, s0 ** 2 / det), 1).reshape(n, 2, 2)
    boxsize = np.full(n, (2 * half + 1) ** 2, dtype=np.int64)
    return dict(boxsize=boxsize, sp=sp, ep=ep, mean=mean, lam=lam.astype(np.float32),
                opac=rng.uniform(0.002, 0.02, (n, 1)).astype(np.float32),
                l_d=rng.uniform(0.2, 1.0, (n, 3)).astype(np.float32), W=W, H=H)


def _sub_scene(case, j):
    """Everything the outputs of Gaussian j depend on: the Gaussians whose boxes meet box j, in depth order, with
    their boxes clipped to box j (the pixel lists inside box j are complete, nothing outside it matters to j)."""
    sp, ep = case["sp"], case["ep"]
    hit = np.flatnonzero((sp[:, 0] <= ep[j, 0]) & (ep[:, 0] >= sp[j, 0]) & (sp[:, 1] <= ep[j, 1]) & (ep[:, 1] >= sp[j, 1]))
    sps, eps = np.maximum(sp[hit], sp[j]), np.minimum(ep[hit], ep[j])
    sub = dict(boxsize=((eps[:, 0] - sps[:, 0] + 1).astype(np.int64) * (eps[:, 1] - sps[:, 1] + 1)), sp=sps, ep=eps,
               mean=case["mean"][hit], lam=case["lam"][hit], opac=case["opac"][hit], l_d=case["l_d"][hit],
               W=case["W"], H=case["H"], grad_image=case["grad_image"])
    return sub, int(np.searchsorted(hit, j))


@pytest.mark.gpu
def test_a_view_of_more_than_2_31_elements_is_one_pass_gpu(monkeypatch):
    """Row a6 (gs_model.py:428, :582-594, :611-615, :726-730): the reference cuts a view into chunks of 2**29 elements
    and carries T per pixel between them; the tile route never forms elements, so a view of 2.26e9 elements (more
    than any int32-indexed op could take) is ONE pass.  Checked at the north star's tolerance against the fp64
    oracle on complete sub-scenes: for three Gaussians (front, middle, back of the depth order) the image over the
    Gaussian's box and its four gradients depend only on the Gaussians meeting that box — the oracle evaluates
    exactly those.  The list route (int32 element indices, like the reference's ops) refuses the view."""
    from simplegaussiansplat_tk71_b200 import compositor
    from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F

    case = _deep_view()
    n = case["boxsize"].size
    assert int(case["boxsize"].sum()) > 2 ** 31
    rng = np.random.default_rng(5)
    case["grad_image"] = rng.uniform(0.1, 1.0, (case["H"] + 1, case["W"] + 1, 3)).astype(np.float32)
    monkeypatch.setattr(compositor, "ROUTE", "tiles")
    img, g_mean, g_lam, g_opac, g_l = _run(case, "cuda")
    assert np.isfinite(img).all() and img.min() >= 0.0 and img.max() > 0.5
    for j in (3, n // 2, n - 2):
        sub, k = _sub_scene(case, j)
        (ref_img, ref_m, ref_L, ref_o, ref_l), (sc_img, sc_m, sc_L, sc_o, sc_l) = _oracle(sub)
        (x0, y0), (x1, y1) = case["sp"][j], case["ep"][j]
        box = np.s_[y0:y1 + 1, x0:x1 + 1]
        for name, a, b, sc in (("image", img[box], ref_img[box], sc_img[box]),
                               ("grad_mean", g_mean[j], ref_m[k], sc_m[k]), ("grad_lambda", g_lam[j], ref_L[k], sc_L[k]),
                               ("grad_opacity", g_opac[j], ref_o[k], sc_o[k]), ("grad_l", g_l[j], ref_l[k], sc_l[k])):
            a = np.asarray(a, np.float64).reshape(b.shape)
            bound = ATOL + RTOL * np.asarray(sc).reshape(b.shape)
            err = np.abs(a - b)
            assert np.all(err <= bound), (j, name, float((err / bound).max()), float(err.max()))
    monkeypatch.setattr(compositor, "ROUTE", "lists")
    with pytest.raises(RuntimeError, match="2\\*\\*31"):
        _run(case, "cuda")
