"""CPU: the compositor oracle (oracle/compositor_oracle.py) against fixtures produced by the reference's own
`custom_autograd_grouped_cumprod` (gs_model.py:477-820) run on CPU — tests/golden/make_compositor_fixture.py."""
import os

import numpy as np
import pytest

FIX = os.path.join(os.path.dirname(__file__), "golden", "compositor_fixture.npz")
CASES = ["small", "dense", "wide", "opaque"]


def load_case(f, name):
    g = lambda k: f[f"{name}/{k}"]  # noqa: E731
    W, H = (int(v) for v in g("WH"))
    return dict(boxsize=g("boxsize"), sp=g("startpoint"), ep=g("endpoint"), mean=g("mean"), lam=g("lam"),
                opac=g("opacity"), l_d=g("l_d"), W=W, H=H, image=g("image"), grad_image=g("grad_image"),
                grad_mean=g("grad_mean"), grad_lambda=g("grad_lambda"), grad_opacity=g("grad_opacity"),
                grad_l=g("grad_l"))


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_function(name):
    """The reference's own fp32 Function (the fixture) lies within the north star's tolerance of the fp64 oracle,
    |fixture - fp64| <= 1e-6 + 1e-5 * scale with scale = sum of |terms| (worst measured ratio 0.16): the oracle
    is pinned, and the same bound is what the native compositor is held to (tests/test_compositor.py)."""
    from oracle import compositor_oracle as co

    c = load_case(np.load(FIX), name)
    img, cache = co.forward(c["boxsize"], c["sp"], c["ep"], c["mean"], c["lam"], c["opac"], c["l_d"], c["W"], c["H"])
    assert img.shape == c["image"].shape == (c["H"] + 1, c["W"] + 1, 3)
    grads, sc = co.backward(cache, c["grad_image"], scales=True)
    ref = (img,) + tuple(grads)
    scales = (co.image_scale(cache, c["W"], c["H"]),) + tuple(sc)
    fix = (c["image"], c["grad_mean"], c["grad_lambda"], c["grad_opacity"], c["grad_l"])
    for out, e, plain, cond in co.error_table(fix, ref, scales):
        assert cond <= 1.0, (out, e, plain, cond)
    # without cancellation the scale IS |ref|: the image meets the plain form of the tolerance too
    assert co.error_table(fix, ref, scales)[0][2] <= 1.0


def test_chunked_reference_render_differs_only_by_its_boundary_carry():
    """with several chunks the reference drops one (1-alpha) factor at each chunk boundary.
    The single-chunk image is the intended result; the fixture keeps the 3-chunk image to document the size
    of that effect (it is NOT what the native compositor reproduces)."""
    f = np.load(FIX)
    one, three = f["dense/image"], f["dense/image_3chunks"]
    assert one.shape == three.shape
    assert np.all(three >= one - 1e-4)          # a missing (1-alpha) <= 1 factor can only brighten
    assert np.abs(three - one).max() > 1e-3      # and it is visible


@pytest.mark.parametrize("name", ["small", "wide", "dense"])
def test_tile_binning_visits_every_pixel_list_in_reference_order(name):
    """The fused route bins (tile, Gaussian) pairs instead of sorting elements.  Walking a tile's pairs in their
    stable order must give, for every pixel, exactly the Gaussian sequence of the reference's sorted element list
    (expansion + stable sort by pixel key, gs_model.py:538-548) — nothing missing, nothing twice, same order."""
    from oracle import compositor_oracle as co
    from oracle import tile_oracle as to

    c = load_case(np.load(FIX), name)
    gid, px, py = co.expand(c["boxsize"], c["sp"], c["ep"])
    key = (py * 10000 + px).astype(np.int64)
    order = np.argsort(key, kind="stable")
    ref = {}
    for k, g in zip(key[order].tolist(), gid[order].tolist()):
        ref.setdefault(k, []).append(g)
    got = to.pixel_lists_from_pairs(c["sp"], c["ep"], c["W"], c["H"])
    assert got == ref
    # and the piece plan covers every tile's list exactly once
    ntx, nty = to.num_tiles(c["W"], c["H"])
    tiles, gids, toff = to.tile_pairs(c["sp"], c["ep"], c["W"], c["H"])
    assert toff[-1] == len(tiles)
    _, start, _ = to.sort_by_tile(tiles, gids, ntx * nty)
    for piece in (32, 128):
        pstart, ptile = to.piece_plan(start, piece)
        assert pstart[-1] == len(ptile) and np.all(np.diff(pstart) >= 1)
        covered = 0
        for p, t in enumerate(ptile.tolist()):
            k = p - pstart[t]
            lo = start[t] + k * piece
            hi = min(start[t + 1], lo + piece)
            assert lo <= hi and (hi > lo or start[t] == start[t + 1])
            covered += hi - lo
        assert covered == len(tiles)


def test_tile_binning_random_boxes_and_image_sizes():
    """Random image sizes (partial tiles, single rows / columns) and random clamped boxes, including one-pixel and
    whole-image ones: the tile walk order equals the stable pixel sort for every pixel."""
    from oracle import compositor_oracle as co
    from oracle import tile_oracle as to

    rng = np.random.default_rng(2024)
    for trial in range(60):
        W, H = int(rng.integers(0, 41)), int(rng.integers(0, 23))
        n = int(rng.integers(1, 40))
        x0 = rng.integers(0, W + 1, n)
        y0 = rng.integers(0, H + 1, n)
        x1 = np.minimum(W, x0 + rng.integers(0, W + 2, n) * (rng.uniform(size=n) < 0.7))
        y1 = np.minimum(H, y0 + rng.integers(0, H + 2, n) * (rng.uniform(size=n) < 0.7))
        sp = np.stack([x0, y0], 1).astype(np.int32)
        ep = np.stack([x1, y1], 1).astype(np.int32)
        if trial % 7 == 0:
            sp[0], ep[0] = (0, 0), (W, H)
        boxsize = ((ep[:, 0] - sp[:, 0] + 1) * (ep[:, 1] - sp[:, 1] + 1)).astype(np.int64)
        gid, px, py = co.expand(boxsize, sp, ep)
        key = (py * 10000 + px).astype(np.int64)
        order = np.argsort(key, kind="stable")
        ref = {}
        for k, g in zip(key[order].tolist(), gid[order].tolist()):
            ref.setdefault(k, []).append(g)
        assert to.pixel_lists_from_pairs(sp, ep, W, H) == ref, (trial, W, H)
        tiles, gids, toff = to.tile_pairs(sp, ep, W, H)
        ntx, nty = to.num_tiles(W, H)
        assert tiles.size == 0 or (tiles.min() >= 0 and tiles.max() < ntx * nty)
        assert np.array_equal(np.diff(toff), np.bincount(gids, minlength=n))
        # the loop-free restatement used for large views (with some empty boxes thrown in): the same arrays
        ep2 = ep.copy()
        ep2[::5, 0] = sp[::5, 0] - 1
        for e in (ep, ep2):
            for a, b in zip(to.tile_pairs(sp, e, W, H), to.tile_pairs_np(sp, e, W, H)):
                assert a.dtype == b.dtype and np.array_equal(a, b)
