"""Row a10 (gs_model.py:402-449): per-view culling/clamping and the chunker, on CPU."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split_by_cumsum_parallel_matches_reference_golden():
    """tests/golden/split_by_cumsum.json was produced by the reference's Utilities.split_by_cumsum_parallel
    (uitility.py:478-488) — see tests/golden/make_split_fixture.py."""
    import json

    from simplegaussiansplat_tk71_b200 import views

    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "split_by_cumsum.json")))
    assert len(cases) >= 4
    for c in cases:
        got = views.split_by_cumsum_parallel(torch.tensor(c["x"], dtype=torch.float32), c["limit"])
        assert got.tolist() == c["counts"], c


def test_visible_boxes_cull_and_inclusive_clamp():
    from simplegaussiansplat_tk71_b200 import views

    W, H = 10, 8
    mean = torch.tensor([[5, 4], [0, 0], [12, 4], [5, 4], [5, 4], [-3, 4]], dtype=torch.int32)
    half = torch.tensor([[2, 1], [3, 3], [3, 1], [0, 2], [2, 1], [3, 1]], dtype=torch.int32)
    z = torch.tensor([1.0, 2.0, 3.0, 4.0, -1.0, 1.0])
    mask, sp, ep, bs = views.visible_boxes(mean, half, z, W, H)
    # 0 visible; 1 visible (touches the corner); 2 visible (12-3 < 10); 3 zero box; 4 behind the camera; 5 outside (-3+3 > 0 fails)
    assert mask.tolist() == [True, True, True, False, False, False]
    assert sp.tolist() == [[3, 3], [0, 0], [9, 3]]
    assert ep.tolist() == [[7, 5], [3, 3], [10, 5]]          # clamped to the INCLUSIVE W, H
    assert bs.tolist() == [15, 16, 6]
    assert views.chunk_ends(bs).tolist() == [3]               # far below 2**29 elements: one chunk


VIEW_FIX = os.path.join(ROOT, "tests", "golden", "view_loop_fixture.npz")


@pytest.mark.parametrize("tag", ["a", "b"])
def test_view_preparation_matches_the_reference_loop(tag):
    """tests/golden/view_loop_fixture.npz holds what the reference's OWN per-view loop (gs_model.py:399-454,
    executed from the reference file by tests/golden/make_view_loop_fixture.py) handed to the compositor for every
    view: the integer side — visibility mask, corners, box sizes, chunk ends — must be reproduced bit for bit."""
    from simplegaussiansplat_tk71_b200 import views

    f = np.load(VIEW_FIX)
    W, H = (int(v) for v in f[f"{tag}/WH"])
    t = lambda k: torch.from_numpy(f[f"{tag}/in/{k}"])  # noqa: E731
    mean_pixel, box, z = t("mean_pixel"), t("box"), t("mean_camera")[:, :, 2].clone()
    empty = int(f[f"{tag}/empty_view"][0])
    if empty >= 0:
        z[empty] = -1.0
    call = 0
    kept = []
    for v in range(mean_pixel.shape[0]):
        mask, sp, ep, boxsize = views.visible_boxes(mean_pixel[v], box[v], z[v], W, H)
        if sp.shape[0] == 0:
            continue                                   # gs_model.py:414-417
        kept.append(v)
        g = lambda k: f[f"{tag}/call{call}/{k}"]  # noqa: E731
        assert np.array_equal(sp.numpy(), g("sp")) and np.array_equal(ep.numpy(), g("ep"))
        assert np.array_equal(boxsize.numpy(), g("boxsize"))
        assert np.array_equal(views.chunk_ends(boxsize).numpy(), g("batch"))
        assert np.array_equal(mean_pixel[v][mask].numpy(), g("mean"))
        assert np.array_equal(t("lam")[v][mask].numpy(), g("lam"))
        assert np.array_equal(t("opac")[v][mask].numpy(), g("opac"))
        call += 1
    assert call == int(f[f"{tag}/n_calls"][0])
    assert kept == f[f"{tag}/kept_samples"].tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_render_views_reproduces_the_reference_image_batch_gpu(tag):
    """views.render_views (cull + clamp + chunker + native compositor + the reference's final reshape, :454) against
    the image batch the reference's own loop and Function produced for the same inputs (fp32, CPU): within twice the
    north star's tolerance (both are fp32 evaluations of the same sums; scale = the images themselves, all terms
    are positive)."""
    from simplegaussiansplat_tk71_b200 import views

    f = np.load(VIEW_FIX)
    W, H = (int(v) for v in f[f"{tag}/WH"])
    t = lambda k: torch.from_numpy(f[f"{tag}/in/{k}"]).cuda()  # noqa: E731
    z = t("mean_camera")[:, :, 2].clone()
    empty = int(f[f"{tag}/empty_view"][0])
    if empty >= 0:
        z[empty] = -1.0
    imgs = views.render_views(t("mean_pixel"), t("box"), z, t("lam"), t("opac"), t("l_d"), W, H)
    want = f[f"{tag}/images"]
    assert tuple(imgs.shape) == want.shape
    err = np.abs(imgs.cpu().numpy().astype(np.float64) - want)
    assert np.all(err <= 2 * (1e-6 + 1e-5 * np.abs(want))), float((err / (1e-6 + 1e-5 * np.abs(want))).max())


@pytest.mark.gpu
def test_render_views_shapes_and_skip_empty_gpu():
    from simplegaussiansplat_tk71_b200 import views

    dev = "cuda"
    W, H, n = 32, 24, 50
    g = torch.Generator().manual_seed(0)
    mean = torch.stack((torch.randint(0, W, (2, n), generator=g), torch.randint(0, H, (2, n), generator=g)), 2).int()
    half = torch.randint(1, 4, (2, n, 2), generator=g).int()
    z = torch.rand(2, n, generator=g) + 0.1
    z[1] = -1.0                                               # second view: nothing in front of the camera
    lam = torch.eye(2).repeat(2, n, 1, 1) * 0.3
    o = torch.full((2, n, 1), 0.5)
    l = torch.rand(2, n, 3, generator=g)
    imgs = views.render_views(mean.to(dev), half.to(dev), z.to(dev), lam.to(dev), o.to(dev), l.to(dev), W, H)
    assert imgs.shape == (1, 3, H, W) and float(imgs.sum()) > 0


def test_compositor_passes_ready_tensors_through_and_converts_the_rest():
    """Host logic of the compositor's input preparation: tensors that are already fp32 / int32 and contiguous go to the
    kernels as they are (no copy, no autograd node), everything else is converted; box corners end up 8-byte aligned."""
    import torch

    from simplegaussiansplat_tk71_b200 import compositor as c

    a = torch.arange(12, dtype=torch.float32).reshape(6, 2).requires_grad_(True)
    assert c._f32(a, (6, 2)) is a
    b = c._f32(a.double(), (12,))
    assert b.dtype == torch.float32 and b.shape == (12,) and not b.requires_grad and b.is_contiguous()
    t = c._f32(a.detach().t(), (2, 6))                       # not contiguous: copied
    assert t.is_contiguous() and torch.equal(t, a.detach().t())
    p = torch.arange(16, dtype=torch.int32).reshape(8, 2)
    assert c._i32_pairs(p) is p
    q = c._i32_pairs(p.long())
    assert q.dtype == torch.int32 and torch.equal(q, p) and q.data_ptr() % 8 == 0
    odd = torch.arange(17, dtype=torch.int32)[1:].reshape(8, 2)   # 4-byte offset into its storage
    r = c._i32_pairs(odd)
    assert r.data_ptr() % 8 == 0 and torch.equal(r, odd)
