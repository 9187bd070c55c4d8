"""Row a10 (gs_model.py:402-449): per-view culling/clamping and the chunker, on CPU."""
import os
import sys

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
