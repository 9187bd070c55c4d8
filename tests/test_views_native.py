"""gcp_views_step (include/gcp_abi.h): a whole batch of views — render, loss gradient, backward, the views' gradients
scatter-added into the parameters' gradient arrays — in one native call, against the same batch driven view by view
through the drop-in Function `custom_autograd_grouped_cumprod` (the per-view loop of gs_model.py:402-449 followed by
what autograd does with the reference's `param[mask]` selections)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scene(n_views, W, H, n, seed0=40):
    from simplegaussiansplat_tk71_b200 import workloads as wl

    return [wl.splat_view_device(W, H, n, seed=seed0 + v, device="cuda") for v in range(n_views)]


def _by_function(views, W, H, n_param, gIs=None, targets=None):
    """View after view through the autograd Function; gradients index_add-ed into parameter-sized arrays."""
    from simplegaussiansplat_tk71_b200 import compositor
    from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F

    compositor.ROUTE = "tiles"
    dev = views[0].startpoint.device
    g = [torch.zeros(n_param, 2, device=dev), torch.zeros(n_param, 4, device=dev), torch.zeros(n_param, device=dev),
         torch.zeros(n_param, 3, device=dev)]
    loss = 0.0
    imgs = []
    for v, sc in enumerate(views):
        m, lam, o, l = (sc.mean.float().requires_grad_(True), sc.lam.clone().requires_grad_(True),
                        sc.opacity.clone().requires_grad_(True), sc.l_d.clone().requires_grad_(True))
        img = F.apply(sc.boxsize, torch.tensor([sc.n]), sc.startpoint, sc.endpoint, m, lam, o, l, W, H)
        if targets is not None:
            d = img - targets[v]
            loss += float((d * d).mean())
            img.backward(d.detach() * (2.0 / d.numel()))
        else:
            img.backward(gIs[v])
        idx = sc.index.long()
        g[0].index_add_(0, idx, m.grad)
        g[1].index_add_(0, idx, lam.grad.reshape(-1, 4))
        g[2].index_add_(0, idx, o.grad.reshape(-1))
        g[3].index_add_(0, idx, l.grad)
        imgs.append(img.detach())
    return g, loss, imgs


@pytest.mark.parametrize("lanes", [1, 2, 3, 4])
def test_native_batch_equals_the_view_by_view_function_gpu(lanes):
    from simplegaussiansplat_tk71_b200.views import NativeViewBatch

    W, H, n_param = 320, 200, 20_000
    views = _scene(5, W, H, n_param)
    gIs = [torch.rand(H + 1, W + 1, 3, device="cuda") + 0.1 for _ in views]
    want, _, imgs = _by_function(views, W, H, n_param, gIs=gIs)
    batch = NativeViewBatch(views, W, H, grad_images=gIs, lanes=lanes, keep_images=True)
    got = [torch.zeros_like(t) for t in want]
    for rep in range(3):                      # repeated steps reuse arenas, events and streams
        for t in got:
            t.zero_()
        batch.step(got[0], got[1], got[2], got[3])
        torch.cuda.synchronize()
        assert batch.finish()
        # same kernels, same order of every sum: bit for bit
        for a, b, name in zip(got, want, ("mean", "lambda", "opacity", "l")):
            assert torch.equal(a, b), (name, lanes, rep, float((a - b).abs().max()))
        for a, b in zip(batch.images, imgs):
            assert torch.equal(a, b)
    assert batch.launches > 0


def test_native_batch_mse_loss_and_overflow_retry_gpu():
    from simplegaussiansplat_tk71_b200.views import NativeViewBatch

    W, H, n_param = 256, 144, 12_000
    views = _scene(4, W, H, n_param, seed0=70)
    targets = [torch.rand(H + 1, W + 1, 3, device="cuda") for _ in views]
    want, want_loss, _ = _by_function(views, W, H, n_param, targets=targets)
    batch = NativeViewBatch(views, W, H, targets=targets, lanes=2)
    got = [torch.zeros_like(t) for t in want]
    loss = torch.zeros(1, device="cuda")
    # a pair capacity that is too small for every view: all of them are skipped, nothing is added, finish() says so
    batch._arenas(64)
    batch.step(got[0], got[1], got[2], got[3], loss)
    torch.cuda.synchronize()
    assert not batch.finish(), "an overflowing step must be reported"
    assert all(float(t.abs().sum()) == 0.0 for t in got) and float(loss) == 0.0
    assert batch.cap >= int(batch.totals_np[:4].max())
    batch.step(got[0], got[1], got[2], got[3], loss)
    torch.cuda.synchronize()
    assert batch.finish()
    assert float(loss) == pytest.approx(want_loss, rel=1e-5)
    for a, b, name in zip(got, want, ("mean", "lambda", "opacity", "l")):
        scale = float(b.abs().max())
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6 * max(scale, 1.0)), (name, float((a - b).abs().max()), scale)


def test_native_batch_with_an_empty_view_gpu():
    from simplegaussiansplat_tk71_b200 import workloads as wl
    from simplegaussiansplat_tk71_b200.views import NativeViewBatch

    W, H, n_param = 64, 40, 500
    views = _scene(3, W, H, n_param, seed0=90)
    e = views[1]
    views[1] = wl.SplatView("empty", e.boxsize[:0], e.startpoint[:0], e.endpoint[:0], e.mean[:0], e.lam[:0], e.opacity[:0],
                            e.l_d[:0], W, H, e.index[:0])
    gIs = [torch.rand(H + 1, W + 1, 3, device="cuda") for _ in views]
    want, _, _ = _by_function([views[0], views[2]], W, H, n_param, gIs=[gIs[0], gIs[2]])
    batch = NativeViewBatch(views, W, H, grad_images=gIs, lanes=2, keep_images=True)
    got = [torch.zeros_like(t) for t in want]
    batch.step(got[0], got[1], got[2], got[3])
    torch.cuda.synchronize()
    assert batch.finish()
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    assert float(batch.images[1].abs().sum()) == 0.0     # a view without Gaussians is a black image


def test_tail_split_and_rank_shards_sum_to_the_single_batch_gpu():
    """(a) gcp_views_step_split: the views in front of the tail add into the main arrays, the tail into its own, the
    event in between fires; main + tail equals the unsplit batch up to the order of the final addition.
    (b) C5 parity (BASELINE.json configs[4]): the buckets of two 'ranks' (views 0,2,4 / 1,3,5: views_for_rank) summed
    — what the all-reduce computes — equal the bucket one rank builds from all six views, to fp32 rounding."""
    from simplegaussiansplat_tk71_b200.views import NativeViewBatch, views_for_rank

    W, H, n_param = 320, 200, 20_000
    views = _scene(6, W, H, n_param, seed0=120)
    gIs = [torch.rand(H + 1, W + 1, 3, device="cuda") + 0.1 for _ in views]
    shapes = ((n_param, 2), (n_param, 4), (n_param,), (n_param, 3))
    zeros = lambda: [torch.zeros(s, device="cuda") for s in shapes]  # noqa: E731

    whole = zeros()
    b = NativeViewBatch(views, W, H, grad_images=gIs, lanes=3)
    b.step(*whole)
    torch.cuda.synchronize()
    assert b.finish()

    main, tail = zeros(), zeros()
    ev = torch.cuda.Event()
    ev.record()
    side = torch.cuda.Stream()
    b.step(*main, tail=(4, tail, ev))
    side.wait_event(ev)
    with torch.cuda.stream(side):
        snapshot = [t.clone() for t in main]        # what an all-reduce started on the event would read
    torch.cuda.synchronize()
    assert b.finish()
    head = zeros()
    b4 = NativeViewBatch(views[:4], W, H, grad_images=gIs[:4], lanes=3)
    b4.step(*head)
    torch.cuda.synchronize()
    for a, s_, h in zip(main, snapshot, head):
        assert torch.equal(a, h) and torch.equal(s_, h)        # complete at the event, untouched by the tail
    for w, a, t in zip(whole, main, tail):
        assert float(t.abs().max()) > 0
        scale = float(w.abs().max())
        assert torch.allclose(a + t, w, rtol=1e-5, atol=1e-6 * max(scale, 1.0))

    ranks = []
    for r in range(2):
        mine = views_for_rank(len(views), r, 2)
        acc = zeros()
        br = NativeViewBatch([views[i] for i in mine], W, H, grad_images=[gIs[i] for i in mine], lanes=2)
        br.step(*acc)
        torch.cuda.synchronize()
        assert br.finish()
        ranks.append(acc)
    for w, a, c in zip(whole, ranks[0], ranks[1]):
        scale = float(w.abs().max())
        assert torch.allclose(a + c, w, rtol=1e-5, atol=1e-6 * max(scale, 1.0))
