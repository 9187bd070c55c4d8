"""B200-native compositor behind the reference's autograd contract.

Mirror of `custom_autograd_grouped_cumprod` (/root/reference/gs_model.py:477-820):

    image = custom_autograd_grouped_cumprod.apply(boxsize, batch, startpoint, endpoint, mean,
                                                  variance_inverse, opacity, l_d, image_width, image_height)

same ten positional inputs (gs_model.py:449, :666), same `(H+1, W+1, 3)` output (:505), gradients for
`mean, variance_inverse, opacity, l_d` only (:820).  What is computed is the reference's result for a view that
fits one chunk (sum(boxsize) <= 2**29, gs_model.py:428):

    T_i = prod_{j<i, same pixel} (1 - alpha_j),   alpha = opacity * exp(-1/2 (r-m) Lambda (r-m)^T)
    image[y, x] = sum_i T_i alpha_i l_i           (elements whose inclusive product is 0 contribute nothing, :575)

but not how.  Two routes, `ROUTE` below, both through the C ABI (include/gcp_abi.h) and both checked against the
fixtures the reference's own Function produced (tests/test_compositor.py):

  "tiles" (default, csrc/gcp_tile.cu) — the fused route: the per-pixel scan is evaluated one pixel per lane, the
  running T in a register, without materialising the element lists.  Two or three C-ABI calls per view, all of
  them working inside two pooled arenas (nothing is allocated per view but the image and the four gradients):
    forward   gcp_view_plan      (tile, Gaussian) pair counts per Gaussian and per 8x4-pixel tile, their offsets, the
                                 pair total (the one number the host needs: it sizes the pair arena)
              gcp_view_render    packed records; the pairs dropped into their tiles and every tile's list put in
                                 Gaussian (= depth) order; one warp per piece of a tile's list: alpha, T, colour; a
                                 checkpoint of T every 8 pairs is all the backward keeps (16 B per pair)
              gcp_view_forward   = plan + render in one call on a pooled arena's capacity (SPECULATE below)
    backward  gcp_view_backward  per 8 pairs T is recomputed from its checkpoint, then the list is walked in
                                 reverse: U_i, dL/dalpha_i = T_i (<dL/dI, l_i> - U_i), the moments of g*dalpha over
                                 each pair's pixels (:733-766), the pairs of a Gaussian summed in pair order (:776-783)
  A whole batch of views in one call: views.NativeViewBatch (gcp_views_step).

  "lists" (csrc/gcp_splat.cu + the scan ops a1 / a3) — the element-list route:
    forward   gcp_splat_pack     per-Gaussian tables -> two 32-byte records (one L2 sector per gather)
              gcp_splat_place    boxes -> the pixel-sorted (key, Gaussian id) element list, built directly by a
                                 counting placement per image row — no sort of the N elements
                                 (alternative kept: gcp_splat_expand + gcp_splat_sort)  (:480-482, :538-548)
              gcp_splat_alpha    x = 1 - opacity*g in sorted order                      (:493-495, :533-535)
              gcp_cumprod_fwd    the segmented scan (op a1)                             (:551)
              gcp_splat_color    exclusive T from the inclusive scan (no division) and
                                 the per-pixel colour sum                               (:562, :498-514)
    backward  gcp_splat_bwd_w    w_k = <dL/dI, alpha_k l_k>, shifted by one in its list
              gcp_cumprod_bwd    division-free T_k*U_k (op a3)                          (replaces :716-722)
              gcp_splat_bwd_elem (dalpha, d) per element, written at its Gaussian-major
                                 position (the un-sort, without a permutation array)
              gcp_splat_bwd_reduce per-element gradients summed over each box (small boxes 8 lanes
                                 each, large ones in 1024-element pieces): no float atomics,
                                 deterministic                                          (:733-783)

There is no un-sort (:555), no flip + second sorted pass (:716-722), no chunk loop (:675, :792), no forward
recompute in the backward (:799) and no division by 1-alpha (:736,:747,:757): with
U_i = w_{i+1} + (1-alpha_{i+1}) U_{i+1},  dL/dalpha_i = T_i <dL/dI, l_i> - T_i U_i  stays exact as alpha -> 1.

`batch` (chunk ends) is accepted and ignored: the scan carries across any length, so a view is one pass.  The
reference's chunked result differs from its own single-chunk result by one (1-alpha) factor per chunk
boundary (tests/golden keeps a fixture of it); the single-chunk result
is the one reproduced here.  No CPU path: CPU tensors raise.
"""
from __future__ import annotations

import torch

from . import _lib, ops

KEY_STRIDE = 10000  # pixel key = y*10000 + x  (gs_model.py:541)
# True: build the sorted element list with the sort-free counting placement (gcp_splat_place);
# False: expand to N (key, gid) pairs and radix-sort them (gcp_splat_expand + gcp_splat_sort).  Same output, bit for bit.
USE_PLACEMENT = True
# "tiles": the fused route (csrc/gcp_tile.cu) — one warp per 8x4-pixel tile walks the tile's Gaussians in depth order
#          with the running T of its pixels in registers; no element list, no float atomics (bitwise reproducible).
# "lists": the element-list route — placement, alpha, the scan ops a1/a3 (gcp_cumprod_fwd/bwd), colour, un-sort.
# Same image and gradients within fp32 rounding (tests/test_compositor.py runs every case through both).
ROUTE = "tiles"


def _p(t):
    return t.data_ptr()


_scratch = {}


def _scratch_bytes(dev, tag: str, nbytes: int) -> torch.Tensor:
    """Grow-only scratch per (device, stream, purpose): the kernels' temporaries are not re-allocated every view
    (keeps the caching allocator from splitting and re-growing the large per-view blocks)."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream, tag)
    t = _scratch.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(nbytes, 1 << 20) * 5 // 4, dtype=torch.uint8, device=dev)
        _scratch[key] = t
    return t


_plans = {}        # (boxsize ptr, startpoint ptr, endpoint ptr, n) -> _Plan, at most _MAX_PLANS entries
_MAX_PLANS = 4
_side_streams = {}


class _Plan:
    __slots__ = ("boxsize", "startpoint", "endpoint", "sp", "ep", "offs", "host", "event", "route", "arena")


def _aligned(t: torch.Tensor) -> torch.Tensor:
    """Contiguous and 16-byte aligned (the kernels read the tables with vector loads; a sliced view may not be)."""
    t = t.contiguous()
    return t.clone() if t.data_ptr() % 16 else t


def _totals_to_host(totals, dev):
    """Queue the copy of a prologue's totals into pinned host memory right behind it; returns (host tensor, event).
    The host waits on the event only after it has queued everything that does not depend on the numbers."""
    host = torch.empty(totals.numel(), dtype=torch.int64, pin_memory=True)
    host.copy_(totals, non_blocking=True)
    event = torch.cuda.Event()
    event.record(torch.cuda.current_stream(dev))
    return host, event


# ---- tile route: pooled arenas ------------------------------------------------------------------------------
# A view lives in two device buffers: the PLAN arena (sized by the number of Gaussians and tiles: counts, offsets,
# packed records, work lists) and the PAIR arena (sized by the number of (tile, Gaussian) pairs: the tile-ordered
# pair list, T checkpoints, piece state, gradient partials).  Both come from per-(device, stream) free lists and
# go back when the view's autograd state is dropped, so a training loop allocates them once.
class _PlanArena:
    __slots__ = ("buf", "totals", "totals_np", "busy")

    def __init__(self, dev, nbytes):
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.totals = torch.zeros(2, dtype=torch.int64).pin_memory()   # [0] = pair count: the plan kernel writes here
        self.totals_np = self.totals.numpy()
        self.busy = None    # event behind the last one-call forward that used this arena (see _forward_speculative)


class _PairArena:
    __slots__ = ("buf", "cap")

    def __init__(self, dev, cap, nbytes):
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.cap = cap


_free_plan = {}   # (device, stream) -> [_PlanArena]
_free_pair = {}   # (device, stream) -> [_PairArena]


def _pool_key(dev):
    return (dev.index, torch.cuda.current_stream(dev).cuda_stream)


def _take_plan_arena(L, dev, n, W, H, key=None) -> _PlanArena:
    need = int(L.gcp_view_plan_bytes(n, W, H))
    free = _free_plan.setdefault(key or _pool_key(dev), [])
    for i, a in enumerate(free):
        if a.buf.numel() >= need:
            return free.pop(i)
    if free:
        free.pop()   # too small for this view: let it go instead of hoarding
    return _PlanArena(dev, need + need // 8)


def _take_pair_arena(L, dev, pairs, W, H) -> _PairArena:
    free = _free_pair.setdefault(_pool_key(dev), [])
    for i, a in enumerate(free):
        # (the bytes a capacity needs also depend on the image size and on gcp_tile_set_piece_pairs)
        if a.cap >= pairs and a.buf.numel() >= int(L.gcp_view_pair_bytes(a.cap, W, H)):
            return free.pop(i)
    if free:
        free.pop()
    cap = pairs + pairs // 4 + 4096      # head room: the next views of a scene fit without a new block
    if cap >= 2 ** 31 - 64:
        cap = max(pairs, 16)
    return _PairArena(dev, cap, int(L.gcp_view_pair_bytes(cap, W, H)))


class _TileView:
    """Tile route: the two arenas of a rendered view, held until its autograd state is dropped."""
    __slots__ = ("n", "P", "W", "H", "plan", "pairs", "key", "piece", "keep")

    def __del__(self):
        try:
            if self.plan is not None:
                _free_plan.setdefault(self.key, []).append(self.plan)
            if self.pairs is not None:
                _free_pair.setdefault(self.key, []).append(self.pairs)
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass


def _plan_tiles(L, dev, startpoint, endpoint, n, W, H, key=None):
    """Queue the plan of a view (pair counts and offsets) on the current stream; returns (sp, ep, arena, event)."""
    stream = torch.cuda.current_stream(dev)
    sp, ep = _i32_pairs(startpoint), _i32_pairs(endpoint)
    arena = _take_plan_arena(L, dev, n, W, H, key)
    _lib.check(L.gcp_view_plan(_p(sp), _p(ep), n, W, H, _p(arena.buf), arena.buf.numel(), _p(arena.totals),
                               stream.cuda_stream), "gcp_view_plan")
    event = torch.cuda.Event()
    event.record(stream)
    return sp, ep, arena, event


def _prologue(L, dev, boxsize, startpoint, endpoint, n):
    """List route: element offsets per Gaussian and (cell, Gaussian) pair offsets (placement cells: one image row x
    2^S pixels; a box contributes rows x strips-it-touches pairs) in one call, on the current stream."""
    stream = torch.cuda.current_stream(dev).cuda_stream
    sp = startpoint.to(torch.int32).contiguous()
    ep = endpoint.to(torch.int32).contiguous()
    offs = torch.empty((2, n + 1), dtype=torch.int64, device=dev)
    totals = torch.empty(2, dtype=torch.int64, device=dev)
    temp = _scratch_bytes(dev, "prepare", int(L.gcp_splat_prepare_bytes(n)))
    _lib.check(L.gcp_splat_prepare(_p(boxsize.to(torch.int64).contiguous()), _p(sp), _p(ep), n, _p(offs[0]),
                                   _p(offs[1]), _p(totals), _p(temp), temp.numel(), stream), "gcp_splat_prepare")
    return sp, ep, offs, totals


def plan_view(boxsize, startpoint, endpoint, image_width=None, image_height=None) -> None:
    """Optional: queue a view's prologue (offsets and the element / pair counts) on a side stream ahead of its
    `custom_autograd_grouped_cumprod.apply(...)`.  The one host sync of a view then finds its two numbers already
    in pinned host memory instead of draining the device queue: a driver that renders many views (views.py,
    bench.py) plans view i+1 before it runs view i.  The plan is consumed by the next apply() that receives the
    SAME three tensors (and, for the tile route, the same image size — without it that route cannot plan and
    apply() runs the prologue itself); the tensors must not be modified in between."""
    dev = startpoint.device
    if dev.type != "cuda":
        raise RuntimeError("plan_view needs CUDA tensors")
    L = _lib.lib()
    n = boxsize.numel()
    tiles = ROUTE == "tiles"
    if tiles and (image_width is None or image_height is None):
        return
    with torch.cuda.device(dev):
        side = _side_streams.get(dev.index)
        if side is None:
            side = _side_streams[dev.index] = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))  # the inputs may still be in production
        pl = _Plan()
        pl.boxsize, pl.startpoint, pl.endpoint = boxsize, startpoint, endpoint  # keep the addresses alive
        pl.arena = None
        key = _pool_key(dev)   # the arena belongs to the stream that will render the view
        with torch.cuda.stream(side):
            if tiles:
                pl.sp, pl.ep, pl.arena, pl.event = _plan_tiles(L, dev, startpoint, endpoint, n, int(image_width),
                                                               int(image_height), key)
                pl.offs = pl.host = None
                pl.route = ("tiles", int(image_width), int(image_height))
            else:
                pl.sp, pl.ep, pl.offs, totals = _prologue(L, dev, boxsize, startpoint, endpoint, n)
                pl.route = ("lists",)
                pl.host, pl.event = _totals_to_host(totals, dev)
    while len(_plans) >= _MAX_PLANS:
        old = _plans.pop(next(iter(_plans)))
        if old.arena is not None:
            old.event.synchronize()
            _free_plan.setdefault(key, []).append(old.arena)
    _plans[(_p(boxsize), _p(startpoint), _p(endpoint), n)] = pl


class _View:
    """Device state of one rendered view kept for the backward (16 B per element + the per-Gaussian tables)."""
    __slots__ = ("n", "N", "W", "H", "key_s", "gid_s", "x_s", "incl", "mean", "lam", "opac", "l_d", "sp", "ep",
                 "goff", "rec_a", "rec_b", "seg_off", "cstart", "pgid", "btab", "P")


def _f32(t, shape=None):
    """fp32, contiguous.  The usual case — already both — costs two attribute reads: the kernels only take the data
    pointer, so neither detach() nor the reshape is needed then (four tensors per view: ~15 us of host time that the
    device would spend idle in front of a view's first kernel)."""
    if t.dtype is torch.float32 and t.is_contiguous():
        return t
    t = t.detach().to(torch.float32)
    if shape is not None:
        t = t.reshape(shape)
    return t.contiguous()


def _i32_pairs(t):
    """int32, contiguous, 8-byte aligned (the box corners are read as int2)."""
    if t.dtype is not torch.int32 or not t.is_contiguous():
        t = t.to(torch.int32).contiguous()
    return t.clone() if t.data_ptr() % 8 else t


def _render_forward_tiles(boxsize, startpoint, endpoint, mean, lam, opacity, l_d, W, H, keep=True) -> tuple:
    dev = startpoint.device
    L = _lib.lib()
    v = _TileView()
    v.plan = v.pairs = None
    v.W, v.H = W, H
    v.piece = int(L.gcp_tile_piece_pairs())
    v.keep = bool(keep)
    n = boxsize.numel()
    v.n = n
    with torch.cuda.device(dev):
        cur = torch.cuda.current_stream(dev)
        stream = cur.cuda_stream
        v.key = _pool_key(dev)
        plan = _plans.pop((_p(boxsize), _p(startpoint), _p(endpoint), n), None)
        if plan is not None and plan.route != ("tiles", W, H):
            if plan.arena is not None:
                plan.event.synchronize()
                _free_plan.setdefault(v.key, []).append(plan.arena)
            plan = None
        mean_, lam_, opac_, l_ = _f32(mean, (n, 2)), _f32(lam, (n, 4)), _f32(opacity, (n,)), _f32(l_d, (n, 3))
        image = torch.empty((H + 1, W + 1, 3), dtype=torch.float32, device=dev)  # every pixel is written by its lane
        if plan is None and SPECULATE and _free_pair.get(v.key):
            # A pair arena of an earlier view is at hand: plan AND render are queued in one call on its capacity, the
            # host then reads the pair count the plan kernel drops into pinned memory.  The device never idles between
            # the plan and the render waiting for the host (that gap was 60-90 us of a 0.4 ms forward); a view that
            # turns out larger than the arena was not rendered at all (every kernel checks) and is redone below.
            if _forward_speculative(L, v, dev, startpoint, endpoint, mean_, lam_, opac_, l_, n, W, H, keep, image, cur):
                return image, v
        if plan is not None:
            cur.wait_event(plan.event)
            sp, ep, v.plan, event = plan.sp, plan.ep, plan.arena, plan.event
            for t_ in (sp, ep):
                t_.record_stream(cur)
        else:
            sp, ep, v.plan, event = _plan_tiles(L, dev, startpoint, endpoint, n, W, H)
        # the one host sync of a view (like the reference's .item() at uitility.py:348): the plan kernel wrote the
        # pair count straight into pinned memory
        event.synchronize()
        v.P = int(v.plan.totals_np[0])
        if v.P >= 2 ** 31 - 64:
            raise RuntimeError("a view is limited to 2**31 (tile, Gaussian) pairs")
        v.pairs = _take_pair_arena(L, dev, v.P, W, H)
        _lib.check(L.gcp_view_render(_p(sp), _p(ep), _p(mean_), _p(lam_), _p(opac_), _p(l_), n, W, H, _p(v.plan.buf),
                                     v.plan.buf.numel(), _p(v.pairs.buf), v.pairs.buf.numel(), v.pairs.cap,
                                     1 if keep else 0, _p(image), stream), "gcp_view_render")
    return image, v


# True: a view whose plan was not queued ahead (plan_view) is rendered on the capacity of a pooled pair arena
# without waiting for its pair count first (see _render_forward_tiles).  Exact either way.
SPECULATE = True


def _forward_speculative(L, v, dev, startpoint, endpoint, mean_, lam_, opac_, l_, n, W, H, keep, image, cur) -> bool:
    import time

    sp, ep = _i32_pairs(startpoint), _i32_pairs(endpoint)
    free = _free_pair[v.key]
    pairs = max(free, key=lambda a: a.cap)
    if pairs.buf.numel() < int(L.gcp_view_pair_bytes(pairs.cap, W, H)):
        return False        # sized for another image / piece length
    free.remove(pairs)
    v.plan = _take_plan_arena(L, dev, n, W, H)
    v.pairs = pairs
    # The pinned count word is written by the DEVICE, asynchronously: a plan kernel of the arena's previous view that
    # is still queued (a host running ahead of the device, e.g. a no_grad render loop) would drop ITS count there
    # after the host has reset the word.  The previous one-call forward on this arena must have finished first.
    if v.plan.busy is not None:
        v.plan.busy.synchronize()
    v.plan.totals_np[0] = -1
    _lib.check(L.gcp_view_forward(_p(sp), _p(ep), _p(mean_), _p(lam_), _p(opac_), _p(l_), n, W, H, _p(v.plan.buf),
                                  v.plan.buf.numel(), _p(pairs.buf), pairs.buf.numel(), pairs.cap, 1 if keep else 0,
                                  _p(image), _p(v.plan.totals), cur.cuda_stream), "gcp_view_forward")
    v.plan.busy = torch.cuda.Event()
    v.plan.busy.record(cur)
    # the count arrives a few microseconds after the plan's scan kernel: poll the pinned word, fall back to a sync
    t0 = time.perf_counter()
    while v.plan.totals_np[0] < 0:
        if time.perf_counter() - t0 > 2e-3:
            cur.synchronize()
            break
    v.P = int(v.plan.totals_np[0])
    if v.P >= 2 ** 31 - 64:
        raise RuntimeError("a view is limited to 2**31 (tile, Gaussian) pairs")
    if v.P <= pairs.cap:
        return True
    # too small: nothing was rendered.  Let the arena go and run the view again on one that fits.
    cur.synchronize()
    v.pairs = _take_pair_arena(L, dev, v.P, W, H)
    _lib.check(L.gcp_view_forward(_p(sp), _p(ep), _p(mean_), _p(lam_), _p(opac_), _p(l_), n, W, H, _p(v.plan.buf),
                                  v.plan.buf.numel(), _p(v.pairs.buf), v.pairs.buf.numel(), v.pairs.cap,
                                  1 if keep else 0, _p(image), _p(v.plan.totals), cur.cuda_stream), "gcp_view_forward")
    v.plan.busy = torch.cuda.Event()
    v.plan.busy.record(cur)
    return True


def _render_backward_tiles(v: _TileView, grad_image):
    dev = grad_image.device
    n = v.n
    L = _lib.lib()
    if not v.keep:
        raise RuntimeError("this view was rendered without keeping T (no input required a gradient)")
    if int(L.gcp_tile_piece_pairs()) != v.piece:
        raise RuntimeError("gcp_tile_set_piece_pairs changed between the forward and the backward of a view")
    g_mean = torch.empty((n, 2), dtype=torch.float32, device=dev)
    g_lam = torch.empty((n, 4), dtype=torch.float32, device=dev)
    g_opac = torch.empty((n,), dtype=torch.float32, device=dev)
    g_l = torch.empty((n, 3), dtype=torch.float32, device=dev)
    gI = _f32(grad_image)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(L.gcp_view_backward(_p(v.plan.buf), v.plan.buf.numel(), _p(v.pairs.buf), v.pairs.buf.numel(),
                                       v.pairs.cap, _p(gI), n, v.W, v.H, _p(g_mean), _p(g_lam), _p(g_opac), _p(g_l),
                                       stream), "gcp_view_backward")
    return g_mean, g_lam, g_opac, g_l


def _render_forward(boxsize, startpoint, endpoint, mean, lam, opacity, l_d, W, H, keep=True) -> tuple:
    """keep=False: a render no backward will follow (tile route: the per-pair T is not stored)."""
    dev = startpoint.device
    if dev.type != "cuda":
        raise RuntimeError("custom_autograd_grouped_cumprod needs CUDA tensors (there is no CPU path)")
    if ROUTE == "tiles":
        return _render_forward_tiles(boxsize, startpoint, endpoint, mean, lam, opacity, l_d, W, H, keep)
    if ROUTE != "lists":
        raise ValueError(f"unknown compositor route {ROUTE!r}")
    L = _lib.lib()
    v = _View()
    v.W, v.H = W, H
    v.seg_off = v.cstart = v.pgid = v.btab = None
    v.P = 0
    n = boxsize.numel()
    v.n = n
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        plan = _plans.pop((_p(boxsize), _p(startpoint), _p(endpoint), n), None)
        if plan is not None and plan.route != ("lists",):
            plan = None
        if plan is not None:
            # planned ahead (plan_view): the prologue ran on a side stream, its totals are in pinned memory
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(plan.event)
            v.sp, v.ep, offs = plan.sp, plan.ep, plan.offs
            host, event = plan.host, plan.event
            for t_ in (v.sp, v.ep, offs):
                t_.record_stream(cur)
        else:
            # then one host sync per view for the two totals, like the reference's .item() at uitility.py:348
            v.sp, v.ep, offs, totals = _prologue(L, dev, boxsize, startpoint, endpoint, n)
            host, event = _totals_to_host(totals, dev)
        goff, poff = offs[0], offs[1]
        v.goff = goff
        v.mean = mean.detach().to(torch.float32).contiguous()
        v.lam = lam.detach().to(torch.float32).reshape(n, 4).contiguous()
        v.opac = opacity.detach().to(torch.float32).reshape(n).contiguous()
        v.l_d = l_d.detach().to(torch.float32).contiguous()
        image = torch.zeros((H + 1, W + 1, 3), dtype=torch.float32, device=dev)
        sp, ep = v.sp, v.ep
        # per-Gaussian tables as 32-byte records: one L2 sector per gather in the per-element kernels
        # (queued before the host waits for the totals, so the device has work meanwhile)
        v.rec_a = torch.empty((n, 8), dtype=torch.float32, device=dev)
        v.rec_b = torch.empty((n, 8), dtype=torch.int32, device=dev)
        _lib.check(L.gcp_splat_pack(_p(v.mean), _p(v.lam), _p(v.opac), _p(v.l_d), _p(sp), _p(ep), _p(goff), n,
                                    _p(v.rec_a), _p(v.rec_b), stream), "gcp_splat_pack")
        event.synchronize()
        N, P = host.tolist()
        v.N = N
        if N == 0:
            v.key_s = v.gid_s = v.x_s = v.incl = None
            return image, v
        if N >= 2 ** 31:
            raise RuntimeError("a view is limited to 2**31-1 elements (the reference ops index with int32)")
        v.key_s = torch.empty(N, dtype=torch.int32, device=dev)
        v.gid_s = torch.empty(N, dtype=torch.int32, device=dev)
        if USE_PLACEMENT:
            seg_off = torch.empty((H + 1) * (W + 1) + 1, dtype=torch.int32, device=dev)
            temp = _scratch_bytes(dev, "place", int(L.gcp_splat_place_bytes(P, W, H)))
            if L.gcp_splat_long_lists(P, W, H):
                # long pixel lists: the backward walks the placement's cells again, keep the pair list with the view
                v.seg_off = seg_off
                v.cstart = torch.empty(int(L.gcp_splat_num_cells(W, H)) + 1, dtype=torch.int32, device=dev)
                v.pgid = torch.empty(max(P, 1), dtype=torch.int32, device=dev)
                v.btab = torch.empty(int(L.gcp_splat_batch_table_ints(P, W, H)), dtype=torch.int32, device=dev)
                v.P = P
                cs, pg, bt = _p(v.cstart), _p(v.pgid), _p(v.btab)
            else:
                cs = pg = bt = None
            _lib.check(L.gcp_splat_place(_p(sp), _p(ep), _p(poff), n, P, W, H, _p(v.key_s), _p(v.gid_s), _p(seg_off),
                                         cs, pg, bt, _p(temp), temp.numel(), stream), "gcp_splat_place")
            del temp, seg_off
        else:
            key = torch.empty(N, dtype=torch.int32, device=dev)
            gid = torch.empty(N, dtype=torch.int32, device=dev)
            _lib.check(L.gcp_splat_expand(_p(sp), _p(ep), _p(goff), n, N, _p(key), _p(gid), stream),
                       "gcp_splat_expand")
            temp = torch.empty(int(L.gcp_splat_sort_bytes(N)), dtype=torch.uint8, device=dev)
            _lib.check(L.gcp_splat_sort(_p(key), _p(gid), _p(v.key_s), _p(v.gid_s), N, H * KEY_STRIDE + W, _p(temp),
                                        temp.numel(), stream), "gcp_splat_sort")
            del key, gid, temp
        v.x_s = torch.empty(N, dtype=torch.float32, device=dev)
        _lib.check(L.gcp_splat_alpha(_p(v.key_s), _p(v.gid_s), _p(v.rec_a), N, _p(v.x_s), stream),
                   "gcp_splat_alpha")
        v.incl = torch.empty_like(v.x_s)
        ops.grouped_cumprod_forward(v.x_s, v.key_s, v.incl)
        _lib.check(L.gcp_splat_color(_p(v.incl), _p(v.x_s), _p(v.key_s), _p(v.gid_s), _p(v.rec_b), N, W, _p(image),
                                     stream), "gcp_splat_color")
    return image, v


def _render_backward(v, grad_image):
    if isinstance(v, _TileView):
        return _render_backward_tiles(v, grad_image)
    dev = grad_image.device
    n = v.n
    if v.N == 0:
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)  # noqa: E731
        return z(n, 2), z(n, 4), z(n), z(n, 3)
    g_mean = torch.empty((n, 2), dtype=torch.float32, device=dev)
    g_lam = torch.empty((n, 4), dtype=torch.float32, device=dev)
    g_opac = torch.empty((n,), dtype=torch.float32, device=dev)
    g_l = torch.empty((n, 3), dtype=torch.float32, device=dev)
    L = _lib.lib()
    gI = _f32(grad_image)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        gshift = torch.empty_like(v.x_s)
        _lib.check(L.gcp_splat_bwd_w(_p(v.incl), _p(v.x_s), _p(v.key_s), _p(v.gid_s), _p(v.rec_b), _p(gI), v.N, v.W,
                                     _p(gshift), stream), "gcp_splat_bwd_w")
        tu = torch.empty_like(v.x_s)
        # the sorted pixel keys serve as segment ids: the backward op only compares neighbours
        # (inv_len is implied by them, include/gcp_abi.h)
        ops.grouped_cumprod_backward(v.x_s, v.incl, gshift, v.key_s, tu,
                                     torch.empty(0, dtype=torch.int32, device=dev))
        elem = gshift.new_empty((v.N, 2))
        del gshift
        if v.pgid is not None:
            _lib.check(L.gcp_splat_bwd_elem_cells(_p(v.incl), _p(v.x_s), _p(tu), _p(v.rec_b), _p(gI), _p(v.seg_off),
                                                  _p(v.cstart), _p(v.pgid), _p(v.btab), v.P, v.W, v.H, _p(elem),
                                                  stream), "gcp_splat_bwd_elem_cells")
        else:
            _lib.check(L.gcp_splat_bwd_elem(_p(v.incl), _p(v.x_s), _p(tu), _p(v.key_s), _p(v.gid_s), _p(v.rec_b),
                                            _p(gI), v.N, v.W, _p(elem), stream), "gcp_splat_bwd_elem")
        temp = _scratch_bytes(dev, "reduce", int(L.gcp_splat_bwd_reduce_bytes(v.N, n)))
        _lib.check(L.gcp_splat_bwd_reduce(_p(elem), _p(v.sp), _p(v.ep), _p(v.goff), _p(v.mean), _p(v.lam),
                                          _p(v.opac), _p(v.l_d), v.N, n, _p(g_mean), _p(g_lam), _p(g_opac),
                                          _p(g_l), _p(temp), temp.numel(), stream), "gcp_splat_bwd_reduce")
    return g_mean, g_lam, g_opac, g_l


class custom_autograd_grouped_cumprod(torch.autograd.Function):
    @staticmethod
    def forward(ctx, boxsize, batch, startpoint, endpoint, mean, variance_inverse, opacity, l_d, image_width,
                image_height):
        with torch.no_grad():
            image, view = _render_forward(boxsize, startpoint, endpoint, mean, variance_inverse, opacity, l_d,
                                          int(image_width), int(image_height), keep=any(ctx.needs_input_grad))
        ctx.view = view
        ctx.shapes = (mean.dtype, variance_inverse.shape, opacity.shape)
        return image

    @staticmethod
    def backward(ctx, grad_image):
        with torch.no_grad():
            g_mean, g_lam, g_opac, g_l = _render_backward(ctx.view, grad_image)
        mdt, lshape, oshape = ctx.shapes
        return (None, None, None, None, g_mean.to(mdt), g_lam.reshape(lshape), g_opac.reshape(oshape), g_l, None,
                None)
