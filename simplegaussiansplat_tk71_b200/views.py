"""View sharding across ranks.

Views are the only data-parallel axis of the reference: `for batch_i in range(shape_image)` renders them
one after another and they never interact (gs_model.py:402-449) until autograd sums the per-Gaussian
parameter gradients of the batch (gs_control.py:180-185).  One process per GPU owns views
rank, rank+world, ...; the compositing scan needs no collective.  The only exchange of a training step is
the sum of the flattened Gaussian-parameter gradient bucket, an all-reduce over NCCL (NVLink/NVSwitch).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

# floats per Gaussian in the reference's parameter set (gs_model.py:151-158):
# mean 3 + quaternion 4 + scale 3 + opacity 1 + SH colour 27
PARAM_FLOATS_PER_GAUSSIAN = 38


def views_for_rank(num_views: int, rank: int, world: int) -> list:
    """Round-robin ownership: rank r renders views r, r+world, ..."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, num_views, world))


def views_for_rank_balanced(costs, rank: int, world: int) -> list:
    """Cost-aware ownership: every rank computes the same assignment from the same per-view costs (e.g. the pair
    counts `NativeViewBatch.totals_np` of the previous step): views in order of decreasing cost, each to the rank
    with the least work so far (ties to the lower rank).  The step of a data-parallel batch ends with the slowest
    rank; with round-robin the ranks' totals differ by the spread of 64/world random views."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    order = sorted(range(len(costs)), key=lambda v: (-float(costs[v]), v))
    load = [0.0] * world
    mine = []
    for v in order:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += float(costs[v])
        if r == rank:
            mine.append(v)
    return sorted(mine)


def aggregate_throughput(elements_local: float, ms_local: float, device=None):
    """(total elements over ranks, max time over ranks).  Works on nccl (cuda tensors) and gloo (cpu)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(elements_local), float(ms_local)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    n = torch.tensor([float(elements_local)], dtype=torch.float64, device=dev)
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=dev)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(n.item()), float(t.item())


def allreduce_param_grads(bucket: torch.Tensor) -> torch.Tensor:
    """Sum the flat f32[n_gaussians * 38] gradient bucket over ranks, in place (one collective per step)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
    return bucket


# --------------------------------------------------------------------------------------------
# Per-view preparation of the compositor inputs (row a10 of SURVEY.md §8): the body of the reference's
# per-view loop, gs_model.py:402-449, as functions.  Plain torch device ops (K-sized, not element-sized).
# --------------------------------------------------------------------------------------------
def visible_boxes(mean_pixel, box_half, z, width: int, height: int):
    """Cull and clamp the depth-sorted Gaussians of ONE view.

    mean_pixel i32[n,2], box_half i32/f32[n,2] (3-sigma half widths), z f32[n] camera depth, all already in
    z order (gs_model.py:356-365).  Returns (mask bool[n], startpoint i32[m,2], endpoint i32[m,2], boxsize i64[m]):
      mask       z > 0, box != 0, box intersects the image                      (gs_model.py:405-406)
      corners    mean -/+ half, clamped to the INCLUSIVE range [0,W] x [0,H]     (:419-423)
      boxsize    prod(end - start + 1)                                          (:424)
    """
    mx, my = mean_pixel[:, 0], mean_pixel[:, 1]
    bx, by = box_half[:, 0], box_half[:, 1]
    mask = (z > 0) & (bx != 0) & (mx - bx < width) & (mx + bx > 0) & (my - by < height) & (my + by > 0)
    mx, my, bx, by = mx[mask], my[mask], bx[mask], by[mask]
    sp = torch.stack(((mx - bx).clamp(min=0, max=width), (my - by).clamp(min=0, max=height)), 1)
    ep = torch.stack(((mx + bx).clamp(min=0, max=width), (my + by).clamp(min=0, max=height)), 1)
    boxsize = torch.prod((ep - sp + 1).to(torch.int64), dim=1)
    return mask, sp.to(torch.int32), ep.to(torch.int32), boxsize


def split_by_cumsum_parallel(x: torch.Tensor, limit: float) -> torch.Tensor:
    """Number of items per chunk when a new chunk starts each time the running sum passes a multiple of `limit`
    (uitility.py:478-488): counts of floor(cumsum(x) / limit)."""
    group_id = torch.floor_divide(torch.cumsum(x, dim=0), limit)
    _, counts = torch.unique(group_id, return_counts=True)
    return counts


def chunk_ends(boxsize: torch.Tensor) -> torch.Tensor:
    """The `batch` argument of the compositor as the reference builds it (gs_model.py:428): chunks of at most
    2**29 elements.  The native compositor renders a view in one pass and ignores it; it is produced only so that
    callers written against the reference keep working."""
    return torch.cumsum(split_by_cumsum_parallel(boxsize / 1024, (1024 ** 3 * 6 / 12) / 1024), dim=0)


_ONE_CHUNK = torch.zeros(1, dtype=torch.int64)   # placeholder for the ignored `batch` argument


def render_views(mean_pixel, box_half, z, lam, opacity, l_d, width: int, height: int):
    """The per-view loop of gs_model.py:402-454 over V views whose Gaussians are already z-sorted:
    inputs [V,n,...]; returns images [V',3,H,W] (views without any visible Gaussian are skipped, :414-417) —
    including the reference's final `[:,1:,1:,:].reshape(-1,3,H,W)` (a reshape, not a permute, :454)."""
    from .compositor import custom_autograd_grouped_cumprod as F, plan_view

    out = []
    V = mean_pixel.shape[0]
    nxt = visible_boxes(mean_pixel[0], box_half[0], z[0], width, height) if V else None
    for v in range(V):
        mask, sp, ep, boxsize = nxt
        if v + 1 < V:
            # cull the next view now and queue its prologue on a side stream: its element count is then already
            # on the host when its turn comes (compositor.plan_view)
            nxt = visible_boxes(mean_pixel[v + 1], box_half[v + 1], z[v + 1], width, height)
            if nxt[1].shape[0]:
                plan_view(nxt[3], nxt[1], nxt[2], width, height)
        if sp.shape[0] == 0:
            continue
        # `batch` (the chunk ends, gs_model.py:428) is accepted and ignored by the native compositor — a view is one
        # pass — so it is not computed here: chunk_ends() costs a cumsum, a torch.unique and a host sync per view
        out.append(F.apply(boxsize, _ONE_CHUNK, sp, ep, mean_pixel[v][mask], lam[v][mask], opacity[v][mask],
                           l_d[v][mask], width, height))
    if not out:
        return torch.zeros((0, 3, height, width), device=mean_pixel.device)
    return torch.stack(out, dim=0)[:, 1:, 1:, :].reshape(-1, 3, height, width)


# --------------------------------------------------------------------------------------------
# The same loop as ONE native call per training step (include/gcp_abi.h: gcp_views_step): render, loss gradient
# and backward of every view of this rank, the views' gradients summed into the parameters' gradient arrays.
# The Python loop above costs ~0.85 ms of host time per view (autograd Function, allocations, ctypes calls) against
# ~0.65 ms of device time; here the host enqueues a whole step in a few milliseconds and never waits.
# --------------------------------------------------------------------------------------------
class NativeViewBatch:
    """The views one rank renders in a training step, prepared once.

    views: sequence of objects with .startpoint .endpoint .mean .lam .opacity .l_d (a workloads.SplatView, or the
    per-view slices gs_model.py:405-425 produces) and optionally .index (i32[n], the parameter row of each of the
    view's Gaussians: the reference's boolean-mask selection as row numbers).  targets: per-view images
    f32[H+1,W+1,3] for the built-in mean-squared-error loss, or grad_images: per-view dL/d image.
    step(...) enqueues the whole batch on the current stream; finish() (after the caller synchronised) tells
    whether every view fitted the pair arenas — if not they have been enlarged and the step must be repeated."""

    def __init__(self, views, width: int, height: int, targets=None, grad_images=None, lanes: int = 4,
                 keep_images: bool = False):
        import ctypes

        from . import _lib

        if (targets is None) == (grad_images is None):
            raise ValueError("give either targets (built-in MSE loss) or grad_images")
        self.L = L = _lib.lib()
        self.W, self.H, self.lanes = int(width), int(height), int(lanes)
        self.V = len(views)
        dev = views[0].startpoint.device if self.V else torch.device("cuda")
        if dev.type != "cuda":
            raise RuntimeError("NativeViewBatch needs CUDA tensors (there is no CPU path)")
        self.dev = dev
        f32 = lambda t, shape: t.detach().to(torch.float32).reshape(shape).contiguous()  # noqa: E731
        self._keep = []          # every tensor a descriptor points at
        self.desc = (_lib.ViewDesc * max(self.V, 1))()
        shape = (self.H + 1, self.W + 1, 3)
        n_img = self.V if keep_images else min(self.lanes, max(self.V, 1))
        self.images = [torch.empty(shape, dtype=torch.float32, device=dev) for _ in range(n_img)]
        self.gscratch = [torch.empty(shape, dtype=torch.float32, device=dev) for _ in range(self.lanes)] \
            if targets is not None else []
        self.n_max = 1
        for v, sc in enumerate(views):
            n = int(sc.startpoint.shape[0])
            self.n_max = max(self.n_max, n)
            sp = sc.startpoint.to(torch.int32).contiguous()
            ep = sc.endpoint.to(torch.int32).contiguous()
            t = [sp, ep, f32(sc.mean, (n, 2)), f32(sc.lam, (n, 4)), f32(sc.opacity, (n,)), f32(sc.l_d, (n, 3))]
            idx = getattr(sc, "index", None)
            if idx is not None:
                idx = idx.to(device=dev, dtype=torch.int32).contiguous()
                if idx.numel() != n:
                    raise ValueError("index must name one parameter row per Gaussian of the view")
            d = self.desc[v]
            d.sp, d.ep, d.mean, d.lam, d.opac, d.l_d = (x.data_ptr() for x in t)
            d.index = idx.data_ptr() if idx is not None else None
            if targets is not None:
                tg = f32(targets[v], shape)
                d.target, d.grad_image = tg.data_ptr(), self.gscratch[v % self.lanes].data_ptr()
            else:
                tg = f32(grad_images[v], shape)
                d.target, d.grad_image = None, tg.data_ptr()
            d.image = self.images[v if keep_images else v % self.lanes].data_ptr()
            d.n = n
            self._keep.append((t, idx, tg))
        ctx = ctypes.c_void_p()
        _lib.check(L.gcp_views_ctx_create(self.lanes, ctypes.byref(ctx)), "gcp_views_ctx_create")
        self.ctx = ctx
        self.totals = torch.zeros(max(self.V, 1), dtype=torch.int64).pin_memory()
        self.totals_np = self.totals.numpy()
        self.plan_bytes = int(L.gcp_view_plan_bytes(self.n_max, self.W, self.H))
        self.plans = [torch.empty(self.plan_bytes, dtype=torch.uint8, device=dev) for _ in range(self.lanes)]
        self.cap = 0
        self.pairs = []
        self.launches = 0

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                torch.cuda.synchronize(self.dev)
                self.L.gcp_views_ctx_destroy(self.ctx)
                self.ctx = None
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    def _arenas(self, cap: int):
        import ctypes

        self.cap = int(cap)
        self.pair_bytes = int(self.L.gcp_view_pair_bytes(self.cap, self.W, self.H))
        self.pairs = [torch.empty(self.pair_bytes, dtype=torch.uint8, device=self.dev) for _ in range(self.lanes)]
        vp = ctypes.c_void_p
        self._plan_ptrs = (vp * self.lanes)(*[p.data_ptr() for p in self.plans])
        self._pair_ptrs = (vp * self.lanes)(*[p.data_ptr() for p in self.pairs])

    def size(self):
        """One pass of the plan kernels over all views (pair counts only) and ONE host sync: sizes the pair arenas
        for the largest view plus 10 %.  step() calls it when there are no arenas yet."""
        from . import _lib

        stream = torch.cuda.current_stream(self.dev)
        for v in range(self.V):
            d = self.desc[v]
            _lib.check(self.L.gcp_view_plan(d.sp, d.ep, d.n, self.W, self.H, self.plans[0].data_ptr(), self.plan_bytes,
                                            self.totals.data_ptr() + 8 * v, stream.cuda_stream), "gcp_view_plan")
        stream.synchronize()
        most = int(self.totals_np[: self.V].max()) if self.V else 0
        self._arenas(most + most // 10 + 4096)

    def step(self, g_mean, g_lam, g_opac, g_l, loss=None, tail=None):
        """Enqueue render + loss gradient + backward of all views on the current stream; the views' gradients are
        ADDED into g_mean f32[*,2], g_lam f32[*,4], g_opac f32[*], g_l f32[*,3] (rows = desc.index).  Returns at
        once; call finish() after synchronising.

        tail = (first_tail_view, (t_mean, t_lam, t_opac, t_l), event): the views from first_tail_view on add into
        the second set of arrays, and `event` (a torch.cuda.Event that has been recorded at least once, or None) is
        recorded as soon as the main arrays are complete — the caller's all-reduce of the main bucket can start
        there, beside the tail views (gcp_views_step_split)."""
        import ctypes

        from . import _lib

        if not self.pairs:
            self.size()
        split = None
        if tail is not None:
            first, arrays, event = tail
            split = _lib.ViewsSplit(int(first), arrays[0].data_ptr(), arrays[1].data_ptr(), arrays[2].data_ptr(),
                                    arrays[3].data_ptr(), event.cuda_event if event is not None else None)
        with torch.cuda.device(self.dev):
            stream = torch.cuda.current_stream(self.dev).cuda_stream
            _lib.check(self.L.gcp_views_step_split(self.ctx, self.desc, self.V, self.W, self.H, self._plan_ptrs,
                                                   self.plan_bytes, self._pair_ptrs, self.pair_bytes, self.cap,
                                                   g_mean.data_ptr(), g_lam.data_ptr(), g_opac.data_ptr(),
                                                   g_l.data_ptr(), loss.data_ptr() if loss is not None else None,
                                                   self.totals.data_ptr(),
                                                   ctypes.byref(split) if split is not None else None, stream),
                       "gcp_views_step_split")
        self.launches = int(self.L.gcp_view_last_launch_count())

    def finish(self) -> bool:
        """After the step's stream was synchronised: True when every view fitted its arena.  False: the arenas have
        been enlarged; zero the gradients and run the step again (views that did not fit were skipped)."""
        most = int(self.totals_np[: self.V].max()) if self.V else 0
        if most <= self.cap:
            return True
        self._arenas(most + most // 10 + 4096)
        return False
