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
        out.append(F.apply(boxsize, chunk_ends(boxsize), sp, ep, mean_pixel[v][mask], lam[v][mask], opacity[v][mask],
                           l_d[v][mask], width, height))
    if not out:
        return torch.zeros((0, 3, height, width), device=mean_pixel.device)
    return torch.stack(out, dim=0)[:, 1:, 1:, :].reshape(-1, 3, height, width)
