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
