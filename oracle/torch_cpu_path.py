"""TEST INFRASTRUCTURE — not product code.

"Reference PyTorch grouped_cumprod fwd+bwd on CPU" (BASELINE.json configs[0]).  The reference ships
no runnable CPU implementation of its ops, only commented-out attempts (uitility.py:369-379 with
torch_scatter, :382-428 with log/cumsum/exp).  This is the pure-PyTorch restatement BASELINE.md B1
describes: segments bucketed by power-of-two length, padded with ones, torch.cumprod(dim=1), and
torch autograd for the backward.  Semantics = grouped_cumprod_forward.cu:17-23 for sorted keys.
Only bench.py's cpu_baseline leg and tests/ import it.
"""
from __future__ import annotations

import torch


def _plan(key: torch.Tensor):
    n = key.numel()
    head = torch.ones(n, dtype=torch.bool)
    head[1:] = key[1:] != key[:-1]
    starts = torch.nonzero(head).flatten()
    lengths = torch.diff(starts, append=torch.tensor([n]))
    seg_id = torch.cumsum(head.to(torch.int64), 0) - 1
    pos = torch.arange(n) - starts[seg_id]
    width = torch.pow(2, torch.ceil(torch.log2(lengths.clamp(min=1).double())).long())
    buckets = []
    for w in torch.unique(width).tolist():
        segs = torch.nonzero(width == w).flatten()
        row_of_seg = torch.full((starts.numel(),), -1, dtype=torch.int64)
        row_of_seg[segs] = torch.arange(segs.numel())
        elem = torch.nonzero(width[seg_id] == w).flatten()
        buckets.append((int(w), segs.numel(), elem, row_of_seg[seg_id[elem]], pos[elem]))
    return buckets


def grouped_cumprod_fwd_bwd(x: torch.Tensor, key: torch.Tensor, grad_out: torch.Tensor, plan=None):
    """Returns (y, grad_x, plan).  x/grad_out f32[N], key i32[N] (CPU tensors)."""
    if plan is None:
        plan = _plan(key)
    xr = x.detach().clone().requires_grad_(True)
    y = torch.empty_like(x)
    loss = None
    for w, rows, elem, r, c in plan:
        pad = torch.ones(rows, w, dtype=x.dtype)
        pad = pad.index_put((r, c), xr[elem])
        cp = torch.cumprod(pad, dim=1)
        vals = cp[r, c]
        y[elem] = vals.detach()
        part = (vals * grad_out[elem]).sum()
        loss = part if loss is None else loss + part
    loss.backward()
    return y, xr.grad, plan
