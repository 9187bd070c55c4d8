"""Autograd wrapper the reference intended but never shipped (cuda_kernel.py:4 says
"kernel call and autograd graph construction class"; the ops themselves are autograd-free and are
only ever called under torch.no_grad(), gs_model.py:545)."""
from __future__ import annotations

import torch

from . import ops


class GroupedCumprod(torch.autograd.Function):
    """y = segmented inclusive cumprod of x over adjacent runs of `inv` (dense int32 segment ids).

    forward  -> grouped_cumprod_forward
    backward -> grouped_cumprod_backward (division-free, exact at x == 0)
    """

    @staticmethod
    def forward(ctx, x: torch.Tensor, inv: torch.Tensor, inv_len: torch.Tensor):
        x = x.contiguous()
        y = torch.empty_like(x)
        ops.grouped_cumprod_forward(x, inv, y)
        ctx.save_for_backward(x, y, inv, inv_len)
        return y

    @staticmethod
    def backward(ctx, grad_y: torch.Tensor):
        x, y, inv, inv_len = ctx.saved_tensors
        grad_x = torch.empty_like(x)
        ops.grouped_cumprod_backward(x, y, grad_y.contiguous(), inv, grad_x, inv_len)
        return grad_x, None, None


def grouped_cumprod(x: torch.Tensor, inv: torch.Tensor, inv_len: torch.Tensor) -> torch.Tensor:
    return GroupedCumprod.apply(x, inv, inv_len)
