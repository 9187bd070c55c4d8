"""TEST INFRASTRUCTURE — not product code.

Plain-PyTorch fp32 formulation of the compositor, kept as the "torch fp32 reference of the same op" for the
native kernels of simplegaussiansplat_tk71_b200/compositor.py: the same division-free algorithm written with
torch device ops around two scan-op calls (`ops`, injected by the tests: the oracle on CPU, the CUDA ops on
GPU).  It was the first working version of the product path and is what the native version was checked
against while it was built.  Only tests/ import it.

Original header:

B200-native compositor behind the reference's autograd contract.

Mirror of `custom_autograd_grouped_cumprod` (/root/reference/gs_model.py:477-820):

    image = custom_autograd_grouped_cumprod.apply(boxsize, batch, startpoint, endpoint, mean,
                                                  variance_inverse, opacity, l_d, image_width, image_height)

same ten positional inputs (gs_model.py:449, :666), same `(H+1, W+1, 3)` output (:505), gradients for
`mean, variance_inverse, opacity, l_d` only (:820).  What is computed is the reference's result for a view that
fits one chunk (sum(boxsize) <= 2**29, gs_model.py:428):

    T_i = prod_{j<i, same pixel} (1 - alpha_j),   alpha = opacity * exp(-1/2 (r-m) Lambda (r-m)^T)
    image[y, x] = sum_i T_i alpha_i l_i           (elements whose inclusive product is 0 contribute nothing, :575)

but not how: there is no inclusive->exclusive division (:562), no un-sort (:555), no second sorted pass with
flips for the backward (:716-722), no chunk loop (:675, :792) and no recompute of the forward in the backward
(:799).  The backward is division-free: with w_k = <dL/dI_pixel, alpha_k l_k>,

    dL/dalpha_i = T_i <dL/dI, l_i> - T_i U_i,     U_i = w_{i+1} + (1-alpha_{i+1}) U_{i+1}

and T_i U_i is exactly what `grouped_cumprod_backward` returns for grad_out = w shifted by one inside each
pixel list, so alpha -> 1 stays exact (the reference divides by 1-alpha at :736,:747,:757).

`batch` (chunk ends) is accepted and ignored: the scan kernels carry across any length, so a view is one
pass.  The reference's chunked result differs from its own single-chunk result by one (1-alpha) factor per
chunk boundary (tests/test_compositor_oracle.py keeps a fixture of it); the single-chunk
result is the one reproduced here.

Expansion, the stable key sort and the per-Gaussian reduction are plain torch device ops in this round
(plumbing around the two scan launches); SURVEY.md §8f ranks 2-3 replace them next.
"""
from __future__ import annotations

import torch

ops = None  # injected by the tests (oracle-backed on CPU, simplegaussiansplat_tk71_b200.ops on GPU)

KEY_STRIDE = 10000  # pixel key = y*10000 + x  (gs_model.py:541)


class ElementPlan:
    """Integer side of one view: the element list in pixel-sorted order (bit-exact contract).

    gid_s   i64[N]  Gaussian of each element, sorted by (pixel key, depth)  — depth order = Gaussian index
    px_s/py_s       pixel of each element
    key_s   i32[N]  y*10000+x, non-decreasing
    inv     i32[N]  dense segment (pixel list) id;  seg_end i32[K] exclusive ends  (cuda_test.py:21,27 layout)
    head/tail bool[N]
    """

    def __init__(self, boxsize, startpoint, endpoint):
        dev = startpoint.device
        boxsize = boxsize.to(torch.int64)
        n = boxsize.numel()
        N = int(boxsize.sum().item())
        self.n, self.N = n, N
        gid = torch.repeat_interleave(torch.arange(n, device=dev), boxsize, output_size=N)
        goff = torch.cumsum(boxsize, 0) - boxsize
        local = torch.arange(N, device=dev) - goff[gid]
        sx = startpoint[:, 0].to(torch.int64)
        sy = startpoint[:, 1].to(torch.int64)
        w = endpoint[:, 0].to(torch.int64) - sx + 1
        wg = w[gid]
        px = sx[gid] + local % wg                       # make_rect_points_parallel, uitility.py:336-366
        py = sy[gid] + torch.div(local, wg, rounding_mode="floor")
        key = (py * KEY_STRIDE + px).to(torch.int32)
        key_s, perm = torch.sort(key, stable=True)      # explicitly stable (the reference relies on it, :547)
        self.key_s = key_s
        self.gid_s = gid[perm]
        self.px_s = px[perm]
        self.py_s = py[perm]
        head = torch.ones(N, dtype=torch.bool, device=dev)
        if N > 1:
            head[1:] = key_s[1:] != key_s[:-1]
        self.head = head
        tail = torch.ones(N, dtype=torch.bool, device=dev)
        if N > 1:
            tail[:-1] = head[1:]
        self.tail = tail
        self.inv = (torch.cumsum(head.to(torch.int32), 0) - 1).to(torch.int32)
        self.seg_end = (torch.nonzero(tail).flatten() + 1).to(torch.int32)


def _element_values(plan: ElementPlan, mean, lam, opacity, l_d):
    gid = plan.gid_s
    m = mean.to(torch.float32)
    d0 = plan.px_s.to(torch.float32) - m[gid, 0]
    d1 = plan.py_s.to(torch.float32) - m[gid, 1]
    L = lam[gid]
    X0 = d0 * L[:, 0, 0] + d1 * L[:, 1, 0]              # (r-m) Lambda, gs_model.py:495,:745
    X1 = d0 * L[:, 0, 1] + d1 * L[:, 1, 1]
    g = torch.exp(-0.5 * (X0 * d0 + X1 * d1))
    o = opacity.reshape(-1)[gid]
    alpha = o * g
    x = (1.0 - alpha).contiguous()                      # anti_opacity, gs_model.py:535
    return d0, d1, X0, X1, g, o, alpha, x


class custom_autograd_grouped_cumprod(torch.autograd.Function):
    @staticmethod
    def forward(ctx, boxsize, batch, startpoint, endpoint, mean, variance_inverse, opacity, l_d, image_width,
                image_height):
        with torch.no_grad():
            W, H = int(image_width), int(image_height)
            plan = ElementPlan(boxsize, startpoint, endpoint)
            image = torch.zeros((H + 1, W + 1, 3), dtype=torch.float32, device=startpoint.device)
            ctx.plan = plan
            ctx.WH = (W, H)
            ctx.save_for_backward(mean, variance_inverse, opacity, l_d)
            if plan.N == 0:
                return image
            _, _, _, _, _, _, alpha, x = _element_values(plan, mean, variance_inverse, opacity, l_d)
            incl = torch.empty_like(x)
            ops.grouped_cumprod_forward(x, plan.key_s, incl)            # a1
            T = torch.where(plan.head, torch.ones_like(incl), torch.roll(incl, 1))   # exclusive, no division
            alive = incl != 0                                           # gs_model.py:575-578
            ta = torch.where(alive, T * alpha, torch.zeros_like(T))
            contrib = ta[:, None] * l_d[plan.gid_s]
            image.view(-1, 3).index_add_(0, plan.py_s * (W + 1) + plan.px_s, contrib)   # C = sum T alpha l
            return image

    @staticmethod
    def backward(ctx, grad_image):
        with torch.no_grad():
            mean, lam, opacity, l_d = ctx.saved_tensors
            plan = ctx.plan
            W, H = ctx.WH
            n = plan.n
            gm = torch.zeros((n, 2), dtype=torch.float32, device=grad_image.device)
            gL = torch.zeros((n, 4), dtype=torch.float32, device=grad_image.device)
            go = torch.zeros((n, 1), dtype=torch.float32, device=grad_image.device)
            gl = torch.zeros((n, 3), dtype=torch.float32, device=grad_image.device)
            if plan.N > 0:
                gid = plan.gid_s
                d0, d1, X0, X1, g, o, alpha, x = _element_values(plan, mean, lam, opacity, l_d)
                incl = torch.empty_like(x)
                ops.grouped_cumprod_forward(x, plan.key_s, incl)
                T = torch.where(plan.head, torch.ones_like(incl), torch.roll(incl, 1))
                alive = incl != 0
                l = l_d[gid]
                pg = grad_image.reshape(-1, 3)[plan.py_s * (W + 1) + plan.px_s]
                pgl = (pg * l).sum(1)                                    # <dL/dI, l_i>
                zero = torch.zeros_like(T)
                w = torch.where(alive, alpha * pgl, zero)                # w_k = <dL/dI, alpha_k l_k>
                gshift = torch.where(plan.tail, zero, torch.roll(w, -1)).contiguous()
                TU = torch.empty_like(x)
                ops.grouped_cumprod_backward(x, incl, gshift, plan.inv, TU, plan.seg_end)   # a3: T_i * U_i
                dalpha = torch.where(alive, T * pgl - TU, zero)
                d = T * w                                                # <dL/dI, p_i>
                coef = alpha * dalpha
                go.index_add_(0, gid, (g * dalpha)[:, None])
                gl.index_add_(0, gid, d[:, None] / l)                    # the reference's d / l (gs_model.py:763-766)
                gm.index_add_(0, gid, torch.stack((coef * X0, coef * X1), 1))
                hc = -0.5 * coef
                gL.index_add_(0, gid, torch.stack((hc * d0 * d0, hc * d0 * d1, hc * d1 * d0, hc * d1 * d1), 1))
            return (None, None, None, None, gm.to(mean.dtype), gL.reshape(n, 2, 2), go.reshape(opacity.shape), gl,
                    None, None)
