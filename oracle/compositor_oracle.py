"""TEST INFRASTRUCTURE — not product code.

numpy (fp64) restatement of the reference compositor `custom_autograd_grouped_cumprod`
(/root/reference/gs_model.py:477-820) for a view that fits ONE chunk (sum(boxsize) <= 2^29, gs_model.py:428),
i.e. without the chunk carry of :611-615.  Each step names the reference lines it follows.

Pinned by tests/golden/compositor_fixture.npz, which the reference Function itself produced on CPU
(tests/golden/make_compositor_fixture.py).  Only tests/ and bench.py's cpu_baseline leg import this.
"""
from __future__ import annotations

import numpy as np

from . import oracle as orc


def _scatter_sum(index, values, size):
    """out[index[i]] += values[i] in element order (what np.add.at does, column by column through np.bincount:
    the same sequential fp64 sums, ~20x faster)."""
    values = np.asarray(values, dtype=np.float64)
    if values.ndim == 1:
        return np.bincount(index, weights=values, minlength=size)
    return np.stack([np.bincount(index, weights=values[:, k], minlength=size) for k in range(values.shape[1])], 1)


def expand(boxsize, sp, ep):
    """make_rect_points_parallel (uitility.py:336-366): element order = Gaussian-major, row-major inside a box."""
    boxsize = np.asarray(boxsize, dtype=np.int64)
    n = boxsize.shape[0]
    N = int(boxsize.sum())
    gid = np.repeat(np.arange(n, dtype=np.int64), boxsize)
    goff = np.cumsum(boxsize) - boxsize
    local = np.arange(N, dtype=np.int64) - goff[gid]
    w = (ep[:, 0].astype(np.int64) - sp[:, 0] + 1)
    px = sp[gid, 0].astype(np.int64) + local % w[gid]
    py = sp[gid, 1].astype(np.int64) + local // w[gid]
    return gid, px, py


def forward(boxsize, sp, ep, mean, lam, opac, l_d, W, H):
    """gs_model.py:598-624 (_forward_batch) + :666-692 (forward).  Returns (image[H+1,W+1,3] f64, cache)."""
    gid, px, py = expand(boxsize, sp, ep)
    mean = np.asarray(mean, dtype=np.float64)
    lam = np.asarray(lam, dtype=np.float64)
    o = np.asarray(opac, dtype=np.float64).reshape(-1)[gid]
    l = np.asarray(l_d, dtype=np.float64)[gid]
    d0 = px - mean[gid, 0]
    d1 = py - mean[gid, 1]
    L = lam[gid]
    # :495  exp(-0.5 * (r-m) Lambda (r-m)^T), row vector times matrix times column vector
    q0 = d0 * L[:, 0, 0] + d1 * L[:, 1, 0]
    q1 = d0 * L[:, 0, 1] + d1 * L[:, 1, 1]
    g = np.exp(-0.5 * (q0 * d0 + q1 * d1))
    x = 1.0 - o * g                                   # :535 anti_opacity
    key = (py * 10000 + px).astype(np.int32)          # :541
    order = np.argsort(key, kind="stable")            # :547 (torch.sort on CUDA is stable)
    ks = key[order]
    # :551 inclusive cumprod in fp32 — it decides which elements survive `T != 0` (:575-578)
    x32 = (np.float32(1.0) - np.asarray(opac, np.float32).reshape(-1)[gid] * g.astype(np.float32)).astype(np.float32)
    incl32 = orc.cumprod_fwd(x32[order], ks, np.float32)
    alive_s = incl32 != 0.0
    incl64 = orc.cumprod_fwd(x[order].astype(np.float32), ks)  # fp64 accumulate of the fp32 inputs
    xs = x[order]
    T_s = np.where(alive_s, incl64 / np.where(xs == 0, 1.0, xs), 0.0)   # :562 inclusive -> exclusive
    inv_order = np.empty_like(order)
    inv_order[order] = np.arange(order.size)
    T = T_s[inv_order]
    alive = alive_s[inv_order]
    p = np.where(alive[:, None], T[:, None] * l * (o * g)[:, None], 0.0)   # :500
    image = _scatter_sum(py * (W + 1) + px, p, (H + 1) * (W + 1)).reshape(H + 1, W + 1, 3)   # :510-514
    cache = dict(gid=gid, px=px, py=py, o=o, l=l, g=g, x=x, d0=d0, d1=d1, L=L, key=key, order=order,
                 inv_order=inv_order, alive=alive, p=p, n=len(boxsize))
    return image, cache


def backward(cache, grad_image, scales=False):
    """gs_model.py:627-663 (_backward_batch), :733-766 (grad_*), :776-783 (scatter to Gaussians).

    scales=True additionally returns, for each of the four gradients, the sum of the ABSOLUTE values of the terms
    it is made of (every product, suffix sum and box sum evaluated with |.|): the condition-aware scale the fp32
    tolerance is stated against, |got - ref| <= 1e-6 + 1e-5 * scale (scale >= |ref|; the two coincide when no
    cancellation occurs).  The ops' own tests use the same form (tests/gpu_util.assert_close, scale=...)."""
    c = cache
    alive = c["alive"]
    pg = np.asarray(grad_image, dtype=np.float64)[c["py"], c["px"], :]      # :703-706
    d = np.where(alive, (pg * c["p"]).sum(1), 0.0)                           # :710-712
    order, inv_order = c["order"], c["inv_order"]
    ks = c["key"][order]
    ds = d[order]
    # :716-722 strict suffix sum inside each pixel: (total of the segment) - (inclusive prefix)
    incl = orc.cumsum_fwd(ds.astype(np.float32), ks)  # fp64 accumulate
    # use fp64 directly for the oracle: segment totals
    head = np.r_[True, ks[1:] != ks[:-1]]
    seg = np.cumsum(head) - 1
    nseg = int(seg[-1]) + 1 if seg.size else 0
    tot = _scatter_sum(seg, ds, nseg)
    pref = np.zeros_like(ds)
    # inclusive prefix in fp64
    csum = np.cumsum(ds)
    start_val = np.r_[0.0, csum[:-1]][head][seg]
    pref = csum - start_val
    s_sorted = tot[seg] - pref
    s = s_sorted[inv_order]
    o, l, g, x = c["o"], c["l"], c["g"], c["x"]
    xsafe = np.where(alive, x, 1.0)
    go = np.where(alive, -(g / xsafe) * s + d / o, 0.0)                      # :733-740
    gl = np.where(alive[:, None], d[:, None] / l, 0.0)                       # :763-766 (d / l, as the reference)
    coef = np.where(alive, -((o * g) / xsafe) * s + d, 0.0)
    X0 = c["d0"] * c["L"][:, 0, 0] + c["d1"] * c["L"][:, 1, 0]               # (r-m) Lambda  :745
    X1 = c["d0"] * c["L"][:, 0, 1] + c["d1"] * c["L"][:, 1, 1]
    gm = np.stack((coef * X0, coef * X1), 1)                                 # :743-750
    coefL = np.where(alive, 0.5 * ((o * g) / xsafe) * s - 0.5 * d, 0.0)      # :753-760
    dd = np.stack((c["d0"] * c["d0"], c["d0"] * c["d1"], c["d1"] * c["d0"], c["d1"] * c["d1"]), 1)
    gL = coefL[:, None] * dd
    n = c["n"]
    gid = c["gid"]
    out_m, out_L = _scatter_sum(gid, gm, n), _scatter_sum(gid, gL, n)
    out_o, out_l = _scatter_sum(gid, go[:, None], n), _scatter_sum(gid, gl, n)
    del incl
    if not scales:
        return out_m, out_L.reshape(n, 2, 2), out_o, out_l
    d_abs = np.where(alive, (np.abs(pg) * np.abs(c["p"])).sum(1), 0.0)
    das = d_abs[order]
    tot_a = _scatter_sum(seg, das, nseg)
    csum_a = np.cumsum(das)
    s_abs = (tot_a[seg] - (csum_a - np.r_[0.0, csum_a[:-1]][head][seg]))[inv_order]
    s_abs = np.maximum(s_abs, 0.0)
    go_a = np.where(alive, (g / np.abs(xsafe)) * s_abs + d_abs / np.abs(o), 0.0)
    gl_a = np.where(alive[:, None], d_abs[:, None] / np.abs(l), 0.0)
    coef_a = np.where(alive, ((np.abs(o) * g) / np.abs(xsafe)) * s_abs + d_abs, 0.0)
    gm_a = np.stack((coef_a * np.abs(X0), coef_a * np.abs(X1)), 1)
    gL_a = 0.5 * coef_a[:, None] * np.abs(dd)
    sc_m, sc_L = _scatter_sum(gid, gm_a, n), _scatter_sum(gid, gL_a, n)
    sc_o, sc_l = _scatter_sum(gid, go_a[:, None], n), _scatter_sum(gid, gl_a, n)
    return (out_m, out_L.reshape(n, 2, 2), out_o, out_l), (sc_m, sc_L.reshape(n, 2, 2), sc_o, sc_l)


def image_scale(cache, W, H):
    """Sum of |terms| of every pixel colour (equals the image where all colours are positive)."""
    return _scatter_sum(cache["py"] * (W + 1) + cache["px"], np.abs(cache["p"]),
                        (H + 1) * (W + 1)).reshape(H + 1, W + 1, 3)


def error_table(got, ref, scales, rtol=1e-5, atol=1e-6):
    """Per output: max |err|, max |err| / (atol + rtol |ref|) and max |err| / (atol + rtol scale).  A ratio <= 1
    is a pass of the north star's tolerance in its plain / its condition-aware form."""
    rows = []
    for name, a, b, sc in zip(("image", "grad_mean", "grad_lambda", "grad_opacity", "grad_l"), got, ref, scales):
        a = np.asarray(a, np.float64).reshape(np.asarray(b).shape)
        b = np.asarray(b, np.float64)
        err = np.abs(a - b)
        rows.append((name, float(err.max()) if err.size else 0.0,
                     float((err / (atol + rtol * np.abs(b))).max()) if err.size else 0.0,
                     float((err / (atol + rtol * np.asarray(sc, np.float64).reshape(b.shape))).max()) if err.size else 0.0))
    return rows
