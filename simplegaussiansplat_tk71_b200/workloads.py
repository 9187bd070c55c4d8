"""Synthetic element lists for the compositing scan (BASELINE.json configs).

An *element* is one Gaussian x pixel pair after box expansion; a *segment* is the
depth-sorted list of elements of one pixel (reference: gs_model.py:544-566).  The
generators below produce exactly what the reference hands to its ops at
gs_model.py:551/:553 and cuda_test.py:23/:29:

  x        f32[N]  "anti opacity" 1 - o*g  (gs_model.py:533-535)
  key      i32[N]  pixel key y*10000 + x   (gs_model.py:538-541), sorted, equal inside a segment
  inv      i32[N]  dense segment id 0..K-1 (cuda_test.py:21)
  seg_end  i32[K]  exclusive end offset of each segment (cuda_test.py:27)
  grad_out f32[N]  upstream gradient

Segment lengths come from numpy (K values); per-element values from a seeded torch
generator on the target device, so the 4K config never has to cross PCIe.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

# fit of the bundled opacity.pt logits
OPACITY_LOGIT_MEAN = 1.735
OPACITY_LOGIT_STD = 1.432


@dataclass
class ElementList:
    name: str
    x: torch.Tensor
    key: torch.Tensor
    inv: torch.Tensor
    seg_end: torch.Tensor
    grad_out: torch.Tensor
    width: int
    height: int

    @property
    def n(self) -> int:
        return int(self.x.numel())

    @property
    def k(self) -> int:
        return int(self.seg_end.numel())

    def to(self, device) -> "ElementList":
        return ElementList(self.name, *(t.to(device) for t in
                                        (self.x, self.key, self.inv, self.seg_end, self.grad_out)),
                           self.width, self.height)


def lengths_c1(n: int = 1 << 20, k: int = 1 << 16, seed: int = 0) -> np.ndarray:
    """C1: lognormal(ln 8, 1.0) scaled to sum n, floored, min 1, remainder +1 over the first segments."""
    rng = np.random.default_rng(seed)
    raw = rng.lognormal(math.log(8.0), 1.0, k)
    L = np.maximum(1, np.floor(raw * (n / raw.sum()))).astype(np.int64)
    diff = n - int(L.sum())
    if diff > 0:
        q, r = divmod(diff, k)
        L += q
        L[:r] += 1
    elif diff < 0:
        # take the excess back from the longest segments, never below 1
        order = np.argsort(-L)
        i = 0
        while diff < 0:
            j = order[i % k]
            if L[j] > 1:
                L[j] -= 1
                diff += 1
            i += 1
    assert int(L.sum()) == n and L.min() >= 1
    return L


def lengths_lognormal(k: int, median: float, seed: int) -> np.ndarray:
    """C3/C4 body: L = max(1, round(lognormal(ln median, 1.0)))."""
    rng = np.random.default_rng(seed)
    return np.maximum(1, np.rint(rng.lognormal(math.log(median), 1.0, k))).astype(np.int64)


def lengths_c4(k: int, seed: int = 2160, deep: int = 512,
               deep_lo: int = 8192, deep_hi: int = 262144):
    """C4: lognormal(ln 30) body plus `deep` segments U[deep_lo, deep_hi] at rng-chosen pixels."""
    rng = np.random.default_rng(seed)
    L = np.maximum(1, np.rint(rng.lognormal(math.log(30.0), 1.0, k))).astype(np.int64)
    pos = rng.choice(k, size=min(deep, k), replace=False)
    L[pos] = rng.integers(deep_lo, deep_hi + 1, size=pos.size)
    is_deep = np.zeros(k, dtype=bool)
    is_deep[pos] = True
    return L, is_deep


def pixel_keys(k: int, width: int) -> np.ndarray:
    """key = y*10000 + x for the k-th pixel in row-major order (gs_model.py:541)."""
    p = np.arange(k, dtype=np.int64)
    return ((p // width) * 10000 + (p % width)).astype(np.int32)


def build(name: str, lengths: np.ndarray, width: int, height: int, seed: int,
          device="cpu", alpha_scale_per_seg: np.ndarray | None = None,
          zero_frac: float = 0.0) -> ElementList:
    """Expand per-segment lengths into the five op inputs on `device`."""
    k = int(lengths.shape[0])
    n = int(lengths.sum())
    assert n < 2 ** 31, "reference ops index with int32 (grouped_cumprod_backward.cu:52)"
    dev = torch.device(device)
    L = torch.from_numpy(lengths).to(dev)
    seg_end64 = torch.cumsum(L, 0)
    inv = torch.repeat_interleave(torch.arange(k, dtype=torch.int32, device=dev), L,
                                  output_size=n)
    pk = torch.from_numpy(pixel_keys(k, width)).to(dev)
    key = pk[inv.long()] if k > 0 else torch.empty(0, dtype=torch.int32, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    logit = torch.empty(n, dtype=torch.float32, device=dev).normal_(
        OPACITY_LOGIT_MEAN, OPACITY_LOGIT_STD, generator=gen)
    o = torch.sigmoid(logit)
    del logit
    u = torch.empty(n, dtype=torch.float32, device=dev).uniform_(0.0, 1.0, generator=gen)
    alpha = o * torch.exp(-4.5 * u)
    del o, u
    if alpha_scale_per_seg is not None:
        sc = torch.from_numpy(alpha_scale_per_seg.astype(np.float32)).to(dev)
        alpha = alpha * sc[inv.long()]
    x = (1.0 - alpha).contiguous()
    del alpha
    if zero_frac > 0.0 and n > 0:
        m = torch.empty(n, dtype=torch.float32, device=dev).uniform_(0.0, 1.0, generator=gen) < zero_frac
        x[m] = 0.0
    grad_out = torch.empty(n, dtype=torch.float32, device=dev).normal_(0.0, 1.0, generator=gen)
    return ElementList(name, x, key.contiguous(), inv.contiguous(),
                       seg_end64.to(torch.int32).contiguous(), grad_out, width, height)


def c1(device="cpu", zeros: bool = False) -> ElementList:
    """BASELINE.json configs[0]: 1 Mi elements / 64 Ki heavy-tailed segments."""
    L = lengths_c1()
    return build("C1 1Mi/64Ki lognormal(ln8,1)", L, 256, 256, 0, device,
                 zero_frac=1e-5 if zeros else 0.0)


def c3(device="cpu", view: int = 0, scale: float = 1.0) -> ElementList:
    """BASELINE.json configs[2], scan-only route: 1920x1080 pixels, lognormal(ln 20, 1) lengths.

    `scale` < 1 shrinks the pixel grid (rows) for CPU-sized parity cases; `view` shifts the seed
    (C5: seed = 1080 + view id).
    """
    w, h = 1920, max(1, int(round(1080 * scale)))
    L = lengths_lognormal(w * h, 20.0, 1080 + view)
    return build(f"C3 {w}x{h} lognormal(ln20,1) view{view}", L, w, h, 1080 + view, device)


def c4(device="cpu", scale: float = 1.0, deep: int = 512) -> ElementList:
    """BASELINE.json configs[3], scan-only route: 3840x2160 + `deep` long segments (look-back path)."""
    w, h = 3840, max(1, int(round(2160 * scale)))
    L, is_deep = lengths_c4(w * h, 2160, deep=max(1, int(round(deep * scale))) if scale < 1 else deep)
    sc = np.where(is_deep, 1e-3, 1.0)
    return build(f"C4 {w}x{h} lognormal(ln30,1)+deep", L, w, h, 2160, device, alpha_scale_per_seg=sc)


def algorithmic_bytes(n: int, k: int) -> dict:
    """fwd 12 B/elem, bwd 16 B/elem + 4 B/segment."""
    return {"fwd": 12 * n, "bwd": 16 * n + 4 * k, "fwd_bwd": 28 * n + 4 * k}


# --------------------------------------------------------------------------------------------
# Splat route (SURVEY.md §8d, C3/C4 "splat route"): per-Gaussian inputs of the compositor
# `custom_autograd_grouped_cumprod.apply(boxsize, batch, startpoint, endpoint, mean, Lambda, opacity, l_d, W, H)`
# exactly as gs_model.py:405-449 prepares them (depth order = index order, integer pixel means,
# inclusive box corners clamped to [0,W]x[0,H]).
# --------------------------------------------------------------------------------------------
@dataclass
class SplatView:
    name: str
    boxsize: torch.Tensor      # i64[n]
    startpoint: torch.Tensor   # i32[n,2]
    endpoint: torch.Tensor     # i32[n,2]
    mean: torch.Tensor         # i32[n,2] (mean_pixel is cast to int32 at gs_model.py:361)
    lam: torch.Tensor          # f32[n,2,2]
    opacity: torch.Tensor      # f32[n,1]
    l_d: torch.Tensor          # f32[n,3]
    width: int
    height: int
    index: torch.Tensor = None  # i32[n] row of every visible Gaussian in the scene's parameter arrays (the
    #                             reference's boolean-mask selection, gs_model.py:405-413, as row numbers), or None

    @property
    def n(self) -> int:
        return int(self.boxsize.numel())

    @property
    def elements(self) -> int:
        # cached: the reduction + .item() is a full device sync, which a timed loop must not pay per view
        c = self.__dict__.get("_elements")
        if c is None:
            c = int(self.boxsize.sum().item())
            self.__dict__["_elements"] = c
        return c


def splat_view(width: int = 1920, height: int = 1080, n: int = 1_000_000, seed: int = 1080, device="cpu",
               clusters: int = 256) -> SplatView:
    rng = np.random.default_rng(seed)
    n_cl = int(0.7 * n)
    cc = np.stack((rng.uniform(0, width, clusters), rng.uniform(0, height, clusters)), 1)
    sig = 0.05 * min(width, height)
    which = rng.integers(0, clusters, n_cl)
    c1 = cc[which] + rng.normal(0.0, sig, (n_cl, 2))
    c2 = np.stack((rng.uniform(-0.05 * width, 1.05 * width, n - n_cl),
                   rng.uniform(-0.05 * height, 1.05 * height, n - n_cl)), 1)
    ctr = np.concatenate((c1, c2), 0)
    rng.shuffle(ctr, axis=0)                                   # depth order is independent of position
    hw_max = 10 * 0.04 * math.sqrt(width * height)             # gs_model.py:364-365
    hw = np.clip(np.floor(rng.lognormal(math.log(2.0), 0.9, (n, 2))), 1, hw_max)
    mean = np.floor(ctr).astype(np.int64)
    vis = (mean[:, 0] - hw[:, 0] < width) & (mean[:, 0] + hw[:, 0] > 0) & \
          (mean[:, 1] - hw[:, 1] < height) & (mean[:, 1] + hw[:, 1] > 0)       # gs_model.py:406
    mean, hw = mean[vis], hw[vis].astype(np.int64)
    sp = np.stack((np.clip(mean[:, 0] - hw[:, 0], 0, width), np.clip(mean[:, 1] - hw[:, 1], 0, height)), 1)
    ep = np.stack((np.clip(mean[:, 0] + hw[:, 0], 0, width), np.clip(mean[:, 1] + hw[:, 1], 0, height)), 1)
    boxsize = np.prod(ep - sp + 1, axis=1)
    m = mean.shape[0]
    sx = np.maximum(hw[:, 0] / 3.0, 0.5)                       # the box is the 3-sigma extent (gs_model.py:327-332)
    sy = np.maximum(hw[:, 1] / 3.0, 0.5)
    rho = rng.uniform(-0.6, 0.6, m)
    det = (1 - rho * rho)
    lam = np.zeros((m, 2, 2), np.float32)
    lam[:, 0, 0] = 1.0 / (sx * sx * det)
    lam[:, 1, 1] = 1.0 / (sy * sy * det)
    lam[:, 0, 1] = lam[:, 1, 0] = -rho / (sx * sy * det)
    opac = 1.0 / (1.0 + np.exp(-rng.normal(OPACITY_LOGIT_MEAN, OPACITY_LOGIT_STD, (m, 1))))
    l_d = rng.uniform(0.05, 0.95, (m, 3))
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dt)  # noqa: E731
    return SplatView(f"splat {width}x{height} n={m} seed{seed}", t(boxsize, torch.int64), t(sp, torch.int32),
                     t(ep, torch.int32), t(mean, torch.int32), t(lam, torch.float32), t(opac, torch.float32),
                     t(l_d, torch.float32), width, height)


def splat_view_device(width: int = 1920, height: int = 1080, n: int = 1_000_000, seed: int = 1080, device="cuda",
                      clusters: int = 256) -> SplatView:
    """The same view distribution as splat_view (SURVEY.md §8d, C3 splat route), drawn with torch's generator ON
    the device: a 64-view training batch (C5: seed = 1080 + view id) is ready in milliseconds instead of minutes of
    host numpy.  Also returns `index`, the parameter row of every visible Gaussian (rows of the n-Gaussian scene)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    U = lambda *shape: torch.rand(*shape, generator=g, device=dev)      # noqa: E731
    N = lambda *shape: torch.randn(*shape, generator=g, device=dev)     # noqa: E731
    n_cl = int(0.7 * n)
    cc = torch.stack((U(clusters) * width, U(clusters) * height), 1)
    which = torch.randint(0, clusters, (n_cl,), generator=g, device=dev)
    c1 = cc[which] + N(n_cl, 2) * (0.05 * min(width, height))
    c2 = torch.stack(((U(n - n_cl) * 1.1 - 0.05) * width, (U(n - n_cl) * 1.1 - 0.05) * height), 1)
    ctr = torch.cat((c1, c2), 0)[torch.randperm(n, generator=g, device=dev)]   # depth order independent of position
    hw_max = 10 * 0.04 * math.sqrt(width * height)                             # gs_model.py:364-365
    hw = torch.exp(N(n, 2) * 0.9 + math.log(2.0)).floor().clamp(1, hw_max)
    mean = ctr.floor().to(torch.int64)
    hwi = hw.to(torch.int64)
    vis = (mean[:, 0] - hwi[:, 0] < width) & (mean[:, 0] + hwi[:, 0] > 0) & \
          (mean[:, 1] - hwi[:, 1] < height) & (mean[:, 1] + hwi[:, 1] > 0)     # gs_model.py:406
    index = torch.nonzero(vis).reshape(-1).to(torch.int32)
    mean, hwi, hw = mean[vis], hwi[vis], hw[vis]
    sp = torch.stack(((mean[:, 0] - hwi[:, 0]).clamp(0, width), (mean[:, 1] - hwi[:, 1]).clamp(0, height)), 1)
    ep = torch.stack(((mean[:, 0] + hwi[:, 0]).clamp(0, width), (mean[:, 1] + hwi[:, 1]).clamp(0, height)), 1)
    boxsize = torch.prod(ep - sp + 1, dim=1)
    m = mean.shape[0]
    sx = (hw[:, 0] / 3.0).clamp(min=0.5)                                       # the box is the 3-sigma extent
    sy = (hw[:, 1] / 3.0).clamp(min=0.5)
    rho = U(m) * 1.2 - 0.6
    det = 1 - rho * rho
    lam = torch.empty((m, 2, 2), dtype=torch.float32, device=dev)
    lam[:, 0, 0] = 1.0 / (sx * sx * det)
    lam[:, 1, 1] = 1.0 / (sy * sy * det)
    lam[:, 0, 1] = lam[:, 1, 0] = -rho / (sx * sy * det)
    opac = torch.sigmoid(N(m, 1) * OPACITY_LOGIT_STD + OPACITY_LOGIT_MEAN)
    l_d = U(m, 3) * 0.9 + 0.05
    return SplatView(f"splat {width}x{height} n={m} seed{seed} (device rng)", boxsize, sp.to(torch.int32),
                     ep.to(torch.int32), mean.to(torch.int32), lam, opac, l_d, width, height, index)


# --------------------------------------------------------------------------------------------
# C2 — BASELINE.json configs[1]: the scene bundled with the reference (trained opacity.pt, COLMAP points and
# intrinsics; data/bundled_scene.npz, derived by tools/make_bundled_scene.py).  mean.pt / color.pt / images.bin
# are missing from the reference snapshot, so the rest is synthesised deterministically as SURVEY.md §8d
# prescribes, and the projection follows gs_model.py:289-365 for isotropic Gaussians.
# --------------------------------------------------------------------------------------------
def bundled_views(device="cpu", n_views: int = 3):
    """Returns a list of SplatView (one per synthesised pose), Gaussians already z-sorted, culled and clamped."""
    import os

    from . import views as vw

    f = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "bundled_scene.npz"))
    logits = f["opacity_logits"].astype(np.float32)
    pts, nn3 = f["points"], f["nn3"]
    fx, fy, cx, cy, W, H = f["intrinsics"]
    W, H = int(W), int(H)
    n = logits.shape[0]
    rng = np.random.default_rng(0)
    reps = -(-n // pts.shape[0])
    mean = (np.tile(pts, (reps, 1))[:n] + rng.normal(0.0, 0.05, (n, 3))).astype(np.float32)   # tiled + jitter
    scale = (np.tile(nn3, reps)[:n] / 4.0).astype(np.float32)        # exp(variance_scale): mean 3-NN distance / 4
    centroid = pts.mean(0)
    out = []
    dev = torch.device(device)
    for v in range(n_views):
        th = 2 * math.pi * v / n_views
        cam = centroid + 6.0 * np.array([math.cos(th), 0.0, math.sin(th)])
        fwd = (centroid - cam) / np.linalg.norm(centroid - cam)
        right = np.cross(fwd, np.array([0.0, -1.0, 0.0]))
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        R = np.stack((right, down, fwd), 0).astype(np.float32)      # world -> camera, z forward
        pc = (mean - cam.astype(np.float32)) @ R.T                   # gs_model.py:289-290
        z = pc[:, 2]
        zc = np.maximum(z, 1e-2)                                     # :294 clamp_min(1e-2)
        px = fx * pc[:, 0] / zc + cx
        py = fy * pc[:, 1] / zc + cy
        # Sigma_px = s^2 J J^T + 1e-6 I with J = [[fx/z, 0, -fx X/z^2], [0, fy/z, -fy Y/z^2]]   (:308-321)
        j00, j02 = fx / zc, -fx * pc[:, 0] / (zc * zc)
        j11, j12 = fy / zc, -fy * pc[:, 1] / (zc * zc)
        s2 = scale * scale
        a = s2 * (j00 * j00 + j02 * j02) + 1e-6
        b = s2 * (j02 * j12)
        d = s2 * (j11 * j11 + j12 * j12) + 1e-6
        det = a * d - b * b
        lam = np.stack((d / det, -b / det, -b / det, a / det), 1).reshape(-1, 2, 2).astype(np.float32)   # :353
        half = 3.0 * np.sqrt(np.stack((a, d), 1))                    # :332  3*sqrt(V^2 |lambda|) = 3*sqrt(diag)
        half = np.minimum(half, 10 * 0.04 * math.sqrt(W * H)).astype(np.int32)                            # :364-365
        order = np.argsort(z, kind="stable")                        # :356 z-sort
        t = lambda arr, dt: torch.from_numpy(np.ascontiguousarray(arr[order])).to(device=dev, dtype=dt)  # noqa: E731
        mean_px = t(np.stack((px, py), 1).clip(-2e6, 2e6), torch.float32).to(torch.int32)                # :361 int32 cast
        mask, sp, ep, boxsize = vw.visible_boxes(mean_px, t(half, torch.int32), t(z, torch.float32), W, H)
        opac = torch.sigmoid(t(logits[:, None], torch.float32))[mask]
        m = int(mask.sum())
        out.append(SplatView(f"C2 bundled scene (poses synthesised) view{v} {W}x{H} n={m}", boxsize, sp, ep,
                             mean_px[mask], t(lam, torch.float32)[mask], opac,
                             torch.full((m, 3), 0.499, dtype=torch.float32, device=dev), W, H))
    return out
