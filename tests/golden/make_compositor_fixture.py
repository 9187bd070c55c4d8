"""Generates tests/golden/compositor_fixture.npz by running the REFERENCE's own compositor
`custom_autograd_grouped_cumprod` (/root/reference/gs_model.py:477-820) — unmodified, imported from where
it lies — forward and backward on small seeded scenes, on the CPU of THIS container.

How a CUDA-only, not-importable-as-shipped module is made to run here without editing it:
  * stub modules for imports that are missing / unused on this path: `sh_utility` (absent from the
    reference repo, gs_model.py:9), `kornia`, `kornia.metrics`, `pycolmap` (not installed);
  * a stand-in module named `grouped_cumprod` whose two forward ops are the CPU oracle
    (oracle/gcp_oracle.c, fp32 sequential) — the reference calls them at gs_model.py:551/:553;
  * the `device="cuda"` literals (gs_model.py:505,:691,:728,:771,:792) are redirected to the CPU by wrapping
    the torch factory functions for the duration of the call;
  * `torch.sort(inv)` at gs_model.py:547 asks for no stability and gets it on CUDA only because the CUDA
    radix sort is stable — the per-pixel DEPTH ORDER depends on it.  torch's CPU sort is
    not stable, so the call is wrapped to pass stable=True: that reproduces what the reference does on a GPU.

Run:  python tests/golden/make_compositor_fixture.py
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

REF = "/root/reference"


def import_reference():
    for name in ("kornia", "kornia.metrics", "pycolmap", "sh_utility"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["kornia"].metrics = sys.modules["kornia.metrics"]
    sys.modules["sh_utility"].eval_sh = lambda *a, **k: None
    gc = types.ModuleType("grouped_cumprod")

    def _fwd(x, key, y):
        y.copy_(torch.from_numpy(orc.cumprod_fwd(x.numpy(), key.numpy(), np.float32)))

    def _sum(x, key, y):
        y.copy_(torch.from_numpy(orc.cumsum_fwd(x.numpy(), key.numpy(), np.float32)))

    gc.grouped_cumprod_forward = _fwd
    gc.grouped_cumsum_forward = _sum
    sys.modules["grouped_cumprod"] = gc
    sys.path.insert(0, REF)
    import gs_model

    return gs_model


@contextlib.contextmanager
def cuda_literals_to_cpu():
    names = ["zeros", "arange", "tensor", "ones", "empty", "full"]
    saved = {n: getattr(torch, n) for n in names}

    def wrap(fn):
        def f(*a, **k):
            if k.get("device", None) == "cuda":
                k["device"] = "cpu"
            return fn(*a, **k)

        return f

    for n in names:
        setattr(torch, n, wrap(saved[n]))
    sort_orig = torch.sort

    def stable_sort(*a, **k):
        k.setdefault("stable", True)
        return sort_orig(*a, **k)

    torch.sort = stable_sort
    try:
        yield
    finally:
        torch.sort = sort_orig
        for n in names:
            setattr(torch, n, saved[n])


def make_scene(seed, W, H, n, max_half, opaque=0):
    """Depth-sorted (= index order) Gaussians with integer pixel means, as gs_model.py:356-425 hands them over."""
    g = torch.Generator().manual_seed(seed)
    cx = torch.randint(0, W, (n,), generator=g)
    cy = torch.randint(0, H, (n,), generator=g)
    hw = torch.randint(1, max_half + 1, (n, 2), generator=g)
    mean = torch.stack((cx, cy), 1).to(torch.int32)
    sp = torch.stack(((cx - hw[:, 0]).clamp(0, W), (cy - hw[:, 1]).clamp(0, H)), 1).to(torch.int32)
    ep = torch.stack(((cx + hw[:, 0]).clamp(0, W), (cy + hw[:, 1]).clamp(0, H)), 1).to(torch.int32)
    boxsize = torch.prod((ep - sp + 1).to(torch.int64), dim=1)
    A = torch.randn(n, 2, 2, generator=g) * 0.3
    lam = A @ A.transpose(1, 2) + 0.05 * torch.eye(2)
    opac = torch.sigmoid(torch.randn(n, 1, generator=g) * 1.432 + 1.735)
    if opaque:
        opac[torch.randperm(n, generator=g)[:opaque]] = 1.0   # alpha == 1 exactly at the centre pixel
    l_d = torch.rand(n, 3, generator=g) * 0.9 + 0.05
    return boxsize, sp, ep, mean, lam, opac, l_d


def run_reference(F, scene, W, H, chunks=1, seed=0):
    boxsize, sp, ep, mean, lam, opac, l_d = scene
    n = boxsize.numel()
    ends = [n] if chunks == 1 else [int(round(n * (i + 1) / chunks)) for i in range(chunks)]
    batch = torch.tensor(ends)
    meanf = mean.clone().float().requires_grad_(True)
    lam = lam.clone().requires_grad_(True)
    opac = opac.clone().requires_grad_(True)
    l_d = l_d.clone().requires_grad_(True)
    with cuda_literals_to_cpu():
        img = F.apply(boxsize, batch, sp, ep, meanf, lam, opac, l_d, torch.tensor(W), torch.tensor(H))
        g = torch.Generator().manual_seed(1000 + seed)
        gI = torch.rand(img.shape, generator=g) * 0.9 + 0.1     # strictly positive: the reference's
        (img * gI).sum().backward()                              # "drop sums == 0" quirk stays inactive
    return {"image": img.detach().numpy(), "grad_image": gI.numpy(), "grad_mean": meanf.grad.numpy(),
            "grad_lambda": lam.grad.numpy(), "grad_opacity": opac.grad.numpy(), "grad_l": l_d.grad.numpy()}


def main():
    gs_model = import_reference()
    F = gs_model.custom_autograd_grouped_cumprod
    cases = {
        "small": dict(seed=1, W=24, H=16, n=12, max_half=4),
        "dense": dict(seed=2, W=40, H=30, n=300, max_half=6),       # ~30 Gaussians per pixel
        "wide": dict(seed=3, W=64, H=48, n=80, max_half=20),        # big boxes, clamped at the borders
        "opaque": dict(seed=4, W=32, H=24, n=120, max_half=5, opaque=10),  # alpha == 1: zero-dropping path
    }
    out = {}
    for name, c in cases.items():
        scene = make_scene(c["seed"], c["W"], c["H"], c["n"], c["max_half"], c.get("opaque", 0))
        res = run_reference(F, scene, c["W"], c["H"], 1, c["seed"])
        boxsize, sp, ep, mean, lam, opac, l_d = scene
        for k, v in dict(boxsize=boxsize, startpoint=sp, endpoint=ep, mean=mean, lam=lam, opacity=opac, l_d=l_d).items():
            out[f"{name}/{k}"] = v.numpy()
        out[f"{name}/WH"] = np.array([c["W"], c["H"]])
        for k, v in res.items():
            out[f"{name}/{k}"] = v
        print(name, "N =", int(boxsize.sum()), "image sum", float(res["image"].sum()))
    # the same "dense" scene rendered in 3 chunks: documents the reference's chunk-boundary carry
    c = cases["dense"]
    scene = make_scene(c["seed"], c["W"], c["H"], c["n"], c["max_half"])
    res3 = run_reference(F, scene, c["W"], c["H"], 3, c["seed"])
    out["dense/image_3chunks"] = res3["image"]
    path = os.path.join(ROOT, "tests", "golden", "compositor_fixture.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
