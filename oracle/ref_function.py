"""TEST / BASELINE INFRASTRUCTURE — not product code.

Loader for the reference's OWN compositor, `custom_autograd_grouped_cumprod` (/root/reference/gs_model.py:477-820),
so that it can run UNMODIFIED on the B200 box with a chosen `grouped_cumprod` extension module behind it:

  * `fetch()`   copies gs_model.py and uitility.py (the only two files the Function needs, gs_model.py:7) from
                /root/reference into baseline/_ref/ — git-ignored, NOT gpurun-ignored, so they travel to the GPU box
                like oracle/_ref/*.so but never enter this repository's history.  A no-op where /root/reference
                does not exist (the GPU box: the files are already there).
  * `load()`    imports baseline/_ref/gs_model.py with stub modules for the imports that are missing from the
                reference snapshot or from this image and are not used on this path (`sh_utility` gs_model.py:9,
                `kornia.metrics` :5, `pycolmap` uitility.py:1) and returns the module.  No source edit.
  * `function(ops)`  the Function with `grouped_cumprod` = "ref" (the reference's four native sources compiled
                unchanged by oracle/build_ref.sh -> oracle/_ref/grouped_cumprod_ref.so) or "dropin" (this repo's
                module `grouped_cumprod`, the product).  gs_model.py binds the extension as a module global
                (`import grouped_cumprod`, gs_model.py:8; used at :551, :553), so the swap is one attribute.

Users: tests/test_reference_function.py (-m gpu: the two must agree) and bench.py's reference legs
(`splat_step.reference_function_ms` / `..._with_dropin_ms`).  Nothing under simplegaussiansplat_tk71_b200/ imports this.
"""
from __future__ import annotations

import importlib.util
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference"
REF_DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ("gs_model.py", "uitility.py")

_module = None


def fetch() -> bool:
    """Copy the two reference files into baseline/_ref/ when the reference tree is present.  True if they exist."""
    if os.path.isdir(REF_SRC):
        os.makedirs(REF_DST, exist_ok=True)
        for f in FILES:
            src, dst = os.path.join(REF_SRC, f), os.path.join(REF_DST, f)
            if not os.path.exists(dst) or os.path.getmtime(src) > os.path.getmtime(dst):
                shutil.copyfile(src, dst)
    return all(os.path.exists(os.path.join(REF_DST, f)) for f in FILES)


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DST, f)) for f in FILES)


def reference_ops():
    """oracle/_ref/grouped_cumprod_ref.so (the reference's CUDA ops, built unchanged) or None."""
    path = os.path.join(ROOT, "oracle", "_ref", "grouped_cumprod_ref.so")
    if not os.path.exists(path):
        return None
    if "grouped_cumprod_ref" in sys.modules:
        return sys.modules["grouped_cumprod_ref"]
    import torch  # noqa: F401  (libtorch must be loaded before the extension)

    spec = importlib.util.spec_from_file_location("grouped_cumprod_ref", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["grouped_cumprod_ref"] = mod
    return mod


def load():
    """The reference's gs_model module, imported from baseline/_ref (stubs for its unused, missing imports)."""
    global _module
    if _module is not None:
        return _module
    if not available():
        raise RuntimeError("baseline/_ref/gs_model.py is missing: run oracle.ref_function.fetch() where "
                           "/root/reference exists (__graft_entry__.build() does)")
    for name in ("kornia", "kornia.metrics", "pycolmap", "sh_utility"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["kornia"].metrics = sys.modules["kornia.metrics"]
    if not hasattr(sys.modules["sh_utility"], "eval_sh"):
        sys.modules["sh_utility"].eval_sh = lambda *a, **k: None
    if "grouped_cumprod" not in sys.modules:
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        import grouped_cumprod  # noqa: F401  (the drop-in module at the repo root)
    # import under a private name: `uitility` / `gs_model` must come from baseline/_ref, not from sys.path order
    saved = sys.path[:]
    sys.path.insert(0, REF_DST)
    try:
        spec = importlib.util.spec_from_file_location("gs_model", os.path.join(REF_DST, "gs_model.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["gs_model"] = mod
        spec.loader.exec_module(mod)
    finally:
        sys.path[:] = saved
    _module = mod
    return mod


def function(ops: str = "ref"):
    """The reference Function bound to the chosen extension ops: "ref" or "dropin"."""
    mod = load()
    if ops == "ref":
        ext = reference_ops()
        if ext is None:
            raise RuntimeError("oracle/_ref/grouped_cumprod_ref.so is not built (oracle/build_ref.sh)")
    elif ops == "dropin":
        import grouped_cumprod as ext
    else:
        raise ValueError(ops)
    mod.grouped_cumprod = ext
    return mod.custom_autograd_grouped_cumprod


def run(ops: str, scene, W: int, H: int, grad_image=None, chunks: int = 1):
    """One forward (+ backward when grad_image is given) of the reference Function on CUDA tensors.
    scene = (boxsize, startpoint, endpoint, mean, lam, opacity, l_d).  Returns (image, grads dict or None)."""
    import torch

    F = function(ops)
    boxsize, sp, ep, mean, lam, opac, l_d = scene
    n = boxsize.numel()
    ends = [n] if chunks == 1 else [int(round(n * (i + 1) / chunks)) for i in range(chunks)]
    batch = torch.tensor(ends)
    meanf = mean.detach().clone().float().requires_grad_(grad_image is not None)
    lam_ = lam.detach().clone().requires_grad_(grad_image is not None)
    opac_ = opac.detach().clone().requires_grad_(grad_image is not None)
    l_ = l_d.detach().clone().requires_grad_(grad_image is not None)
    img = F.apply(boxsize, batch, sp, ep, meanf, lam_, opac_, l_, torch.tensor(W), torch.tensor(H))
    if grad_image is None:
        return img, None
    img.backward(grad_image)
    return img.detach(), {"grad_mean": meanf.grad, "grad_lambda": lam_.grad, "grad_opacity": opac_.grad,
                          "grad_l": l_.grad}
