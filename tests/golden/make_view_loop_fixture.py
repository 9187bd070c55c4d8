"""Generates tests/golden/view_loop_fixture.npz by executing the REFERENCE's own per-view loop — the statements of
/root/reference/gs_model.py:399-454 (cull, clamp, boxsize, chunker, compositor call, stack + reshape) — read from the
reference file at generation time (nothing of it is stored in this repository), with
  * its inputs (`*_zsort` tensors of gs_model.py:356-365) built here from a seeded scene,
  * `custom_autograd_grouped_cumprod` = the reference's own Function (imported unmodified, the C oracle standing in
    for its CUDA ops — see make_compositor_fixture.py), wrapped to record the arguments of every call,
  * `Utilities` = the reference's own class (uitility.py), `Utilities.gpu_mem` silenced.
The fixture holds, per view, what the loop handed to the compositor (boxsize, chunk ends, corners, per-view slices)
and the final image batch: row a10 of SURVEY.md §8 pinned to the reference's own code.

Run:  python tests/golden/make_view_loop_fixture.py
"""
import os
import sys
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_compositor_fixture as mk  # noqa: E402

FIRST, LAST = 399, 454          # gs_model.py lines of the loop (inclusive)


def make_inputs(seed, V, n, W, H):
    """z-sorted per-view tensors as gs_model.py:356-365 leaves them: [V, n, ...]."""
    g = torch.Generator().manual_seed(seed)
    z = torch.rand(V, n, generator=g) * 4.0 - 0.4                      # some Gaussians behind the camera
    z, _ = torch.sort(z, dim=1)
    mean_camera = torch.stack((torch.zeros(V, n), torch.zeros(V, n), z), 2)
    mean_pixel = torch.stack((torch.randint(-6, W + 6, (V, n), generator=g),
                              torch.randint(-6, H + 6, (V, n), generator=g)), 2).to(torch.int32)
    box = torch.randint(0, 6, (V, n, 2), generator=g).to(torch.int32)  # some zero-width boxes
    A = torch.randn(V, n, 2, 2, generator=g) * 0.3
    lam = A @ A.transpose(2, 3) + 0.05 * torch.eye(2)
    opac = torch.sigmoid(torch.randn(V, n, 1, generator=g) * 1.432 + 1.735)
    l_d = torch.rand(V, n, 3, generator=g) * 0.9 + 0.05
    z_inv = torch.argsort(torch.argsort(torch.rand(V, n, generator=g), dim=1), dim=1)
    return mean_camera, mean_pixel, box, lam, opac, l_d, z_inv


def run_reference_loop(gs_model, inputs, W, H, empty_view=None):
    mean_camera, mean_pixel, box, lam, opac, l_d, z_inv = inputs
    V, n = mean_pixel.shape[:2]
    if empty_view is not None:
        mean_camera = mean_camera.clone()
        mean_camera[empty_view, :, 2] = -1.0
    lines = open(os.path.join(mk.REF, "gs_model.py"), encoding="utf-8").read().split("\n")[FIRST - 1:LAST]
    src = textwrap.dedent("\n".join(lines))
    calls = []
    Fref = gs_model.custom_autograd_grouped_cumprod

    class Recorder:
        @staticmethod
        def apply(*args):
            calls.append([a.clone() if torch.is_tensor(a) else a for a in args])
            return Fref.apply(*[a.to(torch.float32) if (torch.is_tensor(a) and i == 4) else a
                                for i, a in enumerate(args)])

    util = gs_model.Utilities
    util.gpu_mem = staticmethod(lambda *a, **k: None)
    env = dict(torch=torch, Utilities=util, custom_autograd_grouped_cumprod=Recorder, shape_image=V,
               shape_gausian=n, shape_width=torch.tensor(W), shape_height=torch.tensor(H), mean_camera_zsort=mean_camera,
               gausian_boxsize_zsort=box, mean_pixel_zsort=mean_pixel, L_d_zsort=l_d, opacity_zsort=opac,
               variance_inverse_zsort=lam, z_inverse_index=z_inv, image_sample=[f"img{v}" for v in range(V)])
    with mk.cuda_literals_to_cpu(), torch.no_grad():
        exec(compile(src, "gs_model.py[399:454]", "exec"), env)
    return calls, env["pixel_image_batch"], env["image_sample"], env["grad_iter"]


def main():
    gs_model = mk.import_reference()
    out = {}
    for tag, (seed, V, n, W, H, empty) in {"a": (3, 3, 60, 40, 28, None), "b": (4, 4, 150, 64, 36, 2)}.items():
        inputs = make_inputs(seed, V, n, W, H)
        calls, images, samples, grad_iter = run_reference_loop(gs_model, inputs, W, H, empty)
        names = ("mean_camera", "mean_pixel", "box", "lam", "opac", "l_d", "z_inv")
        for k, v in zip(names, inputs):
            out[f"{tag}/in/{k}"] = v.numpy()
        out[f"{tag}/WH"] = np.array([W, H])
        out[f"{tag}/empty_view"] = np.array([-1 if empty is None else empty])
        out[f"{tag}/n_calls"] = np.array([len(calls)])
        out[f"{tag}/images"] = images.numpy()
        out[f"{tag}/kept_samples"] = np.array([int(s_[3:]) for s_ in samples if s_ != ""])   # gs_model.py:456
        out[f"{tag}/grad_iter"] = grad_iter.numpy()
        for c, args in enumerate(calls):
            for k, a in zip(("boxsize", "batch", "sp", "ep", "mean", "lam", "opac", "l_d"), args[:8]):
                out[f"{tag}/call{c}/{k}"] = a.numpy()
        print(tag, "views", V, "calls", len(calls), "images", tuple(images.shape), "kept", samples)
    path = os.path.join(HERE, "view_loop_fixture.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
