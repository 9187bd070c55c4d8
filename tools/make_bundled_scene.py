"""Derives simplegaussiansplat_tk71_b200/data/bundled_scene.npz from the data files bundled with the reference
(BASELINE.json configs[1], SURVEY.md §8d "C2"):

  /root/reference/opacity.pt                      trained opacity logits, [514361, 1] f32 (a pickled nn.Parameter)
  /root/reference/colmap/sparse/0/points3D.bin    10 409 COLMAP points (xyz)
  /root/reference/colmap/sparse/0/cameras.bin     100 OPENCV cameras, 640x427 (intrinsics of camera 1 are used)

mean.pt / color.pt / images.bin are missing from the reference snapshot (.MISSING_LARGE_BLOBS), so the scene
is completed deterministically at load time (workloads.bundled_views): means = the COLMAP points tiled with
jitter, scale = mean 3-NN distance / 4, identity rotation, constant colour, three synthesised look-at poses.
Stored: the opacity logits (f16: a benchmark input, +-1e-3 is irrelevant), the points, their mean 3-NN
distance (uitility.py:68-78), and the intrinsics.  Run:  python tools/make_bundled_scene.py
"""
import os
import struct

import numpy as np
import torch
from scipy.spatial import cKDTree

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read_points3d(path):
    out = []
    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        for _ in range(n):
            f.read(8)                                   # point id
            out.append(struct.unpack("<3d", f.read(24)))
            f.read(3 + 8)                               # rgb, error
            (tl,) = struct.unpack("<Q", f.read(8))
            f.read(8 * tl)                              # track
    return np.asarray(out, dtype=np.float32)


def read_camera1(path):
    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        cid, model, w, h = struct.unpack("<IiQQ", f.read(24))
        nparams = {0: 3, 1: 4, 2: 4, 3: 5, 4: 8}[model]   # SIMPLE_PINHOLE, PINHOLE, SIMPLE_RADIAL, RADIAL, OPENCV
        p = struct.unpack(f"<{nparams}d", f.read(8 * nparams))
    fx, fy, cx, cy = (p[0], p[1], p[2], p[3]) if model in (1, 4) else (p[0], p[0], p[1], p[2])
    return np.array([fx, fy, cx, cy, w, h], dtype=np.float64), n


op = torch.load(os.path.join(REF, "opacity.pt"), map_location="cpu", weights_only=True)
logits = op.detach().reshape(-1).numpy().astype(np.float16)
pts = read_points3d(os.path.join(REF, "colmap", "sparse", "0", "points3D.bin"))
intr, ncam = read_camera1(os.path.join(REF, "colmap", "sparse", "0", "cameras.bin"))
d, _ = cKDTree(pts).query(pts, k=4)                     # self + 3 nearest neighbours
nn3 = d[:, 1:].mean(1).astype(np.float32)
out = os.path.join(ROOT, "simplegaussiansplat_tk71_b200", "data", "bundled_scene.npz")
np.savez_compressed(out, opacity_logits=logits, points=pts, nn3=nn3, intrinsics=intr)
print("wrote", out, os.path.getsize(out), "bytes;", logits.shape, pts.shape, "cameras", ncam, "intrinsics", intr,
      "logit mean/std", float(logits.astype(np.float32).mean()), float(logits.astype(np.float32).std()))
