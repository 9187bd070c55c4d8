"""Error table of the native compositor against the fp64 oracle (oracle/compositor_oracle.py): for every output
of every scene and route, max |err|, max |err| / (1e-6 + 1e-5 |ref|) ("plain") and max |err| / (1e-6 + 1e-5 scale)
("cond", scale = sum of |terms|, oracle.backward(scales=True)).  A ratio <= 1 passes the north star's tolerance.
GPU tool (runs the product path); the table goes to DESIGN.md §6 and profiles/.

  python tools/compositor_errors.py [--big]     # --big adds a 1080p / 200 k-Gaussian view
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from oracle import compositor_oracle as co  # noqa: E402
from simplegaussiansplat_tk71_b200 import compositor, workloads as wl  # noqa: E402
from test_compositor import _run  # noqa: E402
from test_compositor_oracle import CASES, FIX, load_case  # noqa: E402


def oracle_case(case):
    img, cache = co.forward(case["boxsize"], case["sp"], case["ep"], case["mean"], case["lam"], case["opac"],
                            case["l_d"], case["W"], case["H"])
    grads, sc = co.backward(cache, case["grad_image"], scales=True)
    return (img,) + tuple(grads), (co.image_scale(cache, case["W"], case["H"]),) + tuple(sc)


def scene_case(sc, k=None, seed=9):
    k = sc.n if k is None else k
    rng = np.random.default_rng(seed)
    c = lambda t: t[:k].cpu().numpy()  # noqa: E731
    return dict(boxsize=c(sc.boxsize), sp=c(sc.startpoint), ep=c(sc.endpoint), mean=c(sc.mean), lam=c(sc.lam),
                opac=c(sc.opacity), l_d=c(sc.l_d), W=sc.width, H=sc.height,
                grad_image=rng.uniform(0.1, 1.0, (sc.height + 1, sc.width + 1, 3)).astype(np.float32))


def main():
    import make_compositor_fixture as mk

    f = np.load(FIX)
    scenes = [(f"fixture/{n}", load_case(f, n)) for n in CASES]
    b, sp, ep, mean, lam, opac, l_d = mk.make_scene(11, 160, 120, 4000, 9, opaque=20)
    scenes.append(("160x120 n=4000", dict(boxsize=b.numpy(), sp=sp.numpy(), ep=ep.numpy(), mean=mean.numpy(),
                                          lam=lam.numpy(), opac=opac.numpy(), l_d=l_d.numpy(), W=160, H=120,
                                          grad_image=np.random.default_rng(5).uniform(0.1, 1.0, (121, 161, 3)).astype(np.float32))))
    scenes.append(("C2 view1 front 4000", scene_case(wl.bundled_views("cpu", n_views=2)[1], 4000)))
    if "--big" in sys.argv:
        scenes.append(("1080p n=200k", scene_case(wl.splat_view(1920, 1080, 200_000, seed=1080, device="cpu"))))
    print(f"{'scene':22s} {'route':6s} {'output':13s} {'max|err|':>10s} {'plain':>8s} {'cond':>8s}")
    worst = {}
    for name, case in scenes:
        ref, scales = oracle_case(case)
        for route in ("tiles", "lists"):
            compositor.ROUTE = route
            got = _run(case, "cuda")
            for out, e, plain, cond in co.error_table(got, ref, scales):
                print(f"{name:22s} {route:6s} {out:13s} {e:10.3e} {plain:8.2f} {cond:8.2f}")
                worst[(route, out)] = max(worst.get((route, out), 0.0), cond)
    print("worst condition-aware ratio per (route, output):", {f"{r}/{o}": round(v, 3) for (r, o), v in worst.items()})


if __name__ == "__main__":
    main()
