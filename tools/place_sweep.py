import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from simplegaussiansplat_tk71_b200 import workloads as wl, compositor, _lib
v = wl.splat_view(1920, 1080, 1_000_000, device="cuda")
L = _lib.lib()
def t_fwd(reps=5):
    ts=[]
    for _ in range(reps+2):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); compositor._render_forward(v.boxsize, v.startpoint, v.endpoint, v.mean.float(), v.lam, v.opacity, v.l_d, v.width, v.height); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts=sorted(ts[2:]); return ts[len(ts)//2]
compositor.USE_PLACEMENT=False; print("expand+sort fwd ms", round(t_fwd(),3))
compositor.USE_PLACEMENT=True
for blocks in (37, 74, 148, 222, 296, 444, 592, 1184):
    L.gcp_splat_set_fill_blocks(blocks); print("placement fill_blocks", blocks, "fwd ms", round(t_fwd(),3))
L.gcp_splat_set_fill_blocks(0)
for thr in (8, 0):  # 0 = long-list kernels (warp-per-list keys, transposed fill) forced on the 1080p scene
    L.gcp_splat_set_long_list_threshold(thr); print("long-list threshold", thr, "fwd ms", round(t_fwd(), 3))
L.gcp_splat_set_long_list_threshold(8)
v = wl.bundled_views("cuda")[1]
for thr in (8, 1 << 20):  # 1<<20 = short-list kernels forced on the bundled scene
    L.gcp_splat_set_long_list_threshold(thr); print("bundled scene: long-list threshold", thr, "fwd ms", round(t_fwd(), 3))
