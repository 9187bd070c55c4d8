import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_function as rf
from simplegaussiansplat_tk71_b200 import workloads as wl
from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F
dev = torch.device("cuda")
sc = wl.splat_view(480, 270, 62_500, seed=1080, device=dev)
scene = (sc.boxsize, sc.startpoint, sc.endpoint, sc.mean, sc.lam, sc.opacity, sc.l_d)
gI = torch.rand(sc.height + 1, sc.width + 1, 3, device=dev) + 0.1
a = rf.run("ref", scene, sc.width, sc.height, gI)
b = rf.run("dropin", scene, sc.width, sc.height, gI)
m, lam, o, l = (sc.mean.float().requires_grad_(True), sc.lam.clone().requires_grad_(True), sc.opacity.clone().requires_grad_(True), sc.l_d.clone().requires_grad_(True))
img = F.apply(sc.boxsize, torch.tensor([sc.n]), sc.startpoint, sc.endpoint, m, lam, o, l, sc.width, sc.height)
img.backward(gI)
print("sums", float(a[0].sum()), float(b[0].sum()), float(img.sum()))
print("diff", float((a[0]-b[0]).abs().max()), float((a[0]-img).abs().max()))
print("grad diff", float((a[1]["grad_opacity"].reshape(-1)-o.grad.reshape(-1)).abs().max()), float(a[1]["grad_opacity"].abs().max()))
