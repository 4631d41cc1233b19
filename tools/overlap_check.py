"""Bring-up check: gradients with the side-stream wgrad overlap on vs off (same inputs), repeated."""
import os, sys, torch
sys.path.insert(0, '/root/repo')
from vae_npvc_b200 import vcc2016_vae_arch
from vae_npvc_b200.engine import Engine
arch = vcc2016_vae_arch()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
os.environ["NPVC_OVERLAP"] = "0"; e0 = Engine(arch, "cuda:0")
os.environ["NPVC_OVERLAP"] = "1"; e1 = Engine(arch, "cuda:0")
g = torch.Generator().manual_seed(1)
x = (torch.rand(n, 513, generator=g) * 2 - 1).cuda(); y = torch.randint(0, 10, (n,), generator=g).cuda()
eps = torch.randn(n, 128, generator=g).cuda()
theta = e0.init_theta(0, 0.1)
g0 = torch.empty_like(theta); g1 = torch.empty_like(theta)
e0.loss_fwd_bwd(theta, x, y, eps, grad=g0, outputs=False); torch.cuda.synchronize()
bad = 0
for it in range(30):
    g1.fill_(float('nan'))
    e1.loss_fwd_bwd(theta, x, y, eps, grad=g1, outputs=False); torch.cuda.synchronize()
    for t in e1.table:
        a = g0[t["offset"]:t["offset"] + t["size"]]; b = g1[t["offset"]:t["offset"] + t["size"]]
        err = float((a - b).abs().max() / (a.abs().max() + 1e-30))
        if not (err < 1e-3):
            bad += 1
            if bad < 30: print("iter", it, t["name"], "rel", err, "nan", bool(torch.isnan(b).any()))
print("bad", bad)
