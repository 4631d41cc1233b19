"""Bring-up check: the 30-step training scenario of tests/test_gpu_parity.py::test_plugin_surface_trains, loss per step."""
import os, sys, torch
sys.path.insert(0, '/root/repo')
from importlib import import_module
from vae_npvc_b200 import vcc2016_vae_arch
arch = vcc2016_vae_arch()
MODEL = getattr(import_module("model.vae"), "ConvVAE"); TRAINER = getattr(import_module("trainer.vae"), "VAETrainer")
a = dict(arch); a["training"] = dict(arch["training"], max_iter=30, lr=float(sys.argv[1]) if len(sys.argv) > 1 else 1e-3)
torch.manual_seed(int(sys.argv[2]) if len(sys.argv) > 2 else 0)      # the sampler's eps stream
machine = MODEL(a)
g = torch.Generator(device="cpu").manual_seed(0)
image = (torch.rand(64, 1, 513, 1, generator=g) * 2 - 1).cuda(); label = torch.randint(0, 10, (64,), generator=g).cuda()
loss = machine.loss(image, label)
tr = TRAINER(loss, a, None, None)
out = []
for i in range(30):
    lo = tr.opt['g']()
    out.append(float(lo[0]))
    gr = tr._state['grad']
    if not bool(torch.isfinite(gr).all()) or not bool(torch.isfinite(lo).all()):
        print("step", i, "losses", lo.tolist())
        for t in machine.engine.table:
            b = gr[t["offset"]:t["offset"] + t["size"]]
            nb = int((~torch.isfinite(b)).sum())
            if nb: print("   non-finite grad", t["name"], nb, "of", t["size"])
        for nm in ("mu", "lv", "z", "xh", "hz", "c_e0", "c_e1", "c_e2", "c_e3", "c_e4", "c_g0", "c_g1", "c_g2", "da_g2", "da_g1", "da_g0", "dz", "da_e4", "da_e3", "da_e2", "da_e1", "da_e0", "rstd_e0", "rstd_g2"):
            try:
                bb = machine.engine.debug_buffer(nm, 64)
                nb = int((~torch.isfinite(bb)).sum())
                if nb: print("   non-finite buffer", nm, nb, "of", bb.numel())
            except Exception as e:
                print("   (no buffer %s: %s)" % (nm, e))
        break
print(" ".join("%.1f" % v for v in out))
