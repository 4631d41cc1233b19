"""GPU bring-up helper: run one fwd+bwd at small n and compare EVERY workspace buffer and every
gradient tensor with the numpy plan interpreter (tests/plan_interp.py).  Not part of the product."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import convvae_ref as R
from vae_npvc_b200 import vcc2016_vae_arch
from vae_npvc_b200.engine import Engine
import plan_interp as PI

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
arch = vcc2016_vae_arch()
if len(sys.argv) > 2:                                   # an alternative architecture of tests/conftest.py
    import conftest, copy
    arch = copy.deepcopy(conftest.ALT_ARCHS[sys.argv[2]])
eng = Engine(arch, "cuda:0")
P = R.init_params(arch, 0)
theta64 = R.flatten_params(arch, P, np.float64)
x, y, eps = R.make_inputs(arch, n)
plan = eng.handle.plan()
tables = {k: eng.handle.plan_table(k) for k in ("pack_src", "pack16_src", "unpack_ptr", "unpack_idx")}
it = PI.Interp(plan, tables, theta64.astype(np.float32).astype(np.float64), n, x.astype(np.float32), y, eps.astype(np.float32))
ref = it.loss_fwd_bwd()
dev = eng.device
theta = torch.tensor(theta64.astype(np.float32), device=dev)
grad = torch.empty_like(theta)
out = eng.loss_fwd_bwd(theta, torch.tensor(x, dtype=torch.float32, device=dev), torch.tensor(y, device=dev),
                       torch.tensor(eps, dtype=torch.float32, device=dev), grad=grad)
torch.cuda.synchronize()
print("losses gpu", out["losses"].cpu().tolist(), "ref", [ref["G"], ref["D_KL"], ref["logP"]])
def rel(a, b):
    b = np.nan_to_num(b); a = np.nan_to_num(a)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-30)
aw = eng.debug_buffer("arena_w", n).cpu().numpy()
n32 = plan["aw16_off"]
print("%-12s rel %.3e" % ("arena_w", rel(aw[:n32], it.aw[:n32])))
def bf16_to_f64(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32).astype(np.float64)
if plan["aw16_count"]:
    a16 = bf16_to_f64(aw[n32:].view(np.uint16)[:plan["aw16_count"]])
    print("%-12s rel %.3e" % ("arena16", rel(a16, it.aw16)))
for b in plan["bufs"]:
    if b["name"] == "acc": continue
    g = eng.debug_buffer(b["name"], n).cpu().numpy()
    r = it.bufs[it.buf_index[b["name"]]]
    if b.get("split"):
        pf = b["per_frame"]
        u = g.view(np.uint16).reshape(-1, 2 * pf)
        g = (bf16_to_f64(u[:, :pf]) + bf16_to_f64(u[:, pf:])).reshape(-1)
    g = g.astype(np.float64)
    print("%-12s rel %.3e   |ref| %.3e" % (b["name"], rel(g, r), np.nanmax(np.abs(r))))
gg = grad.cpu().numpy().astype(np.float64)
for t in eng.table:
    a = gg[t["offset"]:t["offset"] + t["size"]]; b = ref["grad"][t["offset"]:t["offset"] + t["size"]]
    print("grad %-45s rel %.3e" % (t["name"], rel(a, b)))
