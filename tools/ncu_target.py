"""One fwd+bwd pass at cfg2 size with no warm-up: the process ncu wraps for --set full captures
(launch order of the tcgen05 kernels inside a pass: conv_e1..e4, heads, merge, convT_g0, g1, g3,
wgrad_g_last, dgrad_g_last, ...)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_npvc_b200 import vcc2016_vae_arch
from vae_npvc_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
eng = Engine(vcc2016_vae_arch(), "cuda:0")
g = torch.Generator().manual_seed(1)
x = (torch.rand(n, 513, generator=g) * 2 - 1).cuda(); y = torch.randint(0, 10, (n,), generator=g).cuda()
eps = torch.randn(n, 128, generator=g).cuda()
theta = eng.init_theta(0, 0.1); grad = torch.empty_like(theta)
out = eng.loss_fwd_bwd(theta, x, y, eps, grad=grad, outputs=False)
if len(sys.argv) > 2 and sys.argv[2] == "adam":            # the whole training step: + TF-form Adam on the flat buffers
    m = torch.zeros_like(theta); v = torch.zeros_like(theta)
    eng.adam_step(theta, grad, m, v, 1, 1e-4, 0.5, 0.999)
torch.cuda.synchronize()
print(out["losses"].tolist())
