"""cfg5 (BASELINE.json configs[4]): the conv stacks of architecture-vawgan-vcc2016.json at 32 x 256 = 8,192 frames,
fwd + bwd, CUDA-event timed.  The encoder / generator stacks of that JSON are the ConvVAE ones; its discriminator
stack (kernels 7 / 7 / 115, stride 3, 16 / 32 / 64 channels) runs as the encoder of the path
(tests/conftest.py::ALT_ARCHS["vawgan_d_stack"]).  Only the conv stacks are specifiable: the VAWGAN model code is
absent from the reference snapshot.  Prints frames/s of both stacks."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import ALT_ARCHS                       # noqa: E402
from vae_npvc_b200 import vcc2016_vae_arch           # noqa: E402
from vae_npvc_b200.engine import Engine              # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32 * 256
g = torch.Generator().manual_seed(1)
x = (torch.rand(n, 513, generator=g) * 2 - 1).cuda(); y = torch.randint(0, 10, (n,), generator=g).cuda()
eps = torch.randn(n, 128, generator=g).cuda()
for name, arch in (("generator + encoder stacks (== ConvVAE)", vcc2016_vae_arch()), ("discriminator stack as the encoder", ALT_ARCHS["vawgan_d_stack"])):
    eng = Engine(arch, "cuda:0")
    theta = eng.init_theta(0, perturb=0.1); grad = torch.empty_like(theta)
    for _ in range(3):
        eng.loss_fwd_bwd(theta, x, y, eps, grad=grad, outputs=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 20
    e0.record()
    for _ in range(K):
        eng.loss_fwd_bwd(theta, x, y, eps, grad=grad, outputs=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print("cfg5 %s: %d frames fwd+bwd in %.3f ms = %.3f M frames/s" % (name, n, ms, n / ms / 1e3), flush=True)
