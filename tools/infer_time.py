"""cfg3 (BASELINE.json configs[2]): inference-only encode -> z (= mu) -> decode at 256 x 512 = 131,072 frames,
through the plugin surface (model.vae.ConvVAE.encode / .decode), CUDA-event timed.  Prints frames/s."""
import sys, torch
sys.path.insert(0, '/root/repo')
from importlib import import_module
from vae_npvc_b200 import vcc2016_vae_arch
arch = vcc2016_vae_arch()
M = import_module('model.vae').ConvVAE(arch)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
g = torch.Generator().manual_seed(1)
x = (torch.rand(n, 1, 513, 1, generator=g) * 2 - 1).cuda(); y = torch.randint(0, 10, (n,), generator=g).cuda()
for _ in range(3):
    xh = M.decode(M.encode(x), y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
e0.record()
for _ in range(K):
    xh = M.decode(M.encode(x), y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print("cfg3 inference: %d frames in %.3f ms = %.3f M frames/s (x %.0f MB in, xh %.0f MB out)" % (n, ms, n / ms / 1e3, n * 2052 / 1e6, n * 2052 / 1e6))
