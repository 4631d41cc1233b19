"""Times the training step (fwd + bwd + Adam through the plugin) with CUDA events: ms/step at cfg2."""
import sys, torch
sys.path.insert(0, '/root/repo')
from importlib import import_module
from vae_npvc_b200 import vcc2016_vae_arch
arch = vcc2016_vae_arch()
M = import_module('model.vae').ConvVAE(arch); T = import_module('trainer.vae').VAETrainer
n = 16384
g = torch.Generator().manual_seed(1)
x = (torch.rand(n, 513, generator=g) * 2 - 1).cuda(); y = torch.randint(0, 10, (n,), generator=g).cuda()
tr = T(M.loss(x, y), arch, None, None)
for i in range(5): tr.opt['g'](x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20): tr.opt['g'](x, y)
e1.record(); torch.cuda.synchronize()
print('ms/step %.4f' % (e0.elapsed_time(e1) / 20))
