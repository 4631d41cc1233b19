import sys, torch, json
sys.path.insert(0,'/root/repo')
from importlib import import_module
from vae_npvc_b200 import vcc2016_vae_arch
arch = vcc2016_vae_arch()
M = import_module('model.vae').ConvVAE(arch); T = import_module('trainer.vae').VAETrainer
n=int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = torch.Generator().manual_seed(1)
x = (torch.rand(n,513,generator=g)*2-1).cuda(); y = torch.randint(0,10,(n,),generator=g).cuda()
tr = T(M.loss(x,y), arch, None, None)
tr.use_graph = False          # per-op CUDA events need eager launches
for i in range(3): tr.opt['g'](x,y)
M.engine.handle.profile_enable(True)
for i in range(5): tr.opt['g'](x,y)
torch.cuda.synchronize()
p = M.engine.handle.profile()
tot = sum(q['ms'] for q in p)
print('total ms/step', tot/5)
for q in sorted(p, key=lambda q:-q['ms']): print('%-16s %8.4f ms  rows=%8d K=%5d N=%5d' % (q['name'], q['ms']/5, q['rows']//max(q['calls'],1), q['K'], q['N']))
