"""What compute-sanitizer wraps: training passes (default routing of small and benchmark-size batches, in-kernel sampler,
device-side Adam) and the inference path (Layernorm epilogue) -- every kernel family of the library at least once.
    compute-sanitizer --tool memcheck python tools/sanitizer_target.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_npvc_b200 import vcc2016_vae_arch                    # noqa: E402
from vae_npvc_b200.engine import Engine                       # noqa: E402

eng = Engine(vcc2016_vae_arch(), "cuda:0")
theta = eng.init_theta(0, 0.1); grad = torch.empty_like(theta)
m = torch.zeros_like(theta); v = torch.zeros_like(theta)
state = eng.new_step_state(3)
g = torch.Generator().manual_seed(1)
for n in (37, 300, 16384):
    x = (torch.rand(n, 513, generator=g) * 2 - 1).cuda(); y = torch.randint(0, 10, (n,), generator=g).cuda()
    out = eng.loss_fwd_bwd(theta, x, y, None, grad=grad, state=state)
    eng.adam_step(theta, grad, m, v, state, 1e-4, 0.5, 0.999)
    mu, lv = eng.encode(theta, x); xh = eng.decode(theta, mu, y)
    torch.cuda.synchronize()
    print(n, out["losses"].tolist(), bool(torch.isfinite(xh).all()), state.tolist(), flush=True)
