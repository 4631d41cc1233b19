"""Bring-up check of the CTA-pair (cta_group::2) form of the forward kernel: same inputs through a default engine
and through an engine created with NPVC_PAIR / NPVC_PAIR_OPS, outputs and gradients compared, per-op and
per-step times of both printed.  Every line is flushed as it is produced (a trap in the pair kernel ends the
process: what was printed before it tells how far it got).

    python tools/pair_check.py [n_frames] [ops] [NPVC_X=v ...]
        ops: comma-separated op names for the pair form, "wide" (every BN >= 128 layer) or "default" (the library's rule)
        NPVC_X=v: further switches for the second engine only, e.g. NPVC_WGRAD_PAIR=2 (cta_group::2 weight-gradient
        kernel, 256-column N tiles), NPVC_STREAMS=2 (two half-batches on two streams), NPVC_PAIR_TRIM=1
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_npvc_b200 import vcc2016_vae_arch          # noqa: E402
from vae_npvc_b200.engine import Engine             # noqa: E402


def say(*a):
    print(*a, flush=True)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    ops = sys.argv[2] if len(sys.argv) > 2 else "convT_g3"
    extra = dict(a.split("=", 1) for a in sys.argv[3:])
    wgrad = extra.get("NPVC_WGRAD_PAIR", "")
    arch = vcc2016_vae_arch()
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(n, 513, generator=g) * 2 - 1).cuda()
    y = torch.randint(0, 10, (n,), generator=g).cuda()
    eps = torch.randn(n, 128, generator=g).cuda()

    os.environ["NPVC_PAIR"] = "0"
    base = Engine(arch, "cuda:0")                      # single-CTA form everywhere
    if ops == "wide":
        os.environ["NPVC_PAIR"] = "2"                  # every BN >= 128 layer
    elif ops == "default":
        os.environ["NPVC_PAIR"] = "1"                  # the library's own shape rule
    else:
        os.environ["NPVC_PAIR"] = "1"; os.environ["NPVC_PAIR_OPS"] = ops
    os.environ.update(extra)
    pair = Engine(arch, "cuda:0")
    for k in ["NPVC_PAIR", "NPVC_PAIR_OPS"] + list(extra):
        os.environ.pop(k, None)
    theta = base.init_theta(0, perturb=0.1)

    def run(eng, grad):
        out = eng.loss_fwd_bwd(theta, x, y, eps, grad=grad)
        torch.cuda.synchronize()
        return out

    t0 = time.time()
    gb = torch.empty_like(theta); ob = run(base, gb)
    say("base ok  losses", ob["losses"].tolist(), "%.2fs" % (time.time() - t0))
    gp = torch.empty_like(theta); op = run(pair, gp)
    say("pair ok  losses", op["losses"].tolist(), "ops =", ops, extra)

    def rel(a, b):
        return float((a - b).abs().max() / (b.abs().max() + 1e-30))
    for k in ("mu", "lv", "z", "xh"):
        say("  %-3s pair vs base: rel %.3e  finite %s" % (k, rel(op[k], ob[k]), bool(torch.isfinite(op[k]).all())))
    say("  grad pair vs base: rel %.3e  finite %s" % (rel(gp, gb), bool(torch.isfinite(gp).all())))

    def step_ms(eng, grad, iters=10):
        for _ in range(3):
            eng.loss_fwd_bwd(theta, x, y, eps, grad=grad, outputs=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            eng.loss_fwd_bwd(theta, x, y, eps, grad=grad, outputs=False)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    say("fwd+bwd ms/pass: base %.4f  pair %.4f" % (step_ms(base, gb), step_ms(pair, gp)))

    def op_ms(eng, grad, iters=5):
        eng.handle.profile_enable(True)
        for _ in range(iters):
            eng.loss_fwd_bwd(theta, x, y, eps, grad=grad, outputs=False)
        torch.cuda.synchronize()
        p = eng.handle.profile(); eng.handle.profile_enable(False)
        return {q["name"]: q["ms"] / iters for q in p}
    pb, pp = op_ms(base, gb), op_ms(pair, gp)
    for name in sorted(pb, key=lambda k: -pb[k]):
        if abs(pb[name] - pp.get(name, 0.0)) > 0.01 * pb[name] + 0.002 or name in ops.split(",") or (wgrad and name.startswith("wgrad")):
            say("  %-16s base %.4f ms  pair %.4f ms" % (name, pb[name], pp.get(name, float("nan"))))
    say("sum of ops: base %.4f  pair %.4f" % (sum(pb.values()), sum(pp.values())))


if __name__ == "__main__":
    main()
