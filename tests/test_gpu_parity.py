"""GPU parity tests: the CUDA path, called through the C-ABI, against the fp64 oracle on the same
seeded inputs.  Tolerances (written here, SURVEY 8c / north_star): per-tensor max|a-b|/max|b|
<= 1e-4 for z, mu, logsigma^2, xh; <= 1e-3 for gradients and post-Adam parameters."""
import os

import numpy as np
import pytest
import torch

from oracle import convvae_ref as R
from parity_util import check_branches, lrelu_branches

pytestmark = pytest.mark.gpu
TOL_OUT, TOL_GRAD = 1e-4, 1e-3
GOLD = os.path.join(os.path.dirname(__file__), "golden", "convvae_vcc2016_n4.npz")


def rel(a, b):
    a = a.detach().cpu().numpy().astype(np.float64) if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


@pytest.fixture(scope="module")
def eng(arch):
    from vae_npvc_b200.engine import Engine
    return Engine(arch, "cuda:0")


def _dev_inputs(eng, arch, n, seed=1, n_speakers=None):
    x, y, eps = R.make_inputs(arch, n, seed=seed, n_speakers=n_speakers)
    d = eng.device
    return (x, y, eps, torch.tensor(x, dtype=torch.float32, device=d), torch.tensor(y, device=d),
            torch.tensor(eps, dtype=torch.float32, device=d))


def _check_against_oracle(eng, arch, n, seed=1, n_speakers=None):
    P = R.init_params(arch, 0)
    x, y, eps, xd, yd, ed = _dev_inputs(eng, arch, n, seed, n_speakers)
    theta = torch.tensor(R.flatten_params(arch, P), device=eng.device)
    grad = torch.full_like(theta, float("nan"))
    out = eng.loss_fwd_bwd(theta, xd, yd, ed, grad=grad)
    torch.cuda.synchronize()
    # oracle sees the same fp32-rounded inputs and weights
    P32 = {k: np.asarray(v, np.float32).astype(np.float64) for k, v in P.items()}
    fwd = R.forward(arch, P32, x.astype(np.float32), y, eps.astype(np.float32), with_acts=True)
    # lrelu' is discontinuous at 0: the oracle differentiates with the branches the CUDA path took, which
    # may differ from its own only where the pre-activation is within fp32 rounding distance of zero
    pos = lrelu_branches(lambda name: eng.debug_buffer(name, n).cpu().numpy(), arch, P32, n)
    check_branches(pos, fwd["acts"])
    ref = R.forward(arch, P32, x.astype(np.float32), y, eps.astype(np.float32), with_grads=True, lrelu_pos=pos)
    for k in ("z", "mu", "lv", "xh"):
        assert rel(out[k], ref[k]) <= TOL_OUT, (k, rel(out[k], ref[k]))
    lo = out["losses"].cpu().numpy()
    for i, k in enumerate(("G", "D_KL", "logP")):
        assert abs(lo[i] - ref[k]) <= 1e-5 * abs(ref[k]) + 1e-6, k
    gref = R.flatten_params(arch, ref["grads"], np.float64)
    g = grad.cpu().numpy().astype(np.float64)
    assert np.isfinite(g).all()
    for t in eng.table:
        sl = slice(t["offset"], t["offset"] + t["size"])
        den = np.abs(gref[sl]).max()
        if den == 0:
            assert np.abs(g[sl]).max() == 0, t["name"]
        else:
            assert np.abs(g[sl] - gref[sl]).max() / den <= TOL_GRAD, t["name"]
    return theta, grad, gref, P32


@pytest.mark.parametrize("n", [1, 8, 37])
def test_loss_fwd_bwd_matches_oracle(eng, arch, n):
    _check_against_oracle(eng, arch, n)


def test_single_speaker_batch(eng, arch):
    """cfg1 shape: all labels 0 -> the other 9 embedding rows get exactly zero gradient."""
    _check_against_oracle(eng, arch, 16, n_speakers=1)


def test_golden_fixture(eng, arch):
    g = np.load(GOLD)
    P = R.init_params(arch, 0)
    x, y, eps, xd, yd, ed = _dev_inputs(eng, arch, int(g["n"]))
    theta = torch.tensor(R.flatten_params(arch, P), device=eng.device)
    grad = torch.empty_like(theta)
    out = eng.loss_fwd_bwd(theta, xd, yd, ed, grad=grad)
    for k in ("z", "mu", "lv", "xh"):
        assert rel(out[k], g[k]) <= TOL_OUT, k
    assert rel(grad[::int(g["stride"])], g["grad_sample"]) <= TOL_GRAD
    m = torch.zeros_like(theta); v = torch.zeros_like(theta)
    eng.adam_step(theta, grad, m, v, 1, 1e-4, 0.5, 0.999)
    assert rel(theta[::int(g["stride"])], g["adam_theta_sample"]) <= TOL_GRAD


def test_adam_steps_match_tf_form(eng, arch):
    """The Adam kernel against the TF-form closed form on IDENTICAL gradients (<= 1e-6), then the
    end-to-end post-Adam parameters against the oracle's own gradients (<= 1e-3: Adam's m/sqrt(v)
    normalisation turns noise-level differences of near-zero gradients into +-lr steps)."""
    theta, grad, gref, P32 = _check_against_oracle(eng, arch, 8)
    g_gpu = grad.cpu().numpy().astype(np.float64)
    th = R.flatten_params(arch, P32, np.float64); m = np.zeros_like(th); v = np.zeros_like(th)
    th2, m2, v2 = th.copy(), m.copy(), v.copy()
    md = torch.zeros_like(theta); vd = torch.zeros_like(theta)
    for t in (1, 2, 3):
        eng.adam_step(theta, grad, md, vd, t, 1e-4, 0.5, 0.999, 1e-8, 0.5)       # grad_scale 1/2 (two ranks)
        th, m, v = R.adam_step(th, 0.5 * g_gpu, m, v, t, 1e-4, 0.5, 0.999)
        th2, m2, v2 = R.adam_step(th2, 0.5 * gref, m2, v2, t, 1e-4, 0.5, 0.999)
    # v: (1 - beta2) is formed in fp32 from fp32(0.999) exactly as TF's ApplyAdam does (1.3e-5 off the real 0.001)
    assert rel(theta, th) <= 1e-6 and rel(md, m) <= 1e-6 and rel(vd, v) <= 5e-5
    assert rel(theta, th2) <= TOL_GRAD


def test_chunked_equals_unchunked(arch, eng):
    from vae_npvc_b200.engine import Engine
    small = Engine(arch, "cuda:0", max_chunk=16)
    P = R.init_params(arch, 0)
    x, y, eps, xd, yd, ed = _dev_inputs(eng, arch, 40)
    theta = torch.tensor(R.flatten_params(arch, P), device=eng.device)
    g1 = torch.empty_like(theta); g2 = torch.empty_like(theta)
    o1 = eng.loss_fwd_bwd(theta, xd, yd, ed, grad=g1)
    o2 = small.loss_fwd_bwd(theta, xd, yd, ed, grad=g2)
    for k in ("z", "mu", "lv", "xh"):
        assert torch.equal(o1[k], o2[k]), k                      # per-frame results are bit-identical
    assert rel(o2["losses"], o1["losses"].cpu().numpy()) < 1e-6
    assert rel(g2, g1.cpu().numpy()) < 1e-4                      # summation order over frames differs


def test_encode_decode_match_oracle(eng, arch):
    """convert.py path: encode -> mu (no sampling), decode(z, y_target)."""
    P = R.init_params(arch, 0)
    P32 = {k: np.asarray(v, np.float32).astype(np.float64) for k, v in P.items()}
    x, y, eps, xd, yd, ed = _dev_inputs(eng, arch, 23)
    theta = torch.tensor(R.flatten_params(arch, P), device=eng.device)
    mu, lv = eng.encode(theta, xd)
    mu_ref, lv_ref = R.encode(arch, P32, x.astype(np.float32))
    assert rel(mu, mu_ref) <= TOL_OUT and rel(lv, lv_ref) <= TOL_OUT
    yt = torch.full_like(yd, 9)
    xh = eng.decode(theta, mu, yt)
    xh_ref = R.decode(arch, P32, mu.cpu().numpy().astype(np.float64), np.full(23, 9))
    assert rel(xh, xh_ref.reshape(23, -1)) <= TOL_OUT
    z = eng.sample(mu, lv, ed)
    assert rel(z, mu_ref + eps.astype(np.float32) * np.sqrt(np.exp(lv_ref))) <= TOL_OUT


def test_full_size_properties(eng, arch):
    """cfg2 size (N = 64*256 = 16,384): size-independent properties instead of an oracle run --
    frame permutation equivariance (F5/F6), finite outputs, losses = mean of per-frame terms,
    and agreement of a 64-frame slice with the oracle."""
    n = 16384
    g = torch.Generator(device="cpu").manual_seed(5)
    x = (torch.rand(n, 513, generator=g) * 2 - 1).cuda(); y = torch.randint(0, 10, (n,), generator=g).cuda()
    eps = torch.randn(n, 128, generator=g).cuda()
    P = R.init_params(arch, 0)
    theta = torch.tensor(R.flatten_params(arch, P), device=eng.device)
    grad = torch.empty_like(theta)
    out = eng.loss_fwd_bwd(theta, x, y, eps, grad=grad)
    assert all(torch.isfinite(out[k]).all() for k in ("z", "mu", "lv", "xh")) and torch.isfinite(grad).all()
    perm = torch.randperm(n, generator=g).cuda()
    g2 = torch.empty_like(theta)
    out2 = eng.loss_fwd_bwd(theta, x[perm].contiguous(), y[perm].contiguous(), eps[perm].contiguous(), grad=g2)
    for k in ("z", "mu", "xh"):
        assert torch.equal(out2[k], out[k][perm]), k
    assert rel(g2, grad.cpu().numpy()) < 2e-4
    # losses are means of per-frame terms
    xh, mu, lv = out["xh"].double(), out["mu"].double(), out["lv"].double()
    c = float(np.float32(1.0) + np.float32(1e-6))
    logp = (-0.5 * (R.LOG_2PI + (x.double() - xh) ** 2 / c)).sum(1).mean()
    kl = (0.5 * (-lv + (lv.exp() + mu * mu) / c - 1)).sum(1).mean()
    lo = out["losses"].double().cpu()
    assert abs(lo[2] - logp.cpu()) < 1e-5 * abs(logp.cpu()) and abs(lo[1] - kl.cpu()) < 1e-5 * abs(kl.cpu())
    # a slice against the oracle
    P32 = {k: np.asarray(v, np.float32).astype(np.float64) for k, v in P.items()}
    sl = slice(5000, 5064)
    ref = R.forward(arch, P32, x[sl].cpu().numpy(), y[sl].cpu().numpy(), eps[sl].cpu().numpy())
    for k in ("z", "mu", "lv", "xh"):
        assert rel(out[k][sl], ref[k]) <= TOL_OUT, k


def _oracle_chunked(arch, P32, x, y, eps, pos, chunk=1024):
    """fp64 oracle outputs and gradient of the batch mean, accumulated over chunks of frames (the loss is a mean of
    per-frame terms: grad = sum_c (n_c / n) grad_c); `pos` = the lrelu branches the implementation took."""
    n = x.shape[0]
    outs = {k: [] for k in ("z", "mu", "lv", "xh")}
    gsum, G = None, 0.0
    for c0 in range(0, n, chunk):
        sl = slice(c0, min(n, c0 + chunk))
        ref = R.forward(arch, P32, x[sl], y[sl], eps[sl], with_grads=True, lrelu_pos={k: v[sl] for k, v in pos.items()})
        w = (sl.stop - sl.start) / n
        g = R.flatten_params(arch, ref["grads"], np.float64) * w
        gsum = g if gsum is None else gsum + g
        G += float(ref["G"]) * w
        for k in outs:
            outs[k].append(ref[k])
    return {k: np.concatenate(v) for k, v in outs.items()}, gsum, G


@pytest.mark.parametrize("n,n_speakers", [(16384, None), (2048, 1)])
def test_benchmark_sizes_against_the_oracle(eng, arch, n, n_speakers):
    """The sizes the metric is quoted on, with the default routing of those sizes (CTA-pair tiles, split weight
    gradients, fused first layer): cfg2 (64 x 256 frames, 10 speakers) and cfg1 (16 x 128 frames, one speaker).
    Every output and all 44 gradient tensors against the fp64 oracle, which is run in chunks of 1024 frames."""
    P = R.init_params(arch, 0)
    x, y, eps, xd, yd, ed = _dev_inputs(eng, arch, n, seed=3, n_speakers=n_speakers)
    theta = torch.tensor(R.flatten_params(arch, P), device=eng.device)
    grad = torch.full_like(theta, float("nan"))
    out = eng.loss_fwd_bwd(theta, xd, yd, ed, grad=grad)
    torch.cuda.synchronize()
    P32 = {k: np.asarray(v, np.float32).astype(np.float64) for k, v in P.items()}
    x32, e32 = x.astype(np.float32), eps.astype(np.float32)
    pos = lrelu_branches(lambda name: eng.debug_buffer(name, n).cpu().numpy(), arch, P32, n)
    ref, gref, G = _oracle_chunked(arch, P32, x32, y, e32, pos)
    for k in ("z", "mu", "lv", "xh"):
        assert rel(out[k], ref[k]) <= TOL_OUT, (k, rel(out[k], ref[k]))
    assert abs(float(out["losses"][0]) - G) <= 1e-5 * abs(G)
    g = grad.cpu().numpy().astype(np.float64)
    assert np.isfinite(g).all()
    worst = 0.0
    for t in eng.table:
        sl = slice(t["offset"], t["offset"] + t["size"])
        den = np.abs(gref[sl]).max()
        if den == 0:
            assert np.abs(g[sl]).max() == 0, t["name"]
        else:
            e = np.abs(g[sl] - gref[sl]).max() / den
            worst = max(worst, e)
            assert e <= TOL_GRAD, (t["name"], e)
    print("n=%d worst per-tensor gradient error %.2e" % (n, worst))


@pytest.mark.parametrize("switch", ["NPVC_FUSE=0", "NPVC_FUSE_LN_TRAIN=1"])
def test_fusion_switches_match_oracle(arch, monkeypatch, switch):
    """NPVC_FUSE=0: every plan op as its own kernel (the path shapes without a fused kernel take).
    NPVC_FUSE_LN_TRAIN=1: the Layernorm epilogue of the forward kernel (default for inference passes only) in a training
    pass, where it also stores the raw conv output for the backward."""
    from vae_npvc_b200.engine import Engine
    k, v = switch.split("=")
    monkeypatch.setenv(k, v)
    e2 = Engine(arch, "cuda:0")
    monkeypatch.delenv(k)
    for n in (37, 300):
        _check_against_oracle(e2, arch, n)


def test_cfg3_inference_size(eng, arch):
    """cfg3 (convert.py path): encode -> mu -> decode at N = 256*512 = 131,072 frames, processed in
    16,384-frame chunks inside the library; chunk boundaries must be invisible and a slice that
    straddles one must match the oracle."""
    n = 131072
    g = torch.Generator(device="cpu").manual_seed(9)
    x = (torch.rand(n, 513, generator=g) * 2 - 1).cuda(); y = torch.randint(0, 10, (n,), generator=g).cuda()
    P = R.init_params(arch, 0)
    theta = torch.tensor(R.flatten_params(arch, P), device=eng.device)
    mu, lv = eng.encode(theta, x)
    xh = eng.decode(theta, mu, y)
    assert mu.shape == (n, 128) and xh.shape == (n, 513) and torch.isfinite(xh).all() and torch.isfinite(mu).all()
    sl = slice(16384 - 20, 16384 + 20)                      # straddles the first chunk boundary
    mu2, _ = eng.encode(theta, x[sl].contiguous())
    xh2 = eng.decode(theta, mu2, y[sl].contiguous())
    assert torch.equal(mu2, mu[sl]) and torch.equal(xh2, xh[sl])
    P32 = {k: np.asarray(v, np.float32).astype(np.float64) for k, v in P.items()}
    mu_ref, _ = R.encode(arch, P32, x[sl].cpu().numpy())
    xh_ref = R.decode(arch, P32, mu_ref, y[sl].cpu().numpy()).reshape(40, -1)
    assert rel(mu[sl], mu_ref) <= TOL_OUT and rel(xh[sl], xh_ref) <= TOL_OUT


def test_alternative_architectures_on_gpu(alt_arch):
    """The CUDA path on architectures other than VCC2016 (different tile / routing decisions)."""
    from vae_npvc_b200.engine import Engine
    e2 = Engine(alt_arch, "cuda:0")
    _check_against_oracle(e2, alt_arch, 29)


def test_gradients_stay_finite_over_many_passes(eng, arch):
    """Regression: an operand tile may never multiply bits from outside its own frame, not even by a zero
    weight (NaN * 0).  The parity-split dgrad's last tap once read 16 bytes past the last frame's plane:
    ~1.5 % of the passes, depending on the neighbouring buffer's bits, produced NaN gradients."""
    n = 8
    theta = eng.init_theta(0, 0.1)
    grad = torch.empty_like(theta)
    g = torch.Generator(device="cpu").manual_seed(7)
    for _ in range(150):
        x = (torch.rand(n, 513, generator=g) * 2 - 1).to(eng.device)
        y = torch.randint(0, arch["y_dim"], (n,), generator=g).to(eng.device)
        eps = torch.randn(n, arch["z_dim"], generator=g).to(eng.device)
        out = eng.loss_fwd_bwd(theta, x, y, eps, grad=grad, outputs=False)
        assert bool(torch.isfinite(grad).all()) and bool(torch.isfinite(out["losses"]).all())


@pytest.mark.parametrize("which", ["vcc2016"] + sorted(__import__("conftest").ALT_ARCHS))
def test_poisoned_workspace_matches_oracle(arch, monkeypatch, which):
    """NPVC_DEBUG_POISON=1 fills the whole workspace with 0xFF bytes (NaN bit patterns in fp32 and bf16) before every
    pass: any operand tile, halo row or pad that the pass reads without having written it itself poisons the outputs
    deterministically (the NaN * 0 rule of DESIGN.md, not left to whatever bits torch.empty returned)."""
    import copy
    from conftest import ALT_ARCHS
    from vae_npvc_b200.engine import Engine
    a = arch if which == "vcc2016" else copy.deepcopy(ALT_ARCHS[which])
    e2 = Engine(a, "cuda:0")
    monkeypatch.setenv("NPVC_DEBUG_POISON", "1")
    for n in ((1, 37, 300) if which == "vcc2016" else (29,)):
        _check_against_oracle(e2, a, n)
    if which == "vcc2016":                                   # benchmark size: finite everywhere (routing of large batches)
        n = 16384
        g = torch.Generator(device="cpu").manual_seed(21)
        x = (torch.rand(n, 513, generator=g) * 2 - 1).cuda(); y = torch.randint(0, 10, (n,), generator=g).cuda()
        eps = torch.randn(n, 128, generator=g).cuda()
        theta = e2.init_theta(0, 0.1); grad = torch.empty_like(theta)
        out = e2.loss_fwd_bwd(theta, x, y, eps, grad=grad)
        assert all(bool(torch.isfinite(out[k]).all()) for k in ("z", "mu", "lv", "xh", "losses")) and bool(torch.isfinite(grad).all())
        mu, lv = e2.encode(theta, x[:5000].contiguous())     # inference layout + Layernorm epilogue
        xh = e2.decode(theta, mu, y[:5000].contiguous())
        assert bool(torch.isfinite(mu).all()) and bool(torch.isfinite(xh).all())


def test_tanhize_and_record_reader(eng):
    g = torch.Generator(device="cpu").manual_seed(3)
    xmin = torch.randn(513, generator=g) - 3; xmax = xmin + 1 + torch.rand(513, generator=g)
    rec = torch.randn(50, 1029, generator=g); rec[:, -1] = torch.randint(0, 10, (50,), generator=g).float()
    x, y = eng.unpack_records(rec.cuda(), 513, xmin.cuda(), xmax.cuda())
    ref = R.tanhize_forward(rec[:, :513].double().numpy(), xmin.double().numpy(), xmax.double().numpy())
    assert rel(x, ref) < 1e-6 and torch.equal(y.cpu(), rec[:, -1].long())
    back = eng.tanhize_backward(x, xmin.cuda(), xmax.cuda())
    assert rel(back, R.tanhize_backward(ref, xmin.double().numpy(), xmax.double().numpy())) < 1e-6
    fwd = eng.tanhize_forward(rec[:, :513].contiguous().cuda(), xmin.cuda(), xmax.cuda())
    assert torch.equal(fwd, x)


def test_plugin_surface_trains(arch, tmp_path):
    """main.py call shape: MODEL(arch) -> loss -> TRAINER(loss, arch, args, dirs).train(...)."""
    from importlib import import_module
    MODEL = getattr(import_module("model.vae"), "ConvVAE")
    TRAINER = getattr(import_module("trainer.vae"), "VAETrainer")
    # lr 3e-4: at 1e-3 Adam(beta1 = 0.5) on 64 frames diverges around step 25 (exp(logsigma^2) overflows to NaN
    # at a step that depends on the atomics' summation order) -- the check is "the plugin trains", not a stability test
    a = dict(arch); a["training"] = dict(arch["training"], max_iter=30, lr=3e-4)
    machine = MODEL(a)
    g = torch.Generator(device="cpu").manual_seed(0)
    image = (torch.rand(64, 1, 513, 1, generator=g) * 2 - 1).cuda(); label = torch.randint(0, 10, (64,), generator=g).cuda()
    loss = machine.loss(image, label)
    assert set(loss.keys()) == {"G", "D_KL", "logP"}
    g0 = float(loss["G"])
    dirs = {"logdir": str(tmp_path / "train"), "logdir_root": str(tmp_path), "restore_from": str(tmp_path / "train")}
    trainer = TRAINER(loss, a, None, dirs)
    trainer.train(nIter=a["training"]["max_iter"], machine=machine)
    g1 = float(machine.loss(image, label)["G"])
    assert trainer.global_step == 30 and g1 < g0
    assert os.path.exists(os.path.join(dirs["logdir"], "model.ckpt-30"))
    z = machine.encode(image)
    xh = machine.decode(z, label)
    assert z.shape == (64, 128) and xh.shape == (64, 513, 1, 1) and machine.generate == machine.decode


def test_cta_pair_form_is_bit_identical_to_single_cta(arch, monkeypatch):
    """The cta_group::2 form of the forward kernel (two CTAs, one 256 x BN MMA; umma_gemm.cuh PAIR) against the
    single-CTA form on every layer wide enough for it, at cfg2 size: per-frame outputs must be bit-identical
    (same products, same accumulation order), gradients equal up to the atomics' summation order."""
    from vae_npvc_b200.engine import Engine
    n = 16384
    g = torch.Generator(device="cpu").manual_seed(11)
    x = (torch.rand(n, 513, generator=g) * 2 - 1).cuda(); y = torch.randint(0, 10, (n,), generator=g).cuda()
    eps = torch.randn(n, 128, generator=g).cuda()
    monkeypatch.setenv("NPVC_PAIR", "0"); single = Engine(arch, "cuda:0")
    monkeypatch.setenv("NPVC_PAIR", "2"); pair = Engine(arch, "cuda:0")       # every BN >= 128 layer
    monkeypatch.delenv("NPVC_PAIR")
    theta = single.init_theta(0, 0.1)
    g1 = torch.empty_like(theta); g2 = torch.empty_like(theta)
    o1 = single.loss_fwd_bwd(theta, x, y, eps, grad=g1)
    o2 = pair.loss_fwd_bwd(theta, x, y, eps, grad=g2)
    for k in ("z", "mu", "lv", "xh"):
        assert torch.equal(o1[k], o2[k]), k
    assert torch.isfinite(g2).all() and rel(g2, g1.cpu().numpy()) < 1e-5


EXPERIMENTS = {                    # A/B switches of the library: every form they select must reproduce the default engine
    "wgrad_pair": {"NPVC_WGRAD_PAIR": "1"}, "wgrad_pair_256": {"NPVC_WGRAD_PAIR": "2"}, "wgrad_single": {"NPVC_WGRAD_PAIR": "0"},
    "no_merge": {"NPVC_UMMA_MERGE": "0"}, "no_pdl": {"NPVC_PDL": "0"}, "no_resident_weights": {"NPVC_UMMA_BRES": "0"}, "one_mma_issuer": {"NPVC_UMMA_DUAL": "0"},
    "ln_bwd_prefetch": {"NPVC_PREFETCH_MIN": "0", "NPVC_PREFETCH_E0_MIN": "0"},
}


@pytest.mark.parametrize("name", sorted(EXPERIMENTS))
@pytest.mark.parametrize("n", [300, 16384])
def test_experimental_switch_matches_default(arch, monkeypatch, name, n):
    """Every experimental switch must reproduce the default engine: per-frame outputs bit-identical (same products in the
    same order per output element), gradients equal up to the atomics' summation order."""
    from vae_npvc_b200.engine import Engine
    g = torch.Generator(device="cpu").manual_seed(13)
    x = (torch.rand(n, 513, generator=g) * 2 - 1).cuda(); y = torch.randint(0, 10, (n,), generator=g).cuda()
    eps = torch.randn(n, 128, generator=g).cuda()
    base = Engine(arch, "cuda:0")
    for k, v in EXPERIMENTS[name].items():
        monkeypatch.setenv(k, v)
    exp = Engine(arch, "cuda:0")
    for k in EXPERIMENTS[name]:
        monkeypatch.delenv(k)
    theta = base.init_theta(0, 0.1)
    g1 = torch.empty_like(theta); g2 = torch.empty_like(theta)
    o1 = base.loss_fwd_bwd(theta, x, y, eps, grad=g1)
    o2 = exp.loss_fwd_bwd(theta, x, y, eps, grad=g2)
    for k in ("z", "mu", "lv", "xh"):
        assert torch.equal(o1[k], o2[k]), k
    assert torch.isfinite(g2).all() and rel(g2, g1.cpu().numpy()) < 1e-5
    assert rel(o2["losses"], o1["losses"].cpu().numpy()) < 1e-6
