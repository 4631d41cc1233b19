"""GPU tests of the device-resident training state: the in-kernel N(0,1) sampler (Philox4x32-10 + Box-Muller), the
device-side Adam step count, and the CUDA-graph replay of a whole training step."""
import numpy as np
import pytest
import torch

from oracle import philox

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(arch):
    from vae_npvc_b200.engine import Engine
    return Engine(arch, "cuda:0")


def _batch(n, seed=1):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.rand(n, 513, generator=g) * 2 - 1).cuda(), torch.randint(0, 10, (n,), generator=g).cuda()


def test_sampler_matches_the_numpy_philox_and_is_standard_normal(eng):
    st = eng.new_step_state(seed=12345)
    st[1] = 7                                           # pass counter
    n = 8192
    e = eng.normal_draw(st, n, frame_offset=1 << 33).cpu().numpy().astype(np.float64)     # frame index beyond 32 bits
    ref = philox.normal_draw(12345, 7, 1 << 33, n, 128)
    assert np.abs(e - ref).max() < 2e-6                 # fp32 logf / sqrtf / cospif against float64
    v = e.reshape(-1)
    assert abs(v.mean()) < 5e-3 and abs(v.var() - 1.0) < 5e-3
    assert abs((v ** 3).mean()) < 2e-2 and abs((v ** 4).mean() - 3.0) < 5e-2
    # Kolmogorov-Smirnov against the normal CDF (n = 2^20: critical value at 1e-3 is 1.95 / sqrt(n) = 1.9e-3)
    from math import sqrt
    from scipy.special import ndtr
    s = np.sort(v); cdf = ndtr(s); k = np.arange(1, s.size + 1) / s.size
    assert max(np.abs(cdf - k).max(), np.abs(cdf - k + 1.0 / s.size).max()) < 1.95 / sqrt(s.size)
    # dims and frames are uncorrelated
    m = e[:, :64]; c = np.corrcoef(m.T); np.fill_diagonal(c, 0.0)
    assert np.abs(c).max() < 0.06


def test_train_pass_draws_what_normal_draw_reports_for_any_chunking(arch, eng):
    from vae_npvc_b200.engine import Engine
    n = 300
    x, y = _batch(n)
    theta = eng.init_theta(0, 0.1)
    st = eng.new_step_state(seed=99); st[1] = 4; st[2] = 4
    eps = eng.normal_draw(st, n, frame_offset=1000)
    g1 = torch.empty_like(theta); g2 = torch.empty_like(theta); g3 = torch.empty_like(theta)
    o1 = eng.loss_fwd_bwd(theta, x, y, eps, grad=g1)                                   # explicit draw
    o2 = eng.loss_fwd_bwd(theta, x, y, None, grad=g2, state=st, frame_offset=1000)     # in-kernel draw
    assert st.tolist()[1:3] == [5, 5]                                                  # pass and step counters advanced on the device
    for k in ("z", "mu", "lv", "xh"):
        assert torch.equal(o1[k], o2[k]), k
    assert float((g1 - g2).abs().max() / g1.abs().max()) < 1e-5
    small = Engine(arch, "cuda:0", max_chunk=64)                                       # 5 chunks: frame indices continue across chunks
    st2 = eng.new_step_state(seed=99); st2[1] = 4
    o3 = small.loss_fwd_bwd(theta, x, y, None, grad=g3, state=st2, frame_offset=1000)
    assert torch.equal(o3["z"], o1["z"]) and float((g1 - g3).abs().max() / g1.abs().max()) < 1e-4
    o4 = eng.loss_fwd_bwd(theta, x, y, None, grad=None, state=st, frame_offset=1000)   # forward only: a new pass, no step
    assert st.tolist()[1:3] == [6, 5] and not torch.equal(o4["z"], o1["z"])


def test_device_side_adam_step_equals_host_side(eng):
    theta = eng.init_theta(0, 0.1)
    g = torch.Generator(device="cpu").manual_seed(3)
    grad = torch.randn(theta.numel(), generator=g).cuda() * 1e-2
    ta, ma, va = theta.clone(), torch.zeros_like(theta), torch.zeros_like(theta)
    tb, mb, vb = theta.clone(), torch.zeros_like(theta), torch.zeros_like(theta)
    st = eng.new_step_state(0)
    for t in (1, 2, 3, 50):
        st[2] = t
        eng.adam_step(ta, grad, ma, va, t, 1e-4, 0.5, 0.999, 1e-8, 0.25)
        eng.adam_step(tb, grad, mb, vb, st, 1e-4, 0.5, 0.999, 1e-8, 0.25)
        assert torch.equal(ma, mb) and torch.equal(va, vb)
        assert float((ta - tb).abs().max()) <= 1e-9 + 1e-7 * float(ta.abs().max())     # lr_t: host double vs device double pow


@pytest.mark.parametrize("n", [16, 2048])
def test_graph_replay_equals_eager_steps(arch, monkeypatch, n):
    """trainer.vae.VAETrainer captures the training step in a CUDA graph after two eager steps: 8 steps with and without
    the graph from the same start must agree (the gradients' atomics make it 'up to summation order')."""
    from importlib import import_module
    MODEL = getattr(import_module("model.vae"), "ConvVAE")
    TRAINER = getattr(import_module("trainer.vae"), "VAETrainer")
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("NPVC_GRAPH", mode)
        machine = MODEL(arch, seed=0)
        machine.theta.copy_(machine.engine.init_theta(0, perturb=0.1))
        tr = TRAINER(machine.loss(*_batch(n, 1)), arch, None, None)
        assert tr.use_graph == (mode == "1")
        losses = []
        for i in range(8):
            losses.append(tr.opt["g"](*_batch(n, 10 + i)).clone())
        torch.cuda.synchronize()
        assert machine.state.tolist()[2] == 8 and tr.global_step == 8
        res[mode] = (machine.theta.clone(), torch.stack(losses))
        if mode == "1":
            assert any(g["graph"] is not None for g in tr._state["graphs"].values())
    monkeypatch.delenv("NPVC_GRAPH")
    # Adam normalises by sqrt(v): a parameter whose gradient is at noise level moves by up to lr per step in either
    # direction, so 8 steps at lr = 1e-4 may differ by ~1e-3 absolute there; the losses pin the rest
    d = float((res["0"][0] - res["1"][0]).abs().max())
    assert d < 8 * 1e-4 * 1.5, d
    assert float((res["0"][1] - res["1"][1]).abs().max() / res["0"][1].abs().max()) < 1e-4
    assert bool(torch.isfinite(res["1"][1]).all())


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_engine_on_a_non_current_device(arch):
    """Engine(device='cuda:1') while cuda:0 is current: every library call must run with the engine's device current (its
    device tables, kernel attributes and launches are bound to the current device) and on that device's stream."""
    from oracle import convvae_ref as R
    from vae_npvc_b200.engine import Engine
    torch.cuda.set_device(0)
    e1 = Engine(arch, "cuda:1")
    P = R.init_params(arch, 0)
    x, y, eps = R.make_inputs(arch, 8)
    d = torch.device("cuda:1")
    theta = torch.tensor(R.flatten_params(arch, P), device=d)
    grad = torch.empty_like(theta)
    out = e1.loss_fwd_bwd(theta, torch.tensor(x, dtype=torch.float32, device=d), torch.tensor(y, device=d),
                          torch.tensor(eps, dtype=torch.float32, device=d), grad=grad)
    torch.cuda.synchronize(d)
    assert torch.cuda.current_device() == 0 and out["xh"].device == d
    ref = R.forward(arch, P, x, y, eps)
    assert float(np.abs(out["xh"].cpu().numpy() - ref["xh"]).max() / np.abs(ref["xh"]).max()) < 1e-4
    assert bool(torch.isfinite(grad).all())
