"""CPU tests of the oracle itself (the reference has no tests: PARITY UNPINNED -- these pin the
restatement against an independent second formulation, finite differences, closed forms and the
committed golden fixture)."""
import math
import os

import numpy as np
import pytest

from oracle import convvae_loops as L
from oracle import convvae_ref as R
from oracle import detrand

GOLD = os.path.join(os.path.dirname(__file__), "golden", "convvae_vcc2016_n4.npz")


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / (np.abs(np.asarray(b)).max() + 1e-300))


@pytest.fixture(scope="module")
def setup(arch):
    P = R.init_params(arch, 0)
    x, y, eps = R.make_inputs(arch, 4)
    return P, x, y, eps, R.forward(arch, P, x, y, eps, with_grads=True)


def test_param_inventory(arch):
    specs = R.param_specs(arch)
    assert len(specs) == 44 and R.n_params(arch) == 939162
    assert specs[0][0] == "y_embedding/y_emb" and specs[-1][0] == "Generator/conv2d_transpose_3/bias"
    assert R.enc_geometry(arch)[3][6:] == (3, 3) and R.gen_geometry(arch)[0][6] == 3


def test_two_formulations_agree(arch, setup):
    P, x, y, eps, a = setup
    b = L.forward(arch, P, x, y, eps)
    for k in ("mu", "lv", "z", "xh", "D_KL", "logP", "G"):
        assert rel(a[k], b[k]) < 1e-12, k


def test_golden_fixture(arch, setup):
    P, x, y, eps, a = setup
    g = np.load(GOLD)
    theta = R.flatten_params(arch, P, np.float64)
    assert abs(theta.sum() - g["theta_sum"]) < 1e-9 and np.allclose(theta[::int(g["stride"])], g["theta_sample"], rtol=0, atol=1e-15)
    assert abs(x.sum() - g["x_sum"]) < 1e-9 and abs(eps.sum() - g["eps_sum"]) < 1e-9 and (y == g["y"]).all()
    for k in ("mu", "lv", "z", "xh", "D_KL", "logP", "G"):
        assert rel(a[k], g[k]) < 1e-11, k
    flat_g = R.flatten_params(arch, a["grads"], np.float64)
    assert rel(flat_g[::int(g["stride"])], g["grad_sample"]) < 1e-10
    t1, _, _ = R.adam_step(theta, flat_g, 0.0, 0.0, 1, 1e-4, 0.5, 0.999)
    assert rel(t1[::int(g["stride"])], g["adam_theta_sample"]) < 1e-12


def test_finite_difference_gradients(arch, setup):
    P, x, y, eps, a = setup
    rs = np.random.RandomState(0)
    names = ["Encoder/Conv2d-1/Conv2d-1/kernel", "Generator/conv2d_transpose_3/kernel", "y_embedding/y_emb",
             "Encoder/Conv2d-2/layernorm.scale", "Generator/BiasAdd/biases", "Generator/conv2d_transpose/kernel",
             "Encoder/dense_1/kernel", "Generator/ConvT-LN1.offset"]
    for n in names:
        g = a["grads"][n]
        idx = tuple(int(rs.randint(0, s)) for s in g.shape)
        # y_emb rows not in the batch have exactly zero grad; pick a used row
        if n == "y_embedding/y_emb":
            idx = (int(y[0]), idx[1])
        h = 1e-6
        vals = []
        for sgn in (+1, -1):
            Pp = dict(P); w = P[n].copy(); w[idx] += sgn * h; Pp[n] = w
            vals.append(L.forward(arch, Pp, x, y, eps)["G"])
        fd = (vals[0] - vals[1]) / (2 * h)
        assert abs(fd - g[idx]) <= 2e-4 * max(1.0, abs(g[idx])), (n, fd, g[idx])


def test_frames_are_independent(arch, setup):
    """F5/F6: permuting frames permutes outputs (LN statistics never couple frames)."""
    P, x, y, eps, a = setup
    perm = np.array([2, 0, 3, 1])
    b = R.forward(arch, P, x[perm], y[perm], eps[perm])
    for k in ("mu", "z", "xh"):
        assert rel(b[k], a[k][perm]) < 1e-12


def test_kl_properties():
    import torch
    mu = torch.zeros(3, 8, dtype=torch.float64); lv = torch.zeros(3, 8, dtype=torch.float64)
    k0 = R.kld(mu, lv)                       # = 0 up to the (1 + 1e-6) divisor
    assert float(k0.abs().max()) < 8 * 1e-6
    mu = torch.randn(5, 8, dtype=torch.float64); lv = torch.randn(5, 8, dtype=torch.float64)
    assert float(R.kld(mu, lv).min()) > -1e-5


def test_transposed_conv_is_adjoint_of_conv():
    rs = np.random.RandomState(1)
    for (H, k, s, ci, co) in [(57, 7, 3, 4, 6), (19, 9, 3, 5, 3), (12, 7, 1, 2, 2)]:
        W = rs.randn(k, 1, ci, co)                    # conv HWIO
        x = rs.randn(2, ci, H * s)                    # conv input length s*H -> output H
        yv = rs.randn(2, co, H)
        cx = L.conv_same(x, W, np.zeros(co), s)
        # conv2d_transpose kernel layout [k,1,Cout_T,Cin_T] with Cout_T = ci, Cin_T = co
        ct = L.convT_same(yv, W, np.zeros(ci), s)
        if (k - s) % 2 == 0:
            assert abs((cx * yv).sum() - (x * ct).sum()) < 1e-9 * max(1.0, abs((cx * yv).sum()))


def test_same_padding_lengths(arch):
    for (ci, co, k, s, H, Ho, pl, pr) in R.enc_geometry(arch):
        assert Ho == math.ceil(H / s) and pl + pr == max((Ho - 1) * s + k - H, 0) and pl == (pl + pr) // 2


def test_adam_step_closed_form():
    g = np.array([0.5, -2.0, 0.0]); th = np.array([1.0, 1.0, 1.0])
    t1, m1, v1 = R.adam_step(th, g, 0.0, 0.0, 1, lr=1e-4, b1=0.5, b2=0.999)
    # t=1: m = (1-b1) g, v = (1-b2) g^2, lr_t = lr sqrt(1-b2)/(1-b1) -> step = lr * g/(|g| + eps*sqrt(..)) ~ lr*sign(g)
    assert np.allclose(th - t1, [1e-4, -1e-4, 0.0], rtol=1e-4, atol=1e-12)


def test_tanhize_roundtrip():
    rs = np.random.RandomState(0)
    xmin = rs.randn(513) - 3; xmax = xmin + 1 + rs.rand(513)
    x = xmin + (xmax - xmin) * rs.rand(7, 513)
    t = R.tanhize_forward(x, xmin, xmax)
    assert t.min() >= -1 and t.max() <= 1
    assert np.allclose(R.tanhize_backward(t, xmin, xmax), x)
    assert (R.tanhize_forward(xmax + 5, xmin, xmax) == 1).all() and (R.tanhize_forward(xmin - 5, xmin, xmax) == -1).all()


def test_detrand_is_stable():
    assert detrand.bits(0, 3).tolist() == detrand.bits(0, 3).tolist()
    u = detrand.uniform01(7, 1000)
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.05
    nrm = detrand.normal(3, (20000,))
    assert abs(nrm.mean()) < 0.03 and abs(nrm.std() - 1) < 0.03


def test_philox_known_answers():
    """oracle/philox.py against the Random123 known-answer vectors of Philox4x32-10 (kat_vectors)."""
    from oracle import philox
    z = np.zeros(1, np.uint32); f = np.full(1, 0xFFFFFFFF, np.uint32)
    assert [int(v[0]) for v in philox.philox4x32_10((z, z, z, z), (0, 0))] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert [int(v[0]) for v in philox.philox4x32_10((f, f, f, f), (0xFFFFFFFF, 0xFFFFFFFF))] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    pi = [np.full(1, v, np.uint32) for v in (0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344)]
    assert [int(v[0]) for v in philox.philox4x32_10(pi, (0xA4093822, 0x299F31D0))] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    e = philox.normal_draw(7, 3, 1000, 2048, 128)
    assert abs(e.mean()) < 0.01 and abs(e.std() - 1.0) < 0.01 and np.isfinite(e).all()
    # a draw depends on (seed, pass, frame, dim) only
    assert np.array_equal(philox.normal_draw(7, 3, 1500, 8, 128), e[500:508])
    assert not np.array_equal(philox.normal_draw(7, 4, 1000, 8, 128), e[:8])
