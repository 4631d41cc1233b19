"""Shared helpers of the parity tests (test infrastructure)."""
import numpy as np

from oracle import convvae_ref as R


def lrelu_branches(getbuf, arch, P, n):
    """The lrelu branch (pre-activation >= 0) an implementation took for every Layernorm output,
    recomputed the way its backward does -- from the raw conv output, mean and rstd it kept in its
    workspace (getbuf(name) -> flat array) -- and returned in the oracle's NCHW layout.
    lrelu' is discontinuous at 0, so gradient parity is checked against the oracle differentiating
    with THESE branches (oracle.convvae_ref._LreluWithBranch); check_branches() verifies that they
    differ from the oracle's own only within rounding distance of the kink."""
    layers = [("enc%d" % i, "e%d" % i, "Encoder/Conv2d-%d/layernorm" % i, g[1], g[5]) for i, g in enumerate(R.enc_geometry(arch))]
    layers += [("gen%d" % i, "g%d" % i, "Generator/ConvT-LN%d" % i, g[1], g[5]) for i, g in enumerate(R.gen_geometry(arch)[:-1])]
    pos = {}
    for key, tag, pname, co, ho in layers:
        c = np.asarray(getbuf("c_" + tag), np.float32)[:n * ho * co].reshape(n, ho, co)     # channels-last
        mean = np.asarray(getbuf("mean_" + tag), np.float32)[:n].reshape(n, 1, 1)
        rstd = np.asarray(getbuf("rstd_" + tag), np.float32)[:n].reshape(n, 1, 1)
        gamma = np.asarray(P[pname + ".scale"], np.float32).reshape(1, 1, co)
        beta = np.asarray(P[pname + ".offset"], np.float32).reshape(1, 1, co)
        u = ((c - mean) * rstd) * gamma + beta
        pos[key] = np.ascontiguousarray((u >= 0).transpose(0, 2, 1))[..., None]             # [n, C, H, 1]
    return pos


def check_branches(pos, acts, tol=1e-4):
    """Branches may differ from the oracle's only where |pre-activation| < tol (and only rarely)."""
    for k, m in pos.items():
        a = acts[k]
        u = np.where(a >= 0, a, a / R.LEAK)
        flipped = m != (a >= 0)
        assert flipped.sum() <= max(2, 1e-4 * a.size) and (np.abs(u[flipped]) < tol).all(), (k, int(flipped.sum()))
