"""Regenerates tests/golden/convvae_vcc2016_n4.npz from the oracle (formulation 1, float64).

The reference ships no golden vectors (PARITY UNPINNED, see oracle/__init__.py); these fixtures
pin the ORACLE so that later edits to it (or to the deterministic generators) cannot drift
silently.  Inputs/weights are regenerated from oracle/detrand.py; only outputs are stored.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import convvae_ref as R  # noqa: E402
from vae_npvc_b200 import vcc2016_vae_arch  # noqa: E402

N = 4


def build():
    arch = vcc2016_vae_arch()
    P = R.init_params(arch, 0)
    x, y, eps = R.make_inputs(arch, N)
    out = R.forward(arch, P, x, y, eps, with_grads=True)
    theta = R.flatten_params(arch, P, np.float64)
    g = R.flatten_params(arch, out["grads"], np.float64)
    t1, m1, v1 = R.adam_step(theta, g, 0.0, 0.0, 1, 1e-4, 0.5, 0.999)
    stride = 997                       # sparse sample of the flat vectors (prime stride)
    fx = {
        "n": N, "mu": out["mu"], "lv": out["lv"], "z": out["z"], "xh": out["xh"],
        "D_KL": out["D_KL"], "logP": out["logP"], "G": out["G"],
        "theta_sum": theta.sum(), "theta_abs_sum": np.abs(theta).sum(), "theta_sample": theta[::stride],
        "x_sum": x.sum(), "eps_sum": eps.sum(), "y": y,
        "grad_sample": g[::stride], "grad_abs_sum": np.abs(g).sum(),
        "grad_tensor_sums": np.array([out["grads"][n].sum() for n, *_ in R.param_specs(arch)]),
        "adam_theta_sample": t1[::stride], "stride": stride,
    }
    return fx


if __name__ == "__main__":
    fx = build()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "convvae_vcc2016_n4.npz")
    np.savez_compressed(path, **fx)
    print("wrote", path, os.path.getsize(path), "bytes")
