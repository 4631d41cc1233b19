"""Oracle formulation 2: explicit numpy tap loops (no library convolution, no autograd).

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see ``oracle/__init__.py``).

Independent restatement of the same reference functions as ``convvae_ref`` from the
*definitions* of the TF ops:
  conv2d SAME        out[o,j]          = b[o] + sum_{c,k} xpad[c, s*j + k] * W[k,0,c,o]
  conv2d_transpose   full[o, s*i + k] += x[c,i] * W[k,0,o,c];  out = full[:, crop_l : crop_l + s*H]
and used to cross-check formulation 1 (fp64 agreement ~1e-12) and to drive finite-difference
gradient checks.  Cites: ``model/vae.py:72-130``, ``util/layers.py:10-66,147-183``.
"""
import numpy as np

from .convvae_ref import (LEAK, LN_EPS, LOG_2PI, ONE_PLUS_EPS, _convt_name, enc_geometry,
                          gen_geometry)


def _ln(x, scale, offset):
    # x [N, C, H]; stats over (C, H) per frame, biased variance
    m = x.mean(axis=(1, 2), keepdims=True)
    v = ((x - m) ** 2).mean(axis=(1, 2), keepdims=True)
    return (x - m) / np.sqrt(v + LN_EPS) * scale.reshape(1, -1, 1) + offset.reshape(1, -1, 1)


def _lrelu(x):
    return np.maximum(x, LEAK * x)


def conv_same(x, W, b, s):
    """x [N,Cin,H]; W [k,1,Cin,Cout]."""
    N, Cin, H = x.shape
    k, _, _, Cout = W.shape
    Ho = -(-H // s)
    pt = max((Ho - 1) * s + k - H, 0)
    pl = pt // 2
    xp = np.zeros((N, Cin, H + pt), x.dtype)
    xp[:, :, pl:pl + H] = x
    out = np.zeros((N, Cout, Ho), x.dtype)
    for kk in range(k):
        cols = xp[:, :, kk:kk + s * (Ho - 1) + 1:s]                # [N, Cin, Ho]
        out += np.einsum("nch,co->noh", cols, W[kk, 0])
    return out + b.reshape(1, -1, 1)


def convT_same(x, W, b, s):
    """x [N,Cin,H]; W [k,1,Cout,Cin] -> [N,Cout,s*H]."""
    N, Cin, H = x.shape
    k, _, Cout, _ = W.shape
    full = np.zeros((N, Cout, s * (H - 1) + k), x.dtype)
    for kk in range(k):
        full[:, :, kk:kk + s * (H - 1) + 1:s] += np.einsum("nci,oc->noi", x, W[kk, 0])
    cl = max(k - s, 0) // 2
    return full[:, :, cl:cl + s * H] + b.reshape(1, -1, 1)


def forward(arch, params, x, y, eps):
    """Same contract as ``convvae_ref.forward`` (float64 numpy in / out, no grads)."""
    P = {k: np.asarray(v, np.float64) for k, v in params.items()}
    a = np.asarray(x, np.float64)[:, None, :]
    for i, (ci, co, k, s, H, Ho, pl, pr) in enumerate(enc_geometry(arch)):
        p = "Encoder/Conv2d-%d" % i
        a = conv_same(a, P["%s/Conv2d-%d/kernel" % (p, i)], P["%s/Conv2d-%d/bias" % (p, i)], s)
        a = _lrelu(_ln(a, P[p + "/layernorm.scale"], P[p + "/layernorm.offset"]))
    f = a.reshape(a.shape[0], -1)
    mu = f @ P["Encoder/dense/kernel"] + P["Encoder/dense/bias"]
    lv = f @ P["Encoder/dense_1/kernel"] + P["Encoder/dense_1/bias"]
    z = mu + np.asarray(eps, np.float64) * np.sqrt(np.exp(lv))
    h, w, c = arch["generator"]["hwc"]
    e = P["y_embedding/y_emb"][np.asarray(y)]
    a = (z @ P["Generator/fully_connected/weights"] + P["Generator/fully_connected/biases"]
         + e @ P["Generator/fully_connected_1/weights"] + P["Generator/fully_connected_1/biases"]
         + P["Generator/BiasAdd/biases"]).reshape(-1, c, h)
    gg = gen_geometry(arch)
    for i, (ci, co, k, s, H, Ho, cl) in enumerate(gg):
        nm = _convt_name(i)
        a = convT_same(a, P[nm + "/kernel"], P[nm + "/bias"], s)
        if i < len(gg) - 1:
            a = _lrelu(_ln(a, P["Generator/ConvT-LN%d.scale" % i], P["Generator/ConvT-LN%d.offset" % i]))
    xh = a.reshape(a.shape[0], -1)
    D_KL = (0.5 * (-lv + (np.exp(lv) + mu * mu) / ONE_PLUS_EPS - 1.0)).sum(-1).mean()
    xf = np.asarray(x, np.float64)
    logP = (-0.5 * (LOG_2PI + (xf - xh) ** 2 / ONE_PLUS_EPS)).sum(-1).mean()
    return {"mu": mu, "lv": lv, "z": z, "xh": xh, "D_KL": D_KL, "logP": logP, "G": -logP + D_KL}
