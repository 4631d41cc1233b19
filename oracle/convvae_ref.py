"""Oracle formulation 1: torch-CPU restatement of the reference ConvVAE graph.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``) -- PARITY UNPINNED by the reference.

Restates, op for op, in NCHW with library convolutions + autograd:
  * ``model/vae.py:72-82``   _encoder  (5x conv2d_nchw_layernorm -> flatten -> 2 dense heads)
  * ``model/vae.py:84-103``  _generator (embedding lookup, _merge, reshape, 4x conv2d_transpose,
                             Layernorm + lrelu on all but the last)
  * ``model/vae.py:106-130`` loss  (GaussianSampleLayer, GaussianKLD, GaussianLogDensity, means)
  * ``model/vae.py:139-145`` encode / decode
  * ``util/layers.py:10-44`` Layernorm (moments over axes [1,2,3], per-channel scale/offset,
                             eps 1e-5), ``:47-66`` conv2d_nchw_layernorm, ``:147-183`` lrelu,
                             GaussianSampleLayer, GaussianLogDensity, GaussianKLD (EPSILON 1e-6)
  * ``trainer/vae.py:15-24`` tf.train.AdamOptimizer(lr, beta1, beta2).minimize over ALL
                             trainable variables (TF-form Adam, eps=1e-8 outside the sqrt)
  * ``analyzer.py:75-87``    Tanhize

TF-1.x semantics restated here (TensorFlow is not installable in this image):
  * tf.layers.conv2d(padding='same', channels_first): H_out = ceil(H/s),
    pad_total = max((H_out-1)*s + k - H, 0), pad_left = pad_total // 2; kernel HWIO [k,1,Cin,Cout]
  * tf.layers.conv2d_transpose(padding='same'): H_out = s*H, == conv_transpose2d(padding=(k-s)//2);
    kernel [k,1,Cout,Cin]
  * tf.nn.moments: biased variance
  * variable creation order == tf.trainable_variables() order (``param_specs``)
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import detrand

LN_EPS = 1e-5
LEAK = 0.02
# util/layers.py:7  EPSILON = tf.constant(1e-6, tf.float32); used as exp(0) + EPSILON in fp32
ONE_PLUS_EPS = float(np.float32(1.0) + np.float32(1e-6))
LOG_2PI = float(np.log(np.float32(2.0) * np.float32(math.pi)))


# --------------------------------------------------------------------------------------------
# parameters: names / shapes / order / initialisers
# --------------------------------------------------------------------------------------------
def _convt_name(i):
    return "Generator/conv2d_transpose" + ("" if i == 0 else "_%d" % i)


def enc_geometry(arch):
    """[(Cin, Cout, k, s, H_in, H_out, pad_l, pad_r)] for the encoder (SAME, channels_first)."""
    H = arch["hwc"][0]
    cin = arch["hwc"][2]
    out = []
    enc = arch["encoder"]
    for o, k, s in zip(enc["output"], enc["kernel"], enc["stride"]):
        k, s = k[0], s[0]
        Ho = -(-H // s)
        pt = max((Ho - 1) * s + k - H, 0)
        out.append((cin, o, k, s, H, Ho, pt // 2, pt - pt // 2))
        H, cin = Ho, o
    return out


def gen_geometry(arch):
    """[(Cin, Cout, k, s, H_in, H_out, crop_l)] for the generator (SAME transposed conv)."""
    g = arch["generator"]
    H, _, cin = g["hwc"]
    out = []
    for o, k, s in zip(g["output"], g["kernel"], g["stride"]):
        k, s = k[0], s[0]
        pt = max(k - s, 0)
        out.append((cin, o, k, s, H, H * s, pt // 2))
        H, cin = H * s, o
    return out


def param_specs(arch):
    """[(tf_name, shape, fan_in, fan_out, kind)] in tf.trainable_variables() creation order.

    kind in {'glorot', 'zeros', 'ones'} (TF default initialisers on this path).
    """
    z = arch["z_dim"]
    specs = [("y_embedding/y_emb", (arch["y_dim"], z), arch["y_dim"], z, "glorot")]
    eg = enc_geometry(arch)
    for i, (ci, co, k, s, H, Ho, pl, pr) in enumerate(eg):
        p = "Encoder/Conv2d-%d" % i
        specs += [
            ("%s/Conv2d-%d/kernel" % (p, i), (k, 1, ci, co), k * ci, k * co, "glorot"),
            ("%s/Conv2d-%d/bias" % (p, i), (co,), 0, 0, "zeros"),
            ("%s/layernorm.offset" % p, (co, 1, 1), 0, 0, "zeros"),
            ("%s/layernorm.scale" % p, (co, 1, 1), 0, 0, "ones"),
        ]
    flat = eg[-1][1] * eg[-1][5]
    for nm in ("Encoder/dense", "Encoder/dense_1"):
        specs += [(nm + "/kernel", (flat, z), flat, z, "glorot"), (nm + "/bias", (z,), 0, 0, "zeros")]
    gh, gw, gc = arch["generator"]["hwc"]
    fo = gh * gw * gc
    for nm in ("Generator/fully_connected", "Generator/fully_connected_1"):
        specs += [(nm + "/weights", (z, fo), z, fo, "glorot"), (nm + "/biases", (fo,), 0, 0, "zeros")]
    specs.append(("Generator/BiasAdd/biases", (fo,), 0, 0, "zeros"))
    gg = gen_geometry(arch)
    for i, (ci, co, k, s, H, Ho, cl) in enumerate(gg):
        nm = _convt_name(i)
        specs += [
            (nm + "/kernel", (k, 1, co, ci), k * co, k * ci, "glorot"),
            (nm + "/bias", (co,), 0, 0, "zeros"),
        ]
        if i < len(gg) - 1:
            specs += [
                ("Generator/ConvT-LN%d.offset" % i, (co, 1, 1), 0, 0, "zeros"),
                ("Generator/ConvT-LN%d.scale" % i, (co, 1, 1), 0, 0, "ones"),
            ]
    return specs


def n_params(arch):
    return sum(int(np.prod(s[1])) for s in param_specs(arch))


def init_params(arch, seed=0, perturb=0.1):
    """Deterministic float64 parameters: glorot-uniform kernels; biases / LN offsets
    U(-perturb, perturb) and LN scales 1 + U(-perturb, perturb) so those paths are exercised
    (TF would start them at exactly 0 / 1)."""
    params = {}
    for j, (name, shape, fi, fo, kind) in enumerate(param_specs(arch)):
        if kind == "glorot":
            lim = math.sqrt(6.0 / (fi + fo))
            v = detrand.uniform(seed, shape, -lim, lim, stream=j)
        else:
            base = 1.0 if kind == "ones" else 0.0
            v = base + detrand.uniform(seed, shape, -perturb, perturb, stream=j)
        params[name] = v
    return params


def flatten_params(arch, params, dtype=np.float32):
    return np.concatenate([np.asarray(params[n], dtype=np.float64).reshape(-1)
                           for n, *_ in param_specs(arch)]).astype(dtype)


def unflatten_params(arch, flat):
    out, o = {}, 0
    for name, shape, *_ in param_specs(arch):
        n = int(np.prod(shape))
        out[name] = np.asarray(flat[o:o + n]).reshape(shape)
        o += n
    assert o == len(flat)
    return out


def make_inputs(arch, n, seed=1, eps_seed=2, n_speakers=None):
    """x ~ U(-1,1) [n,513], y ~ U{0..y_dim-1}, eps ~ N(0,1) [n, z_dim] (float64 / int64)."""
    H = arch["hwc"][0]
    x = detrand.uniform(seed, (n, H), -1.0, 1.0, stream=0)
    y = detrand.integers(seed, (n,), n_speakers or arch["y_dim"], stream=1)
    eps = detrand.normal(eps_seed, (n, arch["z_dim"]), stream=0)
    return x, y, eps


# --------------------------------------------------------------------------------------------
# forward graph
# --------------------------------------------------------------------------------------------
def _layernorm(x, scale, offset):
    """util/layers.py:10-44: moments over [1,2,3] (biased), per-channel scale/offset [C,1,1]."""
    m = x.mean(dim=(1, 2, 3), keepdim=True)
    v = ((x - m) ** 2).mean(dim=(1, 2, 3), keepdim=True)
    return (x - m) * torch.rsqrt(v + LN_EPS) * scale.reshape(1, -1, 1, 1) + offset.reshape(1, -1, 1, 1)


class _LreluWithBranch(torch.autograd.Function):
    """lrelu whose BACKWARD takes the branch (slope 1 or leak) from a given mask instead of from its own
    input.  lrelu' is discontinuous at 0: a pre-activation closer to 0 than the fp32 rounding error of
    the forward legitimately takes the other branch in an fp32 implementation (TensorFlow included)
    and moves that element's gradient by 0.98*dy.  Gradient-parity tests therefore hand the oracle the
    branches the implementation under test actually took (and separately check that they differ from
    the oracle's own only where |pre-activation| is tiny)."""

    @staticmethod
    def forward(ctx, x, pos):
        ctx.save_for_backward(pos)
        return torch.maximum(x, LEAK * x)

    @staticmethod
    def backward(ctx, g):
        (pos,) = ctx.saved_tensors
        return g * torch.where(pos, torch.ones_like(g), torch.full_like(g, LEAK)), None


def _lrelu(x, pos=None):
    """util/layers.py:147-149: tf.maximum(x, leak*x), leak=0.02.  pos: see _LreluWithBranch."""
    if pos is None:
        return torch.maximum(x, LEAK * x)
    return _LreluWithBranch.apply(x, torch.as_tensor(pos, dtype=torch.bool).reshape(x.shape))


def toeplitz_index(k, H):
    """Index table for the k-tap stride-1 SAME transposed conv written as a dense matrix:
    out[o] = sum_i x[i] * W[o - i + (k-1)//2]; entries outside [0,k) are masked."""
    o = np.arange(H)[None, :]
    i = np.arange(H)[:, None]
    t = o - i + (k - 1) // 2
    return np.clip(t, 0, k - 1), ((t >= 0) & (t < k))


def encoder(P, arch, x, acts=None, lrelu_pos=None):
    """model/vae.py:72-82.  x: [N,1,513,1] -> (mu, lv) [N, z]."""
    for i, (ci, co, k, s, H, Ho, pl, pr) in enumerate(enc_geometry(arch)):
        p = "Encoder/Conv2d-%d" % i
        W = P["%s/Conv2d-%d/kernel" % (p, i)]                      # [k,1,Cin,Cout] HWIO
        b = P["%s/Conv2d-%d/bias" % (p, i)]
        x = F.pad(x, (0, 0, pl, pr))                               # SAME padding on H
        x = F.conv2d(x, W.permute(3, 2, 0, 1), b, stride=(s, 1))   # -> OIHW
        x = _layernorm(x, P[p + "/layernorm.scale"], P[p + "/layernorm.offset"])
        x = _lrelu(x, None if lrelu_pos is None else lrelu_pos.get("enc%d" % i))
        if acts is not None:
            acts["enc%d" % i] = x
    f = x.reshape(x.shape[0], -1)                                  # slim.flatten of NCHW: c*H + h
    mu = f @ P["Encoder/dense/kernel"] + P["Encoder/dense/bias"]
    lv = f @ P["Encoder/dense_1/kernel"] + P["Encoder/dense_1/bias"]
    return mu, lv


def generator(P, arch, z, y, acts=None, lrelu_pos=None):
    """model/vae.py:84-103.  z [N,z], y [N] int64 -> xh NCHW [N,1,513,1]."""
    g = arch["generator"]
    h, w, c = g["hwc"]
    e = P["y_embedding/y_emb"][y]                                  # embedding_lookup
    x = (z @ P["Generator/fully_connected/weights"] + P["Generator/fully_connected/biases"]
         + e @ P["Generator/fully_connected_1/weights"] + P["Generator/fully_connected_1/biases"]
         + P["Generator/BiasAdd/biases"])                          # _merge: sum of FCs + bias_add
    x = x.reshape(-1, c, h, w)
    if acts is not None:
        acts["merge"] = x
    gg = gen_geometry(arch)
    for i, (ci, co, k, s, H, Ho, cl) in enumerate(gg):
        nm = _convt_name(i)
        W = P[nm + "/kernel"]                                      # [k,1,Cout,Cin]
        b = P[nm + "/bias"]
        if s == 1 and k > 64:
            # stride-1 wide kernel: same op as conv_transpose2d(padding=(k-1)//2), written as a
            # dense Toeplitz matmul (conv_transpose2d backward on CPU crawls for 1025 taps)
            idx, mask = toeplitz_index(k, H)
            idx_t = torch.as_tensor(idx)
            mask_t = torch.as_tensor(mask, dtype=W.dtype)
            T = W[:, 0][idx_t] * mask_t[:, :, None, None]          # [i, o_pos, Cout, Cin]
            xin = x[:, :, :, 0]                                    # [N, Cin, H]
            out = torch.einsum("nci,ipoc->nop", xin, T)            # [N, Cout, H]
            x = out.unsqueeze(-1) + b.reshape(1, -1, 1, 1)
        else:
            # SAME transposed conv = full transposed conv cropped to s*H from crop_left = (k-s)//2
            # (== padding=(k-s)//2 when k-s is even, as in the reference architecture)
            full = F.conv_transpose2d(x, W.permute(3, 2, 0, 1), None, stride=(s, 1))
            x = full[:, :, cl:cl + s * H, :] + b.reshape(1, -1, 1, 1)
        if i < len(gg) - 1:
            x = _layernorm(x, P["Generator/ConvT-LN%d.scale" % i], P["Generator/ConvT-LN%d.offset" % i])
            x = _lrelu(x, None if lrelu_pos is None else lrelu_pos.get("gen%d" % i))
        if acts is not None:
            acts["gen%d" % i] = x
    return x


def kld(mu, lv):
    """util/layers.py:170-183 with mu2 = lv2 = 0: per-frame sum over z."""
    return (0.5 * (-lv + (torch.exp(lv) + mu * mu) / ONE_PLUS_EPS - 1.0)).sum(-1)


def log_density(x, xh):
    """util/layers.py:159-167 with log_var = 0: per-frame sum over the 513 bins."""
    return (-0.5 * (LOG_2PI + (x - xh) ** 2 / ONE_PLUS_EPS)).sum(-1)


def _as_torch(params, dtype, requires_grad=False):
    P = {}
    for k, v in params.items():
        t = torch.tensor(np.asarray(v), dtype=dtype)
        t.requires_grad_(requires_grad)
        P[k] = t
    return P


def forward(arch, params, x, y, eps, dtype=torch.float64, with_grads=False, with_acts=False, lrelu_pos=None):
    """model/vae.py:106-130 ``loss``.  x [N,513]; y [N] int64; eps [N,z] (the N(0,1) draw of
    GaussianSampleLayer made explicit).  Returns dict of numpy arrays.
    lrelu_pos: optional {"enc<i>" / "gen<i>": bool [N,C,H,1]} branch masks for the lrelu BACKWARD
    (see _LreluWithBranch); the forward values never depend on it."""
    P = _as_torch(params, dtype, requires_grad=with_grads)
    xt = torch.tensor(np.asarray(x), dtype=dtype).reshape(-1, 1, arch["hwc"][0], 1)
    yt = torch.tensor(np.asarray(y), dtype=torch.int64)
    et = torch.tensor(np.asarray(eps), dtype=dtype)
    acts = {} if with_acts else None
    mu, lv = encoder(P, arch, xt, acts, lrelu_pos)
    z = mu + et * torch.sqrt(torch.exp(lv))                        # util/layers.py:152-156
    xh = generator(P, arch, z, yt, acts, lrelu_pos)
    D_KL = kld(mu, lv).mean()
    logP = log_density(xt.reshape(xt.shape[0], -1), xh.reshape(xh.shape[0], -1)).mean()
    G = -logP + D_KL
    out = {"mu": mu, "lv": lv, "z": z, "xh": xh.reshape(xh.shape[0], -1),
           "D_KL": D_KL, "logP": logP, "G": G}
    res = {k: v.detach().numpy() for k, v in out.items()}
    if with_acts:
        res["acts"] = {k: v.detach().numpy() for k, v in acts.items()}
    if with_grads:
        names = list(P.keys())
        gs = torch.autograd.grad(G, [P[n] for n in names], allow_unused=False)
        res["grads"] = {n: g.numpy() for n, g in zip(names, gs)}
    return res


def encode(arch, params, x, dtype=torch.float64):
    """model/vae.py:139-141: z_mu only (no sampling)."""
    P = _as_torch(params, dtype)
    xt = torch.tensor(np.asarray(x), dtype=dtype).reshape(-1, 1, arch["hwc"][0], 1)
    mu, lv = encoder(P, arch, xt)
    return mu.numpy(), lv.numpy()


def decode(arch, params, z, y, dtype=torch.float64):
    """model/vae.py:143-145: generator + nchw_to_nhwc -> [N,513,1,1]."""
    P = _as_torch(params, dtype)
    xh = generator(P, arch, torch.tensor(np.asarray(z), dtype=dtype),
                   torch.tensor(np.asarray(y), dtype=torch.int64))
    return xh.permute(0, 2, 3, 1).numpy()


# --------------------------------------------------------------------------------------------
# optimiser / normaliser
# --------------------------------------------------------------------------------------------
def adam_step(theta, grad, m, v, t, lr=1e-4, b1=0.5, b2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer (trainer/vae.py:16-24), TF form, step t >= 1:
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMA; theta -= lr_t*m/(sqrt(v)+eps)."""
    lr_t = lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    m = b1 * m + (1.0 - b1) * grad
    v = b2 * v + (1.0 - b2) * grad * grad
    theta = theta - lr_t * m / (np.sqrt(v) + eps)
    return theta, m, v


def tanhize_forward(x, xmin, xmax):
    """analyzer.py:82-84."""
    return np.clip((x - xmin) / (xmax - xmin), 0.0, 1.0) * 2.0 - 1.0


def tanhize_backward(x, xmin, xmax):
    """analyzer.py:86-87."""
    return (x * 0.5 + 0.5) * (xmax - xmin) + xmin
