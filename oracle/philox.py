"""Philox4x32-10 + Box-Muller in numpy: the checker of the in-kernel N(0,1) sampler (TEST INFRASTRUCTURE ONLY).

The reference draws the noise of GaussianSampleLayer with tf.random_normal (util/layers.py:154), an unseeded
Philox stream that cannot be reproduced; the product defines its own counter layout (include/npvc_b200.h,
npvc_train_fwd_bwd): key = (seed_lo, seed_hi ^ draws_hi), counter = (dim, frame_lo, frame_hi, draws_lo), output
words 0 / 1 -> u1 = ((w0 >> 8) + 1) 2^-24, u2 = (w1 >> 8) 2^-24, eps = sqrt(-2 ln u1) cos(2 pi u2).
Philox4x32-10: Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(counter, key):
    """counter: 4 uint32 arrays, key: 2 uint32 scalars / arrays -> 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, np.uint32) for c in counter]
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            h0, l0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            h1, l1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = h1 ^ c1 ^ k0, l1, h0 ^ c3 ^ k1, l0
            k0 = np.uint32(k0 + W0); k1 = np.uint32(k1 + W1)
    return c0, c1, c2, c3


def normal_draw(seed, draws, frame0, n, z):
    """eps[n, z] (float64) of the pass numbered `draws` over frames frame0 .. frame0 + n - 1."""
    frame = (np.uint64(frame0) + np.arange(n, dtype=np.uint64))[:, None] * np.ones((1, z), np.uint64)
    dim = np.arange(z, dtype=np.uint32)[None, :] * np.ones((n, 1), np.uint32)
    seed, draws = int(seed) & (2 ** 64 - 1), int(draws) & (2 ** 64 - 1)
    ctr = (dim, (frame & np.uint64(0xFFFFFFFF)).astype(np.uint32), (frame >> np.uint64(32)).astype(np.uint32),
           np.full((n, z), draws & 0xFFFFFFFF, np.uint32))
    w = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) ^ (draws >> 32)))
    u1 = ((w[0] >> np.uint32(8)).astype(np.float64) + 1.0) * 2.0 ** -24
    u2 = (w[1] >> np.uint32(8)).astype(np.float64) * 2.0 ** -24
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
