"""Version-independent deterministic random streams for fixtures (TEST INFRASTRUCTURE).

A counter-based generator (splitmix64 finaliser on ``seed * 2^32 + index``) written in plain
uint64 numpy arithmetic, so that the golden fixtures under ``tests/golden/`` can store only
OUTPUTS and regenerate the inputs/weights bit-identically on any numpy version.
"""
import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_G = np.uint64(0x9E3779B97F4A7C15)


def _mix(z):
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def bits(seed, n, stream=0):
    """n uint64 values for (seed, stream)."""
    with np.errstate(over="ignore"):
        base = _mix(np.uint64(seed) * _G + np.uint64(stream) * np.uint64(0xD1342543DE82EF95))
        idx = np.arange(n, dtype=np.uint64)
        return _mix(base + (idx + np.uint64(1)) * _G)


def uniform01(seed, n, stream=0):
    """float64 in [0, 1) with 53 random bits."""
    return (bits(seed, n, stream) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def uniform(seed, shape, lo=-1.0, hi=1.0, stream=0):
    n = int(np.prod(shape)) if len(shape) else 1
    return (lo + (hi - lo) * uniform01(seed, n, stream)).reshape(shape)


def normal(seed, shape, stream=0):
    """Standard normal via Box-Muller on two independent streams."""
    n = int(np.prod(shape)) if len(shape) else 1
    u1 = 1.0 - uniform01(seed, n, stream * 2 + 101)   # (0, 1]
    u2 = uniform01(seed, n, stream * 2 + 102)
    return (np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)).reshape(shape)


def integers(seed, shape, hi, stream=0):
    n = int(np.prod(shape)) if len(shape) else 1
    return (bits(seed, n, stream) % np.uint64(hi)).astype(np.int64).reshape(shape)
