"""Deterministic synthetic inputs: element i, limb l <- splitmix64(seed + 2i + l) mod p.
numpy generator used by bench.py and tools (the product path shares no code with the test infrastructure)."""
import numpy as np

P = np.uint64(0xFFFFFFFF00000001)


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def fill_ext(seed, n, out=None, chunk=1 << 22, start=0):
    """2n u64 limbs of extension elements start .. start+n-1 of the stream `seed`."""
    out = np.empty(2 * n, np.uint64) if out is None else out
    for s in range(0, 2 * n, chunk):
        e = min(2 * n, s + chunk)
        with np.errstate(over="ignore"):
            v = _splitmix64(np.arange(s, e, dtype=np.uint64) + np.uint64((seed + 2 * start) & 0xFFFFFFFFFFFFFFFF))
        out[s:e] = np.where(v >= P, v - P, v)
    return out


def fill_base(seed, n, start=0):
    """n base elements start .. start+n-1 of the stream `seed` (element i <- splitmix64(seed + 2i) mod p)."""
    with np.errstate(over="ignore"):
        v = _splitmix64(np.arange(2 * start, 2 * (start + n), 2, dtype=np.uint64) + np.uint64(seed))
    return np.where(v >= P, v - P, v)
