"""Shared test helpers: seeded synthetic field elements (SplitMix64 + rejection)."""
import numpy as np

P = 0xFFFFFFFF00000001
GOLDEN_DIR = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden")


def splitmix64(seed: int, count: int) -> np.ndarray:
    """Vectorised SplitMix64 stream: element i is the output for state seed + (i+1)*gamma."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, count + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def field_elems(seed: int, shape, canonical: bool = True) -> np.ndarray:
    """Uniform elements of [0,p) (rejection: a draw >= p is replaced by a fresh draw)."""
    count = int(np.prod(shape))
    out = splitmix64(seed, count)
    if canonical:
        bad = out >= np.uint64(P)
        k = 1
        while bad.any():
            out[bad] = splitmix64(seed ^ (0xA5A5A5A5 * k), int(bad.sum()))
            bad = out >= np.uint64(P)
            k += 1
    return out.reshape(shape)


def hexlist(a):
    return ["%016x" % int(x) for x in np.asarray(a, dtype=np.uint64).reshape(-1)]


def unhex(lst, shape=None):
    a = np.array([int(x, 16) for x in lst], dtype=np.uint64)
    return a if shape is None else a.reshape(shape)
