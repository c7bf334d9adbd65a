"""The wave kernel's integer divisions (pixel seed -> row / column, path number -> tile; render.cu: div_by) go through
host-prepared reciprocals rounded up (api.cu: `(1.0 / d) * (1 + 2^-50)`): `uint32(double(n) * r)` must equal `n / d`
for every 32-bit numerator the renderer can produce.  Same arithmetic in numpy f64."""
import numpy as np


def _div_by(n, d):
    inv_up = (1.0 / np.float64(d)) * (1.0 + 2.0 ** -50)
    return np.floor(n.astype(np.float64) * inv_up).astype(np.uint64)


def test_reciprocal_division_is_exact():
    rng = np.random.Generator(np.random.Philox(7))
    divisors = np.unique(np.concatenate([
        np.array([1, 2, 3, 5, 7, 599, 600, 800, 1200, 3840, 65536, 100 << 16, 4096 << 16, (1 << 30) - 1, 1 << 30], dtype=np.uint64),
        rng.integers(1, 1 << 30, 2000, dtype=np.uint64)]))
    for d in divisors:
        k = rng.integers(0, (1 << 32) // int(d) + 1, 64, dtype=np.uint64)
        n = np.concatenate([k * d, k * d + d - 1, np.maximum(k * d, 1) - 1, rng.integers(0, 1 << 32, 64, dtype=np.uint64),
                            np.array([0, 1, (1 << 31) - 1, 1 << 31, (1 << 32) - 1], dtype=np.uint64)])
        n = n[n < (1 << 32)]
        assert (_div_by(n, d) == n // d).all(), f"divisor {d}"
