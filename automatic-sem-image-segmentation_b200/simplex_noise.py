"""2-D OpenSimplex noise on numpy arrays: the `opensimplex.random_seed()` / `opensimplex.noise2array(x, y)` pair that
WGAN.simulate_masks uses to cluster particles (WassersteinGAN.py:419-425).  HOST code, not on the accelerated path.

The `opensimplex` package (a port of K. Spencer's public-domain OpenSimplex, 2014) is not installable here, so the published
algorithm is restated: a simplectic honeycomb on the stretched square lattice (stretch (1/sqrt(3) - 1)/2, squish
(sqrt(3) - 1)/2), contributions (2 - dx^2 - dy^2)^4 * <gradient, d> of the lattice points around the sample, eight
gradients of length sqrt(29) selected by a 256-entry permutation shuffled with a 64-bit LCG from the seed, the sum divided by
47.  Vectorised over the whole grid.  Parity with the package itself is UNPINNED (it cannot be run here); the tests check the
properties the caller relies on: deterministic per seed, smooth, zero-mean-ish, values inside [-1, 1].
"""
from __future__ import annotations

import time

import numpy as np

_STRETCH = (1.0 / np.sqrt(3.0) - 1.0) / 2.0
_SQUISH = (np.sqrt(3.0) - 1.0) / 2.0
_NORM = 47.0
_GRAD = np.array([5, 2, 2, 5, -5, 2, -2, 5, 5, -2, 2, -5, -5, -2, -2, -5], dtype=np.float64)
_MASK64 = (1 << 64) - 1

_perm = None


def _lcg(s: int) -> int:
    return (s * 6364136223846793005 + 1442695040888963407) & _MASK64


def _signed(s: int) -> int:
    return s - (1 << 64) if s >= (1 << 63) else s


def seed(value: int = 3) -> None:
    """opensimplex.seed(value): builds the permutation table."""
    global _perm
    source = list(range(256))
    perm = [0] * 256
    s = value & _MASK64
    for _ in range(3):
        s = _lcg(s)
    for i in range(255, -1, -1):
        s = _lcg(s)
        r = (_signed(s) + 31) % (i + 1)
        perm[i] = source[r]
        source[r] = source[i]
    _perm = np.array(perm, dtype=np.int64)


def random_seed() -> None:
    """opensimplex.random_seed(): seeds from the clock."""
    seed(time.time_ns())


def _contrib(xsb, ysb, dx, dy):
    attn = 2.0 - dx * dx - dy * dy
    idx = _perm[(_perm[xsb & 0xFF] + ysb) & 0xFF] & 0x0E
    val = _GRAD[idx] * dx + _GRAD[idx + 1] * dy
    a2 = attn * attn
    return np.where(attn > 0, a2 * a2 * val, 0.0)


def noise2array(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Noise on the grid y (rows) x x (columns): result[i, j] = noise2(x[j], y[i]), values in [-1, 1]."""
    if _perm is None:
        seed()
    X, Y = np.meshgrid(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64))
    # skew the input onto the stretched lattice, split into cell origin and position inside the cell
    s_off = (X + Y) * _STRETCH
    xs, ys = X + s_off, Y + s_off
    xsb, ysb = np.floor(xs).astype(np.int64), np.floor(ys).astype(np.int64)
    sq = (xsb + ysb) * _SQUISH
    dx0, dy0 = X - (xsb + sq), Y - (ysb + sq)
    xins, yins = xs - xsb, ys - ysb
    in_sum = xins + yins
    # the two vertices every sample sees: (1, 0) and (0, 1)
    value = _contrib(xsb + 1, ysb, dx0 - 1 - _SQUISH, dy0 - _SQUISH)
    value = value + _contrib(xsb, ysb + 1, dx0 - _SQUISH, dy0 - 1 - _SQUISH)
    lower = in_sum <= 1                                  # inside the triangle at (0, 0); otherwise the one at (1, 1)
    zins = np.where(lower, 1.0 - in_sum, 2.0 - in_sum)
    # the extra vertex: the one beyond the closer edge, or the opposite corner of the rhombus
    far = np.where(lower, (zins > xins) | (zins > yins), (zins < xins) | (zins < yins))
    x_big = xins > yins
    ext_x = np.where(lower, np.where(far, np.where(x_big, 1, -1), 1), np.where(far, np.where(x_big, 2, 0), 0))
    ext_y = np.where(lower, np.where(far, np.where(x_big, -1, 1), 1), np.where(far, np.where(x_big, 0, 2), 0))
    k = ext_x + ext_y                                    # squish terms grow with the lattice distance from the origin
    dx_ext = dx0 - ext_x - k * _SQUISH
    dy_ext = dy0 - ext_y - k * _SQUISH
    # the cell vertex of the sample's own triangle
    own = np.where(lower, 0, 1)
    dx_own = dx0 - own - 2 * own * _SQUISH
    dy_own = dy0 - own - 2 * own * _SQUISH
    value = value + _contrib(xsb + own, ysb + own, dx_own, dy_own)
    value = value + _contrib(xsb + ext_x, ysb + ext_y, dx_ext, dy_ext)
    return value / _NORM
