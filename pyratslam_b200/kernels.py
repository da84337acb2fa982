"""Host-side filter builders (numpy float64) for the pose-cell network.

These produce the same tables as the reference's builder methods
(``ratslam/posecell_network.py:50-59,97-141,194-235``); they are cheap, run once
per network, and stay on the host so that no device libm result can change a
filter.  What is new here is ``separable_dog_factors``: the reference's 7x7x7
difference-of-Gaussians is the difference of two rank-one (separable) Gaussians,
which is how the CUDA kernels evaluate it (42 MAC per cell instead of 343).
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import cbrt

PC_E_SIGMA = 1
PC_I_SIGMA = 2
PC_E_DIM = 7
PC_I_DIM = 5
PC_GLOBAL_INHIB = 0.2
PC_CELL_X_SIZE = 0.2


def _axes(dim, order):
    c = dim // 2
    d = np.arange(dim, dtype=np.float64) - c
    grids = np.meshgrid(*([d] * order), indexing="ij")
    return c, grids


def diff_gaussian(dim_e=PC_E_DIM, dim_i=PC_I_DIM, sigma_e=PC_E_SIGMA, sigma_i=PC_I_SIGMA, order=3):
    """DoG of ``order`` dimensions with side ``max(dim_e, dim_i)``, |sum| normalised to 1.

    Same values as ``PoseCellNetwork.diff_gaussian`` (posecell_network.py:97-141).  The window
    indicators ``center - dim <= idx <= center + dim`` are kept; they are all ones for (7, 5).
    """
    if order not in (1, 2, 3):
        raise ValueError("order must be 1, 2 or 3")
    dim = max(dim_e, dim_i)
    c, grids = _axes(dim, order)
    r2 = sum(g * g for g in grids)
    idx_hi = np.maximum.reduce([g + c for g in grids])
    idx_lo = np.minimum.reduce([g + c for g in grids])
    we = ((idx_hi <= c + dim_e) & (idx_lo >= c - dim_e)).astype(np.float64)
    wi = ((idx_hi <= c + dim_i) & (idx_lo >= c - dim_i)).astype(np.float64)
    f = (we / (sigma_e * math.sqrt(2 * math.pi)) ** order * np.exp(-r2 / (2.0 * sigma_e ** 2))
         - wi / (sigma_i * math.sqrt(2 * math.pi)) ** order * np.exp(-r2 / (2.0 * sigma_i ** 2)))
    return f / abs(f.sum())


def diff_gaussian_separable(dim_e=PC_E_DIM, dim_i=PC_I_DIM, sigma_e=PC_E_SIGMA, sigma_i=PC_I_SIGMA):
    """Cube root of the normalised 1-D DoG (posecell_network.py:194-208; unused by ``update``)."""
    return cbrt(diff_gaussian(dim_e, dim_i, sigma_e, sigma_i, order=1))


def build_kernel(dim, sigma, order=3):
    """Plain normalised Gaussian (posecell_network.py:62-94; unused by ``update``)."""
    _, grids = _axes(dim, order)
    r2 = sum(g * g for g in grids)
    f = 1.0 / (sigma * math.sqrt(2 * math.pi)) * np.exp(-r2 / (2.0 * sigma ** 2))
    return f / abs(f.sum())


def diff_gaussian_offset_2d(sigma_e=PC_E_SIGMA, sigma_i=PC_I_SIGMA, shape=(7, 7), origin=(0, 0)):
    """Cube root of the normalised 2-D DoG whose centre is moved by ``origin`` (posecell_network.py:210-222).

    numpy's default ``meshgrid`` indexing makes ``origin[0]`` act along axis 1, as in the reference.
    """
    u = np.arange(shape[0]) - origin[0] - float(shape[0] // 2)
    v = np.arange(shape[1]) - origin[1] - float(shape[1] // 2)
    uu, vv = np.meshgrid(u, v)
    q = -(uu ** 2) - vv ** 2
    f = (1.0 / (2 * sigma_e ** 2 * np.pi) * np.exp(q / (2 * sigma_e ** 2))
         - 1.0 / (2 * sigma_i ** 2 * np.pi) * np.exp(q / (2 * sigma_i ** 2)))
    return cbrt(f / abs(f.sum()))


def diff_gaussian_offset_1d(sigma_e=PC_E_SIGMA, sigma_i=PC_I_SIGMA, size=7, origin=0):
    """Cube root of the normalised 1-D DoG centred at ``size//2 + origin`` (posecell_network.py:224-235)."""
    d = np.arange(size) - origin - float(size // 2)
    f = (1.0 / (sigma_e * math.sqrt(2 * np.pi)) * np.exp(-np.square(d) / (2 * sigma_e ** 2))
         - 1.0 / (sigma_i * math.sqrt(2 * np.pi)) * np.exp(-np.square(d) / (2 * sigma_i ** 2)))
    return cbrt(f / abs(f.sum()))


def build_diff_gaussian_set_2d(sigma_e=PC_E_SIGMA, sigma_i=PC_I_SIGMA, shape=(7, 7), precision=1):
    """The LUT of posecell_network.py:50-59.

    Keys are tenths of a cell in [-5, 4]^2.  The reference divides the key by ``precision*10`` with
    Python-2 integer division, so the origin is -1 for negative key parts and 0 otherwise.
    """
    lut = {}
    cache = {}
    for x in range(-5 * precision, 5 * precision):
        for y in range(-5 * precision, 5 * precision):
            o = (x // (precision * 10), y // (precision * 10))
            if o not in cache:
                cache[o] = diff_gaussian_offset_2d(sigma_e, sigma_i, shape=shape, origin=o)
            lut[(x, y)] = cache[o]
    return lut


def separable_dog_factors(dim_e=PC_E_DIM, dim_i=PC_I_DIM, sigma_e=PC_E_SIGMA, sigma_i=PC_I_SIGMA):
    """``(ge, gi, aE, aI)`` with ``kernel_3d == aE*ge(x)ge(x)ge - aI*gi(x)gi(x)gi``.

    Holds whenever both window indicators are all ones (true for the reference's (7, 5) because
    ``center +- dim`` spans the whole 7-wide cube, posecell_network.py:105,108).
    """
    dim = max(dim_e, dim_i)
    c = dim // 2
    if not (c - dim_e <= 0 and c - dim_i <= 0):
        raise ValueError("windowed DoG is not separable for dim_e=%d dim_i=%d" % (dim_e, dim_i))
    d = np.arange(dim, dtype=np.float64) - c
    ge = np.exp(-d * d / (2.0 * sigma_e ** 2))
    gi = np.exp(-d * d / (2.0 * sigma_i ** 2))
    ae = 1.0 / (sigma_e * math.sqrt(2 * math.pi)) ** 3
    ai = 1.0 / (sigma_i * math.sqrt(2 * math.pi)) ** 3
    norm = abs(ae * ge.sum() ** 3 - ai * gi.sum() ** 3)
    return ge, gi, ae / norm, ai / norm


def theta_filter_table(og_range, sigma_e=PC_E_SIGMA, sigma_i=PC_I_SIGMA):
    """``[2*og_range+1, 7]`` theta filters for integer origins -og_range..og_range."""
    return np.stack([diff_gaussian_offset_1d(sigma_e, sigma_i, size=7, origin=float(og))
                     for og in range(-og_range, og_range + 1)])
