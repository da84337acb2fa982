"""CPU oracle for the pose-cell network.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates ``ratslam/posecell_network.py`` (class ``PoseCellNetwork``) together
with the parts of ``ratslam/convolution.py`` its ``update`` reaches, in Python 3
with explicit Python-2 arithmetic.  float64 throughout, like the reference
(``posecell_network.py:27,41``).

The three device correlations are expressed with ``scipy.ndimage`` in wrap
mode -- the equivalence the reference itself asserts
(``sandbox/opencl_test2.py:306-334``, ``posecell_network.py:335,290-291,311``)
and that ``tests/golden/make_golden.py`` re-checks by literal emulation of the
OpenCL index arithmetic.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import ndimage
from scipy.special import cbrt

# ratslam/posecell_network.py:7-17 (module constants; PC_DIM_* are unused there)
PC_E_SIGMA = 1
PC_I_SIGMA = 2
PC_E_DIM = 7
PC_I_DIM = 5
PC_GLOBAL_INHIB = 0.2
PC_CELL_X_SIZE = 0.2


def diff_gaussian_3d(dim_e=PC_E_DIM, dim_i=PC_I_DIM, sigma_e=PC_E_SIGMA, sigma_i=PC_I_SIGMA):
    """7x7x7 difference-of-Gaussians, sum normalised to +1.

    Follows ``posecell_network.py:97-113`` (``order==3`` branch): scalar loop,
    ``math.exp``, window indicators that are always 1 for (7,5), then
    ``f /= abs(sum(f))`` with numpy's pairwise sum.
    """
    dim = max(dim_e, dim_i)
    c = float(dim // 2)  # py2: math.floor(7/2) == 3.0
    f = np.empty((dim, dim, dim))
    for x in range(dim):
        for y in range(dim):
            for z in range(dim):
                hi, lo = max(x, y, z), min(x, y, z)
                we = 1 if (hi <= c + dim_e and lo >= c - dim_e) else 0
                wi = 1 if (hi <= c + dim_i and lo >= c - dim_i) else 0
                num = -(x - c) ** 2 - (y - c) ** 2 - (z - c) ** 2
                f[x, y, z] = (
                    we * 1.0 / (sigma_e * math.sqrt(2 * math.pi)) ** 3 * math.exp(num / (2 * sigma_e ** 2))
                    - wi * 1.0 / (sigma_i * math.sqrt(2 * math.pi)) ** 3 * math.exp(num / (2 * sigma_i ** 2))
                )
    f /= abs(np.sum(f.ravel()))
    return f


def diff_gaussian_offset_2d(sigma_e=PC_E_SIGMA, sigma_i=PC_I_SIGMA, shape=(7, 7), origin=(0, 0)):
    """Shifted 2-D DoG, normalised, then element-wise cube root.

    ``posecell_network.py:210-222``.
    """
    x, y = np.meshgrid(np.arange(shape[0]) - origin[0], np.arange(shape[1]) - origin[1])
    c0, c1 = float(shape[0] // 2), float(shape[1] // 2)
    q = -(x - c0) ** 2 - (y - c1) ** 2
    f = (1.0 / (2 * sigma_e ** 2 * np.pi) * np.exp(q / (2 * sigma_e ** 2))
         - 1.0 / (2 * sigma_i ** 2 * np.pi) * np.exp(q / (2 * sigma_i ** 2)))
    f /= abs(np.sum(f.ravel()))
    return cbrt(f)


def diff_gaussian_offset_1d(sigma_e=PC_E_SIGMA, sigma_i=PC_I_SIGMA, size=7, origin=0):
    """Shifted 1-D DoG, normalised, cube-rooted.  ``posecell_network.py:224-235``."""
    x = np.arange(size) - origin
    c = float(size // 2)
    f = (1.0 / (sigma_e * math.sqrt(2 * np.pi)) * np.exp(-np.square(x - c) / (2 * sigma_e ** 2))
         - 1.0 / (sigma_i * math.sqrt(2 * np.pi)) * np.exp(-np.square(x - c) / (2 * sigma_i ** 2)))
    f /= abs(np.sum(f.ravel()))
    return cbrt(f)


def build_filter_lut_2d(sigma_e=PC_E_SIGMA, sigma_i=PC_I_SIGMA, shape=(7, 7), precision=1):
    """Look-up table of 2-D filters keyed by tenths of a cell.

    ``posecell_network.py:50-59``.  The origin handed to the builder is
    ``x / (precision*10)`` evaluated with **Python-2 integer division**, i.e.
    floor division: every negative key collapses to origin -1, every
    non-negative key to origin 0.
    """
    lut = {}
    for x in range(-5 * precision, 5 * precision):
        for y in range(-5 * precision, 5 * precision):
            lut[(x, y)] = diff_gaussian_offset_2d(
                sigma_e, sigma_i, shape=shape,
                origin=(x // (precision * 10), y // (precision * 10)))
    return lut


def path_integration_plan(vtrans, vrot, n_th, vtrans_scale=PC_CELL_X_SIZE):
    """The integer decisions of one path-integration step.

    Returns ``(vt, origins float[2,Th], diff float[2,Th], keys int[Th], og int)`` following
    ``posecell_network.py:252-267,249,304``:  ``vt = vtrans/0.2``, per-heading
    exact offsets, ``numpy.around`` (half to even), LUT key ``int(d_x*10)``
    (truncation toward zero; the x component is used for *both* key parts),
    and the theta origin ``floor(vr+.5)``.
    """
    vrot_scale = 2.0 * np.pi / n_th
    vt = vtrans / vtrans_scale
    vr = vrot / vrot_scale
    mid = n_th // 2
    dir_pc = np.arange(n_th).reshape((1, n_th))
    ex = np.concatenate((vt * np.cos((dir_pc - mid) * vrot_scale),
                         vt * np.sin((dir_pc - mid) * vrot_scale)), axis=0)
    origins = np.around(ex)
    diff = ex - origins
    keys = [int(diff[0, z] * 10) for z in range(n_th)]
    og = math.floor(vr + 0.5)
    return vt, origins, diff, keys, og


class PoseCellNetwork:
    """Restatement of ``posecell_network.py:22-353``."""

    def __init__(self, shape, global_inhibition=PC_GLOBAL_INHIB, **kwargs):
        self.shape = tuple(int(s) for s in shape)
        self.posecells = np.zeros(self.shape)
        self.kernel_3d = diff_gaussian_3d()
        self.global_inhibition = global_inhibition
        self.pc_vtrans_scale = PC_CELL_X_SIZE
        self.pc_vrot_scale = 2.0 * np.pi / self.shape[2]
        self.filter_dict_2d = build_filter_lut_2d()
        self.filter_dict_2d_precision = 10
        self.max_pc = (0, 0, 0)
        self.last_total = 0.0

    # posecell_network.py:322-324.  py2 callers pass math.floor() floats; old
    # numpy truncated float indices.
    def inject(self, energy, loc):
        self.posecells[tuple(int(v) for v in loc)] += energy

    # posecell_network.py:317-319: first maximum in C order.
    def get_pc_max(self):
        x, y, th = np.unravel_index(self.posecells.argmax(), self.posecells.shape)
        return (int(x), int(y), int(th))

    # posecell_network.py:244-250
    def filters_from_origins_approx(self, origins_diff, shape=(7, 7)):
        num = origins_diff.shape[1]
        filters = np.empty((shape[0], shape[1], num))
        prec = self.filter_dict_2d_precision
        for z in range(num):
            k = int(origins_diff[0, z] * prec)
            filters[:, :, z] = self.filter_dict_2d[(k, k)]  # KeyError when k == 5
        return filters

    # posecell_network.py:252-314
    def path_integration(self, vtrans, vrot):
        X, Y, T = self.shape
        vt, origins, diff, _keys, og = path_integration_plan(vtrans, vrot, T, self.pc_vtrans_scale)
        filters = self.filters_from_origins_approx(diff)  # may raise KeyError first, as in the reference
        radius = int(np.ceil(abs(vt)))
        if 3 + radius > min(X, Y):
            # convolution.py:661-675 would convolve never-written memory here.
            raise ValueError("shift radius %d does not fit the grid" % radius)
        # convolution.py:320-340 + 615-694: per-plane correlation of the plane
        # rolled by minus the integer origin, periodic.
        out = np.empty_like(self.posecells)
        for k in range(T):
            ox, oy = int(origins[0, k]), int(origins[1, k])
            plane = np.roll(self.posecells[:, :, k], (-ox, -oy), axis=(0, 1))
            out[:, :, k] = ndimage.correlate(plane, filters[:, :, k], mode="wrap")
        out[out < 0] = 0
        # convolution.py:344-359: 7-tap correlation along theta, periodic.
        f1 = diff_gaussian_offset_1d(size=7, origin=og)
        out = ndimage.correlate1d(out, f1, axis=2, mode="wrap")
        out[out < 0] = 0
        self.posecells = out

    # posecell_network.py:326-353
    def update(self, v=(0.0, 0.0)):
        vtrans, vrot = float(v[0]), float(v[1])
        pc = ndimage.correlate(self.posecells, self.kernel_3d, mode="wrap")
        gi = self.global_inhibition
        pc[pc < gi] = 0
        pc[pc >= gi] -= gi
        total = np.sum(pc.ravel())
        self.last_total = float(total)
        if total != 0:
            pc /= total
        self.posecells = pc
        self.path_integration(vtrans, vrot)
        self.max_pc = self.get_pc_max()
        return self.max_pc


def run_ensemble(shape, gis, odom, inject_at=None, energy=1.0):
    """B independent networks (BASELINE config 4), one after the other.

    ``gis``: float[B] global inhibition per network; ``odom``: float[T,B,2].
    Returns ``(argmax int[T,B,3], final states float64[B,X,Y,Th])``.
    """
    B = len(gis)
    T = odom.shape[0]
    if inject_at is None:
        inject_at = tuple(s // 2 for s in shape)
    amax = np.zeros((T, B, 3), dtype=np.int64)
    states = np.zeros((B,) + tuple(shape))
    for b in range(B):
        net = PoseCellNetwork(shape, global_inhibition=float(gis[b]))
        net.inject(energy, inject_at)
        for t in range(T):
            amax[t, b] = net.update(odom[t, b])
        states[b] = net.posecells
    return amax, states
