"""Generate tests/golden/*.npz by running the reference's own host code (see py2ref.py).

Run once in the build container (``python tests/golden/make_golden.py``); needs
``/root/reference``.  The fixtures are committed; the tests only read them.
Every array is produced by the *reference classes* -- ``PoseCellNetwork``,
``Convolution``, ``ViewTemplates``/``ViewTemplate``, ``ExperienceMap`` -- driven
exactly as ``simulate.py:13-40,56-58`` and ``ros_simulate.py:52-57,67-70,98-105,
125-137,152-166`` drive them.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import py2ref  # noqa: E402
from synth import synth_frames  # noqa: E402


def _t(v):
    return tuple(int(x) for x in v)


def gen_kernels(ref):
    P = ref.posecell_network.PoseCellNetwork((9, 8, 7))
    out = {"kernel_3d": P.kernel_3d}
    lut = P.filter_dict_2d
    keys = sorted(lut.keys())
    out["lut_keys"] = np.array(keys, dtype=np.int64)
    out["lut_filters"] = np.stack([lut[k] for k in keys])
    for og in range(-5, 6):
        out["f1d_og%+d" % og] = P.diff_gaussian_offset_1d(1, 2, size=7, origin=float(og))
    np.savez_compressed(os.path.join(HERE, "kernels.npz"), **out)


def run_pcn(ref, shape, odom, inject=None, keep=()):
    P = ref.posecell_network.PoseCellNetwork(shape)
    mid = tuple(int(math.floor(s / 2)) for s in shape) if inject is None else inject
    P.inject(1, mid)
    amax, states = [], {}
    for s in range(len(odom)):
        amax.append(_t(P.update(odom[s, :].copy())))
        if s in keep:
            states["state_%03d" % s] = P.posecells.copy()
    return np.array(amax, dtype=np.int64), states, P


def gen_simulate(ref):
    data = np.zeros((40, 2))
    data[:, 0] = 3
    data[4:9, 1] = np.pi / 4                      # simulate.py:38-40
    keep = (0, 1, 4, 5, 8, 9, 20, 39)
    amax, states, _ = run_pcn(ref, (50, 50, 10), data, keep=keep)
    np.savez_compressed(os.path.join(HERE, "simulate_50x50x10.npz"), odom=data, argmax=amax, **states)


def gen_ros_grid(ref):
    rng = np.random.default_rng(11)
    T = 30
    lin = rng.uniform(0.0, 3.0, T) / 10.0          # vtrans = linear.x / ODOM_FREQ
    ang = rng.uniform(-1.0, 1.0, T) / 10.0
    odom = np.stack([lin, ang], axis=1)
    keep = (0, 1, 2, 10, 29)
    amax, states, _ = run_pcn(ref, (21, 21, 36), odom, keep=keep)
    np.savez_compressed(os.path.join(HERE, "ros_21x21x36.npz"), odom=odom, argmax=amax, **states)
    # the pi/4 turn of simulate.py kills this grid (og = 5): keep that too
    data = np.zeros((8, 2))
    data[:, 0] = 0.3
    data[4:, 1] = np.pi / 4
    amax, states, _ = run_pcn(ref, (21, 21, 36), data, keep=(3, 4, 5))
    np.savez_compressed(os.path.join(HERE, "ros_21x21x36_dies.npz"), odom=data, argmax=amax, **states)


def gen_odd_grid(ref):
    rng = np.random.default_rng(12)
    T = 12
    odom = np.stack([rng.uniform(-0.55, 0.55, T), rng.uniform(-1.2, 1.2, T)], axis=1)
    keep = tuple(range(T))
    amax, states, _ = run_pcn(ref, (9, 8, 7), odom, inject=(4, 3, 2), keep=keep)
    np.savez_compressed(os.path.join(HERE, "odd_9x8x7.npz"), odom=odom, argmax=amax, **states)
    # a second injection mid-run and a larger odd grid with big shifts
    odom = np.stack([rng.uniform(-2.0, 2.0, T), rng.uniform(-0.5, 0.5, T)], axis=1)
    amax, states, _ = run_pcn(ref, (17, 23, 11), odom, inject=(16, 0, 10), keep=(0, 5, 11))
    np.savez_compressed(os.path.join(HERE, "odd_17x23x11.npz"), odom=odom, argmax=amax, **states)


def gen_keyerror(ref):
    """vtrans values whose scaled offset lands exactly on +0.5 raise KeyError (posecell_network.py:249)."""
    rows = []
    for v in (0.1, 0.5, 0.9, 0.3, 0.7, 0.05, 0.2):
        P = ref.posecell_network.PoseCellNetwork((21, 21, 36))
        P.inject(1, (10, 10, 18))
        try:
            P.update(np.array([v, 0.0]))
            rows.append((v, 0))
        except KeyError:
            rows.append((v, 1))
    np.savez_compressed(os.path.join(HERE, "keyerror.npz"), cases=np.array(rows))


def gen_view_templates(ref):
    VTs = ref.view_templates.ViewTemplates
    vts = VTs(x_range=(32, 96), y_range=(32, 96), x_step=2, y_step=2, im_x=256, im_y=256, match_threshold=45000)
    out = {"mask_rows": np.flatnonzero(vts.mask.any(axis=1)), "mask_cols": np.flatnonzero(vts.mask.any(axis=0)),
           "mask_count": np.array(int(vts.mask.sum())), "shape": np.array(vts.shape)}
    rng = np.random.default_rng(21)
    T = 40
    frames = synth_frames(rng, T)
    idx, created, nlib, best = [], [], [], []
    for t in range(T):
        n0 = len(vts.templates)
        tm = vts.match(frames[t], t % 21, (2 * t) % 21, t % 36)
        idx.append(tm.get_index())
        created.append(len(vts.templates) > n0)
        nlib.append(len(vts.templates))
        sub = frames[t][vts.mask].reshape(vts.shape)
        vals = [int(T_.match(sub)) for T_ in vts.templates[:n0]]
        best.append(min(vals) if vals else -1)
    out.update(frame_seed=np.array(21), n_frames=np.array(T), index=np.array(idx), created=np.array(created),
               n_templates=np.array(nlib), best_score=np.array(best, dtype=np.int64))
    # a score table: 12 stored templates x 6 queries, uint8 and float64
    lib = rng.integers(0, 256, (12, 32, 32), dtype=np.uint8)
    qs = rng.integers(0, 256, (6, 32, 32), dtype=np.uint8)
    qs[0] = lib[3]
    qs[1] = np.roll(lib[5], 3, axis=0)
    qs[2] = np.clip(lib[7].astype(np.int16) - 2, 0, 255).astype(np.uint8)
    VT = ref.view_templates.ViewTemplate
    tbl_u8 = np.array([[int(VT(0, 0, 0, i, lib[i]).match(q)) for i in range(12)] for q in qs], dtype=np.int64)
    libf, qsf = lib.astype(np.float64), qs.astype(np.float64)
    tbl_f = np.array([[float(VT(0, 0, 0, i, libf[i]).match(q)) for i in range(12)] for q in qsf])
    out.update(lib_u8=lib, queries_u8=qs, scores_u8=tbl_u8, scores_f64=tbl_f)
    # threshold equality counts as a match (strict '>' at view_templates.py:67)
    v2 = VTs((32, 96), (32, 96), 2, 2, 256, 256, match_threshold=int(tbl_u8[3].min()))
    f0 = np.zeros((256, 256), np.uint8)
    f1 = np.zeros((256, 256), np.uint8)
    rows = np.arange(33, 96, 2)
    f0[np.ix_(rows, rows)] = lib[int(tbl_u8[3].argmin())]
    f1[np.ix_(rows, rows)] = qs[3]
    a = v2.match(f0, 0, 0, 0).get_index()
    b = v2.match(f1, 0, 0, 0).get_index()
    v3 = VTs((32, 96), (32, 96), 2, 2, 256, 256, match_threshold=int(tbl_u8[3].min()) - 1)
    c = v3.match(f0, 0, 0, 0).get_index()
    d = v3.match(f1, 0, 0, 0).get_index()
    out.update(threshold_case=np.array([a, b, c, d]))
    np.savez_compressed(os.path.join(HERE, "view_templates.npz"), **out)


def gen_experience_map(ref):
    em = ref.experience_map.ExperienceMap()
    rng = np.random.default_rng(31)
    v = np.stack([rng.uniform(0, 0.3, 50), rng.uniform(-0.9, 0.9, 50)], axis=1)
    pts, th = [], []
    for t in range(50):
        em.update(v[t, 0], v[t, 1], (t % 5, t % 7, t % 3))
        pts.append(em.get_current_point())
        th.append(em.accum_delta_th)
    clip_in = np.array([3.2, -3.2, 7.0, -7.0, np.pi, -np.pi, 0.0, 12.6])
    clip_out = np.array([float(ref.experience_map.clip_rad_180(a)) for a in clip_in])
    np.savez_compressed(os.path.join(HERE, "experience_map.npz"), v=v, points=np.array(pts, dtype=np.float64),
                        theta=np.array(th, dtype=np.float64), n=np.array(len(em.experiences)),
                        clip_in=clip_in, clip_out=clip_out)


def gen_replay(ref):
    """The ROS loop (ros_simulate.py) replayed offline on a short synthetic stream."""
    rng = np.random.default_rng(41)
    T = 24
    frames = synth_frames(rng, T)
    odom = np.stack([rng.uniform(0, 3.0, T), rng.uniform(-1, 1, T)], axis=1)
    odom[5] = (0.0005, 0.0002)    # below the gate at ros_simulate.py:128
    odom[11] = (0.0, 0.0)
    pcn = ref.posecell_network.PoseCellNetwork((21, 21, 36))
    pcn.inject(1, (10, 10, 18))
    vts = ref.view_templates.ViewTemplates((32, 96), (32, 96), 2, 2, 256, 256, 45000)
    em = ref.experience_map.ExperienceMap()
    tidx, created, amax, nexp = [], [], [], []
    for t in range(T):
        lin, ang = float(odom[t, 0]), float(odom[t, 1])
        if abs(lin) > 0.001 or abs(ang) > 0.001:
            vtrans, vrot = lin / 10, ang / 10
            pcn.update((vtrans, vrot))
            em.update(vtrans, vrot, pcn.get_pc_max())
        pm = _t(pcn.get_pc_max())
        n0 = len(vts.templates)
        tm = vts.match(frames[t], pm[0], pm[1], pm[2])
        tidx.append(tm.get_index())
        created.append(len(vts.templates) > n0)
        amax.append(pm)
        nexp.append(len(em.experiences))
    np.savez_compressed(os.path.join(HERE, "replay_ros.npz"), frame_seed=np.array(41), n_frames=np.array(T),
                        odom=odom, template=np.array(tidx), created=np.array(created),
                        argmax=np.array(amax, dtype=np.int64), n_exp=np.array(nexp),
                        final_state=pcn.posecells, em_xy=np.array(em.get_points(), dtype=np.float64))


def main():
    ref = py2ref.load_reference()
    gen_kernels(ref)
    gen_simulate(ref)
    gen_ros_grid(ref)
    gen_odd_grid(ref)
    gen_keyerror(ref)
    gen_view_templates(ref)
    gen_experience_map(ref)
    gen_replay(ref)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print("%-28s %8d bytes" % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == "__main__":
    main()
