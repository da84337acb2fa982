"""The oracle against every fixture generated from the reference's own host code.

Fixtures: tests/golden/*.npz, written by tests/golden/make_golden.py (which
executes /root/reference/ratslam/*.py under a Python-2 shim).  Integer results
must be identical; float64 states must agree to 1e-13 relative (observed: 0).
"""
import math

import numpy as np
import pytest

from oracle import drivers, experience_map as oem, posecells as opc, view_templates as ovt
from synth import synth_frames

RTOL64 = 1e-13


def _close(a, b, rtol=RTOL64):
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() <= rtol * scale


def test_kernels(golden):
    g = golden("kernels.npz")
    assert np.array_equal(opc.diff_gaussian_3d(), g["kernel_3d"])
    lut = opc.build_filter_lut_2d()
    keys = [tuple(k) for k in g["lut_keys"].tolist()]
    assert sorted(lut.keys()) == keys
    for k, f in zip(keys, g["lut_filters"]):
        assert np.array_equal(lut[k], f)
    # the LUT collapses to two filters (py2 floor division of the key)
    f0, fm1 = lut[(0, 0)], lut[(-1, -1)]
    for k in keys:
        assert np.array_equal(lut[k], opc.diff_gaussian_offset_2d(origin=(-1 if k[0] < 0 else 0, -1 if k[1] < 0 else 0)))
    assert np.unravel_index(f0.argmax(), f0.shape) == (3, 3)
    assert np.unravel_index(fm1.argmax(), fm1.shape) == (2, 2)
    for og in range(-5, 6):
        assert np.array_equal(opc.diff_gaussian_offset_1d(size=7, origin=float(og)), g["f1d_og%+d" % og])
    # |og| >= 5: no positive lobe -> the network dies; |og| == 4 keeps one positive edge tap
    for og in (-5, 5):
        assert (g["f1d_og%+d" % og] < 0).all()
    assert (g["f1d_og-4"] > 0).tolist() == [True] + [False] * 6
    assert (g["f1d_og+4"] > 0).tolist() == [False] * 6 + [True]


def _replay_pcn(shape, odom, inject=None):
    net = opc.PoseCellNetwork(shape)
    net.inject(1, tuple(s // 2 for s in shape) if inject is None else inject)
    amax, states = [], []
    for s in range(len(odom)):
        amax.append(net.update(odom[s]))
        states.append(net.posecells.copy())
    return np.array(amax), states


@pytest.mark.parametrize("name,shape,inject", [
    ("simulate_50x50x10.npz", (50, 50, 10), None),
    ("ros_21x21x36.npz", (21, 21, 36), None),
    ("ros_21x21x36_dies.npz", (21, 21, 36), None),
    ("odd_9x8x7.npz", (9, 8, 7), (4, 3, 2)),
    ("odd_17x23x11.npz", (17, 23, 11), (16, 0, 10)),
])
def test_posecell_trajectories(golden, name, shape, inject):
    g = golden(name)
    amax, states = _replay_pcn(shape, g["odom"], inject)
    assert np.array_equal(amax, g["argmax"])
    n = 0
    for key in g.files:
        if key.startswith("state_"):
            s = int(key.split("_")[1])
            assert _close(states[s], g[key]), (name, s)
            n += 1
    assert n > 0


def test_simulate_driver_matches(golden):
    g = golden("simulate_50x50x10.npz")
    amax, totals, final = drivers.simulate_run()
    assert np.array_equal(amax, g["argmax"])
    assert _close(final, g["state_039"])
    assert abs(totals[0] - 0.0610825802870828) < 1e-15


def test_dead_network_stays_dead(golden):
    g = golden("ros_21x21x36_dies.npz")
    assert g["state_004"].max() == 0.0 and g["state_003"].max() > 0
    assert (g["argmax"][4:] == 0).all()


def test_keyerror_cases(golden):
    for v, expect in golden("keyerror.npz")["cases"]:
        net = opc.PoseCellNetwork((21, 21, 36))
        net.inject(1, (10, 10, 18))
        if expect:
            with pytest.raises(KeyError):
                net.update((v, 0.0))
        else:
            net.update((v, 0.0))


def test_view_template_mask_and_sequence(golden):
    g = golden("view_templates.npz")
    vts = ovt.ViewTemplates((32, 96), (32, 96), 2, 2, 256, 256, 45000)
    assert tuple(g["shape"]) == vts.shape == (32, 32)
    assert int(g["mask_count"]) == int(vts.mask.sum()) == 1024
    assert np.array_equal(np.flatnonzero(vts.mask.any(axis=1)), g["mask_rows"])
    assert np.array_equal(np.flatnonzero(vts.mask.any(axis=0)), g["mask_cols"])
    T = int(g["n_frames"])
    frames = synth_frames(np.random.default_rng(int(g["frame_seed"])), T)
    for t in range(T):
        n0 = len(vts.templates)
        sub = frames[t][vts.mask].reshape(vts.shape)
        vals = [int(x.match(sub)) for x in vts.templates]
        tm = vts.match(frames[t], t % 21, (2 * t) % 21, t % 36)
        assert tm.get_index() == g["index"][t]
        assert (len(vts.templates) > n0) == bool(g["created"][t])
        assert (min(vals) if vals else -1) == g["best_score"][t]
    assert g["created"].sum() < T and g["created"].sum() > 3  # both branches fired


def test_view_template_score_tables(golden):
    g = golden("view_templates.npz")
    lib, qs = g["lib_u8"], g["queries_u8"]
    for qi, q in enumerate(qs):
        got = [int(ovt.match_score(lib[i], q)) for i in range(len(lib))]
        assert got == g["scores_u8"][qi].tolist()
        assert ovt.library_scores(lib, q).tolist() == got
        gotf = [float(ovt.match_score(lib[i].astype(np.float64), q.astype(np.float64))) for i in range(len(lib))]
        assert gotf == g["scores_f64"][qi].tolist()
    # the uint8 rule is NOT a true SAD: wrap-around makes it one-sided
    true_sad = np.abs(lib[0, 8:24].astype(int) - qs[3, 8:24].astype(int)).sum()
    assert true_sad != ovt.match_scores_all_offsets(lib[0], qs[3])[7]
    # identity used by the CUDA kernel: sum((a-b) mod 256) = sum(a) - sum(b) + 256 * #{a<b}
    a, b = lib[0, 8:24].astype(np.int64), qs[3, 8:24].astype(np.int64)
    assert ovt.match_scores_all_offsets(lib[0], qs[3])[7] == a.sum() - b.sum() + 256 * (a < b).sum()


def test_view_template_threshold_is_strict(golden):
    assert golden("view_templates.npz")["threshold_case"].tolist() == [0, 0, 0, 1]


def test_circular_extension_consistency():
    rng = np.random.default_rng(5)
    lib = rng.integers(0, 256, (7, 32, 32), dtype=np.uint8)
    q = np.roll(lib[4], 11, axis=0)
    s = ovt.library_scores(lib, q, mode="circular")
    assert s[4] == 0 and int(np.argmin(s)) == 4
    assert [int(ovt.circular_score(lib[i], q)) for i in range(7)] == s.tolist()


def test_experience_map(golden):
    g = golden("experience_map.npz")
    em = oem.ExperienceMap()
    for t, (vt, vr) in enumerate(g["v"]):
        em.update(vt, vr, (t % 5, t % 7, t % 3))
        assert em.get_current_point() == tuple(g["points"][t])
        assert em.accum_delta_th == g["theta"][t]
    assert len(em.experiences) == int(g["n"])
    assert [oem.clip_rad_180(a) for a in g["clip_in"]] == g["clip_out"].tolist()


def test_replay_loop(golden):
    g = golden("replay_ros.npz")
    T = int(g["n_frames"])
    frames = synth_frames(np.random.default_rng(int(g["frame_seed"])), T)
    rec = drivers.replay_run(frames, g["odom"])
    assert np.array_equal(rec["template"], g["template"])
    assert np.array_equal(rec["created"], g["created"])
    assert np.array_equal(rec["argmax"], g["argmax"])
    assert np.array_equal(rec["n_exp"], g["n_exp"])
    assert _close(rec["final_state"], g["final_state"])
    assert np.array_equal(rec["em_xy"][-1], g["em_xy"][-1])
