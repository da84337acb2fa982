"""CPU oracle for the two driver loops.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

* ``simulate_run``  -- ``ratslam/simulate.py:13-40,56-58`` without the plotting.
* ``replay_run``    -- the update-loop semantics of ``ratslam/ros_simulate.py``
  (``:52-57,67-70,98-105,125-137,152-166``) as an offline, single-threaded
  replay: odometry update (if above the 0.001 gate), then the image match with
  the fresh arg-max, then nothing else (``ExperienceMap.update`` is called from
  inside the odometry update, ``:137``).
"""
from __future__ import annotations

import math

import numpy as np

from .experience_map import ExperienceMap, LinkedExperienceMap
from .posecells import PoseCellNetwork
from .view_templates import ViewTemplates

POSE_SIZE_SIM = (50, 50, 10)     # simulate.py:9
POSE_SIZE_ROS = (21, 21, 36)     # ros_simulate.py:31
IM_SIZE = (256, 256)             # ros_simulate.py:32
X_RANGE = (32, 96)
Y_RANGE = (32, 96)
X_STEP = 2
Y_STEP = 2
MATCH_THRESHOLD = 45000
ODOM_FREQ = 10


def simulate_data(steps=40):
    """``simulate.py:38-40``."""
    data = np.zeros((steps, 2))
    data[:, 0] = 3
    data[4:9, 1] = np.pi / 4
    return data


def simulate_run(data=None, shape=POSE_SIZE_SIM, keep_states=False):
    """Returns ``(argmax int[T,3], totals float[T], states or final state)``."""
    if data is None:
        data = simulate_data()
    pcn = PoseCellNetwork(shape)
    mid = (math.floor(shape[0] / 2), math.floor(shape[1] / 2), math.floor(shape[2] / 2))
    pcn.inject(1, mid)
    amax = np.zeros((len(data), 3), dtype=np.int64)
    totals = np.zeros(len(data))
    states = []
    for s in range(len(data)):
        amax[s] = pcn.update(data[s, :])
        totals[s] = pcn.last_total
        if keep_states:
            states.append(pcn.posecells.copy())
    return amax, totals, (np.stack(states) if keep_states else pcn.posecells)


def replay_run(frames, odom, shape=POSE_SIZE_ROS, match_threshold=MATCH_THRESHOLD, inject_energy=None,
               experience_links=False):
    """Offline replay of the ROS loop.

    ``frames``: uint8[T,256,256]; ``odom``: float64[T,2] = (linear.x, angular.z).
    ``inject_energy``: if set, the coupling the reference left commented out at ``ros_simulate.py:106-108`` is
    enabled: ``pcn.inject(energy, template_match.location())`` after every match.
    ``experience_links``: use ``LinkedExperienceMap`` (the specification of the links / loop-closure extension), each
    odometry update tagged with the most recent template match.
    Returns a dict of per-frame records.
    """
    pcn = PoseCellNetwork(shape)
    mid = (math.floor(shape[0] / 2), math.floor(shape[1] / 2), math.floor(shape[2] / 2))
    pcn.inject(1, mid)
    vts = ViewTemplates(X_RANGE, Y_RANGE, X_STEP, Y_STEP, IM_SIZE[0], IM_SIZE[1], match_threshold)
    em = LinkedExperienceMap(shape) if experience_links else ExperienceMap()
    last_vt = None
    T = len(frames)
    rec = {
        "template": np.zeros(T, np.int64), "created": np.zeros(T, np.bool_),
        "argmax": np.zeros((T, 3), np.int64), "n_exp": np.zeros(T, np.int64),
        "em_xy": np.zeros((T, 2)),
    }
    for t in range(T):
        lin, ang = float(odom[t, 0]), float(odom[t, 1])
        if abs(lin) > 0.001 or abs(ang) > 0.001:           # ros_simulate.py:128
            vtrans, vrot = lin / ODOM_FREQ, ang / ODOM_FREQ  # :157-158
            pcn.update((vtrans, vrot))                     # :135
            em.update(vtrans, vrot, pcn.get_pc_max(), last_vt if experience_links else None)      # :136-137
        pc_max = pcn.get_pc_max()                          # :103
        n_before = len(vts.templates)
        tm = vts.match(frames[t], pc_max[0], pc_max[1], pc_max[2])  # :104
        if inject_energy is not None:
            pcn.inject(inject_energy, tm.location())       # :106-108 (commented out in the reference)
        last_vt = tm.get_index()
        rec["template"][t] = tm.get_index()
        rec["created"][t] = len(vts.templates) > n_before
        rec["argmax"][t] = pc_max
        rec["n_exp"][t] = len(em.experiences)
        if em.current_exp is not None:
            rec["em_xy"][t] = em.get_current_point()
    rec["em"] = em
    rec["final_state"] = pcn.posecells
    rec["n_templates"] = len(vts.templates)
    return rec
