"""Latency of one PoseCellNetwork.update(): C call (copies + kernels + sync) vs the Python around it."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyratslam_b200 import PoseCellNetwork, _native as nat  # noqa: E402

for shape in [(21, 21, 36), (50, 50, 10)]:
    net = PoseCellNetwork(shape)
    net.inject(1.0, tuple(s // 2 for s in shape))
    rng = np.random.default_rng(0)
    od = np.stack([rng.uniform(0.01, 0.29, 600), rng.uniform(-0.05, 0.05, 600)], axis=1)
    for t in range(50):
        net.update(od[t])
    real = nat._lib
    acc = {"c": 0.0}
    orig = real.prs_pc_step_host_xyz

    def timed_call(*a):
        t0 = time.perf_counter()
        r = orig(*a)
        acc["c"] += time.perf_counter() - t0
        return r

    class L:
        def __getattr__(self, k):
            return timed_call if k == "prs_pc_step_host_xyz" else getattr(real, k)

    nat.lib = lambda: L()
    t0 = time.perf_counter()
    for t in range(50, 550):
        net.update(od[t])
    tot = time.perf_counter() - t0
    nat.lib = lambda: real
    print("%s path=%s: update() %.1f us = C call %.1f us + python %.1f us" %
          (shape, net.path, tot / 500 * 1e6, acc["c"] / 500 * 1e6, (tot - acc["c"]) / 500 * 1e6))
