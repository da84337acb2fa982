"""Where a fused frame's time goes: C call (copies + launches + sync) vs Python bookkeeping."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from synth import synth_frames  # noqa: E402
from pyratslam_b200 import _native as nat, ros_simulate  # noqa: E402

T = 400
frames = synth_frames(np.random.default_rng(1), T)
rng = np.random.default_rng(1)
odom = np.stack([rng.uniform(0, 3, T), rng.uniform(-1, 1, T)], axis=1)
node = ros_simulate.RatslamRos()
for t in range(50):
    node.fused_frame((float(odom[t, 0]), float(odom[t, 1])), frames[t])
orig = nat.lib().prs_frame_run
acc = {"c": 0.0}


def timed_call(*a):
    t0 = time.perf_counter()
    r = orig(*a)
    acc["c"] += time.perf_counter() - t0
    return r


class L:
    def __getattr__(self, k):
        return timed_call if k == "prs_frame_run" else getattr(nat._lib, k)


real = nat._lib
nat.lib = lambda: L()
t0 = time.perf_counter()
for t in range(50, T):
    node.fused_frame((float(odom[t, 0]), float(odom[t, 1])), frames[t])
tot = time.perf_counter() - t0
n = T - 50
print("per frame: total %.1f us, inside prs_frame_run %.1f us, python around it %.1f us"
      % (tot / n * 1e6, acc["c"] / n * 1e6, (tot - acc["c"]) / n * 1e6))
# device-only time of the same sequence (no sync inside): events around 200 fused frames is not possible (the call syncs);
# instead time the pose-cell step alone and the sweep alone
nat.lib = lambda: real
e = node.pcn._ens
od = torch.zeros((1, 2), dtype=torch.float64, device="cuda")
od[0, 0] = 0.13
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(200):
    e.update_async(od)
ev1.record()
torch.cuda.synchronize()
print("pose-cell update (B=1, device, back to back): %.1f us" % (ev0.elapsed_time(ev1) / 200 * 1e3))
