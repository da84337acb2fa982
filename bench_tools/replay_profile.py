"""Frames per second of the replay loop: Python-driven variants against prs_replay_run with 1..8 plans in flight,
and the split of a native replay into the C call and the Python bookkeeping around it.

usage: python bench_tools/replay_profile.py [frames]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from synth import synth_frames  # noqa: E402
from pyratslam_b200 import _native as nat, ros_simulate  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
frames = synth_frames(np.random.default_rng(1), T)
rng = np.random.default_rng(1)
odom = np.stack([rng.uniform(0, 3, T), rng.uniform(-1, 1, T)], axis=1)

ros_simulate.replay(frames[:30], odom[:30], fused=True, pipelined=True)
t0 = time.perf_counter()
ref = ros_simulate.replay(frames, odom, fused=True, pipelined=True)
dt = time.perf_counter() - t0
print("pipelined (Python loop, 2 plans): %.1f us/frame  %.0f frames/s" % (dt / T * 1e6, T / dt))

for n_plans in (1, 2, 4, 8):
    node = ros_simulate.RatslamRos()
    node.replay_native(frames[:30], odom[:30], n_plans=n_plans)     # warm-up: eager frame, capture, replay
    node = ros_simulate.RatslamRos()
    node.replay_native(frames[:30], odom[:30], n_plans=n_plans)
    orig = nat.lib().prs_replay_run
    acc = {"c": 0.0}

    def timed_call(*a):
        t1 = time.perf_counter()
        r = orig(*a)
        acc["c"] += time.perf_counter() - t1
        return r

    real = nat._lib

    class L:
        def __getattr__(self, k):
            return timed_call if k == "prs_replay_run" else getattr(real, k)

    nat.lib = lambda: L()
    t0 = time.perf_counter()
    res = node.replay_native(frames[30:], odom[30:], n_plans=n_plans)
    dt = time.perf_counter() - t0
    nat.lib = lambda: real
    n = T - 30
    same = np.array_equal(res["template_index"], ref["template"][30:])
    print("native n_plans=%d: %.1f us/frame (%.0f frames/s); inside prs_replay_run %.1f us/frame, Python %.1f us/frame  %s"
          % (n_plans, dt / n * 1e6, n / dt, acc["c"] / n * 1e6, (dt - acc["c"]) / n * 1e6, "same records" if same else "MISMATCH"))

# ---- where a frame's time goes: host enqueue cost vs device time of the replayed frame graph
import ctypes  # noqa: E402
import torch  # noqa: E402

node = ros_simulate.RatslamRos()
node.replay_native(frames[:40], odom[:40], n_plans=2)
sl = node._p_slots[0]
sl["frame_np"][...] = frames[5]
sl["odom_np"][:] = (0.13, 0.02)
st = ctypes.c_void_p(node._f_stream.cuda_stream)
L = nat.lib()
for moved in (1, 0):
    for _ in range(5):
        L.prs_frame_launch(sl["plan"], moved, st)
    node._f_stream.synchronize()
    K = 500
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(node._f_stream)
    t0 = time.perf_counter()
    for _ in range(K):
        L.prs_frame_launch(sl["plan"], moved, st)
    t1 = time.perf_counter()
    e1.record(node._f_stream)
    node._f_stream.synchronize()
    print("frame graph moved=%d: host enqueue %.1f us/frame, device %.1f us/frame (500 launches back to back)"
          % (moved, (t1 - t0) / K * 1e6, e0.elapsed_time(e1) / K * 1e3))
