"""Latency of ONE network update per kernel family: device time (update_async, CUDA events) and the time of
the blocking public call (PoseCellNetwork.update, wall clock).  usage: python bench_tools/single_latency.py [steps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyratslam_b200 import PoseCellNetwork  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(2)
od_h = np.stack([rng.uniform(0, 0.3, 64), rng.uniform(-0.05, 0.05, 64)], axis=1)
od = torch.from_numpy(od_h).cuda()
for shape in [(50, 50, 10), (21, 21, 36), (64, 64, 36)]:
    for path in ["cluster", "resident", "tiled", "generic"]:
        net = PoseCellNetwork(shape)
        try:
            net._ens.force_path(path)
        except ValueError:
            continue
        net.inject(1.0, tuple(s // 2 for s in shape))
        for t in range(10):
            net._ens.update_async(od[t:t + 1])
        torch.cuda.synchronize()
        # device time without the Python enqueue cost: 32 consecutive updates captured into one CUDA graph
        # (prs_pc_step issues plain launches when the caller's stream is capturing)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for t in range(32):
                    net._ens.update_async(od[t:t + 1])
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        dev_us = e0.elapsed_time(e1) / 320 * 1e3
        for t in range(10):
            net.update(od_h[t])
        t0 = time.perf_counter()
        for t in range(steps):
            net.update(od_h[t % 64])
        host_us = (time.perf_counter() - t0) / steps * 1e6
        print("shape=%-12s path=%-8s device %.1f us/update   update() %.1f us/call" % ("x".join(map(str, shape)), path, dev_us, host_us))
