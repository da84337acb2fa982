"""Device time per update of a 21x21x36 float32 ensemble as a function of the network count B, per kernel family
(16 dependent updates captured in one CUDA graph, so the Python enqueue cost is not in the number).
  python bench_tools/b_sweep.py [B ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyratslam_b200 import PoseCellEnsemble  # noqa: E402

SHAPE = (21, 21, 36)
Bs = [int(a) for a in sys.argv[1:]] or [1, 8, 13, 16, 32, 64, 74, 100, 128, 148, 200, 296, 444, 512, 592]
rng = np.random.default_rng(2)
for B in Bs:
    od = torch.from_numpy(np.stack([rng.uniform(0, 0.3, (16, B)), rng.uniform(-0.05, 0.05, (16, B))], axis=-1)).cuda()
    row = []
    for path in ["auto", "cluster", "resident", "pair"]:
        ens = PoseCellEnsemble(SHAPE, B, global_inhibition=np.linspace(0.05, 0.25, B))
        try:
            ens.force_path(path)
        except ValueError:
            row.append("%s n/a" % path)
            continue
        if path == "cluster" and B > 300:
            row.append("cluster skipped")
            continue
        ens.inject(1.0, (10, 10, 18))
        for t in range(4):
            ens.update_async(od[t])
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for t in range(16):
                    ens.update_async(od[t])
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 160 * 1e3
        row.append("%s(%s) %7.1f us" % (path, ens.path, us))
    print("B=%4d  " % B + "   ".join(row), flush=True)
