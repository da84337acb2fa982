"""Run the single 256x256x72 network (BASELINE config 3) for a few updates; used under ncu for the per-kernel list."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyratslam_b200 import PoseCellNetwork  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
shape = tuple(int(v) for v in sys.argv[2].split("x")) if len(sys.argv) > 2 else (256, 256, 72)
net = PoseCellNetwork(shape)
net.inject(1.0, tuple(s // 2 for s in shape))
rng = np.random.default_rng(2)
od = torch.from_numpy(np.stack([rng.uniform(0, 0.3, 64), rng.uniform(-0.05, 0.05, 64)], axis=1)).cuda()
for t in range(5):
    net._ens.update_async(od[t:t + 1])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(steps):
    net._ens.update_async(od[t % 64:t % 64 + 1])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print("path=%s shape=%s ms_per_step=%.4f cells_per_s=%.3e" % (net.path, shape, ms, np.prod(shape) / (ms * 1e-3)))
