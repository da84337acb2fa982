"""Device time per update of the active-set path against the dense kernels, same ensemble, same odometry.
python bench_tools/active_time.py 21x21x36 4096 [f32|f64] [replicas]
Replicas: that many ensembles are stepped in turn so that the state streamed per update exceeds L2."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyratslam_b200 import PoseCellEnsemble  # noqa: E402

shape = tuple(int(v) for v in sys.argv[1].split("x"))
B = int(sys.argv[2])
dtype = np.float64 if len(sys.argv) > 3 and sys.argv[3] == "f64" else np.float32
R = int(sys.argv[4]) if len(sys.argv) > 4 else 1
rng = np.random.default_rng(5)
od = torch.from_numpy(np.stack([rng.uniform(0, 0.3, (16, B)), rng.uniform(-0.1, 0.1, (16, B))], axis=-1)).cuda()
cells = B * int(np.prod(shape))
nb = np.dtype(dtype).itemsize
res = {}
for mode in (0, 1, 2):
    ens = [PoseCellEnsemble(shape, B, global_inhibition=np.linspace(0.05, 0.25, B), dtype=dtype, active_set=mode)
           for _ in range(R)]
    for e in ens:
        e.inject(1.0, tuple(s // 2 for s in shape))
    for t in range(6):
        for e in ens:
            e.update_async(od[t])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 12
    e0.record()
    for t in range(steps):
        for e in ens:
            e.update_async(od[(6 + t) % 16])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (steps * R)
    res[mode] = (ms, ens[0]._argmax.cpu().numpy().copy(), ens[0].state.clone())
    nzc = int((ens[0].state != 0).sum().item())
    print("shape=%s B=%d %s active_set=%d path=%s: %.4f ms/update, %.3e cell-updates/s, %.0f GB/s at %d B/cell; %d non-zero "
          "cells (%.1f per network)" % (sys.argv[1], B, np.dtype(dtype).name, mode, ens[0].path, ms, cells / (ms * 1e-3),
                                        2 * nb * cells / (ms * 1e-3) / 1e9, 2 * nb, nzc, nzc / B), flush=True)
    del ens
for m in (1, 2):
    same = np.array_equal(res[m][1], res[0][1])
    d = (res[m][2] - res[0][2]).abs().max().item() / res[0][2].abs().max().item()
    print("active_set=%d vs dense: arg-max identical=%s, state rel diff %.2e, speed-up %.2fx" % (m, same, d, res[0][0] / res[m][0]))
