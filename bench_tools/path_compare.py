"""Device time per update of a 21x21x36 ensemble for every kernel family that supports it.
  python bench_tools/path_compare.py [B ...]      (default: 4096; float32 and float64)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyratslam_b200 import PoseCellEnsemble  # noqa: E402

SHAPE = (21, 21, 36)
NC = 21 * 21 * 36


def time_path(B, dtype, path, steps=30):
    gis = np.linspace(0.05, 0.25, B)
    R = max(1, int(np.ceil(300e6 / (B * NC * np.dtype(dtype).itemsize))))   # replicas: > L2 of distinct state
    R = min(R, 64)
    ens = []
    for _ in range(R):
        e = PoseCellEnsemble(SHAPE, B, global_inhibition=gis, dtype=dtype)
        try:
            e.force_path(path)
        except ValueError:
            return None
        e.inject(1.0, (10, 10, 18))
        ens.append(e)
    rng = np.random.default_rng(3)
    od = torch.from_numpy(np.stack([rng.uniform(0, 0.3, (16, B)), rng.uniform(-0.1, 0.1, (16, B))], axis=-1)).cuda()
    for t in range(5):
        ens[t % R].update_async(od[t % 16])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(steps):
        ens[t % R].update_async(od[t % 16])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, R


if __name__ == "__main__":
    Bs = [int(a) for a in sys.argv[1:]] or [4096]
    for B in Bs:
        for dtype in (np.float32, np.float64):
            for path in ("resident", "pair", "cluster", "generic"):
                if path == "generic" and dtype == np.float32 and B > 512:
                    continue
                r = time_path(B, dtype, path)
                if r is None:
                    continue
                ms, R = r
                print("B=%5d %-8s %-9s %9.4f ms/update  %6.2f us/network  %.3e cell-updates/s  (%d replicas)"
                      % (B, np.dtype(dtype).name, path, ms, ms * 1e3 / B, B * NC / (ms * 1e-3), R), flush=True)
