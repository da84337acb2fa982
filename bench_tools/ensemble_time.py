"""Device time per update of an ensemble of any shape.  python bench_tools/ensemble_time.py 50x50x10 2600 [f32|f64] [path]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyratslam_b200 import PoseCellEnsemble  # noqa: E402

shape = tuple(int(v) for v in sys.argv[1].split("x"))
B = int(sys.argv[2])
dtype = np.float64 if len(sys.argv) > 3 and sys.argv[3] == "f64" else np.float32
ens = PoseCellEnsemble(shape, B, global_inhibition=np.linspace(0.05, 0.25, B), dtype=dtype)
if len(sys.argv) > 4:
    ens.force_path(sys.argv[4])
ens.inject(1.0, tuple(s // 2 for s in shape))
rng = np.random.default_rng(5)
od = torch.from_numpy(np.stack([rng.uniform(0, 0.3, (16, B)), rng.uniform(-0.1, 0.1, (16, B))], axis=-1)).cuda()
for t in range(4):
    ens.update_async(od[t])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 12
e0.record()
for t in range(steps):
    ens.update_async(od[t % 16])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
cells = B * int(np.prod(shape))
nb = np.dtype(dtype).itemsize
print("shape=%s B=%d %s path=%s: %.4f ms/update, %.3e cell-updates/s, %.0f GB/s algorithmic (%.3f of 6537)"
      % (sys.argv[1], B, np.dtype(dtype).name, ens.path, ms, cells / (ms * 1e-3), 2 * nb * cells / (ms * 1e-3) / 1e9,
         2 * nb * cells / (ms * 1e-3) / 1e9 / 6537.3))
