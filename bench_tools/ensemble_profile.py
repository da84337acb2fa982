"""A few updates of the headline ensemble (4096 x 21x21x36) for ncu captures.  python bench_tools/ensemble_profile.py [steps] [f32|f64] [path]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pyratslam_b200 import PoseCellEnsemble  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dtype = np.float64 if len(sys.argv) > 2 and sys.argv[2] == "f64" else np.float32
B = 4096
gis, odom = bench.ensemble_inputs(B, 64, 3)
ens = PoseCellEnsemble(bench.SHAPE, B, global_inhibition=gis, dtype=dtype)
if len(sys.argv) > 3:
    ens.force_path(sys.argv[3])
ens.inject(1.0, (10, 10, 18))
od = torch.from_numpy(odom).cuda()
for t in range(steps):
    ens.update_async(od[t % 64])
torch.cuda.synchronize()
print("path", ens.path)
