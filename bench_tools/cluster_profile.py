import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from pyratslam_b200 import PoseCellNetwork
rng = np.random.default_rng(2)
od = torch.from_numpy(np.stack([rng.uniform(0, 0.3, 64), rng.uniform(-0.05, 0.05, 64)], axis=1)).cuda()
for shape in [(50, 50, 10), (21, 21, 36)]:
    net = PoseCellNetwork(shape)
    net._ens.force_path("cluster")
    net.inject(1.0, tuple(s // 2 for s in shape))
    for t in range(6):
        net._ens.update_async(od[t:t + 1])
    torch.cuda.synchronize()
