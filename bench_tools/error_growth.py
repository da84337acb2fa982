"""Per-update error of network 373 of a 600-network float32 ensemble against the float64 oracle, for the fused dense kernel,
the generic kernels, the active-set kernels and the float64 build: the dynamics of some parameter regimes amplify ANY rounding
error by ~1.3x per update (the float64 path grows 1e-16 -> 1e-13 in 30 updates), so two float32 implementations drift apart
to a few 1e-5 of the peak there although each update is accurate to 1e-7.  python bench_tools/error_growth.py"""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyratslam_b200 import PoseCellEnsemble
from oracle import posecells as opc
shape, B, T = (21, 21, 36), 600, 30
rng = np.random.default_rng(12)
gis = np.linspace(0.05, 0.25, B)
odom = np.stack([rng.uniform(0, 0.3, (T, B)), rng.uniform(-0.1, 0.1, (T, B))], axis=-1)
b = 373
ens = {"resident": PoseCellEnsemble(shape, B, global_inhibition=gis), "active": PoseCellEnsemble(shape, B, global_inhibition=gis, active_set=1),
       "generic": PoseCellEnsemble(shape, B, global_inhibition=gis), "f64": PoseCellEnsemble(shape, B, global_inhibition=gis, dtype=np.float64)}
ens["generic"].force_path("generic")
ref = opc.PoseCellNetwork(shape, global_inhibition=float(gis[b]))
for e in list(ens.values()) + [ref]:
    e.inject(1.0, (10, 10, 18))
for t in range(T):
    ref.update(tuple(odom[t, b]))
    r = ref.posecells
    line = "t=%2d max=%.4f nnz=%d gi=%.4f" % (t, r.max(), (r > 0).sum(), gis[b])
    for n, e in ens.items():
        e.update(odom[t])
        s = e.state[b].permute(1, 2, 0).double().cpu().numpy()
        line += "  %s %.1e" % (n, np.abs(s - r).max() / r.max())
    print(line)
