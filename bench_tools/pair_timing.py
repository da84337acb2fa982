"""Per-stage cycles of the pair kernel measured inside the kernel (clock64 between barriers, CTA 0).
Builds a profiling copy of the library with -DPRS_PAIR_TIMING into bench_tools/_timing/ (on the CPU box: `build`),
then runs the ensemble workload on it (on the GPU box).   python bench_tools/pair_timing.py [build|run] [f32|f64]"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "bench_tools", "_timing")


def build():
    os.makedirs(OUT, exist_ok=True)
    from pyratslam_b200 import build as b
    objs, procs = [], []
    for src in b.SOURCES:
        obj = os.path.join(OUT, src.replace(".cu", ".o"))
        procs.append(subprocess.Popen([b._nvcc()] + b.NVCC_FLAGS + ["-DPRS_PAIR_TIMING", "-c", os.path.join(b.CSRC, src), "-o", obj]))
        objs.append(obj)
    for p in procs:
        assert p.wait() == 0
    lib = os.path.join(OUT, "libpyratslam_b200.so")
    subprocess.check_call([b._nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"])
    for o in objs:
        os.remove(o)
    return lib


if __name__ == "__main__":
    if "build" in sys.argv:
        print(build())
        sys.exit(0)
    from pyratslam_b200 import _native as nat
    nat.LIB_PATH = os.path.join(OUT, "libpyratslam_b200.so")
    import numpy as np
    import torch
    import bench
    from pyratslam_b200 import PoseCellEnsemble
    dtype = np.float64 if "f64" in sys.argv else np.float32
    B = 4096
    gis, odom = bench.ensemble_inputs(B, 64, 3)
    ens = PoseCellEnsemble(bench.SHAPE, B, global_inhibition=gis, dtype=dtype)
    ens.force_path("pair")
    ens.inject(1.0, (10, 10, 18))
    od = torch.from_numpy(odom).cuda()
    for t in range(5):
        ens.update_async(od[t])
    buf = (ctypes.c_ulonglong * 12)()
    L = nat.lib()
    L.prs_debug_pair_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.prs_debug_pair_cycles(buf, 1)
    steps = 20
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(steps):
        ens.update_async(od[t % 64])
    e1.record()
    torch.cuda.synchronize()
    L.prs_debug_pair_cycles(buf, 0)
    ncl = int(os.environ.get("PRS_PAIR_CLUSTERS", "0")) or (148 if dtype == np.float32 else 74)
    nets = steps * ((B + ncl - 1) // ncl)
    names = ["loop tail -> top", "1 theta (global -> E,I)", "2 y pass", "3 x pass + sum", "3b store A2", "4 7x7 + cluster barrier",
             "5 theta + store + max", "5b argmax + cluster barrier"]
    tot = sum(buf[i] for i in range(8))
    print("%s, %.4f ms/update, %d clusters assumed" % (np.dtype(dtype).name, e0.elapsed_time(e1) / steps, ncl))
    for i, n in enumerate(names):
        print("%-30s %8.0f clk/network  %5.1f%%" % (n, buf[i] / nets, 100.0 * buf[i] / tot))
    print("%-30s %8.0f clk/network" % ("total", tot / nets))
