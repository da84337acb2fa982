"""Build variants of the library with extra -D flags (here, on the CPU box) and time the headline ensemble step of each
(on the GPU box).

  python bench_tools/variant_bench.py build name1:-DFOO=1 name2:-DFOO=2,-DBAR   -> bench_tools/_variants/<name>/
  python bench_tools/variant_bench.py run name1 name2 ...                         (one process per variant)
  python bench_tools/variant_bench.py one <name>                                  (used by run)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "bench_tools", "_variants")


def build(name, flags):
    from pyratslam_b200 import build as b
    d = os.path.join(OUT, name)
    os.makedirs(d, exist_ok=True)
    objs, procs = [], []
    for src in b.SOURCES:
        obj = os.path.join(d, src.replace(".cu", ".o"))
        procs.append(subprocess.Popen([b._nvcc()] + b.NVCC_FLAGS + flags + ["-c", os.path.join(b.CSRC, src), "-o", obj]))
        objs.append(obj)
    for p in procs:
        assert p.wait() == 0
    lib = os.path.join(d, "libpyratslam_b200.so")
    subprocess.check_call([b._nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                     "-cudart", "static"])
    for o in objs:
        os.remove(o)
    return lib


def one(name):
    from pyratslam_b200 import _native as nat
    if name != "default":
        nat.LIB_PATH = os.path.join(OUT, name, "libpyratslam_b200.so")
    import torch
    import bench
    from pyratslam_b200 import PoseCellEnsemble
    B = 4096
    gis, odom = bench.ensemble_inputs(B, 64, 3)
    ens = PoseCellEnsemble(bench.SHAPE, B, global_inhibition=gis)
    ens.inject(1.0, (10, 10, 18))
    od = torch.from_numpy(odom).cuda()
    for t in range(10):
        ens.update_async(od[t % 64])
    best = 1e9
    for rep in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(50):
            ens.update_async(od[t % 64])
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 50)
    alive = int((ens.state.reshape(B, -1).sum(1) > 0).sum().item())
    print("%-24s %.4f ms per ensemble update (best of 5 x 50), %d networks alive" % (name, best, alive), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        for spec in sys.argv[2:]:
            name, _, fl = spec.partition(":")
            print(build(name, [f for f in fl.split(",") if f]))
    elif sys.argv[1] == "run":
        for name in sys.argv[2:]:
            subprocess.call([sys.executable, os.path.abspath(__file__), "one", name])
    else:
        one(sys.argv[2])
