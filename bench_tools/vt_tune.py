"""Sweep the tuning knobs of the view-template library sweeps (prs_vt_tune) on one GPU.

usage: python bench_tools/vt_tune.py [queries]
For every variant: the packed key and the per-template scores must equal the register-kernel's (knob 0 = 0 / knob 2 = 0),
then CUDA-event time per query over a library larger than L2 (2^20 uint8 templates bit-sliced, 2^18 float32 templates).
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyratslam_b200 import _native as nat  # noqa: E402


def timed(fn, k):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(k):
        fn(t)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 20
    peak = 6537.3
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    L = nat.lib()
    g = torch.Generator(device="cuda").manual_seed(4)
    key = torch.zeros(1, dtype=torch.int64, device="cuda")
    scratch = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    # ---- uint8, bit-sliced, reference mode
    for n in (() if "--f32-only" in sys.argv else ((1 << 20) - 37, 1 << 20)):  # a ragged size first (correctness), then the benchmark size
        lib = torch.randint(0, 256, (n, 32, 32), dtype=torch.uint8, device="cuda", generator=g)
        qs = torch.randint(0, 256, (8, 32, 32), dtype=torch.uint8, device="cuda", generator=g)
        packed = torch.zeros(int(L.prs_vt_packed_bytes(n)), dtype=torch.uint8, device="cuda")
        nat.check(L.prs_vt_pack_u8(lib.data_ptr(), n, packed.data_ptr(), 0, nat.stream_ptr()))
        del lib
        scores = torch.zeros(n, dtype=torch.int32, device="cuda")

        def sweep(t, sc=None):
            nat.check(L.prs_vt_sweep_packed_u8(packed.data_ptr(), n, qs[t % 8].data_ptr(), 0, 0, key.data_ptr(),
                                               sc.data_ptr() if sc is not None else None, scratch.data_ptr(),
                                               nat.stream_ptr()))
        nat.check(L.prs_vt_tune(0, 0))
        want = []
        for t in range(3):
            sweep(t, scores)
            want.append((int(key.item()), scores.clone()))
        for depth in (0, 2, 4, 8, 34, 44):
            for ctas in ((16,) if depth == 0 else (1,) if depth == 44 else (3, 4, 5, 6)):
                if (depth % 10) * 4 * 2056 * ctas > 225 * 1024:
                    continue
                nat.check(L.prs_vt_tune(0, depth))
                nat.check(L.prs_vt_tune(1, ctas))
                ok = True
                for t in range(3):
                    scores.zero_()
                    sweep(t, scores)
                    ok &= int(key.item()) == want[t][0] and bool(torch.equal(scores, want[t][1]))
                timed(sweep, 3)
                ms = min(timed(sweep, k) for _ in range(3))
                gbs = n * 1024 / (ms * 1e-3) / 1e9
                print("u8 packed ref n=%d depth=%d ctas/SM=%d: %.4f ms/query  %.0f GB/s  frac %.3f  %s"
                      % (n, depth, ctas, ms, gbs, gbs / peak, "OK" if ok else "MISMATCH"), flush=True)
            if n != 1 << 20 and depth == 4:
                break
        del packed, scores
    # ---- float32, reference mode
    for nf in (() if "--u8-only" in sys.argv else ((1 << 18) - 5, 1 << 18)):
        libf = torch.rand((nf, 32, 32), dtype=torch.float32, device="cuda", generator=g) * 255
        qf = torch.rand((32, 32), dtype=torch.float32, device="cuda", generator=g) * 255
        sc = torch.zeros(nf, dtype=torch.float32, device="cuda")

        def sweepf(t, s=None):
            nat.check(L.prs_vt_sweep_f32(libf.data_ptr(), nf, qf.data_ptr(), 0, 0, key.data_ptr(),
                                         s.data_ptr() if s is not None else None, nat.stream_ptr()))
        nat.check(L.prs_vt_tune(2, 0))
        sweepf(0, sc)
        want = (int(key.item()), sc.clone())
        for depth in (0, 1, 2, 3, 4, 11, 12, 13):
            for ctas in ((8,) if depth == 0 else (2, 3, 4, 6, 8, 12)):
                per_cta = depth * 4 * 3848 if depth < 10 else (depth - 10) * 4 * 7688
                if per_cta * ctas > 225 * 1024:
                    continue
                nat.check(L.prs_vt_tune(2, depth))
                nat.check(L.prs_vt_tune(3, ctas))
                sc.zero_()
                sweepf(0, sc)
                if depth < 10:   # same arithmetic in the same order
                    ok = int(key.item()) == want[0] and bool(torch.equal(sc, want[1]))
                else:            # column-pair kernel: another summation order
                    ok = (int(key.item()) & 0xFFFFFFFF) == (want[0] & 0xFFFFFFFF) and \
                        bool(((sc - want[1]).abs() <= 1e-5 * want[1].abs()).all())
                timed(sweepf, 3)
                ms = min(timed(sweepf, k) for _ in range(3))
                gbs = nf * 4096 / (ms * 1e-3) / 1e9
                print("f32 ref n=%d depth=%d ctas/SM=%d: %.4f ms/query  %.0f GB/s  frac %.3f  %s"
                      % (nf, depth, ctas, ms, gbs, gbs / peak, "OK" if ok else "MISMATCH"), flush=True)
            if nf != 1 << 18 and depth == 2:
                break
        del libf, sc


if __name__ == "__main__":
    main()
