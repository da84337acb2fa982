"""Multi-GPU parity check, run under torchrun (one rank per GPU, NCCL):

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_check.py

* template library sharded by contiguous ranges + one device-side MIN exchange of the packed key per query
  (csrc/sharded.cu over CUDA-IPC peer memory; the NCCL all-reduce path is checked as well)
  == numpy.argmin over the whole library (ties -> lowest global index), create-or-match identical on all ranks;
* pose-cell ensemble sharded by network, no collective in the step, results gathered and compared with the oracle.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import posecells as opc  # noqa: E402
from oracle import view_templates as ovt  # noqa: E402
from pyratslam_b200 import PoseCellEnsemble, ShardedViewTemplates  # noqa: E402
from pyratslam_b200.sharding import shard_range  # noqa: E402


def run_checks(rank, world, exchange="auto"):
    """Sharded library and sharded ensemble against the oracle on the ranks of the CURRENT process group.
    Returns the name of the exchange the library used ("fused" or "nccl")."""
    # ---- view templates
    rng = np.random.default_rng(17)
    n = 5003
    lib = rng.integers(0, 256, (n, 32, 32), dtype=np.uint8)
    lo, hi = shard_range(n, rank, world)
    lib[n - 2] = lib[7]                                   # duplicate living in the last shard
    used = None
    for mode in ("ref", "circular"):
        svt = ShardedViewTemplates(lib[lo:hi], lo, match_threshold=45000, mode=mode, exchange=exchange)
        used = svt.exchange
        assert svt.n_total == n
        queries = [np.clip(lib[n - 2].astype(np.int16) - 1, 0, 255).astype(np.uint8),   # tie 7 / n-2 -> 7
                   np.roll(lib[n // 2 + 11], 5, axis=0),                                # lives in a middle/last shard
                   rng.integers(0, 256, (32, 32), dtype=np.uint8)]                      # novel -> created
        for qi, q in enumerate(queries):
            ref = ovt.library_scores(lib, q, mode=mode)
            score, idx = svt.match_key(torch.from_numpy(q).cuda())
            assert (score, idx) == (int(ref.min()), int(np.argmin(ref))), (rank, mode, qi, score, idx)
        batch = svt.match_keys(torch.from_numpy(np.stack(queries)).cuda())     # one exchange for the three queries
        assert batch == [(int(r.min()), int(np.argmin(r))) for r in (ovt.library_scores(lib, q, mode=mode) for q in queries)]
        index, created = svt.match(torch.from_numpy(queries[2]).cuda())
        assert created and index == n and svt.n_total == n + 1
        index, created = svt.match(torch.from_numpy(queries[2]).cuda())     # now it is in the last shard
        assert not created and index == n
        assert svt.match(torch.from_numpy(queries[0]).cuda()) == (7, False)
        # several appends in a row: the owning shard grows slot by slot, indices stay contiguous
        for j in range(3):
            qn = rng.integers(0, 256, (32, 32), dtype=np.uint8)
            assert svt.match(torch.from_numpy(qn).cuda()) == (n + 1 + j, True)
            assert svt.match(torch.from_numpy(qn).cuda()) == (n + 1 + j, False)
        svt.close()
    # float32 profiles (true SAD): min and first index over the shards
    libf = rng.integers(0, 256, (301, 32, 32)).astype(np.float32)
    lo, hi = shard_range(301, rank, world)
    svt = ShardedViewTemplates(libf[lo:hi], lo, match_threshold=2000.0, mode="ref", exchange=exchange)
    qf = libf[200] + 1.0
    ref = ovt.library_scores(libf, qf)
    score, idx = svt.match_key(torch.from_numpy(qf).cuda())
    assert idx == int(np.argmin(ref)) == 200 and abs(score - float(ref.min())) <= 1e-5 * float(ref.min()), (score, idx)
    assert svt.match(torch.from_numpy(qf).cuda()) == (200, False)
    qn = rng.integers(0, 256, (32, 32)).astype(np.float32)
    assert svt.match(torch.from_numpy(qn).cuda()) == (301, True)
    assert svt.match(torch.from_numpy(qn).cuda()) == (301, False)
    svt.close()

    # ---- pose-cell ensemble
    shape, B, T = (21, 21, 36), 64, 8
    gis = np.linspace(0.05, 0.25, B)
    odom = np.stack([rng.uniform(0, 0.3, (T, B)), rng.uniform(-0.1, 0.1, (T, B))], axis=-1)
    lo, hi = shard_range(B, rank, world)
    ens = PoseCellEnsemble(shape, hi - lo, global_inhibition=gis[lo:hi])
    ens.inject(1.0, (10, 10, 18))
    mine = torch.from_numpy(ens.run(odom[:, lo:hi])).cuda()                 # [T, b, 3]
    if world > 1:
        sizes = [shard_range(B, r, world) for r in range(world)]
        parts = [torch.zeros((T, b - a, 3), dtype=torch.int64, device="cuda") for a, b in sizes]
        dist.all_gather(parts, mine)
        got = torch.cat(parts, dim=1).cpu().numpy()
    else:
        got = mine.cpu().numpy()
    pick = [0, B // 2 - 1, B // 2, B - 1]
    want, _ = opc.run_ensemble(shape, gis[pick], odom[:, pick])
    assert np.array_equal(got[:, pick], want)
    if world > 1:
        dist.barrier()
    return used


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    used = run_checks(rank, world, exchange=os.environ.get("PRS_EXCHANGE", "auto"))
    if os.environ.get("PRS_EXCHANGE", "auto") == "auto":
        run_checks(rank, world, exchange="nccl")                          # the torch.distributed path as well
    if rank == 0:
        print("multigpu_check ok: world=%d exchange=%s" % (world, used))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
