"""world_size-2 gloo tests of the multi-GPU host logic (no GPU needed).

The view-template library shards by contiguous template ranges; each rank produces one packed key
from its shard (here: from the oracle, on the CPU) and a MIN all-reduce must give every rank the
answer numpy.argmin gives on the whole library -- including ties (lowest global index wins) and
empty shards.  Pose-cell ensembles shard by network with no exchange; only the partition is checked.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import view_templates as ovt
from pyratslam_b200 import sharding as sh


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        lib = rng.integers(0, 256, (37, 32, 32), dtype=np.uint8)
        lib[30] = lib[5]                      # an exact duplicate: tie between shards
        queries = [np.clip(lib[30].astype(np.int16) - 1, 0, 255).astype(np.uint8),  # ties 5 and 30 -> 5
                   lib[33], rng.integers(0, 256, (32, 32), dtype=np.uint8)]
        out = []
        for n_total in (37, 1):               # n_total=1: rank 1 owns an empty shard
            lo, hi = sh.shard_range(n_total, rank, world)
            for qv in queries:
                if hi > lo:
                    s = ovt.library_scores(lib[lo:hi], qv)
                    j = int(np.argmin(s))
                    key = sh.pack_key(int(s[j]), lo + j)
                    key_t = torch.tensor([key], dtype=torch.int64)
                else:
                    key_t = torch.tensor([-1], dtype=torch.int64)   # what the kernel leaves for n == 0
                sh.reduce_packed_key(key_t)
                score, idx = sh.unpack_key(int(key_t.item()))
                full = ovt.library_scores(lib[:n_total], qv)
                assert score == int(full.min()) and idx == int(np.argmin(full)), (rank, n_total, score, idx)
                out.append(sh.decide(score, idx, n_total, 45000))
            # the batched form (ShardedViewTemplates.match_keys): ONE reduction for all the queries' keys
            keys = []
            for qv in queries:
                if hi > lo:
                    s = ovt.library_scores(lib[lo:hi], qv)
                    j = int(np.argmin(s))
                    keys.append(sh.pack_key(int(s[j]), lo + j))
                else:
                    keys.append(-1)
            kt = torch.tensor(keys, dtype=torch.int64)
            sh.reduce_packed_key(kt)
            for k, qv in zip(kt.tolist(), queries):
                full = ovt.library_scores(lib[:n_total], qv)
                assert sh.unpack_key(k) == (int(full.min()), int(np.argmin(full)))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_packed_key_min_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    assert res[0] == res[1]                  # every rank takes the same create-or-match decision
    assert res[0][0] == (5, False)           # the tie went to the lowest global index
    assert res[0][2][1] is True              # a random query creates a template


def test_shard_ranges_partition():
    for n in (0, 1, 7, 4096, 2 ** 20 + 3):
        for world in (1, 2, 3, 8):
            spans = [sh.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_key_roundtrip_and_decide():
    assert sh.unpack_key(sh.pack_key(771, 123456)) == (771, 123456)
    bits = int(np.array([1234.5], dtype=np.float32).view(np.uint32)[0])
    assert sh.unpack_key(sh.pack_key(bits, 9), is_float=True) == (1234.5, 9)
    assert sh.unpack_key(-1) == (None, -1) and sh.unpack_key(sh.KEY_EMPTY) == (None, -1)
    assert sh.decide(None, -1, 0, 45000) == (0, True)
    assert sh.decide(45000, 3, 10, 45000) == (3, False)      # equality is a match (strict '>')
    assert sh.decide(45001, 3, 10, 45000) == (10, True)
