"""Host-side sharding helpers (no CUDA needed): contiguous ranges and the packed-key MIN reduction.

The view-template library shards by contiguous template index ranges and needs exactly one exchange per
query: a MIN all-reduce of ``(score << 32) | global_index`` -- minimum score, lowest index among equals,
i.e. ``numpy.argmin`` over the concatenated library (``ratslam/view_templates.py:73``).  Pose-cell
ensembles shard by network with no exchange at all.
"""
from __future__ import annotations

import numpy as np
import torch

KEY_EMPTY = (1 << 63) - 1  # what an empty shard contributes (INT64_MAX; never wins a MIN)


def shard_range(n, rank, world):
    """Contiguous ``[lo, hi)`` of ``n`` units owned by ``rank``; sizes differ by at most one."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_key(score_bits, index):
    """``score_bits``: the integer score (uint8 mode) or the IEEE-754 bits of a non-negative float32 score."""
    return (int(score_bits) << 32) | int(index)


def unpack_key(key, is_float=False):
    """``(score, index)``; ``(None, -1)`` for the empty key."""
    key = int(key) & ((1 << 64) - 1)
    if key in (KEY_EMPTY, (1 << 64) - 1):
        return None, -1
    hi, idx = key >> 32, key & 0xFFFFFFFF
    if is_float:
        return float(np.array([hi], dtype=np.uint32).view(np.float32)[0]), int(idx)
    return int(hi), int(idx)


def reduce_packed_key(key, group=None):
    """In-place MIN all-reduce of an int64 key tensor (any device) over ``group``.

    A raw UINT64_MAX from the kernel reads as -1 in int64; it is mapped to ``KEY_EMPTY`` first so
    that an empty shard cannot win.
    """
    import torch.distributed as dist
    key.copy_(torch.where(key < 0, torch.full_like(key, KEY_EMPTY), key))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(key, op=dist.ReduceOp.MIN, group=group)
    return key


def decide(score, index, n_total, threshold):
    """The create-or-match rule of ``ViewTemplates.match`` (``view_templates.py:67``): returns
    ``(template_index, created)``.  Strict ``>``: a score equal to the threshold is a match."""
    if score is None or score > threshold:
        return int(n_total), True
    return int(index), False
