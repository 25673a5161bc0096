"""Image sharding for the multi-GPU path: images (or stereo pairs) are independent given the
shared parameters, so rank r of G owns the contiguous range [r*N/G, (r+1)*N/G) with its
observations, poses and per-image blocks resident on that GPU; the only exchange per evaluation
is one SUM all-reduce of the reduced normal equations."""
from __future__ import annotations


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    if world < 1 or not (0 <= rank < world) or n_items < 0:
        raise ValueError("bad shard request")
    return (rank * n_items) // world, ((rank + 1) * n_items) // world


def reduced_size(num_shared: int) -> int:
    """doubles exchanged per evaluation: J^T J (Ks x Ks), J^T r (Ks) and the cost"""
    return num_shared * num_shared + num_shared + 1
