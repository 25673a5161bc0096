"""CPU, world_size 2 over gloo: the host-side logic of the N>1 path -- contiguous image shards,
one SUM all-reduce of the reduced normal equations, identical result on every rank and equal to
the unsharded reduction.  The per-image blocks come from the oracle here; on the
GPU box the same exchange carries the CUDA kernels' blocks (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _reduced_from_blocks(H, K):
    """A (KxK), g (K), cost from packed per-image blocks with column order [intr(K), pose(6), r]."""
    W = K + 7
    iu = np.triu_indices(W)
    full = np.zeros((W, W))
    full[iu] = H.sum(axis=0)
    full = full + np.triu(full, 1).T
    return np.concatenate([full[:K, :K].ravel(), full[:K, W - 1], [0.5 * full[W - 1, W - 1]]])


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import synthdata as sd
    from oracle.pyoracle import Oracle
    from visgeom_b200.sharding import reduced_size, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = sd.make_mono(sd.EUCM, 37, seed=20245)          # 37 does not divide by 2: ragged shards
    lo, hi = shard_range(d["n_img"], rank, world)
    o = Oracle()
    H = o.evaluate_batch(sd.EUCM, d["intr_init"], d["board"], d["obs"][lo:hi], [d["xi_init"][lo:hi]], [0], [0],
                         want_r=False, want_J=False, want_H=True)["H"]
    buf = torch.from_numpy(_reduced_from_blocks(H, 6))
    assert buf.numel() == reduced_size(6)
    dist.all_reduce(buf)                                # the single collective of an evaluation
    np.save(os.path.join(out_dir, f"red{rank}.npy"), buf.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduction_equals_unsharded(tmp_path, oracle):
    import synthdata as sd
    from visgeom_b200.sharding import shard_range
    assert [shard_range(37, r, 2) for r in range(2)] == [(0, 18), (18, 37)]
    assert [shard_range(200000, r, 8)[1] - shard_range(200000, r, 8)[0] for r in range(8)] == [25000] * 8
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "red0.npy"), np.load(tmp_path / "red1.npy")
    assert (r0 == r1).all()                             # every rank holds the same reduced system
    d = sd.make_mono(sd.EUCM, 37, seed=20245)
    H = oracle.evaluate_batch(sd.EUCM, d["intr_init"], d["board"], d["obs"], [d["xi_init"]], [0], [0],
                              want_r=False, want_J=False, want_H=True)["H"]
    full = _reduced_from_blocks(H, 6)
    assert np.abs(r0 - full).max() <= 1e-12 * np.abs(full).max()
