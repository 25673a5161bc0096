"""Two GPUs: the in-kernel exchange of the reduced normal equations over peer memory (vg_peer.cuh) against one GPU
holding all the images -- evaluation (cost, reduced system) and a whole LM solve.  Skipped on a one-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["VG_ROOT"])
import synthdata as sd
import visgeom_b200 as vg

import traceback
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
def main():
  dev = torch.device("cuda", rank)
  torch.cuda.set_device(dev)
  dist.init_process_group("nccl", device_id=dev)
  n = 64
  d = sd.make_mono(sd.EUCM, n, seed=2026)
  lo, hi = rank * n // world, (rank + 1) * n // world

  def build(P, sl):
      cam = P.add_camera(sd.EUCM, d["intr_init"])
      tr = P.add_transform(d["xi_init"][sl], is_global=False)
      P.add_dataset(cam, d["board"], d["obs"][sl], [tr], [0])
      return cam, tr

  S = vg.Problem(rank)
  build(S, slice(lo, hi))
  mine = torch.frombuffer(bytearray(S.peer_export()), dtype=torch.uint8).to(dev)
  every = torch.empty(world * 64, dtype=torch.uint8, device=dev)
  dist.all_gather_into_tensor(every, mine)
  S.peer_connect(rank, world, bytes(every.cpu().numpy().tobytes()))
  A = vg.Problem(rank)                       # all the images on this GPU
  build(A, slice(0, n))
  for rep in range(3):                       # both slot parities, repeated epochs
      cs, rs = S.evaluate(want_reduced=True)
      ca, ra = A.evaluate(want_reduced=True)
      assert abs(cs - ca) <= 1e-12 * ca, (cs, ca)
      assert np.abs(rs - ra).max() <= 1e-11 * np.abs(ra).max()
  # deferred exchange: evaluate_async only posts, the next launch (or the fetch) forms the sum
  for burst in (1, 2, 5):
      for _ in range(burst):
          S.evaluate_async()
      cs, rs = S.fetch_reduced()
      assert abs(cs - ca) <= 1e-12 * ca, (burst, cs, ca)
      assert np.abs(rs - ra).max() <= 1e-11 * np.abs(ra).max()
  S.evaluate_async()                         # posted, then a synchronous evaluation collects it first
  cs, rs = S.evaluate(want_reduced=True)
  assert abs(cs - ca) <= 1e-12 * ca and np.isfinite(rs).all()
  o = S.default_options(); o.max_num_iterations = 12
  ss, sa = S.solve(o), A.solve(o)
  # (the iteration counts may differ by one at the very end: the last trial steps change the cost by rounding noise)
  assert abs(ss.final_cost - sa.final_cost) <= 1e-9 * sa.final_cost
  assert np.abs(S.camera(0) - A.camera(0)).max() <= 1e-8 * np.abs(A.camera(0)).max()
  assert np.abs(S.transform(0) - A.transform(0)[lo:hi]).max() < 1e-8
  # every rank must hold bit-identical shared parameters (the sums are formed in the same order everywhere)
  t = torch.from_numpy(S.camera(0).copy()).to(dev)
  g = [torch.empty_like(t) for _ in range(world)]
  dist.all_gather(g, t)
  assert all(torch.equal(g[0], x) for x in g)
  dist.barrier()
  if rank == 0:
      print("peer exchange ok")
  dist.destroy_process_group()

try:
    main()
except Exception:
    print("WORKER FAILED rank", rank, traceback.format_exc(), flush=True)
    raise
'''


def test_peer_exchange_matches_single_gpu(gpu, tmp_path):
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs on one NVLink domain")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, VG_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "peer exchange ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def _two_local_ranks(gpu, d, n):
    """Two problems of this process on device 0, connected as ranks 0 / 1 of one exchange group through the addresses
    of their inboxes (vg_problem_peer_connect_local): the protocol itself needs no second GPU."""
    import synthdata as sd
    probs = []
    for r in range(2):
        lo, hi = r * n // 2, (r + 1) * n // 2
        P = gpu.Problem(0)
        cam = P.add_camera(sd.EUCM, d["intr_init"])
        tr = P.add_transform(d["xi_init"][lo:hi], is_global=False)
        P.add_dataset(cam, d["board"], d["obs"][lo:hi], [tr], [0])
        probs.append(P)
    inboxes = [P.peer_inbox() for P in probs]
    for r, P in enumerate(probs):
        P.peer_connect_local(r, inboxes)
    return probs


def test_local_ranks_deferred_exchange_on_one_gpu(gpu):
    """Deferred mode (the kernel's tail posts, the next launch or the fetch collects) with both ranks driven from one
    host thread on one GPU: sums equal to the single problem holding all the images, bit-identical on both ranks."""
    import numpy as np
    import synthdata as sd
    n = 90
    d = sd.make_mono(sd.EUCM, n, seed=2027)
    A = gpu.Problem(0)
    cam = A.add_camera(sd.EUCM, d["intr_init"])
    tr = A.add_transform(d["xi_init"], is_global=False)
    A.add_dataset(cam, d["board"], d["obs"], [tr], [0])
    ca, ra = A.evaluate(want_reduced=True)
    P0, P1 = _two_local_ranks(gpu, d, n)
    for burst in (1, 3, 2):                       # slot parities reused, collection by the next launch's head
        for _ in range(burst):
            P0.evaluate_async(); P1.evaluate_async()
        (c0, r0), (c1, r1) = P0.fetch_reduced(), P1.fetch_reduced()
        assert abs(c0 - ca) <= 1e-12 * ca and np.abs(r0 - ra).max() <= 1e-11 * np.abs(ra).max()
        assert c0 == c1 and (r0 == r1).all()


def test_missing_rank_is_reported_not_nan(gpu):
    """A rank that never posts: the collect gives up after the configured number of polls and the next call that
    synchronises returns VG_ERR_PEER (-6) with a message, instead of handing NaN sums to the LM loop."""
    import synthdata as sd
    n = 40
    d = sd.make_mono(sd.EUCM, n, seed=2028)
    P0, P1 = _two_local_ranks(gpu, d, n)
    P0.set_peer_timeout(20000)                    # a few milliseconds of polling
    P0.evaluate_async()                           # rank 1 never evaluates
    with pytest.raises(gpu.VisgeomError, match=r"error -6: peer exchange 1 timed out on rank 0"):
        P0.fetch_reduced()
    # the exchange group is out of step from here on, as after any lost rank; a fresh pair works
    Q0, Q1 = _two_local_ranks(gpu, d, n)
    Q0.evaluate_async(); Q1.evaluate_async()
    assert Q0.fetch_reduced()[0] == Q1.fetch_reduced()[0]
