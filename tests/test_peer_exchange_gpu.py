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
