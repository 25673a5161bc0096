"""Timings of the rows around the hot path on the GPU, each next to the CPU restatement of
the reference's code (oracle/, all host threads):
  * ICamera point API: vg_project_points_dev (projectPoint + both Jacobians per point), device-resident, CUDA events,
    against the HBM peak (24 B in, 16 + 48 + 16 K + 1 B out per point);
  * per-image initialisation solves: vg_refine_poses (host arrays in, poses out);
  * TransformationPrior / OdometryPrior / OdometryCost functors: vg_eval_* (host arrays in and out);
  * TrajectoryVisualQuality::visualCov: vg_visual_cov (host arrays in and out).
    python tools/aux_timing.py"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synthdata as sd  # noqa: E402
import visgeom_b200 as vg  # noqa: E402
from oracle.pyoracle import Oracle, OracleProblem, _dp, _f64  # noqa: E402


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6458.4


def best(fn, reps=3):
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); t.append(time.perf_counter() - t0)
    return min(t)


def main():
    orc = Oracle()
    cores = orc.max_threads()
    dev = torch.device("cuda:0")
    L = vg.lib()
    peak = peak_gbs()
    print(f"# tools/aux_timing.py  (HBM peak {peak:.0f} GB/s; CPU = oracle restatement on {cores} threads)")
    # ---- ICamera point API -------------------------------------------------------------------------------------------
    for model, name in ((sd.EUCM, "EUCM"), (sd.MEI, "MEI")):
        intr = np.array({sd.EUCM: sd.EUCM_GT_LEFT, sd.MEI: sd.MEI_GT}[model], dtype=np.float64)
        K = len(intr)
        n = 8_000_000
        rng = np.random.default_rng(1)
        X = np.concatenate([rng.uniform(-1, 1, (n, 2)), rng.uniform(0.3, 2.0, (n, 1))], axis=1)
        tX = torch.from_numpy(X).to(dev); ti = torch.from_numpy(intr).to(dev)
        uv = torch.empty(n, 2, dtype=torch.float64, device=dev); dx = torch.empty(n, 6, dtype=torch.float64, device=dev)
        da = torch.empty(n, 2 * K, dtype=torch.float64, device=dev); ok = torch.empty(n, dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream().cuda_stream

        def run():
            assert L.vg_project_points_dev(model, ti.data_ptr(), n, tX.data_ptr(), uv.data_ptr(), dx.data_ptr(), da.data_ptr(),
                                           ok.data_ptr(), st) == 0
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 10
        by = n * (24 + 16 + 48 + 16 * K + 1)
        m = 400_000
        o_uv = np.zeros((m, 2)); o_dx = np.zeros((m, 6)); o_da = np.zeros((m, 2 * K)); o_ok = np.zeros(m, np.uint8)
        orc.lib.vgo_project_batch.argtypes = [C.c_int, C.c_void_p, C.c_long] + [C.c_void_p] * 5 + [C.c_int]
        Xm = np.ascontiguousarray(X[:m])
        cpu = best(lambda: orc.lib.vgo_project_batch(model, intr.ctypes.data, m, Xm.ctypes.data, o_uv.ctypes.data, o_dx.ctypes.data,
                                                     o_da.ctypes.data, o_ok.ctypes.data, cores))
        g = uv[:m].cpu().numpy()
        assert np.abs(g - o_uv).max() < 1e-9
        print(f"project_points {name}: {n} points in {us:.0f} us -> {n / us:.0f} M points/s, {by / us / 1e3:.0f} GB/s algorithmic = "
              f"{by / us / 1e3 / peak:.3f} of the HBM peak; CPU {m / cpu / 1e6:.1f} M points/s -> x{n / us / (m / cpu / 1e6):.0f}")
    # ---- per-image initialisation solves -------------------------------------------------------------------------------
    d = sd.make_mono(sd.EUCM, 10000, seed=20242)
    vg.refine_poses(sd.EUCM, d["intr_gt"], d["board"], d["obs"][:64], d["xi_init"][:64], 25.0)
    t = best(lambda: vg.refine_poses(sd.EUCM, d["intr_gt"], d["board"], d["obs"], d["xi_init"], 25.0))
    x, it, cost, term = vg.refine_poses(sd.EUCM, d["intr_gt"], d["board"], d["obs"], d["xi_init"], 25.0)
    m = 200
    t0 = time.perf_counter()
    worst = 0.0
    for i in range(m):
        O = OracleProblem(orc)
        cam = O.add_camera(sd.EUCM, d["intr_gt"], constant=True)
        tr = O.add_transform(d["xi_init"][i:i + 1], is_global=False)
        ds = O.add_dataset(cam, d["board"], d["obs"][i:i + 1], [tr], [0])
        O.set_loss(ds, 25.0)
        O.solve()
        worst = max(worst, np.abs(x[i] - O.transform(tr)[0]).max())
    cpu = (time.perf_counter() - t0) / m
    assert worst < 1e-7, worst
    print(f"refine_poses: 10000 images x 54 corners (host arrays in, poses out) in {t * 1e3:.2f} ms -> {10000 / t:.0f} images/s, "
          f"mean {it.mean():.1f} LM iterations; CPU (oracle LM, one problem per image, 1 thread) {1 / cpu:.0f} images/s "
          f"-> x{10000 / t * cpu:.0f} per thread, x{10000 / t * cpu / cores:.0f} against {cores} threads")
    # ---- prior functors --------------------------------------------------------------------------------------------------
    n = 200_000
    rng = np.random.default_rng(2)
    stiff = rng.uniform(1, 50, (n, 6)); xp = rng.normal(0, 0.5, (n, 6)); xi = xp + rng.normal(0, 0.05, (n, 6))
    vg.eval_transformation_prior(stiff[:64], xp[:64], xi[:64])
    t = best(lambda: vg.eval_transformation_prior(stiff, xp, xi))
    r = np.zeros((n, 6)); J = np.zeros((n, 36))
    orc.lib.vgo_transformation_prior_batch.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_int]
    cpu = best(lambda: orc.lib.vgo_transformation_prior_batch(n, stiff.ctypes.data, xp.ctypes.data, xi.ctypes.data, r.ctypes.data,
                                                              J.ctypes.data, 1, cores), 2)
    print(f"TransformationPrior: {n} blocks (host in / out) in {t * 1e3:.2f} ms -> {n / t / 1e6:.1f} M blocks/s; CPU (incl. construction) "
          f"{n / cpu / 1e6:.2f} M blocks/s")
    o1 = rng.normal(0, 0.5, (n, 6)); inc = rng.normal(0, 0.05, (n, 6))
    o2 = o1 + inc; x1 = o1 + rng.normal(0, 0.01, (n, 6)); x2 = o2 + rng.normal(0, 0.01, (n, 6))
    vg.eval_odometry_prior(0.1, 0.05, 0.02, o1[:64], o2[:64], x1[:64], x2[:64])
    t = best(lambda: vg.eval_odometry_prior(0.1, 0.05, 0.02, o1, o2, x1, x2))
    J1 = np.zeros((n, 36)); J2 = np.zeros((n, 36))
    orc.lib.vgo_odometry_prior_batch.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double] + [C.c_void_p] * 7 + [C.c_int, C.c_int]
    cpu = best(lambda: orc.lib.vgo_odometry_prior_batch(n, 0.1, 0.05, 0.02, o1.ctypes.data, o2.ctypes.data, x1.ctypes.data, x2.ctypes.data,
                                                        r.ctypes.data, J1.ctypes.data, J2.ctypes.data, 1, cores), 2)
    print(f"OdometryPrior: {n} blocks (host in / out) in {t * 1e3:.2f} ms -> {n / t / 1e6:.1f} M blocks/s; CPU (incl. construction) "
          f"{n / cpu / 1e6:.2f} M blocks/s")
    nb, mlen = 50_000, 40
    blocks = [rng.uniform(0.05, 0.5, (mlen, 2)) for _ in range(nb)]
    ip = np.array([0.1, 0.1, 0.5]); itr = ip * 1.02
    vg.eval_odometry_cost(0.1, 0.05, 0.003, blocks[:64], ip, x1[:64], x2[:64], itr)
    t = best(lambda: vg.eval_odometry_cost(0.1, 0.05, 0.003, blocks, ip, x1[:nb], x2[:nb], itr), 2)
    off = (np.arange(nb + 1) * mlen).astype(np.int32); dq = np.ascontiguousarray(np.concatenate(blocks))
    J3 = np.zeros((nb, 18)); xa = np.ascontiguousarray(x1[:nb]); xb = np.ascontiguousarray(x2[:nb])
    orc.lib.vgo_odometry_cost_batch.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double] + [C.c_void_p] * 10 + [C.c_int, C.c_int]
    cpu = best(lambda: orc.lib.vgo_odometry_cost_batch(nb, 0.1, 0.05, 0.003, off.ctypes.data, dq.ctypes.data, ip.ctypes.data, xa.ctypes.data,
                                                       xb.ctypes.data, itr.ctypes.data, r.ctypes.data, J1.ctypes.data, J2.ctypes.data,
                                                       J3.ctypes.data, 1, cores), 2)
    print(f"OdometryCost: {nb} blocks x {mlen} increments (host in / out, Python packing of the increment lists included) in "
          f"{t * 1e3:.2f} ms -> {nb / t / 1e6:.2f} M blocks/s; CPU (incl. construction) {nb / cpu / 1e6:.3f} M blocks/s")
    # ---- visualCov -------------------------------------------------------------------------------------------------------
    n = 100_000
    board = sd.make_board(9, 6, 0.1)
    xi_board = np.array([0.0, 0.0, 1.0, 0.05, -0.03, 0.02])
    poses = rng.normal(0, 0.05, (n, 6))
    intr = np.array(sd.EUCM_GT_LEFT, dtype=np.float64)
    vg.visual_cov(sd.EUCM, intr, xi_board, board, 0.25, poses[:64])
    t = best(lambda: vg.visual_cov(sd.EUCM, intr, xi_board, board, 0.25, poses))
    m = 5000
    cpu = best(lambda: orc.visual_cov(sd.EUCM, intr, xi_board, board, 0.25, poses[:m]), 2)
    print(f"visualCov: {n} poses x 54 board points (host in / out) in {t * 1e3:.2f} ms -> {n / t / 1e3:.0f} k poses/s; CPU (1 thread) "
          f"{m / cpu / 1e3:.1f} k poses/s -> x{n / t / (m / cpu):.0f} per thread")


if __name__ == "__main__":
    main()
