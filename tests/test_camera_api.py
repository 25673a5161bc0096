"""The ICamera point API (generic_camera.h:36-113; eucm.h, ucm.h, mei.h): projectPoint / projectionJacobian /
intrinsicJacobian / reconstructPoint of the reference's three camera classes, recorded from the reference build
(tests/golden/make_golden.py::camera_api -> reference_camera.npz), against
  * the C oracle (CPU, always), and
  * the CUDA entry points vg_project_points / vg_reconstruct_points and the host ICamera mirror on top of them (GPU).
Tolerance 1e-12 of each quantity's scale (the kernels' reciprocal / rsqrt are 1-2 ulp; north_star allows 1e-6)."""
import os
import subprocess

import numpy as np
import pytest

import synthdata as sd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "reference_camera.npz"))
CASES = ["eucm", "eucm_a04", "ucm", "mei"]
TOL = 1e-12


def case(name):
    g = {k.split("/", 1)[1]: GOLD[k] for k in GOLD.files if k.startswith(name + "/")}
    g["model"] = int(g["model"])
    return g


def close(got, want, what):
    """(UCM / MEI back-projection has no validity test, ucm.h:81-103: far outside the image the reference returns
    true with NaN coordinates -- NaN must come out in the same places)"""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert (np.isnan(got) == np.isnan(want)).all(), f"{what}: NaN pattern differs"
    fin = ~np.isnan(want)
    if not fin.any():
        return
    scale = max(1.0, float(np.abs(want[fin]).max()))
    err = np.abs(got[fin] - want[fin]) / scale
    assert err.max() <= TOL, f"{what}: {err.max():.3e}"


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_camera_vectors(oracle, name):
    g = case(name)
    K = len(g["intr"])
    for i, X in enumerate(g["X"]):
        uv, ok = oracle.project(g["model"], g["intr"], X)
        jx, okx = oracle.projection_jacobian(g["model"], g["intr"], X)
        ja, oka = oracle.intrinsic_jacobian(g["model"], g["intr"], X)
        assert (int(ok) | 2 * int(okx) | 4 * int(oka)) == g["flags"][i]
        if ok:
            close(uv, g["uv"][i], "uv")
            close(jx.ravel(), g["dPdX"][i], "dPdX")
            close(ja.ravel(), g["dPdintr"][i], "dPdintr")
    for i, px in enumerate(g["px"]):
        X, ok = oracle.reconstruct(g["model"], g["intr"], px)
        assert int(ok) == g["rec_ok"][i]
        if ok:
            close(X, g["Xrec"][i], "reconstruct")
    assert g["flags"].max() == 7 and K in (5, 6, 10)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_point_api_matches_reference_vectors(gpu, name):
    g = case(name)
    n, K = len(g["X"]), len(g["intr"])
    uv, ok, dx, da = gpu.project_points(g["model"], g["intr"], g["X"], uv_init=np.full((n, 2), -7.0))
    want_ok = (g["flags"] & 1) == 1
    assert (ok == want_ok).all()
    close(uv[ok], g["uv"][ok], "uv")
    assert (uv[~ok] == -7.0).all()                        # a failed point keeps the caller's value (eucm.h:46-54)
    close(dx.reshape(n, 6), g["dPdX"], "dPdX")           # zero rows where the reference returned false
    close(da.reshape(n, 2 * K), g["dPdintr"], "dPdintr")
    X, rok = gpu.reconstruct_points(g["model"], g["intr"], g["px"], X_init=np.full((len(g["px"]), 3), -3.0))
    assert (rok == (g["rec_ok"] == 1)).all()
    close(X[rok], g["Xrec"][rok], "reconstruct")
    assert (X[~rok] == -3.0).all()
    # projection only, no outputs but the mask, empty input
    uv2, ok2, _, _ = gpu.project_points(g["model"], g["intr"], g["X"], want_jacobians=False)
    assert (ok2 == ok).all() and (uv2[ok] == uv[ok]).all()
    assert gpu.project_points(g["model"], g["intr"], np.zeros((0, 3)))[0].shape == (0, 2)


@pytest.mark.gpu
def test_gpu_point_api_large_batch_against_oracle(gpu, oracle):
    """One million points in one launch (grid-stride over 8 CTAs per SM) against the oracle on a sample."""
    n = 1_000_000
    u = sd.uniform(4242, 1, 3 * n).reshape(n, 3)
    X = np.stack([(2 * u[:, 0] - 1) * 1.2, (2 * u[:, 1] - 1) * 0.8, 0.3 + u[:, 2]], axis=1)
    uv, ok, dx, da = gpu.project_points(sd.MEI, sd.MEI_GT, X)
    assert ok.all()
    for i in range(0, n, 9973):
        o_uv, _ = oracle.project(sd.MEI, sd.MEI_GT, X[i])
        o_dx, _ = oracle.projection_jacobian(sd.MEI, sd.MEI_GT, X[i])
        o_da, _ = oracle.intrinsic_jacobian(sd.MEI, sd.MEI_GT, X[i])
        close(uv[i], o_uv, "uv"); close(dx[i].ravel(), o_dx.ravel(), "dPdX"); close(da[i].ravel(), o_da.ravel(), "dPdintr")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["eucm", "mei"])
def test_host_icamera_mirror_projects_through_the_gpu(gpu, name, tmp_path):
    """include/visgeom_b200/camera.hpp: projectPoint, projectionJacobian, intrinsicJacobian (single-point virtuals),
    projectPointCloud and reconstructPointCloud with masks, from a C++ caller, against the reference vectors."""
    g = case(name)
    exe = str(tmp_path / "host_probe")
    lib_dir = os.path.join(ROOT, "visgeom_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "host_probe.cpp"), "-o", exe, "-L" + lib_dir, "-lvisgeom_b200",
                           "-Wl,-rpath," + lib_dir])
    idx = list(range(0, len(g["X"]), 7))
    X, K = g["X"][idx], len(g["intr"])
    line = f"project {g['model']} {K} " + " ".join(repr(float(x)) for x in g["intr"]) + f" {len(X)} " + \
           " ".join(repr(float(x)) for x in X.ravel())
    r = subprocess.run([exe], input=line + "\n", capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    rows = [np.array([float(x) for x in ln.split()]) for ln in r.stdout.strip().splitlines()]
    m = len(X)
    single, all_ok, cloud, back = rows[:m], rows[m], rows[m + 1:2 * m + 1], rows[2 * m + 1:]
    for k, i in enumerate(idx):
        assert int(single[k][0]) == g["flags"][i]
        if g["flags"][i] & 1:
            close(single[k][1:3], g["uv"][i], "uv")
            close(cloud[k][1:3], g["uv"][i], "cloud uv")
        else:
            assert (single[k][1:3] == -7).all()
        close(single[k][3:9], g["dPdX"][i], "dPdX")
        close(single[k][9:9 + 2 * K], g["dPdintr"][i], "dPdintr")
        assert int(cloud[k][0]) == (g["flags"][i] & 1)
    assert int(all_ok[0]) == int(all(g["flags"][i] & 1 for i in idx))
    # back-projection of the projected points is a ray through the original point (EUCM, points in front of the camera)
    if name == "eucm":
        for k, i in enumerate(idx):
            if g["flags"][i] & 1 and int(back[k][0]):
                ray, P = back[k][1:], g["X"][i]
                assert np.abs(np.cross(ray / np.linalg.norm(ray), P / np.linalg.norm(P))).max() < 1e-9
