"""The calibration front end (vg_calib: JSON problem files -> GenericCameraCalibration mirror -> C ABI), the
reference's `calib` tool (test/calibration/generic_calibration.cpp, unified_calibration.cpp:91-356,632-660).

CPU part: the parsing / configuration errors the reference raises before any solve, and the loud failure
without a GPU.  GPU part: whole problems against the oracle's LM solution of the same data."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synthdata as sd  # noqa: E402
import make_calib_problem as mk  # noqa: E402
from visgeom_b200 import build as vg_build  # noqa: E402


@pytest.fixture(scope="module")
def cli():
    vg_build.build()
    return vg_build.build_cli()


def run(cli, *args, cwd=None):
    return subprocess.run([cli, *args], capture_output=True, text=True, cwd=cwd, timeout=600)


def edit(path, fn):
    prob = json.load(open(path))
    fn(prob)
    json.dump(prob, open(path, "w"))


def test_config_errors_match_the_reference(cli, tmp_path):
    path, _ = mk.write_mono(str(tmp_path / "a"), 4)
    edit(path, lambda p: p["cameras"][0].update(type="pinhole"))
    r = run(cli, path)
    assert r.returncode == 1 and "invalid camera model name" in r.stderr                       # :177
    path, _ = mk.write_mono(str(tmp_path / "b"), 4)
    edit(path, lambda p: p["cameras"][0].update(value=[0.5, 1, 300, 300, 600]))
    assert "invalid number of intrinsic parameters" in run(cli, path).stderr                   # :151
    path, _ = mk.write_mono(str(tmp_path / "c"), 4)
    edit(path, lambda p: p["transformations"][0].update(constant=True))
    assert "is constant but there is no prior" in run(cli, path).stderr                        # :103
    path, _ = mk.write_mono(str(tmp_path / "d"), 4)
    edit(path, lambda p: p["transformations"].append({"name": "x2", "global": False, "constant": False, "prior": False})
         or p["data"][0]["transform_chain"].append({"name": "x2", "direct": True}))
    assert "not one sequences in a transform chain" in run(cli, path).stderr                   # :227
    path, _ = mk.write_mono(str(tmp_path / "e"), 4)
    edit(path, lambda p: p["data"][0].update(init="nope"))
    assert "does not exist, impossible to initialize" in run(cli, path).stderr                 # :433
    path, _ = mk.write_mono(str(tmp_path / "f"), 4)
    edit(path, lambda p: p["data"][0].update(type="images"))                                   # an "images" dataset needs
    r = run(cli, path)                                                                         # object.cols / rows / size
    assert r.returncode == 1 and "object.cols" in r.stderr
    assert "cannot open file" in run(cli, str(tmp_path / "missing.json")).stderr
    assert run(cli).returncode == 2
    # the odometry / prior datasets (unified_calibration.cpp:742-829)
    path, _ = mk.write_odometry(str(tmp_path / "g"), 5)
    edit(path, lambda p: p["data"][0].update(transform="xiBaseCam"))
    assert "is global. Odometry must be a sequence" in run(cli, path).stderr                   # :752-755
    path, _ = mk.write_odometry(str(tmp_path / "h"), 5)
    edit(path, lambda p: p["data"][2].update(transform="xiOdom"))
    assert "must have a prior value" in run(cli, path).stderr                                  # :816-819
    path, _ = mk.write_odometry(str(tmp_path / "i"), 5)
    edit(path, lambda p: p["data"].insert(1, dict(p["data"][0])))
    assert "has already been initialized" in run(cli, path).stderr                             # :781-784
    path, _ = mk.write_odometry(str(tmp_path / "j"), 5)
    edit(path, lambda p: p["data"][0].update(transform="nope"))
    assert "has not been declared" in run(cli, path).stderr                                    # :747-750
    path, _ = mk.write_odometry(str(tmp_path / "k"), 5)
    edit(path, lambda p: p["data"][0].update(type="odometry_intrinsic"))
    assert "is not supported by this engine" in run(cli, path).stderr


def test_fails_loudly_without_a_gpu(cli, tmp_path):
    import visgeom_b200 as vg
    if vg.device_count() > 0:
        pytest.skip("a CUDA device is present")
    path, _ = mk.write_mono(str(tmp_path), 4)
    r = run(cli, path)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


def parse_report(out):
    intr, glob = {}, {}
    sect = None
    for line in out.splitlines():
        if line.startswith("Intrinsic parameters"): sect = "i"; continue
        if line.startswith("Local extrinsic"): sect = "l"; continue
        if line.startswith("Global extrinsic"): sect = "g"; continue
        m = re.match(r"^(\w+) : (.*)$", line)
        if m and sect == "i": intr[m.group(1)] = np.array([float(x) for x in m.group(2).split()])
        if m and sect == "g": glob[m.group(1)] = np.array([float(x) for x in m.group(2).split()])
    return intr, glob


@pytest.mark.gpu
def test_mono_problem_matches_oracle(cli, tmp_path, oracle):
    from oracle.pyoracle import OracleProblem
    n = 24
    path, d = mk.write_mono(str(tmp_path), n, skip=(5,))
    r = run(cli, "--precision", "17", "--out", str(tmp_path) + "/", path)
    assert r.returncode == 0, r.stderr
    intr, _ = parse_report(r.stdout)
    keep = [i for i in range(n) if i != 5]
    O = OracleProblem(oracle)
    cam = O.add_camera(sd.EUCM, d["intr_init"])
    tr = O.add_transform(d["xi_init"][keep], is_global=False)
    O.add_dataset(cam, d["board"], d["obs"][keep], [tr], [0])
    O.solve()
    rel = np.abs(intr["camera1"] - O.camera(cam)) / np.abs(O.camera(cam))
    assert rel.max() < 1e-6, rel          # north_star: final intrinsics <= 1e-6 relative
    # image_error_0.txt: err_u err_v proj_u proj_v tx..rz per corner of every image with a board
    rows = np.loadtxt(str(tmp_path / "image_error_0.txt"))
    assert rows.shape == (len(keep) * d["P"], 10)
    obs = d["obs"][keep].reshape(-1, 2)
    assert np.abs(rows[:, 0:2] + rows[:, 2:4] - obs).max() < 1e-9           # err = detected - projected
    assert np.sqrt((rows[:, 0:2] ** 2).mean()) < 0.2                          # 0.1 px noise
    assert np.abs(rows[:d["P"], 4:10] - O.transform(tr)[0]).max() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("prior", [True, False])
def test_stereo_problem_matches_oracle(cli, tmp_path, oracle, prior):
    from oracle.pyoracle import OracleProblem
    path, s = mk.write_stereo(str(tmp_path), 40, prior=prior)
    r = run(cli, "--precision", "17", "--out", str(tmp_path) + "/", path)
    assert r.returncode == 0, r.stderr
    intr, glob = parse_report(r.stdout)
    O = OracleProblem(oracle)
    c1 = O.add_camera(sd.EUCM, s["intr1_init"]); c2 = O.add_camera(sd.EUCM, s["intr2_init"])
    tb = O.add_transform(s["xi_init"], is_global=False)
    t12 = O.add_transform(s["xi12_init"], is_global=True)
    O.add_dataset(c1, s["board"], s["obs1"], [tb], [0])
    O.add_dataset(c2, s["board"], s["obs2"], [t12, tb], [1, 0])
    O.solve()
    for name, cid in (("camera1", c1), ("camera2", c2)):
        rel = np.abs(intr[name] - O.camera(cid)) / np.abs(O.camera(cid))
        assert rel.max() < 1e-6, (name, rel)
    assert np.abs(glob["xiCam12"] - O.transform(t12)[0]).max() < 1e-6
    assert os.path.exists(str(tmp_path / "image_error_1.txt"))


@pytest.mark.gpu
def test_odometry_problem_matches_oracle(cli, tmp_path, oracle):
    """"odometry" (init + anchor) and "transformation_prior" datasets next to a reprojection dataset."""
    from oracle.pyoracle import OracleProblem
    path, d = mk.write_odometry(str(tmp_path), 30)
    r = run(cli, "--precision", "17", "--out", str(tmp_path) + "/", path)
    assert r.returncode == 0, r.stderr
    intr, glob = parse_report(r.stdout)
    O = OracleProblem(oracle)
    cam = O.add_camera(sd.EUCM, d["intr_init"])
    bc = O.add_transform(d["xi_bc_init"], is_global=True)
    od = O.add_transform(d["odom"], is_global=False)
    wB = O.add_transform(d["xi_wB_init"], is_global=True)
    O.add_dataset(cam, d["board"], d["obs"], [bc, od, wB], d["status"])
    O.add_odometry(od, d["err_v"], d["err_w"], d["lam"], d["odom"])
    O.set_pose_constant(od, 0)
    O.add_transformation_prior(bc, [10, 10, 10, 20, 20, 20])
    so = O.solve()
    m = re.search(r"final cost\s+(\S+)", r.stdout)
    assert abs(float(m.group(1)) - so.final_cost) <= 1e-6 * so.final_cost
    rel = np.abs(intr["camera1"] - O.camera(cam)) / np.abs(O.camera(cam))
    assert rel.max() < 1e-6, rel
    assert np.abs(glob["xiBaseCam"] - O.transform(bc)[0]).max() < 1e-6
    assert np.abs(glob["xiWorldBoard"] - O.transform(wB)[0]).max() < 1e-6


@pytest.mark.gpu
def test_show_outliers_report(cli, tmp_path):
    """"show_outliers" (unified_calibration.cpp:1217-1272): the textual part of the report -- a sample is listed when
    one of its corners is off by 3.6 sigma or a pixel."""
    path, d = mk.write_mono(str(tmp_path), 12)
    corners = json.load(open(str(tmp_path / "corners.json")))
    corners[4][0]["points"][10][0] += 6.0            # one gross outlier in image 4
    json.dump(corners, open(str(tmp_path / "corners.json"), "w"))
    edit(path, lambda p: p["data"][0].update(parameters=["show_outliers"]))
    r = run(cli, "--out", str(tmp_path) + "/", path)
    assert r.returncode == 0, r.stderr
    assert "Sample #4" in r.stdout and "standard deviation : " in r.stdout
    m = re.search(r"err : (\S+)", r.stdout)
    assert m and float(m.group(1)) > 3.0


@pytest.mark.gpu
def test_images_dataset_detects_and_calibrates(cli, tmp_path):
    """Dataset type "images" (unified_calibration.cpp:279-309, 632-647, 992-1062): the front end reads the pictures
    (PGM / PNG), runs the checkerboard detector on the GPU, and calibrates from the corners it found.  Rendered boards:
    the calibration must come back to the intrinsics the pictures were rendered with, from a start 5 % off."""
    n = 14
    path, info = mk.write_images(str(tmp_path), n)
    r = run(cli, "--precision", "17", "--out", str(tmp_path) + "/", path)
    assert r.returncode == 0, r.stderr
    assert "missing.png : ERROR, file not found" in r.stdout                               # :1027-1030
    assert "img_003.png : ERROR, pattern not found" in r.stdout                            # :1034-1038
    m = re.search(r"DETECTION RATE : (\d+) of (\d+) detected", r.stdout)
    assert m and int(m.group(2)) == n + 1 and int(m.group(1)) >= n - 3
    intr, _ = parse_report(r.stdout)
    rel = np.abs(intr["camera1"] - info["intr"]) / np.abs(info["intr"])
    assert rel[2:].max() < 0.02 and rel.max() < 0.25, rel       # focal lengths, centre to 2 %; alpha / beta are weakly observed
    rows = np.loadtxt(str(tmp_path / "image_error_0.txt"))
    assert rows.shape[0] == int(m.group(1)) * 54
    assert np.sqrt((rows[:, 0:2] ** 2).mean()) < 0.25           # sub-pixel corners: residual RMS well under a pixel
    # the corners the calibration used are the detector's: compare with the Python binding on the same pictures
    import visgeom_b200 as vg
    img0, _ = sd.render_board_image(640, 480, seed=20400, model=sd.EUCM, supersample=2)
    found, c = vg.detect_pattern(img0, improve=True)
    assert found and np.abs(rows[:54, 0:2] + rows[:54, 2:4] - c).max() < 1e-9
