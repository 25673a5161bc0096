"""CPU: the host-side mirror of visgeom's Transformation / Quaternion / ICamera (include/visgeom_b200/geometry.hpp,
camera.hpp: what the calibration front end uses once per image for initial poses and the report) against the golden
vectors the REFERENCE build produced (tests/golden/reference_vectors.npz, probe/*) and against the oracle."""
import os
import subprocess

import numpy as np
import pytest

import synthdata as sd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))


@pytest.fixture(scope="module")
def probe(tmp_path_factory, vg):
    exe = str(tmp_path_factory.mktemp("host") / "host_probe")
    lib_dir = os.path.join(ROOT, "visgeom_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "host_probe.cpp"), "-o", exe, "-L" + lib_dir, "-lvisgeom_b200",
                           "-Wl,-rpath," + lib_dir])

    def run(lines):
        r = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, timeout=60)
        assert r.returncode == 0, r.stderr
        return [np.array([float(x) for x in ln.split()]) for ln in r.stdout.strip().splitlines()]
    return run


def fmt(v):
    return " ".join(repr(float(x)) for x in np.ravel(v))


def test_transformation_algebra_matches_reference_vectors(probe):
    ta, tb, comp = GOLD["probe/ta"], GOLD["probe/tb"], GOLD["probe/compose"]
    out = probe([f"compose {k} {fmt(ta[i])} {fmt(tb[i])}" for k in range(3) for i in range(len(ta))])
    got = np.array(out).reshape(3, len(ta), 6)
    assert np.abs(got - comp).max() < 1e-14


def test_rotation_matrix_and_its_inverse_map(probe, oracle):
    v = GOLD["probe/rotvec"]
    out = probe([f"rotmat {fmt(x)}" for x in v])
    for i, o in enumerate(out):
        assert np.abs(o[:9] - GOLD["probe/rotation_matrix"][i]).max() < 1e-15
        th = np.linalg.norm(v[i])
        if 1e-3 < th < 3.0:            # rotationVector(R) goes through sqrt(1 + tr R): singular towards a half turn
            assert np.abs(o[9:] - v[i]).max() < 1e-9
    # the quaternion constructor (7-value transforms of the JSON format, include/json.h:49-53)
    r = np.array([0.3, -0.5, 0.2]); th = np.linalg.norm(r)
    q = np.r_[r / th * np.sin(th / 2), np.cos(th / 2)]
    got = probe([f"quat 1 2 3 {fmt(q)}"])[0]
    assert np.abs(got - np.r_[1, 2, 3, r]).max() < 1e-15


@pytest.mark.parametrize("model,name,intr", [(sd.EUCM, "eucm", sd.EUCM_GT_LEFT), (sd.UCM, "ucm", sd.UCM_GT),
                                             (sd.MEI, "mei", sd.MEI_GT)])
def test_camera_objects_match_reference_vectors(probe, model, name, intr):
    uv, X, ok = GOLD[f"probe/reconstruct_{name}_uv"], GOLD[f"probe/reconstruct_{name}_X"], GOLD[f"probe/reconstruct_{name}_ok"]
    out = probe([f"reconstruct {model} {len(intr)} {fmt(intr)} {fmt(uv[i])}" for i in range(len(uv))])
    for i, o in enumerate(out):
        assert bool(o[0]) == bool(ok[i])
        if ok[i]:
            assert np.abs(o[1:] - X[i]).max() < 1e-14
    b = GOLD[f"probe/bounds_{name}"]
    out = probe([f"bounds {model} {len(intr)} {fmt(intr)} {i}" for i in range(len(intr))])
    assert np.array_equal(np.array(out), b)


# ---- the initialisation half of the front end, on the host alone ("do_not_solve") ------------------------------------
import json
import sys

sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_calib_problem as mk  # noqa: E402


@pytest.fixture(scope="module")
def calib_probe(tmp_path_factory, vg):
    exe = str(tmp_path_factory.mktemp("calib") / "calib_probe")
    lib_dir = os.path.join(ROOT, "visgeom_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(lib_dir, "host"), os.path.join(ROOT, "tests", "calib_probe.cpp"),
                           os.path.join(lib_dir, "host", "calibration.cpp"), "-o", exe, "-L" + lib_dir, "-lvisgeom_b200", "-lz", "-lpthread",
                           "-Wl,-rpath," + lib_dir])

    def run(path):
        r = subprocess.run([exe, path], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        seq, glob, lines, i = {}, {}, r.stdout.splitlines(), 0
        while i < len(lines):
            w = lines[i].split()
            if w and w[0] == "SEQ":
                n = int(w[2])
                seq[w[1]] = np.array([[float(x) for x in lines[i + 1 + k].split()] for k in range(n)])
                i += 1 + n
            elif w and w[0] == "GLOBAL":
                glob[w[1]] = np.array([float(x) for x in lines[i + 1].split()])
                i += 2
            else:
                i += 1
        return seq, glob
    return run


def initial_grid(oracle, model, intr, board, corners, nx=9, ny=6):
    """estimateInitialGrid's closed form (unified_calibration.cpp:1066-1130) restated in numpy on top of the oracle's
    reconstructPoint: four outer corners back-projected, scaled by the board's edge lengths, an orthonormal basis."""
    iUL, iUR, iBL, iBR = 0, nx - 1, nx * (ny - 1), nx * ny - 1
    X = {}
    for k, i in (("UL", iUL), ("UR", iUR), ("BL", iBL), ("BR", iBR)):
        v, ok = oracle.reconstruct(model, intr, corners[i])
        assert ok
        X[k] = v / np.linalg.norm(v)
    n = np.linalg.norm
    sXU = n(board[iUR] - board[iUL]) / n(X["UR"] - X["UL"]); sXB = n(board[iBR] - board[iBL]) / n(X["BR"] - X["BL"])
    sYL = n(board[iBL] - board[iUL]) / n(X["BL"] - X["UL"]); sYR = n(board[iBR] - board[iUR]) / n(X["BR"] - X["UR"])
    pos = X["UL"] * min(sXU, sYL)
    ex = X["UR"] * min(sXU, sYR) - pos
    ey = X["BL"] * min(sXB, sYL) - pos
    ex /= n(ex)
    ey = ey - ex * ex.dot(ey)
    ey /= n(ey)
    R = np.stack([ex, ey, np.cross(ex, ey)], axis=1)
    w = np.sqrt(1 + np.trace(R)) / 2                     # rotationVector(R): quaternion.h:52-59, then :84-98
    q = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (4 * w)
    s = n(q)
    return np.r_[pos, q / s * 2 * np.arctan2(s, w)]


def test_initial_poses_follow_the_reference_closed_form(calib_probe, oracle, tmp_path):
    path, d = mk.write_mono(str(tmp_path / "m"), 8, skip=(2,))
    prob = json.load(open(path)); prob["data"][0]["parameters"] = ["do_not_solve"]; json.dump(prob, open(path, "w"))
    seq, _ = calib_probe(path)
    got = seq["xiCamBoard"]
    assert got.shape == (8, 6)
    for t in range(8):
        if t == 2:
            assert np.array_equal(got[t], [0, 0, 1, 0, 0, 0])           # no board extracted: the default stays (:456)
            continue
        want = initial_grid(oracle, sd.EUCM, d["intr_init"], d["board"], d["obs"][t].reshape(-1, 2))
        assert np.abs(got[t] - want).max() < 1e-12
    # with the true intrinsics the closed form is a starting point near the true poses (it scales the four corner rays by
    # edge-length ratios, which is only approximate for a tilted board -- the solve that follows refines it)
    prob["cameras"][0]["value"] = [float(v) for v in d["intr_gt"]]; json.dump(prob, open(path, "w"))
    got = calib_probe(path)[0]["xiCamBoard"]
    keep = [t for t in range(8) if t != 2]
    assert np.abs(got[keep, 3:] - d["xi_gt"][keep, 3:]).max() < 0.5
    assert np.abs(got[keep, :3] - d["xi_gt"][keep, :3]).max() < 0.25


def test_global_transform_is_unwound_from_the_chain(calib_probe, oracle, tmp_path):
    """Stereo layout without a prior on xiCam12: camera1's dataset initialises the board poses, camera2's dataset
    initialises xiCam12 from its first image through getInitTransform (:311-348): chain [xiCam12 INVERSE, board DIRECT]
    -> xiCam12 = (xi_cam2<-board o xi_cam1<-board^-1)^-1."""
    path, s = mk.write_stereo(str(tmp_path / "s"), 5, prior=False)
    prob = json.load(open(path))
    for ds in prob["data"]:
        ds["parameters"] = ["do_not_solve"]
    json.dump(prob, open(path, "w"))
    seq, glob = calib_probe(path)
    b1 = seq["xiCamBoardStereo"]
    for t in range(5):
        want = initial_grid(oracle, sd.EUCM, s["intr1_init"], s["board"], s["obs1"][t].reshape(-1, 2))
        assert np.abs(b1[t] - want).max() < 1e-12
    g2 = initial_grid(oracle, sd.EUCM, s["intr2_init"], s["board"], s["obs2"][0].reshape(-1, 2))
    x = oracle.compose(g2, b1[0], "compose_inverse")                     # xi o board^-1
    R = oracle.rotation_matrix(-x[3:])
    want = np.r_[-(R @ x[:3]), -x[3:]]                                   # Transformation::inverse, transformation.h:112-119
    assert np.abs(glob["xiCam12"] - want).max() < 1e-12


def test_json_reader_errors_and_value_forms(calib_probe, tmp_path):
    """The Boost.PropertyTree reader the reference uses is replaced by visgeom_b200/host/json.hpp: numbers may come as
    strings (ptree stores everything as text) and with exponents; malformed input, a missing key (ptree's wording) and
    a missing file are reported."""
    path, d = mk.write_mono(str(tmp_path / "j"), 3)
    prob = json.load(open(path))
    prob["data"][0]["parameters"] = ["do_not_solve"]
    json.dump(prob, open(path, "w"))
    base = calib_probe(path)[0]["xiCamBoard"]
    prob["cameras"][0]["value"] = [str(v) for v in prob["cameras"][0]["value"]]       # numbers as strings
    prob["cameras"][0]["value"][2] = "3.0e2"                                             # 300 with an exponent
    p2 = tmp_path / "strings.json"
    p2.write_text(json.dumps(prob))
    assert np.array_equal(calib_probe(str(p2))[0]["xiCamBoard"], base)

    def error_of(text):
        p = tmp_path / "bad.json"
        p.write_text(text)
        with pytest.raises(AssertionError) as e:
            calib_probe(str(p))
        return str(e.value)
    good = json.dumps(prob)
    assert "JSON: " in error_of(good[:len(good) // 2])                                   # truncated
    assert "JSON: " in error_of(good.replace(":", "=", 1))
    del prob["cameras"][0]["type"]
    assert "No such node (type)" in error_of(json.dumps(prob))
    with pytest.raises(AssertionError, match="cannot open file"):
        calib_probe(str(tmp_path / "missing.json"))
