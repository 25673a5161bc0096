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
