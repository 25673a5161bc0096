"""Writes synthetic calibration problems in the reference's JSON format (README.md:36-223 of visgeom,
dataset type "ir_data": pre-extracted corners, unified_calibration.cpp:234-277) for vg_calib / the tests.

  python tools/make_calib_problem.py mono|stereo|odometry|images OUT_DIR [n_images]
"""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synthdata as sd  # noqa: E402


def _object(board, nx=9, ny=6):
    return {"corner_ul": 0, "corner_ur": nx - 1, "corner_bl": nx * (ny - 1), "corner_br": nx * ny - 1,
            "points": [[float(v) for v in p] for p in board]}


def _dataset(camera, chain, init, data_file, board):
    return {"type": "ir_data", "camera": camera,
            "transform_chain": [{"name": n, "direct": bool(d)} for n, d in chain], "init": init,
            "parameters": [], "image_width": sd.IMAGE_W, "image_height": sd.IMAGE_H, "data_file": data_file,
            "object": _object(board)}


def _points(obs_row):
    return [[float(obs_row[2 * i]), float(obs_row[2 * i + 1])] for i in range(len(obs_row) // 2)]


def write_mono(out_dir, n_img=20, model=sd.EUCM, seed=20241, skip=()):
    """One camera, one per-image board pose initialised from the corners (init = the sequence).  Images in
    `skip` carry no entry for the camera (no board extracted)."""
    os.makedirs(out_dir, exist_ok=True)
    d = sd.make_mono(model, n_img, seed=seed)
    data_file = os.path.join(out_dir, "corners.json")
    json.dump([[] if i in skip else [{"camera": "camera1", "points": _points(d["obs"][i])}] for i in range(n_img)],
              open(data_file, "w"))
    prob = {"transformations": [{"name": "xiCamBoard", "global": False, "constant": False, "prior": False}],
            "cameras": [{"name": "camera1", "type": sd.MODEL_NAMES[model], "constant": False,
                         "value": [float(v) for v in d["intr_init"]]}],
            "data": [_dataset("camera1", [("xiCamBoard", True)], "xiCamBoard", data_file, d["board"])]}
    path = os.path.join(out_dir, "problem.json")
    json.dump(prob, open(path, "w"), indent=1)
    return path, d


def write_images(out_dir, n_img=10, model=sd.EUCM, seed=20400, width=640, height=480, improve=True):
    """Dataset type "images" (unified_calibration.cpp:279-309, 632-647): rendered pictures of the 9 x 6 board (PGM and PNG
    files alternating, one JPEG where cv2 can write it, one missing file, one picture without a board), the camera started off its true intrinsics."""
    import numpy as np
    os.makedirs(out_dir, exist_ok=True)
    truth, names = [], []
    for i in range(n_img):
        img, uv = sd.render_board_image(width, height, seed=seed + i, model=model, supersample=2)
        if i == 3:
            img = np.full_like(img, 120)                        # no board on this one
        name = "img_%03d.%s" % (i, "png" if i % 2 else "pgm")
        jpeg = None
        if i == 7:                                              # one JPEG among them, where OpenCV is there to write it
            try:
                import cv2
                jpeg = "img_%03d.jpg" % i
                if not cv2.imwrite(os.path.join(out_dir, jpeg), img, [cv2.IMWRITE_JPEG_QUALITY, 95]):
                    jpeg = None
            except ImportError:
                jpeg = None
        if jpeg:
            name = jpeg
        elif i % 2:
            sd.write_png(os.path.join(out_dir, name), img, filter_type=(i // 2) % 3)
        else:
            sd.write_pgm(os.path.join(out_dir, name), img)
        names.append(name); truth.append(uv)
    names.insert(5, "missing.png")                              # a file that does not exist: reported, skipped
    truth.insert(5, None)
    intr = sd.board_image_intrinsics(width, height, model)
    init = intr.copy()
    init[-4:-2] *= 1.05; init[-2] += 6.0; init[-1] -= 5.0
    prob = {"transformations": [{"name": "xiCamBoard", "global": False, "constant": False, "prior": False}],
            "cameras": [{"name": "camera1", "type": sd.MODEL_NAMES[model], "constant": False, "value": [float(v) for v in init]}],
            "data": [{"type": "images", "camera": "camera1", "transform_chain": [{"name": "xiCamBoard", "direct": True}],
                      "init": "xiCamBoard", "parameters": ["improve_detection"] if improve else [],
                      "object": {"type": "checkboard", "cols": 9, "rows": 6, "size": 0.1},
                      "images": {"prefix": out_dir.rstrip("/") + "/", "names": names}}]}
    path = os.path.join(out_dir, "problem.json")
    json.dump(prob, open(path, "w"), indent=1)
    return path, dict(intr=intr, intr_init=init, truth=truth, names=names)


def write_stereo(out_dir, n_pairs=20, seed=20244, prior=True):
    """The layout of data/calib_stereo_example.json: camera1 sees the board through xiCamBoardStereo, camera2
    through xiCam12^-1 o xiCamBoardStereo; xiCam12 is global, with a prior or (prior=False) initialised from the
    second dataset (unified_calibration.cpp:497-511: first extracted image, then a solve over the dataset)."""
    os.makedirs(out_dir, exist_ok=True)
    s = sd.make_stereo(n_pairs, seed=seed)
    data_file = os.path.join(out_dir, "corners.json")
    json.dump([[{"camera": "camera1", "points": _points(s["obs1"][i])},
                {"camera": "camera2", "points": _points(s["obs2"][i])}] for i in range(n_pairs)], open(data_file, "w"))
    prob = {"transformations": [{"name": "xiCamBoardStereo", "global": False, "constant": False, "prior": False},
                                ({"name": "xiCam12", "global": True, "constant": False, "prior": True,
                                  "value": [float(v) for v in s["xi12_init"]]} if prior else
                                 {"name": "xiCam12", "global": True, "constant": False, "prior": False})],
            "cameras": [{"name": "camera1", "type": "eucm", "constant": False, "value": [float(v) for v in s["intr1_init"]]},
                        {"name": "camera2", "type": "eucm", "constant": False, "value": [float(v) for v in s["intr2_init"]]}],
            "data": [_dataset("camera1", [("xiCamBoardStereo", True)], "xiCamBoardStereo", data_file, s["board"]),
                     _dataset("camera2", [("xiCam12", False), ("xiCamBoardStereo", True)], "none" if prior else "xiCam12",
                              data_file, s["board"])]}
    path = os.path.join(out_dir, "problem.json")
    json.dump(prob, open(path, "w"), indent=1)
    return path, s


def write_odometry(out_dir, n=30, seed=20246, anchor=True, prior=True):
    """A camera on a wheeled base looking at one board (synthdata.make_odometry): the sequence xiOdom takes its
    initial values from an "odometry" dataset (unified_calibration.cpp:742-807), the camera extrinsic carries a
    "transformation_prior" (:808-829); the reprojection dataset's chain is
    [xiBaseCam inverse, xiOdom inverse, xiWorldBoard direct]."""
    os.makedirs(out_dir, exist_ok=True)
    d = sd.make_odometry(n, seed=seed)
    data_file = os.path.join(out_dir, "corners.json")
    json.dump([[{"camera": "camera1", "points": _points(d["obs"][i])}] for i in range(n)], open(data_file, "w"))
    fl = lambda v: [float(x) for x in v]
    data = [{"type": "odometry", "transform": "xiOdom", "err_v": d["err_v"], "err_w": d["err_w"], "lambda": d["lam"],
             "init": True, "anchor": bool(anchor), "value": [fl(x) for x in d["odom"]]},
            _dataset("camera1", [("xiBaseCam", False), ("xiOdom", False), ("xiWorldBoard", True)], "none", data_file, d["board"])]
    if prior:
        data.append({"type": "transformation_prior", "transform": "xiBaseCam", "stiffness": [10, 10, 10, 20, 20, 20]})
    prob = {"transformations": [{"name": "xiOdom", "global": False, "constant": False, "prior": False},
                                {"name": "xiBaseCam", "global": True, "constant": False, "prior": True, "value": fl(d["xi_bc_init"])},
                                {"name": "xiWorldBoard", "global": True, "constant": False, "prior": True, "value": fl(d["xi_wB_init"])}],
            "cameras": [{"name": "camera1", "type": "eucm", "constant": False, "value": fl(d["intr_init"])}],
            "data": data}
    path = os.path.join(out_dir, "problem.json")
    json.dump(prob, open(path, "w"), indent=1)
    return path, d


if __name__ == "__main__":
    kind, out = sys.argv[1], sys.argv[2]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    print({"mono": write_mono, "stereo": write_stereo, "odometry": write_odometry, "images": write_images}[kind](out, n)[0])
