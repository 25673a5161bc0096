"""The front end's image reader (visgeom_b200/host/image_io.hpp; the reference calls cv::imread(name, 0),
unified_calibration.cpp:1025): binary PGM and 8-bit PNG, against the arrays the files were written from and, where cv2
is importable, against files written by OpenCV itself (libpng picks among all five line filters)."""
import os
import subprocess

import numpy as np
import pytest

import synthdata as sd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def imread(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("img") / "image_probe")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(ROOT, "visgeom_b200", "host"), os.path.join(ROOT, "tests", "image_probe.cpp"),
                           "-o", exe, "-lz"])

    def run(path):
        out = subprocess.run([exe, path], capture_output=True, timeout=60, check=True).stdout
        head, _, body = out.partition(b"\n")
        w, h = (int(x) for x in head.split())
        return None if w == 0 else np.frombuffer(body, dtype=np.uint8).reshape(h, w)
    return run


def test_pgm_and_png_round_trip(imread, tmp_path):
    img, _ = sd.render_board_image(173, 131, seed=20250)
    sd.write_pgm(str(tmp_path / "a.pgm"), img)
    assert np.array_equal(imread(str(tmp_path / "a.pgm")), img)
    for f in (0, 1, 2):
        sd.write_png(str(tmp_path / f"f{f}.png"), img, filter_type=f)
        assert np.array_equal(imread(str(tmp_path / f"f{f}.png")), img), f
    rgb = np.random.default_rng(3).integers(0, 256, (37, 45, 3), dtype=np.uint8)
    sd.write_png(str(tmp_path / "rgb.png"), rgb, filter_type=1)
    r64 = rgb.astype(np.int64)
    grey = ((r64[..., 0] * 4899 + r64[..., 1] * 9617 + r64[..., 2] * 1868 + 8192) >> 14).astype(np.uint8)
    assert np.array_equal(imread(str(tmp_path / "rgb.png")), grey)


def test_files_written_by_opencv(imread, tmp_path):
    cv2 = pytest.importorskip("cv2")
    img, _ = sd.render_board_image(320, 240, seed=20251, model=sd.MEI)
    noise = np.random.default_rng(5).integers(0, 256, (61, 83), dtype=np.uint8)
    for name, a in (("board", img), ("noise", noise)):
        for ext in ("png", "pgm"):
            path = str(tmp_path / f"{name}.{ext}")
            assert cv2.imwrite(path, a)
            assert np.array_equal(imread(path), a), path
            assert np.array_equal(cv2.imread(path, 0), a)
    # a 16-bit picture: imread(name, 0) keeps the high byte
    deep = (img.astype(np.uint16) << 8) | np.random.default_rng(6).integers(0, 256, img.shape, dtype=np.uint16)
    path = str(tmp_path / "deep.png")
    assert cv2.imwrite(path, deep)
    assert np.array_equal(imread(path), cv2.imread(path, 0)) and np.array_equal(imread(path), img)
    # and OpenCV reads what the test data writer wrote
    sd.write_png(str(tmp_path / "w.png"), img, filter_type=2)
    assert np.array_equal(cv2.imread(str(tmp_path / "w.png"), 0), img)


def test_unreadable_files_give_an_empty_image(imread, tmp_path):
    assert imread(str(tmp_path / "missing.png")) is None                # cv::imread returns an empty Mat
    (tmp_path / "junk.png").write_bytes(b"not an image at all, just bytes")
    assert imread(str(tmp_path / "junk.png")) is None
    img, _ = sd.render_board_image(64, 48, seed=1)
    sd.write_png(str(tmp_path / "t.png"), img)
    data = (tmp_path / "t.png").read_bytes()
    (tmp_path / "cut.png").write_bytes(data[: len(data) // 2])
    assert imread(str(tmp_path / "cut.png")) is None
    (tmp_path / "cut.pgm").write_bytes(b"P5\n64 48\n255\n" + bytes(100))
    assert imread(str(tmp_path / "cut.pgm")) is None


def test_palette_png(imread, tmp_path):
    """colour type 3: indices into PLTE, converted like a colour picture"""
    import struct
    import zlib
    rng = np.random.default_rng(8)
    pal = rng.integers(0, 256, (16, 3), dtype=np.uint8)
    idx = rng.integers(0, 16, (23, 31), dtype=np.uint8)
    raw = b"".join(b"\x00" + idx[y].tobytes() for y in range(idx.shape[0]))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    path = str(tmp_path / "pal.png")
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 31, 23, 8, 3, 0, 0, 0)) + chunk(b"PLTE", pal.tobytes()) +
                chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))
    p64 = pal.astype(np.int64)
    grey = ((p64[:, 0] * 4899 + p64[:, 1] * 9617 + p64[:, 2] * 1868 + 8192) >> 14).astype(np.uint8)
    got = imread(path)
    assert np.array_equal(got, grey[idx])
    try:
        import cv2
    except ImportError:
        return
    ref = cv2.imread(path, 0)
    assert ref is not None and np.abs(got.astype(int) - ref.astype(int)).max() <= 1      # libpng's own rgb -> grey rounding


def test_jpeg_luminance_matches_opencv(imread, tmp_path):
    """baseline JPEG read as grey: only Y is reconstructed, with libjpeg's integer inverse DCT -- the pixels OpenCV returns.
    Grey and colour files, 4:2:0 / 4:2:2 / 4:4:4 sampling, sizes that are not multiples of the MCU, restart intervals."""
    cv2 = pytest.importorskip("cv2")
    board, _ = sd.render_board_image(317, 243, seed=20252, model=sd.UCM)
    rng = np.random.default_rng(9)
    colour = np.stack([board, np.roll(board, 7, axis=1), rng.integers(0, 256, board.shape, dtype=np.uint8)], axis=-1)
    noise = rng.integers(0, 256, (67, 129), dtype=np.uint8)
    cases = []
    for q in (50, 92, 100):
        cases.append((board, [cv2.IMWRITE_JPEG_QUALITY, q]))
    cases.append((noise, [cv2.IMWRITE_JPEG_QUALITY, 75]))
    cases.append((board, [cv2.IMWRITE_JPEG_QUALITY, 85, cv2.IMWRITE_JPEG_RST_INTERVAL, 5]))
    cases.append((colour, [cv2.IMWRITE_JPEG_QUALITY, 90]))
    cases.append((colour[:131, :200], [cv2.IMWRITE_JPEG_QUALITY, 80, cv2.IMWRITE_JPEG_RST_INTERVAL, 3]))
    for flag in ("IMWRITE_JPEG_SAMPLING_FACTOR_444", "IMWRITE_JPEG_SAMPLING_FACTOR_422", "IMWRITE_JPEG_SAMPLING_FACTOR_420"):
        if hasattr(cv2, flag) and hasattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR"):
            cases.append((colour, [cv2.IMWRITE_JPEG_QUALITY, 88, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, getattr(cv2, flag)]))
    for i, (a, params) in enumerate(cases):
        path = str(tmp_path / f"c{i}.jpg")
        assert cv2.imwrite(path, a, params)
        want = cv2.imread(path, 0)
        got = imread(path)
        assert got is not None and got.shape == want.shape, (i, params)
        assert np.array_equal(got, want), (i, params, int(np.abs(got.astype(int) - want.astype(int)).max()))
    # progressive files are not decoded: an empty image, not garbage
    path = str(tmp_path / "prog.jpg")
    assert cv2.imwrite(path, board, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    assert imread(path) is None
    data = open(str(tmp_path / "c0.jpg"), "rb").read()
    (tmp_path / "cut.jpg").write_bytes(data[: len(data) // 2])
    assert imread(str(tmp_path / "cut.jpg")) is None
