"""Shared helpers of the parity tests."""
import numpy as np

# north_star asks for <= 1e-6 relative; the kernels are held to 1e-9 of each block's scale
RTOL = 1e-9


def assert_close(got, want, name, rtol=RTOL):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, f"{name}: shape {got.shape} vs {want.shape}"
    scale = max(1.0, float(np.abs(want).max()) if want.size else 1.0)
    err = float(np.abs(got - want).max()) if want.size else 0.0
    assert np.isfinite(got).all(), f"{name}: non-finite values"
    assert err <= rtol * scale, f"{name}: max abs err {err:.3e} > {rtol:.1e} * scale {scale:.3e}"
    return err / scale


def compare_eval(gpu_out, ref_out, tag, rtol=RTOL):
    worst = 0.0
    worst = max(worst, assert_close(gpu_out["r"], ref_out["r"], f"{tag}: residual", rtol))
    if ref_out.get("J_intr") is not None:
        worst = max(worst, assert_close(gpu_out["J_intr"], ref_out["J_intr"], f"{tag}: J_intr", rtol))
        for e, (a, b) in enumerate(zip(gpu_out["J_xi"], ref_out["J_xi"])):
            worst = max(worst, assert_close(a, b, f"{tag}: J_xi[{e}]", rtol))
    if ref_out.get("H") is not None:
        worst = max(worst, assert_close_hessian(gpu_out["H"], ref_out["H"], f"{tag}: H", rtol))
    return worst


def unpack_hessian(H):
    """(n, W(W+1)/2) packed upper triangles -> (n, W, W) symmetric matrices."""
    H = np.asarray(H)
    ne = H.shape[1]
    W = int(round((np.sqrt(8 * ne + 1) - 1) / 2))
    assert W * (W + 1) // 2 == ne
    iu = np.triu_indices(W)
    full = np.zeros((H.shape[0], W, W))
    full[:, iu[0], iu[1]] = H
    full[:, iu[1], iu[0]] = H
    return full


def assert_close_hessian(got, want, name, rtol=RTOL):
    """Entries of J^T J span many decades: entry (i,j) is held to rtol * sqrt(H_ii H_jj)."""
    g, w = unpack_hessian(got), unpack_hessian(want)
    assert np.isfinite(g).all(), f"{name}: non-finite values"
    d = np.sqrt(np.maximum(np.einsum("nii->ni", w), 1e-300))
    scale = d[:, :, None] * d[:, None, :]
    rel = np.abs(g - w) / np.maximum(scale, 1e-300)
    worst = float(rel.max()) if rel.size else 0.0
    assert worst <= rtol, f"{name}: max scaled err {worst:.3e} > {rtol:.1e}"
    return worst


def local_maxima_np(resp, avg, R):
    """numpy statement of selectCandidates' scan (corner_detector.cpp:494-534) -> (values (n,), uv (n, 2)) in scan order."""
    h, w = resp.shape
    if h <= 2 * R or w <= 2 * R:
        return np.zeros(0, np.float32), np.zeros((0, 2), np.int32)
    c = resp[R:h - R, R:w - R]
    ok = np.ones(c.shape, bool)
    if np.isfinite(avg):
        ok &= ~(c.astype(np.float64) < avg)
    for j in range(-R, R + 1):
        for i in range(-R, R + 1):
            if (i == 0 and j == 0) or i * i + j * j > R * R + 1:
                continue
            nb = resp[R + j:h - R + j, R + i:w - R + i]
            later = i > 0 or (i == 0 and j > 0)
            ok &= ~((c < nb) | ((c == nb) & (not later)))
    vs, us = np.nonzero(ok)
    return c[vs, us].astype(np.float32), np.stack([us + R, vs + R], 1).astype(np.int32)
