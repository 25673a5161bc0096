"""detector_oracle.py -- CPU restatement (pure Python + numpy) of visgeom's checkerboard detector after the response
stage.  TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing else); the product never touches it.

    CornerDetector::detectPattern            src/calibration/corner_detector.cpp:223-260
    CornerDetector::selectCandidates         :494-609
    CornerDetector::checkCorner              :331-440
    CornerDetector::scaleInvarient           :444-492
    CornerDetector::constructGraph           :612-836
    extractSequence / selectBestOrthogonalChain / selectPattern / verifyDetection   :850-1076
    getCircle / getSamples / centralDifferences / getTransitions / initPoin        :1079-1298
    improveCorners                           :162-198  (+ ceres::GradientProblemSolver, third party, restated)
    SubpixelCorner::Evaluate                 :47-100   (+ ceres::BiCubicInterpolator, third party: Catmull-Rom)
    setZero, comp, normalizePoint            include/calibration/corner_detector.h:144-178
    CurveRasterizer, Polynomial2::Circle     include/utils/curve_rasterizer.h:31-66,168-259
    bilinear                                 include/ocv.h:68-90

The response stage itself (computeResponse :262-329, with OpenCV's fixed-point GaussianBlur) is restated in C
(oracle/corner_oracle.c) and supplies the maps this file works on.  Line-faithful on purpose: plain loops, the
reference's own order of operations and tie-breaking -- including the standard library's heap algorithms
(std::make_heap / std::pop_heap of libstdc++, bits/stl_heap.h), on whose order among equal keys the numbering of the
candidates depends.  Slow (seconds per small image): for fixtures, not for throughput.

Pinned on the reference's own corner_detector.cpp compiled where it lies (oracle/_ref, recorded in
tests/golden/detector.npz): candidates in graph order, grid, initPoin values, SubpixelCorner's cost / gradient, the
refined corners and the minimiser's iteration counts are reproduced exactly (tests/test_detector_oracle.py)."""
from __future__ import annotations

import math

import numpy as np

DOUBLE_MAX = float(np.finfo(np.float64).max)


def sign(x) -> int:
    """include/std.h:74-77: sign(0) = -1"""
    return 2 * int(x > 0) - 1


def c_round(x: float) -> float:
    """C's round(): halves away from zero"""
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


# ---- libstdc++ heap algorithms (bits/stl_heap.h), comparator a.first < b.first ------------------------------------------
def _push_heap(a, hole, top, value):
    parent = (hole - 1) // 2
    while hole > top and a[parent][0] < value[0]:
        a[hole] = a[parent]
        hole = parent
        parent = (hole - 1) // 2
    a[hole] = value


def _adjust_heap(a, hole, length, value):
    top = hole
    child = hole
    while child < (length - 1) // 2:
        child = 2 * (child + 1)
        if a[child][0] < a[child - 1][0]:
            child -= 1
        a[hole] = a[child]
        hole = child
    if (length & 1) == 0 and child == (length - 2) // 2:
        child = 2 * (child + 1)
        a[hole] = a[child - 1]
        hole = child - 1
    _push_heap(a, hole, top, value)


def make_heap(a):
    n = len(a)
    if n < 2:
        return
    parent = (n - 2) // 2
    while True:
        _adjust_heap(a, parent, n, a[parent])
        if parent == 0:
            return
        parent -= 1


def pop_heap(a):
    """moves the largest element to a[-1] (the caller pops it)"""
    n = len(a)
    if n > 1:
        value = a[n - 1]
        a[n - 1] = a[0]
        _adjust_heap(a, 0, n - 1, value)


# ---- curve rasteriser on a circle (curve_rasterizer.h) -------------------------------------------------------------------
class _Circle:
    def __init__(self, u0, v0, r):
        self.kuu, self.kvv, self.kuv = 1.0, 1.0, 0.0
        self.ku, self.kv = -2.0 * u0, -2.0 * v0
        self.k1 = float(u0 * u0 + v0 * v0 - r * r)

    def __call__(self, u, v):
        return (self.kuu * u + self.kuv * v + self.ku) * u + (self.kvv * v + self.kv) * v + self.k1

    def gradu(self, u, v):
        return 2 * self.kuu * u + self.kuv * v + self.ku

    def gradv(self, u, v):
        return self.kuv * u + 2 * self.kvv * v + self.kv


class _Raster:
    def __init__(self, u, v, eu, ev, surf):
        self.u, self.v, self.surf = u, v, surf
        self.fu, self.fv, self.delta = surf.gradu(u, v), surf.gradv(u, v), surf(u, v)
        self.eps = 1 if self.fu * (ev - v) - self.fv * (eu - u) > 0 else -1

    def move_u(self, du):
        if du == 0:
            return
        self.u += du
        fu2 = self.surf.gradu(self.u, self.v)
        self.delta += 0.5 * du * (self.fu + fu2)
        self.fu = fu2
        self.fv = self.surf.gradv(self.u, self.v)

    def move_v(self, dv):
        if dv == 0:
            return
        self.v += dv
        fv2 = self.surf.gradv(self.u, self.v)
        self.delta += 0.5 * dv * (self.fv + fv2)
        self.fv = fv2
        self.fu = self.surf.gradu(self.u, self.v)

    def step(self):
        if abs(self.fu) > abs(self.fv):
            self.move_v(self.eps * sign(self.fu))
            self.move_u(int(-c_round(self.delta / self.fu)))
        else:
            self.move_u(-self.eps * sign(self.fv))
            self.move_v(int(-c_round(self.delta / self.fv)))


class DetectorOracle:
    """One image.  img uint8 (H, W); maps = what computeResponse leaves behind: resp, gradx, grady, imgrad (float32),
    s2 = _src2 (uint8), avg = _avgVal."""

    def __init__(self, img, maps, s2, nx, ny, init_radius):
        self.img = np.asarray(img)
        self.resp, self.gradx, self.grady, self.imgrad = (np.asarray(maps[k]) for k in ("resp", "gradx", "grady", "imgrad"))
        self.avg = float(maps["avg"])
        self.s2 = np.asarray(s2)
        self.rows, self.cols = self.img.shape
        self.nx, self.ny, self.R = nx, ny, init_radius

    # ---- small helpers -----------------------------------------------------------------------------------------------
    def normalize_point(self, u, v):
        return (max(0, min(u, self.cols - 1)), max(0, min(v, self.rows - 1)))

    def get_circle(self, pt, radius):                     # :1079-1108
        if radius == 1:
            du = (1, 1, 0, -1, -1, -1, 0, 1)
            dv = (0, 1, 1, 1, 0, -1, -1, -1)
            return [self.normalize_point(pt[0] + du[i], pt[1] + dv[i]) for i in range(8)]
        res = []
        pt0 = (pt[0] + radius, pt[1])
        raster = _Raster(pt[0] + radius, pt[1], pt[0], pt[1] + radius, _Circle(pt[0], pt[1], radius))
        i = 0
        while True:
            res.append(self.normalize_point(raster.u, raster.v))
            if i > 5 and abs(raster.u - pt0[0]) <= 1 and abs(raster.v - pt0[1]) <= 1:
                break
            i += 1
            raster.step()
        return res

    def get_samples(self, pts):                           # :1110-1119
        return [float(self.img[v, u]) for (u, v) in pts]

    @staticmethod
    def central_differences(s):                           # :1121-1133
        n = len(s)
        res = [s[1] - s[-1]]
        for i in range(1, n - 1):
            res.append(s[i + 1] - s[i - 1])
        res.append(s[0] - s[n - 2])
        return res

    @staticmethod
    def set_zero(t, k):                                   # corner_detector.h:149-166
        n, ref = len(t), t[k]
        i = k
        while True:
            t[i] = 0.0
            i += 1
            if i == n:
                i = 0
            if not t[i] * ref > 0:
                break
        i = k
        while True:
            t[i] = 0.0
            if i == 0:
                i = n
            i -= 1
            if not t[i] * ref > 0:
                break

    @staticmethod
    def _argmax(t, a=0, b=None):                          # std::max_element: the first largest; an empty range returns a
        b = len(t) if b is None else b
        k = a
        for i in range(a + 1, b):
            if t[k] < t[i]:
                k = i
        return k

    @staticmethod
    def _argmin(t, a=0, b=None):
        b = len(t) if b is None else b
        k = a
        for i in range(a + 1, b):
            if t[i] < t[k]:
                k = i
        return k

    # ---- candidate tests ---------------------------------------------------------------------------------------------
    def check_corner(self, pt, check_radius):             # :331-440 (MAX_FAULTS = 0: the first fault ends it)
        A1, A2 = 0.3, 0.5
        for radius in range(check_radius, check_radius + max(3, check_radius)):
            t = self.central_differences(self.get_samples(self.get_circle(pt, radius)))
            n = len(t)
            thresh = n // 2 - 2
            thresh2 = n - thresh
            i1 = self._argmax(t); max1 = t[i1]
            self.set_zero(t, i1)
            i2 = self._argmax(t)
            if t[i2] < max1 * A1:
                return False
            max2 = t[i2]
            self.set_zero(t, i2)
            if abs(i1 - i2) < thresh or abs(i1 - i2) > thresh2:
                return False
            i3 = self._argmax(t)
            if t[i3] > max2 * A2:
                return False
            self.set_zero(t, i3)
            j1 = self._argmin(t); min1 = t[j1]
            if min1 > max1 * -A1:
                return False
            self.set_zero(t, j1)
            j2 = self._argmin(t)
            if t[j2] > min1 * -A1:
                return False
            min2 = t[j2]
            self.set_zero(t, j2)
            if abs(j1 - j2) < thresh or abs(j1 - j2) > thresh2:
                return False
            j3 = self._argmin(t)
            if t[j3] < min2 * A2:
                return False
        return True

    def scale_invariant(self, pt):                        # :444-492
        for radius in range(self.R, self.R * 2):
            acc, norm_acc = 0.0, 1e-10
            for dv in range(-radius, radius + 1):
                for du in range(-radius, radius + 1):
                    sq = float(du * du + dv * dv)
                    if sq > radius * radius + 1 or sq < 1:
                        continue
                    u, v = pt[0] + du, pt[1] + dv
                    if u < 0 or u >= self.cols or v < 0 or v >= self.rows:
                        continue
                    gx, gy = float(self.gradx[v, u]), float(self.grady[v, u])
                    g2 = gx * gx + gy * gy
                    if g2 < 1e-3:
                        continue
                    p = gx * du + gy * dv
                    acc += p * p / sq
                    norm_acc += g2
            if acc / norm_acc < 0.3:
                return True
        return False

    def select_candidates(self):                          # :494-609 -> the candidates in the order constructGraph pops them
        W = self.R
        heap = []
        for v in range(W, self.rows - W):
            for u in range(W, self.cols - W):
                val = float(self.resp[v, u])
                if val < self.avg:
                    continue
                is_max = True
                for j in range(-W, W + 1):
                    for i in range(-W, W + 1):
                        if i == 0 and j == 0:
                            continue
                        if i * i + j * j > W * W + 1:
                            continue
                        nb = float(self.resp[v + j, u + i])
                        if val <= nb:
                            if val == nb and (i > 0 or (i == 0 and j > 0)):
                                continue
                            is_max = False
                            break
                    if not is_max:
                        break
                if is_max:
                    heap.append((val, (u, v)))
        make_heap(heap)
        thresh_heap = list(heap)
        acc = 0.0
        n_ref = self.nx * self.ny
        for _ in range(n_ref):
            if not thresh_heap:
                break
            pop_heap(thresh_heap)
            acc += thresh_heap.pop()[0]
        val_thresh = 0.05 * acc / n_ref
        hyp = []
        i = 0
        while heap and heap[0][0] > val_thresh and i < 10 * self.nx * self.ny:
            pop_heap(heap)
            pt = heap.pop()[1]
            checked = False
            for radius in range(1, self.R + 1):
                if checked:
                    break
                checked = self.check_corner(pt, radius)
            if not checked:
                continue
            if not self.scale_invariant(pt):
                continue
            hyp.append((float(-pt[0] - pt[1]), pt))
            i += 1
        make_heap(hyp)
        out = []
        while hyp:
            pop_heap(hyp)
            out.append(hyp.pop()[1])
        return out

    # ---- transitions ---------------------------------------------------------------------------------------------------
    def get_transitions(self, pt):                        # :1135-1259
        res = []
        detected = False
        best = 0.0
        for radius in range(1, self.R + 2):
            circle = self.get_circle(pt, radius)
            t = self.central_differences(self.get_samples(circle))
            n = len(t)

            def opposite(k, arg):
                l1 = k + n // 4
                l2 = l1 + n // 2
                if l1 < n <= l2:
                    l2 %= n
                    a, b = arg(t, l1, n), arg(t, 0, l2)
                    if arg is self._argmax:
                        return a if t[a] > t[b] else b
                    return a if t[a] < t[b] else b
                return arg(t, l1 % n, l2 % n)
            hi1 = self._argmax(t); hi2 = opposite(hi1, self._argmax)
            lo1 = self._argmin(t); lo2 = opposite(lo1, self._argmin)
            if detected:
                if best > 0.7 * t[hi1]:
                    break
                detected = False
                res = []
            if t[hi2] < 0.4 * t[hi1]:
                continue
            if t[lo2] > 0.4 * t[lo1]:
                continue
            res = [circle[hi1], circle[hi2], circle[lo1], circle[lo2]]
            best = t[hi1]
            detected = True
        return res

    def init_point(self, pt):                             # :1261-1298
        A, Cc, B, D = self.get_transitions(pt)
        m00, m01 = float(A[1] - Cc[1]), float(Cc[0] - A[0])
        m10, m11 = float(B[1] - D[1]), float(D[0] - B[0])
        b0 = A[1] * m01 + A[0] * m00
        b1 = B[1] * m11 + B[0] * m10
        inv_det = 1.0 / (m00 * m11 - m01 * m10)            # Eigen's 2 x 2 inverse: adjugate * (1 / det)
        i00, i01, i10, i11 = m11 * inv_det, -m01 * inv_det, -m10 * inv_det, m00 * inv_det
        return [i00 * b0 + i01 * b1, i10 * b0 + i11 * b1, math.atan2(A[1] - Cc[1], A[0] - Cc[0]),
                math.atan2(B[1] - D[1], B[0] - D[0]), 0.0]

    # ---- graph -------------------------------------------------------------------------------------------------------------
    def bilinear_s2(self, x, y):                          # include/ocv.h:68-90 with T = double on _src2
        u, v = int(x), int(y)                             # truncation towards zero
        dx, dy = x - u, y - v
        dx2 = 1 - dx
        fail = False
        if u < 0:
            fail = True; u = 0
        elif u > self.cols - 2:
            fail = True; u = self.cols - 1
        # `if (fail |= v < 0) v = 0;`: the assignment fires whenever fail is already set
        if fail or v < 0:
            fail = True; v = 0
        elif v > self.rows - 2:
            fail = True; v = self.rows - 1
        if fail:
            return float(self.s2[v, u])
        i00, i01 = float(self.s2[v, u]), float(self.s2[v, u + 1])
        i10, i11 = float(self.s2[v + 1, u]), float(self.s2[v + 1, u + 1])
        return (i11 * dx + i10 * dx2) * dy + (i01 * dx + i00 * dx2) * (1 - dy)

    def construct_graph(self, cand):                      # :612-836 -> (arcs, arc_sign)
        from collections import deque
        idx_map = np.full((self.rows, self.cols), -1, dtype=np.int16)
        fringe = deque()
        n = len(cand)
        arcs = [[] for _ in range(n)]
        arc_sign = {}
        grad_thresh = []
        for i, pt in enumerate(cand):
            grad_thresh.append(DOUBLE_MAX)
            for q in self.get_transitions(pt):
                fringe.append((0, i, q[0], q[1]))
                grad_thresh[-1] = min(float(self.imgrad[q[1], q[0]]) / 2, grad_thresh[-1])
        du8 = (-1, 0, 1, 1, 1, 0, -1, -1)
        dv8 = (-1, -1, -1, 0, 1, 1, 1, 0)
        while fringe and fringe[0][0] < 140:
            t, idx, eu, ev = fringe.popleft()
            if idx_map[ev, eu] != -1:
                continue
            idx_map[ev, eu] = idx
            connected = False
            for k in range(8):
                u2, v2 = eu + du8[k], ev + dv8[k]
                if u2 < 0 or u2 >= self.cols or v2 < 0 or v2 >= self.rows:
                    continue
                idx2 = int(idx_map[v2, u2])
                if idx2 != -1 and idx2 != idx:
                    connected = True
                    if idx2 not in arcs[idx]:
                        ax, ay = float(cand[idx2][0] - cand[idx][0]), float(cand[idx2][1] - cand[idx][1])
                        norm = math.sqrt(ax * ax + ay * ay)
                        nx_, ny_ = ax / norm, ay / norm
                        acc = 0
                        for base in range(1, self.R + 1):
                            sx, sy = nx_ * base, ny_ * base
                            for lam in range(1, 5):
                                f = float(lam) / 5
                                mx, my = cand[idx][0] + ax * f, cand[idx][1] + ay * f
                                acc += sign(self.bilinear_s2(mx - sy, my + sx) - self.bilinear_s2(mx + sy, my - sx))
                        acc = sign(acc)
                        arcs[idx].append(idx2)
                        arcs[idx2].append(idx)
                        arc_sign[(idx, idx2)] = acc
                        arc_sign[(idx2, idx)] = -acc
            if not connected:
                for k in range(8):
                    u2, v2 = eu + du8[k], ev + dv8[k]
                    if u2 < 0 or u2 >= self.cols or v2 < 0 or v2 >= self.rows:
                        continue
                    if idx_map[v2, u2] != -1:
                        continue
                    x1, y1 = float(u2 - cand[idx][0]), float(v2 - cand[idx][1])
                    tt = math.sqrt(x1 * x1 + y1 * y1)
                    xg, yg = float(self.gradx[v2, u2]), float(self.grady[v2, u2])
                    grad_proj = abs(xg * y1 - yg * x1)
                    if tt > 0 and grad_proj / tt < grad_thresh[idx]:
                        continue
                    if tt > 0:
                        c = x1 * xg + y1 * yg
                        s = x1 * yg - y1 * xg
                        angle = abs(math.atan2(s, c))
                        thresh = min(math.pi / 5, tt / 150.0 + 1 / tt)
                        if angle < math.pi / 2 - thresh or angle > math.pi / 2 + thresh:
                            continue
                    fringe.append((t + 1, idx, u2, v2))
        return arcs, arc_sign

    # ---- pattern -------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _inorm(x, y):                                     # Eigen's Matrix<int, 2, 1>::norm(): sqrt converted back to int
        return int(math.sqrt(float(x * x + y * y)))

    def compare_vectors(self, v1, v2):                    # :838-848
        return self._inorm(v1[0] - v2[0], v1[1] - v2[1]) / float(self._inorm(v1[0], v1[1]))

    def extract_sequence(self, cand, arcs, sg, idx0, idx1):       # :850-922
        def d(a, b):
            return (cand[b][0] - cand[a][0], cand[b][1] - cand[a][1])
        d1 = d(idx0, idx1)
        idx2, best = -1, 1.0
        for n2 in arcs[idx1]:
            diff = self.compare_vectors(d1, d(idx1, n2))
            if sg.get((idx1, n2), 0) == sg.get((idx0, idx1), 0):
                continue
            if diff < best:
                best, idx2 = diff, n2
        if idx2 == -1:
            return []
        chain = [idx0, idx1, idx2]
        while True:
            i1, i0 = chain[-1], chain[-2]
            d0 = d(i0, i1)
            nxt, best = -1, 1.0
            for n2 in arcs[i1]:
                if n2 == i0:
                    continue
                if sg.get((i1, n2), 0) == sg.get((i0, i1), 0):
                    continue
                diff = self.compare_vectors(d0, d(i1, n2))
                if diff < best:
                    best, nxt = diff, n2
            if nxt == -1:
                break
            chain.append(nxt)
        return chain

    def best_orthogonal_chain(self, cand, arcs, sg, idx0, idx1, eps, length):     # :930-972
        best_cost, best_chain = 0.3, []
        base = sg.get((idx0, idx1), 0)
        for nx_ in arcs[idx0]:
            if nx_ == idx1:
                continue
            if sg.get((idx0, nx_), 0) == base:
                continue
            d1 = (cand[idx1][0] - cand[idx0][0], cand[idx1][1] - cand[idx0][1])
            d2 = (cand[nx_][0] - cand[idx0][0], cand[nx_][1] - cand[idx0][1])
            chain = self.extract_sequence(cand, arcs, sg, idx0, nx_)
            cost = eps * (d1[0] * d2[1] - d1[1] * d2[0]) / float(self._inorm(*d1) * self._inorm(*d2))
            if cost < best_cost:
                continue
            if len(chain) >= length:
                best_cost, best_chain = cost, chain
        return best_chain[:length]

    def verify(self, cand, arcs, sg, idx):                # :1060-1076
        if len(idx) != self.nx * self.ny:
            return False
        for i in range(1, self.nx):
            if idx[self.nx + i] not in arcs[idx[i]]:
                return False
            chain = self.extract_sequence(cand, arcs, sg, idx[i], idx[self.nx + i])
            if len(chain) < self.ny:
                return False
            for j in range(2, self.ny):
                if chain[j] != idx[j * self.nx + i]:
                    return False
        return True

    def select_pattern(self, cand, arcs, sg):             # :975-1058
        for idx0 in range(len(cand)):
            if len(arcs[idx0]) < 2:
                continue
            chain_y, chain_x = [], []
            for n in arcs[idx0]:
                chain = self.extract_sequence(cand, arcs, sg, idx0, n)
                if len(chain) >= self.ny:
                    chain = chain[:self.ny]
                    chain_x = self.best_orthogonal_chain(cand, arcs, sg, idx0, n, -1, self.nx)
                    if len(chain_x) == self.nx:
                        chain_y = chain
                        break
            if not chain_y or not chain_x:
                continue
            res = list(chain_x)
            for i in range(1, len(chain_y)):
                row = self.best_orthogonal_chain(cand, arcs, sg, chain_y[i], chain_y[i - 1], 1, self.nx)
                if len(row) >= self.nx:
                    res += row[:self.nx]
                else:
                    break
            if len(res) == self.nx * self.ny and self.verify(cand, arcs, sg, res):
                return res
        return []

    def detect(self):
        """one scale of detectPattern (:231-245): (candidates in graph order, pattern indices or [])"""
        cand = self.select_candidates()
        if len(cand) < self.nx * self.ny:
            return cand, []
        arcs, sg = self.construct_graph(cand)
        return cand, self.select_pattern(cand, arcs, sg)


def refinement_reach(grid, nx):
    """improveCorners' radMax (:164-175); grid (n, 2)"""
    g = np.asarray(grid, dtype=np.float64)
    out = []
    for i in range(len(g)):
        r = 7.0
        r = min(r, float(np.linalg.norm(g[i] - g[i - nx if i > nx else i + nx])) * 0.7)
        r = min(r, float(np.linalg.norm(g[i] - g[i - 1 if i > 0 else i + 1])) * 0.7)
        out.append(r)
    return np.array(out)


# ---- SubpixelCorner::Evaluate (:47-100) -------------------------------------------------------------------------------------
def _cubic(p0, p1, p2, p3, x):
    """ceres/cubic_interpolation.h CubicHermiteSpline (Catmull-Rom): value and derivative"""
    a = 0.5 * (-p0 + 3.0 * p1 - 3.0 * p2 + p3)
    b = 0.5 * (2.0 * p0 - 5.0 * p1 + 4.0 * p2 - p3)
    c = 0.5 * (-p0 + p2)
    return p1 + x * (c + x * (b + x * a)), c + x * (2.0 * b + 3.0 * a * x)


def _bicubic(grid, r, c):
    """ceres::BiCubicInterpolator<Grid2D<float>>::Evaluate(r, c) -> f, dfdr, dfdc; Grid2D clamps (include/ceres.h:46-62)"""
    rows, cols = grid.shape
    row, col = math.floor(r), math.floor(c)
    f, df = [], []
    for k in range(4):
        rr = min(max(row - 1 + k, 0), rows - 1)
        p = [float(grid[rr, min(max(col - 1 + j, 0), cols - 1)]) for j in range(4)]
        a, b = _cubic(p[0], p[1], p[2], p[3], c - col)
        f.append(a); df.append(b)
    val, dfdr = _cubic(f[0], f[1], f[2], f[3], r - row)
    dfdc, _ = _cubic(df[0], df[1], df[2], df[3], r - row)
    return val, dfdr, dfdc


def subpixel_evaluate(gradx, grady, prior, length, params, steps=7):
    """cost, gradient (5) of SubpixelCorner(gradx, grady, prior, steps, length) at params"""
    u, v, h = params[0], params[1], params[4]
    step_len = length / steps
    step_vec = []
    for i in range(1, steps + 1):
        step_vec += [-i * step_len, i * step_len]
    cost = 0.1 * ((prior[0] - u) ** 2 + (prior[1] - v) ** 2)
    g = [0.2 * (u - prior[0]), 0.2 * (v - prior[1]), 0.0, 0.0, 0.0]
    for direction in range(2):
        th = 2 + direction
        s, c = math.sin(params[th]), math.cos(params[th])
        flow = 1.0 if direction else -1.0
        for ln in step_vec:
            eta = (1 if ln > 0 else -1) * flow
            ui = u + c * ln - s * h * eta
            vi = v + s * ln + c * h * eta
            fu, fuv, fuu = _bicubic(gradx, vi, ui)
            fv, fvv, fvu = _bicubic(grady, vi, ui)
            cost += eta * (fv * c - fu * s)
            dudth = -s * ln - c * h * eta
            dvdth = c * ln - s * h * eta
            g[0] += eta * (fvu * c - fuu * s)
            g[1] += eta * (fvv * c - fuv * s)
            g[th] += eta * ((fvv * dvdth + fvu * dudth) * c - (fuv * dvdth + fuu * dudth) * s - fu * c - fv * s)
            g[4] += fvv * c * c + fuu * s * s - s * c * (fvu + fuv)
    return cost, np.array(g)


# ---- improveCorners' minimisation (:176-196): ceres::GradientProblemSolver with default options -----------------------------
# Ceres is a third-party dependency absent from /root/reference: its published line-search minimiser is restated here as
# in oracle/shim/ceres/gradient_solver.h (L-BFGS, rank 20, H0 = I, pairs with s.y <= 1e-14 skipped; Wolfe search =
# bracketing + zoom on the cubic through two (value, slope) samples; Ceres' default tolerances).  PARITY UNPINNED against
# Ceres itself for the minimiser's path -- see DESIGN.md section 2.
def _cubic_min(a, b, lo, hi):
    """minimiser over [lo, hi] of the cubic through samples a, b = (x, f, g): midpoint, ends, real parts of p' roots"""
    h = b[0] - a[0]
    d = (b[1] - a[1]) / h
    k3 = (a[2] + b[2] - 2.0 * d) / (h * h)
    k2 = (3.0 * d - 2.0 * a[2] - b[2]) / h

    def P(x):
        t = x - a[0]
        return a[1] + t * (a[2] + t * (k2 + t * k3))
    best_x = 0.5 * (lo + hi)
    best = P(best_x)
    for x in (lo, hi):
        v = P(x)
        if v < best:
            best, best_x = v, x
    qa, qb, qc = 3.0 * k3, 2.0 * k2, a[2]
    roots = []
    if qa != 0.0:
        D = qb * qb - 4.0 * qa * qc
        sD = math.sqrt(abs(D))
        if D >= 0.0:
            if qb >= 0.0:
                roots = [(-qb - sD) / (2.0 * qa), (2.0 * qc) / (-qb - sD)]
            else:
                roots = [(2.0 * qc) / (-qb + sD), (-qb + sD) / (2.0 * qa)]
        else:
            roots = [-qb / (2.0 * qa)] * 2
    elif qb != 0.0:
        roots = [-qc / qb]
    for r in roots:
        x = r + a[0]
        if not (lo <= x <= hi):
            continue
        v = P(x)
        if v < best:
            best, best_x = v, x
    return best_x


def minimize_gradient_problem(fun, x0, max_iter=50, rank=20, f_tol=1e-6, g_tol=1e-10, x_tol=1e-8, min_step=1e-9, c1=1e-4, c2=0.9,
                              expand=10.0, max_trials=20, max_restarts=5):
    """fun(x) -> (cost, gradient).  Returns (x, iterations)."""
    x = [float(v) for v in x0]
    n = len(x)
    f, g = fun(x)
    g = [float(v) for v in g]
    f_prev = 0.0
    S, Y, SY = [], [], []
    it = restarts = 0

    def dot(a, b):
        s = 0.0
        for i in range(n):
            s += a[i] * b[i]
        return s
    if max(abs(v) for v in g) <= g_tol:
        return np.array(x), 0
    while it < max_iter:
        it += 1
        steepest = not S
        d = list(g)
        if not steepest:
            m = len(S)
            alpha = [0.0] * m
            for i in range(m - 1, -1, -1):
                alpha[i] = dot(S[i], d) / SY[i]
                for k in range(n):
                    d[k] -= alpha[i] * Y[i][k]
            for i in range(m):
                beta = dot(Y[i], d) / SY[i]
                for k in range(n):
                    d[k] += S[i][k] * (alpha[i] - beta)
        d = [-v for v in d]
        slope = dot(g, d)
        if not steepest and slope >= 0.0:
            restarts += 1
            if restarts > max_restarts:
                break
            S, Y, SY = [], [], []
            d = [-v for v in g]
            slope = dot(g, d)
            steepest = True
        gmax = max(abs(v) for v in g)
        step0 = min(1.0, 1.0 / gmax) if (it == 1 or steepest) else min(1.0, 2.0 * (f - f_prev) / slope)
        if not step0 > 0.0:
            break
        dmax = max(abs(v) for v in d)

        def phi(a):
            ft, gt = fun([x[i] + a * d[i] for i in range(n)])
            return (a, ft, dot([float(v) for v in gt], d))
        start = (0.0, f, slope)
        prev, cur, lo, hi = start, phi(step0), start, start
        zoom, ok, trials = False, True, 0
        while True:                                        # bracketing
            trials += 1
            if cur[1] > start[1] + c1 * start[2] * cur[0] or (prev[0] > 0.0 and cur[1] > prev[1]):
                zoom, lo, hi = True, prev, cur
                break
            if abs(cur[2]) <= -c2 * start[2]:
                lo = hi = cur
                break
            if cur[2] >= 0.0:
                zoom, lo, hi = True, cur, prev
                break
            if abs(cur[0] - prev[0]) * dmax < min_step:
                ok = False
                break
            if trials >= max_trials:
                lo = cur if cur[1] < lo[1] else lo
                break
            a = _cubic_min(prev, cur, cur[0], cur[0] * expand)
            if a * dmax < min_step:
                ok = False
                break
            prev, cur = cur, phi(a)
        if not ok:
            break
        best = lo
        if zoom:
            if lo[1] > hi[1]:
                lo, hi = hi, lo
            sol = None
            while True:
                if trials >= max_trials or abs(hi[0] - lo[0]) * dmax < min_step:
                    break
                trials += 1
                lb, ub = (lo, hi) if lo[0] < hi[0] else (hi, lo)
                sol = phi(_cubic_min(lb, ub, lb[0], ub[0]))
                if sol[1] > start[1] + c1 * start[2] * sol[0] or sol[1] >= lo[1]:
                    hi = sol
                    continue
                if abs(sol[2]) <= -c2 * start[2]:
                    break
                if sol[2] * (hi[0] - lo[0]) >= 0.0:
                    hi = lo
                lo = sol
            best = lo if (sol is None or sol[1] > lo[1]) else sol
        if not best[0] > 0.0:
            break
        xn = [x[i] + best[0] * d[i] for i in range(n)]
        fn, gn = fun(xn)
        gn = [float(v) for v in gn]
        s = [best[0] * d[i] for i in range(n)]
        y = [gn[i] - g[i] for i in range(n)]
        step_norm = math.sqrt(sum(v * v for v in s))
        x_norm = math.sqrt(sum(v * v for v in xn))
        sy = dot(s, y)
        if sy > 1e-14:
            if len(S) == rank:
                S.pop(0); Y.pop(0); SY.pop(0)
            S.append(s); Y.append(y); SY.append(sy)
        f_prev, f, x, g = f, fn, xn, gn
        if max(abs(v) for v in g) <= g_tol:
            break
        if step_norm <= x_tol * (x_norm + x_tol):
            break
        if abs(f_prev - f) <= f_tol * abs(f_prev):
            break
    return np.array(x), it


def improve_corners(gradx, grady, grid, start, nx):
    """improveCorners (:162-198) on the integer grid with initPoin's start values -> refined (n, 2), iterations (n,)"""
    reach = refinement_reach(grid, nx)
    out, its = [], []
    for i in range(len(grid)):
        prior = (float(grid[i][0]), float(grid[i][1]))
        x, it = minimize_gradient_problem(lambda p: subpixel_evaluate(gradx, grady, prior, float(reach[i]), p), start[i])
        out.append(x[:2]); its.append(it)
    return np.array(out), np.array(its)
