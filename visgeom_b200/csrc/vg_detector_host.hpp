// vg_detector_host.hpp -- the sequential half of the checkerboard detector, host C++:
//   CornerDetector::selectCandidates (after the scan for maxima)   src/calibration/corner_detector.cpp:536-609
//   CornerDetector::checkCorner / scaleInvarient                    :331-440, :444-492
//   CornerDetector::constructGraph                                  :612-836
//   CornerDetector::selectPattern, extractSequence,
//     selectBestOrthogonalChain, verifyDetection                    :850-1076
//   CornerDetector::getCircle / getTransitions / initPoin           :1079-1108, :1135-1259, :1261-1298
// The pixel-parallel stages (the blurs, gradients, response, the scan for local maxima, the sub-pixel refinement) run on
// the GPU (vg_corner.cu, vg_detector.cu); what is here is order-dependent by construction -- a priority heap, a
// breadth-first flood whose result depends on the queue order, walks over a small graph -- and touches a few thousand
// pixels per image, so it runs on host threads, one image per thread, from the two blurred 8-bit images the GPU sends
// back (2 bytes per pixel): the float maps _gradx / _grady / _imgrad are recomputed per access with the very
// operations of computeResponse (:283-290), which is cheaper than moving 12 bytes per pixel over PCIe.
// Results must equal the reference's, candidate for candidate and arc for arc, so the tie-breaking rules of the
// reference are kept on purpose: std::make_heap / pop_heap with a comparator on the key alone, first-maximum searches,
// integer-truncated vector norms (Eigen's Matrix<int>::norm), bilinear()'s coupled clamping, sign(0) = -1.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <utility>
#include <vector>

namespace vg {
namespace det {

struct Pt { int u, v; };
struct Maximum { float val; int u, v; };               // what the GPU's scan emits

// one image and the blurred copies computeResponse made of it
struct Frame {
    const unsigned char *img, *s1, *s2;
    int W, H;
    int px(int v, int u) const { return img[(size_t)v * W + u]; }
    bool interior(int v, int u) const { return u >= 1 && v >= 1 && u < W - 1 && v < H - 1; }
    // sharp gradient before the 0.01 scale (:283-285); zero on the one-pixel border, which the reference leaves unset
    double gxs(int v, int u) const
    {
        const size_t i = (size_t)v * W + u;
        return ((int)s1[i + 1] - (int)s1[i - 1] - 0.3 * ((int)s2[i + 1] - (int)s2[i - 1])) / 2.;
    }
    double gys(int v, int u) const
    {
        const size_t i = (size_t)v * W + u;
        return ((int)s1[i + W] - (int)s1[i - W] - 0.3 * ((int)s2[i + W] - (int)s2[i - W])) / 2.;
    }
    float gradx(int v, int u) const { return interior(v, u) ? (float)(gxs(v, u) * 0.01) : 0.f; }      // :288
    float grady(int v, int u) const { return interior(v, u) ? (float)(gys(v, u) * 0.01) : 0.f; }      // :289
    float imgrad(int v, int u) const                                                                  // :290
    {
        if (!interior(v, u)) return 0.f;
        const double a = gxs(v, u), b = gys(v, u);
        return (float)(std::sqrt(a * a + b * b) * 0.01);
    }
    Pt clamp(int u, int v) const { return Pt{std::max(0, std::min(u, W - 1)), std::max(0, std::min(v, H - 1))}; }   // normalizePoint
    // bilinear<double>(_src2, x, y) (include/ocv.h): truncation, and a clamp of v that also fires when u was clamped
    double s2_bilinear(double x, double y) const
    {
        int u = (int)x, v = (int)y;
        const double dx = x - u, dy = y - v, dx2 = 1 - dx;
        bool out = false;
        if ((out |= u < 0)) u = 0;
        else if ((out |= u > W - 2)) u = W - 1;
        if ((out |= v < 0)) v = 0;
        else if ((out |= v > H - 2)) v = H - 1;
        const size_t i = (size_t)v * W + u;
        if (out) return s2[i];
        const double i00 = s2[i], i01 = s2[i + 1], i10 = s2[i + W], i11 = s2[i + W + 1];
        return (i11 * dx + i10 * dx2) * dy + (i01 * dx + i00 * dx2) * (1 - dy);
    }
};

inline int sgn(double x) { return 2 * int(x > 0) - 1; }          // include/std.h: sign(0) = -1

// ---- circles -------------------------------------------------------------------------------------------------------
// getCircle (:1079-1108): radius 1 is the 8-neighbourhood from (1, 0) through (1, 1); larger radii are walked with the
// implicit-curve rasteriser (include/utils/curve_rasterizer.h) from (r, 0) towards (0, r) until the walk is back within
// one pixel of its start.  The walk depends on the radius only (every quantity is an integer offset from the centre),
// so the offsets are tabulated once; clamping to the image happens per sample.
class CircleTable {
public:
    static constexpr int RMAX = 12;
    CircleTable()
    {
        static const int du[8] = {1, 1, 0, -1, -1, -1, 0, 1}, dv[8] = {0, 1, 1, 1, 0, -1, -1, -1};
        for (int i = 0; i < 8; i++) ring_[1].push_back(Pt{du[i], dv[i]});
        for (int r = 2; r <= RMAX; r++) walk(r, ring_[r]);
    }
    const std::vector<Pt> &ring(int r) const { return ring_[r]; }

private:
    std::vector<Pt> ring_[RMAX + 1];
    // f(u, v) = u^2 + v^2 - r^2, grad = (2u, 2v); state: position, f there, gradient there
    static void walk(const int r, std::vector<Pt> &out)
    {
        int u = r, v = 0;
        double fu = 2.0 * u, fv = 2.0 * v, delta = 0.0;
        const int eps = (fu * (r - v) - fv * (0 - u) > 0) ? 1 : -1;
        auto move_u = [&](int d) {
            if (d == 0) return;
            u += d;
            const double f2 = 2.0 * u;
            delta += 0.5 * d * (fu + f2);
            fu = f2; fv = 2.0 * v;
        };
        auto move_v = [&](int d) {
            if (d == 0) return;
            v += d;
            const double f2 = 2.0 * v;
            delta += 0.5 * d * (fv + f2);
            fv = f2; fu = 2.0 * u;
        };
        for (int i = 0;; i++) {
            out.push_back(Pt{u, v});
            if (i > 5 && std::abs(u - r) <= 1 && std::abs(v) <= 1) break;
            if (std::fabs(fu) > std::fabs(fv)) { move_v(eps * sgn(fu)); move_u((int)-std::round(delta / fu)); }
            else { move_u(-eps * sgn(fv)); move_v((int)-std::round(delta / fv)); }
            if (out.size() > 1024) break;          // cannot happen for a circle; keeps a broken table finite
        }
    }
};

inline const CircleTable &circles() { static const CircleTable t; return t; }

// samples of the image on the circle and their circular central differences (:1110-1133)
struct Ring {
    int n;
    Pt p[128];
    double t[128];
};

inline void ring_transitions(const Frame &F, const Pt c, const int radius, Ring &R)
{
    const std::vector<Pt> &off = circles().ring(radius);
    const int n = R.n = (int)off.size();
    double s[128];
    for (int i = 0; i < n; i++) {
        R.p[i] = F.clamp(c.u + off[i].u, c.v + off[i].v);
        s[i] = F.px(R.p[i].v, R.p[i].u);
    }
    R.t[0] = s[1] - s[n - 1];
    for (int i = 1; i < n - 1; i++) R.t[i] = s[i + 1] - s[i - 1];
    R.t[n - 1] = s[0] - s[n - 2];
}

inline int first_max(const double *t, int a, int b) { int k = a; for (int i = a + 1; i < b; i++) if (t[k] < t[i]) k = i; return k; }
inline int first_min(const double *t, int a, int b) { int k = a; for (int i = a + 1; i < b; i++) if (t[i] < t[k]) k = i; return k; }

// setZero (corner_detector.h:149-166): clears the run of entries around k that share its sign, both ways round the ring
inline void clear_run(double *t, const int n, const int k)
{
    const double ref = t[k];
    int i = k;
    do { t[i] = 0; if (++i == n) i = 0; } while (t[i] * ref > 0);
    i = k;
    do { t[i] = 0; if (i == 0) i = n; i--; } while (t[i] * ref > 0);
}

// ---- candidate tests -----------------------------------------------------------------------------------------------
// checkCorner (:331-440): on every ring of radius checkRadius .. checkRadius + max(3, checkRadius) - 1 the transitions
// must show two strong rises and two strong falls, each pair roughly opposite, and no third one (MAX_FAULTS = 0)
inline bool check_corner(const Frame &F, const Pt c, const int check_radius)
{
    const int r_end = check_radius + std::max(3, check_radius);
    const double A1 = 0.3, A2 = 0.5;
    Ring R;
    for (int radius = check_radius; radius < r_end; radius++) {
        ring_transitions(F, c, radius, R);
        double *t = R.t;
        const int n = R.n, near = n / 2 - 2, far = n - near;
        const int m1 = first_max(t, 0, n);
        const double max1 = t[m1];
        clear_run(t, n, m1);
        const int m2 = first_max(t, 0, n);
        if (t[m2] < max1 * A1) return false;
        const double max2 = t[m2];
        clear_run(t, n, m2);
        if (std::abs(m2 - m1) < near || std::abs(m2 - m1) > far) return false;
        const int m3 = first_max(t, 0, n);
        if (t[m3] > max2 * A2) return false;
        clear_run(t, n, m3);
        const int n1 = first_min(t, 0, n);
        const double min1 = t[n1];
        if (min1 > max1 * -A1) return false;
        clear_run(t, n, n1);
        const int n2 = first_min(t, 0, n);
        if (t[n2] > min1 * -A1) return false;
        const double min2 = t[n2];
        clear_run(t, n, n2);
        if (std::abs(n2 - n1) < near || std::abs(n2 - n1) > far) return false;
        const int n3 = first_min(t, 0, n);
        if (t[n3] < min2 * A2) return false;
    }
    return true;
}

// scaleInvarient (:444-492): on some ring-shaped window of radius R .. 2R - 1 the gradients must be tangential
inline bool scale_invariant(const Frame &F, const Pt c, const int init_radius)
{
    for (int radius = init_radius; radius < init_radius * 2; radius++) {
        double acc = 0, norm_acc = 1e-10;
        for (int dv = -radius; dv <= radius; dv++)
            for (int du = -radius; du <= radius; du++) {
                const double sq = du * du + dv * dv;
                if (sq > radius * radius + 1 || sq < 1) continue;
                const int u = c.u + du, v = c.v + dv;
                if (u < 0 || u >= F.W || v < 0 || v >= F.H) continue;
                const double gx = F.gradx(v, u), gy = F.grady(v, u);
                const double g2 = gx * gx + gy * gy;
                if (g2 < 1e-3) continue;
                const double radial = gx * du + gy * dv;
                acc += radial * radial / sq;
                norm_acc += g2;
            }
        if (acc / norm_acc < 0.3) return true;
    }
    return false;
}

// getTransitions (:1135-1259): the four points (strongest rise, its opposite rise, strongest fall, its opposite fall)
// on the smallest ring where they are balanced, replaced by a larger ring's while the strongest rise keeps growing
inline std::vector<Pt> transitions(const Frame &F, const Pt c, const int init_radius)
{
    std::vector<Pt> res;
    bool found = false;
    double best = 0;
    Ring R;
    // the strongest entry of the half ring opposite k (a quarter turn away on each side)
    auto opposite = [&](const int k, const bool want_max) {
        const int n = R.n;
        int a = k + n / 4, b = a + n / 2;
        auto pick = [&](int x, int y) { return want_max ? first_max(R.t, x, y) : first_min(R.t, x, y); };
        if (a < n && b >= n) {
            b %= n;
            const int i1 = pick(a, n), i2 = pick(0, b);
            return (want_max ? R.t[i1] > R.t[i2] : R.t[i1] < R.t[i2]) ? i1 : i2;
        }
        a %= n; b %= n;
        return pick(a, b);
    };
    for (int radius = 1; radius <= init_radius + 1; radius++) {
        ring_transitions(F, c, radius, R);
        const int hi1 = first_max(R.t, 0, R.n), hi2 = opposite(hi1, true);
        const int lo1 = first_min(R.t, 0, R.n), lo2 = opposite(lo1, false);
        if (found) {
            if (best > 0.7 * R.t[hi1]) break;
            found = false;
            res.clear();
        }
        if (R.t[hi2] < 0.4 * R.t[hi1]) continue;
        if (R.t[lo2] > 0.4 * R.t[lo1]) continue;
        res = {R.p[hi1], R.p[hi2], R.p[lo1], R.p[lo2]};
        best = R.t[hi1];
        found = true;
    }
    return res;
}

// initPoin (:1261-1298): start values of the refinement -- the intersection of the lines through the two rises and
// through the two falls, and the directions of those lines.  Returns false where the reference would read past an empty
// transition list (it never checks; a corner that passed the candidate tests has its four points).
inline bool init_point(const Frame &F, const Pt c, const int init_radius, double *d)
{
    const std::vector<Pt> tr = transitions(F, c, init_radius);
    if (tr.size() != 4) return false;
    const Pt A = tr[0], C = tr[1], B = tr[2], D = tr[3];
    const double m00 = A.v - C.v, m01 = C.u - A.u, m10 = B.v - D.v, m11 = D.u - B.u;
    const double b0 = A.v * m01 + A.u * m00, b1 = B.v * m11 + B.u * m10;
    // Eigen's 2 x 2 inverse: the adjugate times 1 / det, then the product with b
    const double inv_det = 1.0 / (m00 * m11 - m01 * m10);
    const double i00 = m11 * inv_det, i01 = -m01 * inv_det, i10 = -m10 * inv_det, i11 = m00 * inv_det;
    d[0] = i00 * b0 + i01 * b1;
    d[1] = i10 * b0 + i11 * b1;
    d[2] = std::atan2((double)(A.v - C.v), (double)(A.u - C.u));
    d[3] = std::atan2((double)(B.v - D.v), (double)(B.u - D.u));
    d[4] = 0;
    return true;
}

// ---- heap of (key, point) with the reference's comparator: the key alone, so equal keys keep the order the standard
// library's heap operations give them -- the reference's order
struct Keyed { double key; Pt p; };
inline bool keyed_less(const Keyed &a, const Keyed &b) { return a.key < b.key; }

// selectCandidates after the scan (:536-609).  `maxima` in scan order (rows, then columns).  Returns the candidates in
// the order constructGraph pops them (ascending u + v through the heap).
inline std::vector<Pt> select_candidates(const Frame &F, const std::vector<Maximum> &maxima, const int Nx, const int Ny,
                                         const int init_radius)
{
    std::vector<Keyed> heap;
    heap.reserve(maxima.size());
    for (const Maximum &m : maxima) heap.push_back(Keyed{(double)m.val, Pt{m.u, m.v}});
    std::make_heap(heap.begin(), heap.end(), keyed_less);
    // threshold: 5 % of the mean of the Nx Ny strongest
    std::vector<Keyed> top = heap;
    double acc = 0;
    const int n_ref = Nx * Ny;
    for (int i = 0; i < n_ref && !top.empty(); i++) {
        std::pop_heap(top.begin(), top.end(), keyed_less);
        acc += top.back().key;
        top.pop_back();
    }
    const double thresh = 0.05 * acc / n_ref;
    std::vector<Keyed> hyp;
    const int cap = 10 * Nx * Ny;
    int taken = 0;
    while (!heap.empty() && heap.front().key > thresh && taken < cap) {
        std::pop_heap(heap.begin(), heap.end(), keyed_less);
        const Pt p = heap.back().p;
        heap.pop_back();
        bool corner = false;
        for (int radius = 1; radius <= init_radius && !corner; radius++) corner = check_corner(F, p, radius);
        if (!corner) continue;
        if (!scale_invariant(F, p, init_radius)) continue;
        hyp.push_back(Keyed{(double)(-p.u - p.v), p});
        taken++;
    }
    std::make_heap(hyp.begin(), hyp.end(), keyed_less);
    std::vector<Pt> out;
    out.reserve(hyp.size());
    while (!hyp.empty()) {
        std::pop_heap(hyp.begin(), hyp.end(), keyed_less);
        out.push_back(hyp.back().p);
        hyp.pop_back();
    }
    return out;
}

// ---- the graph over the candidates -----------------------------------------------------------------------------------
struct Graph {
    std::vector<Pt> pt;
    std::vector<std::vector<int>> arcs;                 // neighbours in the order they were found
    std::map<std::pair<int, int>, int> sign;            // which side of the arc is the bright one
    int arc_sign(int a, int b) const
    {
        auto it = sign.find(std::make_pair(a, b));
        return it == sign.end() ? 0 : it->second;       // std::map::operator[] of the reference: a missing arc reads 0
    }
    bool linked(int a, int b) const { return std::find(arcs[a].begin(), arcs[a].end(), b) != arcs[a].end(); }
};

// The flood's angle test (:759-770) accepts a pixel at offset (x1, y1) from its candidate when the gradient there is
// within tol(t) = min(pi / 5, t / 150 + 1 / t) of perpendicular to the offset, t = |(x1, y1)|: with c = offset . grad and
// s = offset x grad, |atan2(s, c)| must lie in [pi / 2 - tol, pi / 2 + tol].  That is |c| <= |s| tan(tol) -- one
// multiplication against a table of tan(tol) over the integer t^2 instead of an atan2 per pixel; within 1e-12 of the
// boundary (and for s = 0) the reference's own expression decides, so the outcome is the reference's in every case.
class AngleGate {
public:
    static constexpr int SPAN = 1 << 16;             // t^2 up to 65535: offsets up to 181 pixels (the flood reaches 140 + a ring)
    AngleGate() : tan_tol_(SPAN)
    {
        for (int q = 1; q < SPAN; q++) {
            const double t = std::sqrt((double)q);
            tan_tol_[q] = std::tan(std::min(M_PI / 5, t / 150. + 1 / t));
        }
    }
    // t > 0
    bool accepts(const int t2, const double t, const double c, const double s) const
    {
        if (t2 < SPAN && s != 0) {
            const double lim = std::fabs(s) * tan_tol_[t2], ac = std::fabs(c);
            if (ac > lim * (1 + 1e-12)) return false;
            if (ac < lim * (1 - 1e-12)) return true;
        }
        const double angle = std::fabs(std::atan2(s, c));
        const double tol = std::min(M_PI / 5, t / 150. + 1 / t);
        return !(angle < M_PI / 2 - tol || angle > M_PI / 2 + tol);
    }

private:
    std::vector<double> tan_tol_;
};
inline const AngleGate &angle_gate() { static const AngleGate g; return g; }

// the flood's ownership map: one 16-bit candidate index per pixel (the reference's Mat16s _idxMap).  A flood owns a few
// thousand pixels of the million, so a map is cleared cell by cell after use and handed back to a pool -- the worker
// threads of a pass are short-lived, and a fresh 2 MB map per image would cost as much as a fifth of the flood.
// `judged` remembers, per pixel, the last candidate whose flood has tested it (+(idx + 1): queued, -(idx + 1): refused).
// The reference tests and queues a pixel again from every neighbour that is popped (five times on average); the test is
// a function of pixel and candidate alone, and a second queued copy is popped after the first, when the pixel is owned
// -- a no-op either way -- so skipping the repeats changes neither the ownership map nor the order of the arcs.
struct OwnerMap {
    std::vector<int16_t> cell, judged;
    std::vector<uint32_t> touched, touched_judged;
    void prepare(const size_t n)
    {
        if (cell.size() < n) { cell.assign(n, (int16_t)-1); judged.assign(n, (int16_t)0); }
        touched.clear(); touched_judged.clear();
    }
    void reset()
    {
        for (uint32_t i : touched) cell[i] = -1;
        for (uint32_t i : touched_judged) judged[i] = 0;
        touched.clear(); touched_judged.clear();
    }
};

class OwnerMapPool {
public:
    std::unique_ptr<OwnerMap> take()
    {
        std::lock_guard<std::mutex> g(mu_);
        if (free_.empty()) return std::unique_ptr<OwnerMap>(new OwnerMap());
        std::unique_ptr<OwnerMap> m = std::move(free_.back());
        free_.pop_back();
        return m;
    }
    void give(std::unique_ptr<OwnerMap> m)
    {
        std::lock_guard<std::mutex> g(mu_);
        if (free_.size() < 32) free_.push_back(std::move(m));        // (4 bytes per pixel each: what does not fit is freed)
    }

private:
    std::mutex mu_;
    std::vector<std::unique_ptr<OwnerMap>> free_;
};
inline OwnerMapPool &owner_maps() { static OwnerMapPool p; return p; }

// constructGraph (:612-836): every candidate floods outwards along the edges that leave it (pixels whose gradient is
// strong enough and perpendicular to the ray from the candidate); where two floods touch, the candidates are linked.
inline void construct_graph(const Frame &F, const std::vector<Pt> &cand, const int init_radius, Graph &G)
{
    struct Cell { int t, idx, u, v; };
    const int n = (int)cand.size();
    G.pt = cand;
    G.arcs.assign(n, {});
    G.sign.clear();
    std::unique_ptr<OwnerMap> owner_map_ptr = owner_maps().take();
    OwnerMap &owner_map = *owner_map_ptr;
    owner_map.prepare((size_t)F.W * F.H);
    int16_t *owner = owner_map.cell.data(), *judged = owner_map.judged.data();
    const AngleGate &gate = angle_gate();
    std::vector<double> grad_min(n);
    std::queue<Cell> fringe;
    for (int i = 0; i < n; i++) {
        grad_min[i] = std::numeric_limits<double>::max();
        for (const Pt &q : transitions(F, cand[i], init_radius)) {
            fringe.push(Cell{0, i, q.u, q.v});
            grad_min[i] = std::min((double)F.imgrad(q.v, q.u) / 2, grad_min[i]);
        }
    }
    static const int du8[8] = {-1, 0, 1, 1, 1, 0, -1, -1}, dv8[8] = {-1, -1, -1, 0, 1, 1, 1, 0};
    const int REACH = 140;
    while (!fringe.empty() && fringe.front().t < REACH) {
        const Cell e = fringe.front();
        fringe.pop();
        const size_t at = (size_t)e.v * F.W + e.u;
        if (owner[at] != -1) continue;
        owner[at] = (int16_t)e.idx;
        owner_map.touched.push_back((uint32_t)at);
        bool touched = false;
        for (int k = 0; k < 8; k++) {
            const int u2 = e.u + du8[k], v2 = e.v + dv8[k];
            if (u2 < 0 || u2 >= F.W || v2 < 0 || v2 >= F.H) continue;
            const int other = owner[(size_t)v2 * F.W + u2];
            if (other == -1 || other == e.idx) continue;
            touched = true;
            if (G.linked(e.idx, other)) continue;
            // the side of the arc that is brighter, voted over 4 points of the arc and init_radius distances
            const double ax = G.pt[other].u - G.pt[e.idx].u, ay = G.pt[other].v - G.pt[e.idx].v;
            const double len = std::sqrt(ax * ax + ay * ay), nx = ax / len, ny = ay / len;
            int votes = 0;
            for (int base = 1; base <= init_radius; base++) {
                const double sx = nx * base, sy = ny * base;
                for (int lambda = 1; lambda < 5; lambda++) {
                    const double f = double(lambda) / 5;
                    const double mx = G.pt[e.idx].u + ax * f, my = G.pt[e.idx].v + ay * f;
                    votes += sgn(F.s2_bilinear(mx - sy, my + sx) - F.s2_bilinear(mx + sy, my - sx));
                }
            }
            votes = sgn(votes);
            G.arcs[e.idx].push_back(other);
            G.arcs[other].push_back(e.idx);
            G.sign[std::make_pair(e.idx, other)] = votes;
            G.sign[std::make_pair(other, e.idx)] = -votes;
        }
        if (touched) continue;
        for (int k = 0; k < 8; k++) {
            const int u2 = e.u + du8[k], v2 = e.v + dv8[k];
            if (u2 < 0 || u2 >= F.W || v2 < 0 || v2 >= F.H) continue;
            const size_t at2 = (size_t)v2 * F.W + u2;
            if (owner[at2] != -1) continue;
            const int16_t tag = (int16_t)(e.idx + 1), was = judged[at2];
            if (was == tag || was == -tag) continue;                // this flood has judged the pixel already
            if (was == 0) owner_map.touched_judged.push_back((uint32_t)at2);
            judged[at2] = -tag;
            const int ix = u2 - G.pt[e.idx].u, iy = v2 - G.pt[e.idx].v;
            const double x1 = ix, y1 = iy;
            const double t = std::sqrt(x1 * x1 + y1 * y1);
            const double xg = F.gradx(v2, u2), yg = F.grady(v2, u2);
            const double across = std::fabs(xg * y1 - yg * x1);
            if (t > 0 && across / t < grad_min[e.idx]) continue;
            if (t > 0 && !gate.accepts(ix * ix + iy * iy, t, x1 * xg + y1 * yg, x1 * yg - y1 * xg)) continue;
            judged[at2] = tag;
            fringe.push(Cell{e.t + 1, e.idx, u2, v2});
        }
    }
    owner_map.reset();
    owner_maps().give(std::move(owner_map_ptr));
}

// Eigen's Matrix<int, 2, 1>::norm(): the square root converted back to int
inline int inorm(int x, int y) { return (int)std::sqrt((double)(x * x + y * y)); }
// compareVectors (:838-848)
inline double direction_change(int x1, int y1, int x2, int y2) { return inorm(x1 - x2, y1 - y2) / double(inorm(x1, y1)); }

// extractSequence (:850-922): from the arc a -> b keep stepping to the neighbour that continues it best (relative change
// of the step below 1) across an arc of the opposite sign
inline std::vector<int> extract_sequence(const Graph &G, const int a, const int b)
{
    std::vector<int> chain;
    auto next = [&](const int i0, const int i1, const bool skip_back) {
        const int dx = G.pt[i1].u - G.pt[i0].u, dy = G.pt[i1].v - G.pt[i0].v;
        int best = -1;
        double best_change = 1;
        for (int n2 : G.arcs[i1]) {
            if (skip_back && n2 == i0) continue;
            const double change = direction_change(dx, dy, G.pt[n2].u - G.pt[i1].u, G.pt[n2].v - G.pt[i1].v);
            if (G.arc_sign(i1, n2) == G.arc_sign(i0, i1)) continue;
            if (change < best_change) { best_change = change; best = n2; }
        }
        return best;
    };
    const int c = next(a, b, false);
    if (c == -1) return chain;
    chain = {a, b, c};
    for (;;) {
        const int k = next(chain[chain.size() - 2], chain.back(), true);
        if (k == -1) break;
        chain.push_back(k);
    }
    return chain;
}

// selectBestOrthogonalChain (:930-972): among the chains that leave a across an arc of the other sign, the one most
// perpendicular (on the side eps) to a -> b that has at least `length` corners
inline std::vector<int> best_orthogonal_chain(const Graph &G, const int a, const int b, const int eps, const int length)
{
    double best_cost = 0.3;
    std::vector<int> best;
    const int base = G.arc_sign(a, b);
    for (int nx : G.arcs[a]) {
        if (nx == b) continue;
        if (G.arc_sign(a, nx) == base) continue;
        const int x1 = G.pt[b].u - G.pt[a].u, y1 = G.pt[b].v - G.pt[a].v;
        const int x2 = G.pt[nx].u - G.pt[a].u, y2 = G.pt[nx].v - G.pt[a].v;
        std::vector<int> chain = extract_sequence(G, a, nx);
        const double cost = eps * (x1 * y2 - y1 * x2) / double(inorm(x1, y1) * inorm(x2, y2));
        if (cost < best_cost) continue;
        if ((int)chain.size() >= length) { best_cost = cost; best = chain; }
    }
    while ((int)best.size() > length) best.pop_back();
    return best;
}

// verifyDetection (:1060-1076): every column of the grid must also be a chain of the graph
inline bool verify(const Graph &G, const std::vector<int> &idx, const int Nx, const int Ny)
{
    if ((int)idx.size() != Nx * Ny) return false;
    for (int i = 1; i < Nx; i++) {
        if (!G.linked(idx[i], idx[Nx + i])) return false;
        const std::vector<int> chain = extract_sequence(G, idx[i], idx[Nx + i]);
        if ((int)chain.size() < Ny) return false;
        for (int j = 2; j < Ny; j++)
            if (chain[j] != idx[j * Nx + i]) return false;
    }
    return true;
}

// selectPattern (:975-1058): a corner of the board = a candidate with a chain of Ny along one arc and a perpendicular
// chain of Nx; the rows are then woven from the corners of the first column
inline std::vector<int> select_pattern(const Graph &G, const int Nx, const int Ny)
{
    std::vector<int> res;
    for (int i0 = 0; i0 < (int)G.pt.size(); i0++) {
        if (G.arcs[i0].size() < 2) continue;
        std::vector<int> col, row;
        for (int n : G.arcs[i0]) {
            std::vector<int> chain = extract_sequence(G, i0, n);
            if ((int)chain.size() < Ny) continue;
            chain.resize(Ny);
            row = best_orthogonal_chain(G, i0, n, -1, Nx);
            if ((int)row.size() == Nx) { col = chain; break; }
        }
        if (col.empty() || row.empty()) continue;
        res = row;
        for (size_t i = 1; i < col.size(); i++) {
            const std::vector<int> next_row = best_orthogonal_chain(G, col[i], col[i - 1], 1, Nx);
            if ((int)next_row.size() < Nx) break;
            res.insert(res.end(), next_row.begin(), next_row.begin() + Nx);
        }
        if ((int)res.size() == Nx * Ny && verify(G, res, Nx, Ny)) return res;
        res.clear();
    }
    return res;
}

// one scale of detectPattern after the GPU stages (:234-242): candidates, graph, pattern.  Returns the Nx Ny integer
// corner positions in board order, or nothing.
inline std::vector<Pt> detect_at_scale(const Frame &F, std::vector<Maximum> &maxima, const int Nx, const int Ny,
                                       const int init_radius)
{
    std::sort(maxima.begin(), maxima.end(), [](const Maximum &a, const Maximum &b) { return a.v != b.v ? a.v < b.v : a.u < b.u; });
    const std::vector<Pt> cand = select_candidates(F, maxima, Nx, Ny, init_radius);
    std::vector<Pt> out;
    if ((int)cand.size() < Nx * Ny) return out;
    if (cand.size() > 32766) return out;               // the reference's index map is 16-bit; 10 Nx Ny candidates at most
    Graph G;
    construct_graph(F, cand, init_radius, G);
    const std::vector<int> idx = select_pattern(G, Nx, Ny);
    if ((int)idx.size() != Nx * Ny) return out;
    for (int i : idx) out.push_back(G.pt[i]);
    return out;
}

// improveCorners' first loop (:164-175): the reach of the refinement's sample lines, 70 % of the distance to the
// neighbouring corners, 7 pixels at most
inline void refinement_reach(const std::vector<Pt> &p, const int Nx, double *reach)
{
    auto dist = [&](int a, int b) {
        const double dx = p[a].u - p[b].u, dy = p[a].v - p[b].v;
        return std::sqrt(dx * dx + dy * dy);
    };
    const int n = (int)p.size();
    for (int i = 0; i < n; i++) {
        double r = 7;
        r = std::min(r, dist(i, i > Nx ? i - Nx : i + Nx) * 0.7);
        r = std::min(r, dist(i, i > 0 ? i - 1 : i + 1) * 0.7);
        reach[i] = r;
    }
}

}  // namespace det
}  // namespace vg
