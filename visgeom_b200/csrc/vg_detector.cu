// vg_detector.cu -- the checkerboard detector behind the C ABI:
//   CornerDetector::detectPattern        src/calibration/corner_detector.cpp:223-260
//   CornerDetector::selectCandidates     :494-534  (the scan for local maxima: local_maxima_kernel)
//   SubpixelCorner::Evaluate             :47-100   (subpixel_refine_kernel)
//   CornerDetector::improveCorners       :162-198  (subpixel_refine_kernel: the line-search minimiser of the reference's
//                                                   ceres::GradientProblemSolver with default options, one warp per corner)
// Division of labour: everything that visits every pixel runs on the GPU, batched over images -- the two blurs, the
// gradient maps and the saddle response (vg_corner.cu), the scan for local maxima, and the refinement of the Nx Ny
// corners found, whose 2 x 28 bicubic samples per cost evaluation sit one per lane.  The order-dependent middle
// (candidate tests, flood-fill graph, pattern search: vg_detector_host.hpp) runs on host threads, one image per thread,
// from the two blurred 8-bit images (2 B per pixel back over PCIe) and the list of maxima.
// This translation unit is compiled with -fmad=false: the refinement's arithmetic then rounds operation by operation
// like the reference's host code, so that the two minimisers walk the same path (sin / cos and the order of the 28-term
// sums still differ in the last bit; results are held to a tolerance, see tests/test_detector_gpu.py).
#include "vg_detector.cuh"
#include "vg_detector_host.hpp"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>

namespace vg {
namespace {

// ---- scan for local maxima (:494-534) ---------------------------------------------------------------------------------
// A pixel is kept when its response is not below the image's mean kept response and no pixel of the disc of radius R
// (i^2 + j^2 <= R^2 + 1) holds a larger one; of two equal neighbours the one that comes first column-wise stays.
__global__ void __launch_bounds__(256)
local_maxima_kernel(const float *__restrict__ resp, const double *__restrict__ avg, const int W, const int H, const int R,
                    det::Maximum *__restrict__ out, unsigned int *__restrict__ count, const int cap)
{
    const int u = blockIdx.x * 32 + (threadIdx.x & 31), v = blockIdx.y * 8 + (threadIdx.x >> 5), im = blockIdx.z;
    if (u < R || u >= W - R || v < R || v >= H - R) return;
    const float *r = resp + (size_t)im * W * H;
    const float val = __ldg(r + (size_t)v * W + u);
    if ((double)val < avg[im]) return;
    for (int j = -R; j <= R; j++)
        for (int i = -R; i <= R; i++) {
            if ((i == 0 && j == 0) || i * i + j * j > R * R + 1) continue;
            const float nb = __ldg(r + (size_t)(v + j) * W + u + i);
            if (val <= nb && !(val == nb && (i > 0 || (i == 0 && j > 0)))) return;
        }
    const unsigned int k = atomicAdd(count + im, 1u);
    if (k < (unsigned)cap) out[(size_t)im * cap + k] = det::Maximum{val, u, v};
}

// ---- sub-pixel refinement (:47-100, :162-198) -------------------------------------------------------------------------
struct RefineJob {
    double prior[2];        // the integer corner: SubpixelCorner's _prior
    double x[5];            // start values from initPoin: u, v, the two edge directions, the half width h
    double reach;           // improveCorners' radMax = SubpixelCorner's `length`
    int img;                // slot of the image in the batch
    int pad;
};

constexpr int REFINE_WARPS = 4;
constexpr int LBFGS_RANK = 20;
constexpr int STEPS = 7;            // SubpixelCorner(..., steps = 7, ...)

// Catmull-Rom segment through p1, p2 (ceres/cubic_interpolation.h: CubicHermiteSpline)
__device__ __forceinline__ void spline(const double p0, const double p1, const double p2, const double p3, const double x,
                                       double &f, double &dfdx)
{
    const double a = 0.5 * (-p0 + 3.0 * p1 - 3.0 * p2 + p3);
    const double b = 0.5 * (2.0 * p0 - 5.0 * p1 + 4.0 * p2 - p3);
    const double c = 0.5 * (-p0 + p2);
    f = p1 + x * (c + x * (b + x * a));
    dfdx = c + x * (2.0 * b + 3.0 * a * x);
}

// ceres::BiCubicInterpolator<Grid2D<float>>::Evaluate(r, c): value and both first derivatives; the grid clamps
// coordinates to the image (include/ceres.h: Grid2D::GetValue)
__device__ __forceinline__ void bicubic(const float *__restrict__ g, const int W, const int H, const double r, const double c,
                                        double &f, double &dfdr, double &dfdc)
{
    const int row = (int)floor(r), col = (int)floor(c);
    int cc[4];
#pragma unroll
    for (int j = 0; j < 4; j++) cc[j] = min(max(col - 1 + j, 0), W - 1);
    double fr[4], dfr[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float *line = g + (size_t)min(max(row - 1 + k, 0), H - 1) * W;
        spline((double)__ldg(line + cc[0]), (double)__ldg(line + cc[1]), (double)__ldg(line + cc[2]), (double)__ldg(line + cc[3]),
               c - col, fr[k], dfr[k]);
    }
    double unused;
    spline(fr[0], fr[1], fr[2], fr[3], r - row, f, dfdr);
    spline(dfr[0], dfr[1], dfr[2], dfr[3], r - row, dfdc, unused);
}

__device__ __forceinline__ double warp_sum(double x)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    return x;             // the butterfly leaves the same bits in every lane
}

struct Corner {
    const float *gu, *gv;       // _gradx, _grady of the corner's image
    int W, H, lane;
    double prior_u, prior_v, step_len;
};

// SubpixelCorner::Evaluate: lane l < 28 takes sample l (direction l / 14; steps -1, +1, -2, +2, ... times step_len)
__device__ void evaluate(const Corner &C, const double *x, double &cost, double *grad)
{
    double c_f = 0, c_0 = 0, c_1 = 0, c_2 = 0, c_3 = 0, c_4 = 0;
    if (C.lane < 4 * STEPS) {
        const int dir = C.lane / (2 * STEPS), k = C.lane % (2 * STEPS), idx = (k >> 1) + 1;
        const double len = ((k & 1) ? idx : -idx) * C.step_len;
        double s, c;
        sincos(x[2 + dir], &s, &c);
        const double h = x[4], eta = (len > 0 ? 1 : -1) * (dir ? 1.0 : -1.0);
        const double ui = x[0] + c * len - s * h * eta, vi = x[1] + s * len + c * h * eta;
        double fu, fuv, fuu, fv, fvv, fvu;
        bicubic(C.gu, C.W, C.H, vi, ui, fu, fuv, fuu);
        bicubic(C.gv, C.W, C.H, vi, ui, fv, fvv, fvu);
        const double dudth = -s * len - c * h * eta, dvdth = c * len - s * h * eta;
        c_f = eta * (fv * c - fu * s);
        c_0 = eta * (fvu * c - fuu * s);
        c_1 = eta * (fvv * c - fuv * s);
        const double th = eta * ((fvv * dvdth + fvu * dudth) * c - (fuv * dvdth + fuu * dudth) * s - fu * c - fv * s);
        if (dir) c_3 = th; else c_2 = th;
        c_4 = fvv * c * c + fuu * s * s - s * c * (fvu + fuv);
    }
    const double du = C.prior_u - x[0], dv = C.prior_v - x[1];
    cost = 0.1 * (du * du + dv * dv) + warp_sum(c_f);
    grad[0] = 0.2 * (x[0] - C.prior_u) + warp_sum(c_0);
    grad[1] = 0.2 * (x[1] - C.prior_v) + warp_sum(c_1);
    grad[2] = warp_sum(c_2);
    grad[3] = warp_sum(c_3);
    grad[4] = warp_sum(c_4);
}

struct Trial { double x, f, g; };       // a step along the search direction, the cost there, the slope there

// minimiser over [lo, hi] of the cubic through two (value, slope) samples; candidates are the midpoint, the two ends and
// the real parts of the roots of the derivative (ceres/polynomial.cc: MinimizeInterpolatingPolynomial)
__device__ double cubic_minimiser(const Trial &a, const Trial &b, const double lo, const double hi)
{
    const double h = b.x - a.x, d = (b.f - a.f) / h;
    const double k3 = (a.g + b.g - 2.0 * d) / (h * h), k2 = (3.0 * d - 2.0 * a.g - b.g) / h;
    auto P = [&](const double x) { const double t = x - a.x; return a.f + t * (a.g + t * (k2 + t * k3)); };
    double best_x = 0.5 * (lo + hi), best = P(best_x);
    auto take = [&](const double x) { const double v = P(x); if (v < best) { best = v; best_x = x; } };
    take(lo);
    take(hi);
    const double qa = 3.0 * k3, qb = 2.0 * k2, qc = a.g;
    double r0 = 0, r1 = 0;
    int nr = 0;
    if (qa != 0.0) {
        const double D = qb * qb - 4.0 * qa * qc, sD = sqrt(fabs(D));
        if (D >= 0.0) {
            if (qb >= 0.0) { r0 = (-qb - sD) / (2.0 * qa); r1 = (2.0 * qc) / (-qb - sD); }
            else { r0 = (2.0 * qc) / (-qb + sD); r1 = (-qb + sD) / (2.0 * qa); }
        } else r0 = r1 = -qb / (2.0 * qa);
        nr = 2;
    } else if (qb != 0.0) { r0 = -qc / qb; nr = 1; }
    if (nr > 0) { const double x = r0 + a.x; if (x >= lo && x <= hi) take(x); }
    if (nr > 1) { const double x = r1 + a.x; if (x >= lo && x <= hi) take(x); }
    return best_x;
}

// One warp minimises one corner's SubpixelCorner cost: L-BFGS directions (rank 20, initial Hessian the identity, pairs
// with s.y <= 1e-14 dropped) and a Wolfe line search (bracketing, then zoom, trial steps from the cubic minimiser) --
// ceres::GradientProblemSolver::Options() as improveCorners leaves them.  Every lane carries the same x, gradient and
// search state (the sums come out of a butterfly), so the control flow is uniform across the warp.
__global__ void __launch_bounds__(32 * REFINE_WARPS)
subpixel_refine_kernel(const float *__restrict__ gradx, const float *__restrict__ grady, const int W, const int H,
                       const RefineJob *__restrict__ jobs, const int n_jobs, double *__restrict__ out, int *__restrict__ iters)
{
    __shared__ double sS[REFINE_WARPS][LBFGS_RANK][5], sY[REFINE_WARPS][LBFGS_RANK][5], sSY[REFINE_WARPS][LBFGS_RANK];
    const int wib = threadIdx.x >> 5, job = blockIdx.x * REFINE_WARPS + wib;
    if (job >= n_jobs) return;
    const RefineJob J = jobs[job];
    Corner C;
    C.gu = gradx + (size_t)J.img * W * H; C.gv = grady + (size_t)J.img * W * H;
    C.W = W; C.H = H; C.lane = threadIdx.x & 31;
    C.prior_u = J.prior[0]; C.prior_v = J.prior[1]; C.step_len = J.reach / STEPS;
    double (*S)[5] = sS[wib], (*Y)[5] = sY[wib], *SY = sSY[wib];

    const double F_TOL = 1e-6, G_TOL = 1e-10, X_TOL = 1e-8, MIN_STEP = 1e-9, C1 = 1e-4, C2 = 0.9, EXPAND = 10.0;
    const int MAX_IT = 50, MAX_TRIALS = 20, MAX_RESTARTS = 5;
    double x[5], g[5], dir[5], xt[5], gt[5];
#pragma unroll
    for (int i = 0; i < 5; i++) x[i] = J.x[i];
    double f, f_prev = 0;
    evaluate(C, x, f, g);
    auto inf_norm = [](const double *a) { double m = 0; for (int i = 0; i < 5; i++) m = fmax(m, fabs(a[i])); return m; };
    auto dot5 = [](const double *a, const double *b) { double s = 0; for (int i = 0; i < 5; i++) s += a[i] * b[i]; return s; };
    int it = 0, restarts = 0, hist = 0, head = 0;        // history: `hist` pairs, the oldest at `head`
    if (inf_norm(g) > G_TOL)
        for (;;) {
            if (it >= MAX_IT) break;
            it++;
            // two-loop recursion
            bool steepest = hist == 0;
            for (int i = 0; i < 5; i++) dir[i] = g[i];
            if (!steepest) {
                double alpha[LBFGS_RANK];
                for (int k = hist - 1; k >= 0; k--) {
                    const int q = (head + k) % LBFGS_RANK;
                    alpha[k] = dot5(S[q], dir) / SY[q];
                    for (int i = 0; i < 5; i++) dir[i] -= alpha[k] * Y[q][i];
                }
                for (int k = 0; k < hist; k++) {
                    const int q = (head + k) % LBFGS_RANK;
                    const double beta = dot5(Y[q], dir) / SY[q];
                    for (int i = 0; i < 5; i++) dir[i] += S[q][i] * (alpha[k] - beta);
                }
            }
            for (int i = 0; i < 5; i++) dir[i] = -dir[i];
            double slope = dot5(g, dir);
            if (!steepest && slope >= 0.0) {
                if (++restarts > MAX_RESTARTS) break;
                hist = 0; head = 0;
                for (int i = 0; i < 5; i++) dir[i] = -g[i];
                slope = dot5(g, dir);
                steepest = true;
            }
            const double step0 = (it == 1 || steepest) ? fmin(1.0, 1.0 / inf_norm(g)) : fmin(1.0, 2.0 * (f - f_prev) / slope);
            if (!(step0 > 0.0)) break;
            const double dmax = inf_norm(dir);
            auto probe = [&](const double a) {
                for (int i = 0; i < 5; i++) xt[i] = x[i] + a * dir[i];
                Trial t;
                t.x = a;
                evaluate(C, xt, t.f, gt);
                t.g = dot5(gt, dir);
                return t;
            };
            const Trial start = {0.0, f, slope};
            Trial prev = start, cur = probe(step0), lo = start, hi = start, sol = start;
            bool zoom = false, ok = true;
            int trials = 0;
            for (;;) {          // bracketing
                trials++;
                if (cur.f > start.f + C1 * start.g * cur.x || (prev.x > 0.0 && cur.f > prev.f)) { zoom = true; lo = prev; hi = cur; break; }
                if (fabs(cur.g) <= -C2 * start.g) { lo = hi = cur; break; }
                if (cur.g >= 0.0) { zoom = true; lo = cur; hi = prev; break; }
                if (fabs(cur.x - prev.x) * dmax < MIN_STEP) { ok = false; break; }
                if (trials >= MAX_TRIALS) { if (cur.f < lo.f) lo = cur; break; }
                const double a = cubic_minimiser(prev, cur, cur.x, cur.x * EXPAND);
                if (a * dmax < MIN_STEP) { ok = false; break; }
                prev = cur;
                cur = probe(a);
            }
            if (!ok) break;
            Trial best = lo;
            if (zoom) {
                if (lo.f > hi.f) { const Trial t = lo; lo = hi; hi = t; }
                bool have = false;
                for (;;) {
                    if (trials >= MAX_TRIALS) break;
                    if (fabs(hi.x - lo.x) * dmax < MIN_STEP) break;
                    trials++;
                    const bool up = lo.x < hi.x;
                    const double a = cubic_minimiser(up ? lo : hi, up ? hi : lo, up ? lo.x : hi.x, up ? hi.x : lo.x);
                    sol = probe(a);
                    have = true;
                    if (sol.f > start.f + C1 * start.g * sol.x || sol.f >= lo.f) { hi = sol; continue; }
                    if (fabs(sol.g) <= -C2 * start.g) break;
                    if (sol.g * (hi.x - lo.x) >= 0.0) hi = lo;
                    lo = sol;
                }
                best = (!have || sol.f > lo.f) ? lo : sol;
            }
            if (!(best.x > 0.0)) break;
            // the step
            double fn, s[5], y[5], step2 = 0, x2 = 0;
            for (int i = 0; i < 5; i++) xt[i] = x[i] + best.x * dir[i];
            evaluate(C, xt, fn, gt);
            for (int i = 0; i < 5; i++) { s[i] = best.x * dir[i]; y[i] = gt[i] - g[i]; step2 += s[i] * s[i]; x2 += xt[i] * xt[i]; }
            const double sy = dot5(s, y);
            if (sy > 1e-14) {
                int q;
                if (hist == LBFGS_RANK) { q = head; head = (head + 1) % LBFGS_RANK; }
                else q = (head + hist++) % LBFGS_RANK;
                __syncwarp();
                if (C.lane == 0) {
                    for (int i = 0; i < 5; i++) { S[q][i] = s[i]; Y[q][i] = y[i]; }
                    SY[q] = sy;
                }
                __syncwarp();
            }
            f_prev = f; f = fn;
            for (int i = 0; i < 5; i++) { x[i] = xt[i]; g[i] = gt[i]; }
            if (inf_norm(g) <= G_TOL) break;
            if (sqrt(step2) <= X_TOL * (sqrt(x2) + X_TOL)) break;
            if (fabs(f_prev - f) <= F_TOL * fabs(f_prev)) break;
        }
    if (C.lane == 0) {
        out[2 * job] = x[0];
        out[2 * job + 1] = x[1];
        if (iters) iters[job] = it;
    }
}

// SubpixelCorner::Evaluate alone, for the parity test of the cost and its gradient: one warp per parameter vector
__global__ void __launch_bounds__(32 * REFINE_WARPS)
subpixel_evaluate_kernel(const float *__restrict__ gradx, const float *__restrict__ grady, const int W, const int H,
                         const RefineJob *__restrict__ jobs, const int n_jobs, double *__restrict__ cost, double *__restrict__ grad)
{
    const int job = blockIdx.x * REFINE_WARPS + (threadIdx.x >> 5);
    if (job >= n_jobs) return;
    const RefineJob J = jobs[job];
    Corner C;
    C.gu = gradx + (size_t)J.img * W * H; C.gv = grady + (size_t)J.img * W * H;
    C.W = W; C.H = H; C.lane = threadIdx.x & 31;
    C.prior_u = J.prior[0]; C.prior_v = J.prior[1]; C.step_len = J.reach / STEPS;
    double f, g[5];
    evaluate(C, J.x, f, g);
    if (C.lane == 0) {
        cost[job] = f;
        for (int i = 0; i < 5; i++) grad[5 * job + i] = g[i];
    }
}

// ---- pipeline ---------------------------------------------------------------------------------------------------------
// Images go through in passes of a few images.  A pass = upload, response + scan on the GPU, the blurred images and the
// maxima back, the host stages (one image per thread), the refinement of the grids found (while that pass's gradient
// maps are still on the device), results out.  DEFAULT_LANES passes are in flight at once, each on its own stream and buffers
// and driven by its own host thread, so the copies and kernels of one pass run under the host stages of the others.
// Buffers are kept between calls (grow only).
constexpr int MAX_LANES = 6;
constexpr int DEFAULT_LANES = 4;        // VG_DETECT_LANES overrides (developer knob)

struct Lane {
    unsigned char *img = nullptr, *s1 = nullptr, *s2 = nullptr;
    float *resp = nullptr, *gradx = nullptr, *grady = nullptr;
    double *avg = nullptr, *refined = nullptr;
    unsigned int *count = nullptr;
    det::Maximum *maxima = nullptr;
    RefineJob *jobs = nullptr;
    void *work = nullptr;                // the response kernel's partial sums
    size_t work_bytes = 0;
    unsigned char *h_img = nullptr, *h_s1 = nullptr, *h_s2 = nullptr;     // pinned
    det::Maximum *h_maxima = nullptr;                                     // pinned
    unsigned int *h_count = nullptr;                                      // pinned
    RefineJob *h_jobs = nullptr;                                          // pinned
    double *h_refined = nullptr;                                          // pinned
    cudaStream_t st = nullptr;
    size_t pixels = 0, corners = 0;      // capacity: pixels of a pass, corners of a pass
    int images = 0, cap = 0, dev = -1;   // images of a pass, maxima kept per image

    void release()
    {
        cudaFree(img); cudaFree(s1); cudaFree(s2); cudaFree(resp); cudaFree(gradx); cudaFree(grady); cudaFree(avg);
        cudaFree(refined); cudaFree(count); cudaFree(maxima); cudaFree(jobs); cudaFree(work);
        work = nullptr; work_bytes = 0;
        cudaFreeHost(h_img); cudaFreeHost(h_s1); cudaFreeHost(h_s2); cudaFreeHost(h_maxima); cudaFreeHost(h_count);
        cudaFreeHost(h_jobs); cudaFreeHost(h_refined);
        img = s1 = s2 = h_img = h_s1 = h_s2 = nullptr; resp = gradx = grady = nullptr; avg = refined = h_refined = nullptr;
        count = h_count = nullptr; maxima = h_maxima = nullptr; jobs = h_jobs = nullptr;
        pixels = corners = 0; images = cap = 0;
    }
    cudaError_t reserve_maxima(const int images_, const int cap_)
    {
        cudaFree(maxima); cudaFreeHost(h_maxima);
        maxima = h_maxima = nullptr;
        cudaError_t e = cudaMalloc(&maxima, sizeof(det::Maximum) * (size_t)cap_ * images_);
        if (e == cudaSuccess) e = cudaMallocHost(&h_maxima, sizeof(det::Maximum) * (size_t)cap_ * images_);
        cap = cap_;
        return e;
    }
    cudaError_t reserve(const int device, const int W, const int H, const int images_, const int P, const int cap_)
    {
        const size_t N = (size_t)W * H, wb = corner_response_work_bytes(images_, W, H);
        if (dev == device && pixels >= N * images_ && images >= images_ && corners >= (size_t)P * images_ && cap >= cap_ && st &&
            work_bytes >= wb)
            return cudaSuccess;
        if (st) cudaStreamSynchronize(st);
        release();
        if (dev != device && st) {                   // the stream and the buffers belong to the device they were made on
            int cur = device;
            cudaGetDevice(&cur);
            if (dev >= 0) cudaSetDevice(dev);
            cudaStreamDestroy(st);
            st = nullptr;
            cudaSetDevice(cur);
        }
        dev = device;
        cudaError_t e = st ? cudaSuccess : cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        const size_t px = N * images_, nc = (size_t)P * images_;
        if (e == cudaSuccess) e = cudaMalloc(&img, px);
        if (e == cudaSuccess) e = cudaMalloc(&s1, px);
        if (e == cudaSuccess) e = cudaMalloc(&s2, px);
        if (e == cudaSuccess) e = cudaMalloc(&resp, px * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&gradx, px * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&grady, px * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&avg, sizeof(double) * images_);
        if (e == cudaSuccess) e = cudaMalloc(&count, sizeof(unsigned int) * images_);
        if (e == cudaSuccess) e = cudaMalloc(&jobs, sizeof(RefineJob) * nc);
        if (e == cudaSuccess) e = cudaMalloc(&refined, sizeof(double) * 2 * nc);
        if (e == cudaSuccess) e = cudaMalloc(&work, wb);
        if (e == cudaSuccess) work_bytes = wb;
        if (e == cudaSuccess) e = cudaMallocHost(&h_img, px);
        if (e == cudaSuccess) e = cudaMallocHost(&h_s1, px);
        if (e == cudaSuccess) e = cudaMallocHost(&h_s2, px);
        if (e == cudaSuccess) e = cudaMallocHost(&h_count, sizeof(unsigned int) * images_);
        if (e == cudaSuccess) e = cudaMallocHost(&h_jobs, sizeof(RefineJob) * nc);
        if (e == cudaSuccess) e = cudaMallocHost(&h_refined, sizeof(double) * 2 * nc);
        if (e == cudaSuccess) e = reserve_maxima(images_, cap_);
        if (e == cudaSuccess) { pixels = px; corners = nc; images = images_; }
        return e;
    }
};

struct Pipeline {
    std::mutex mu;                      // one vg_detect_pattern at a time per process
    Lane lane[MAX_LANES];
};
Pipeline &pipeline() { static Pipeline p; return p; }

// The host threads of the pipeline: one pool for all lanes, alive for the process, so that the images of whichever
// pass is ready keep every core busy (a pool per pass would idle at each pass's slowest image and pay its threads'
// start-up sixteen times per call).
class Workers {
public:
    static Workers &get() { static Workers w; return w; }
    // fn(0) .. fn(n - 1) on the pool; returns when all are done.  The caller helps.
    template <typename Fn> void run(const int n, Fn fn)
    {
        if (n <= 0) return;
        Batch b;
        b.fn = [&](int i) { fn(i); };
        b.n = n;
        std::unique_lock<std::mutex> g(mu_);
        ensure_threads();
        queue_.push_back(&b);
        cv_.notify_all();
        process(&b, g);
        b.done_cv.wait(g, [&] { return b.finished == b.n; });
    }

private:
    struct Batch {
        std::function<void(int)> fn;
        int n = 0, next = 0, finished = 0;           // guarded by mu_
        std::condition_variable done_cv;
    };
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Batch *> queue_;                      // batches with unclaimed items
    std::vector<std::thread> threads_;
    bool stop_ = false;

    void ensure_threads()
    {
        if (!threads_.empty()) return;
        const int hw = std::max(2, (int)std::thread::hardware_concurrency());
        for (int t = 0; t < hw - 1; t++) threads_.emplace_back([this] { loop(); });
    }
    // Claim and run items of b until none is left to claim.  mu_ is held on entry and on exit.  A batch lives on its
    // owner's stack until finished == n, so it is only touched while that cannot be the case: under the lock, and
    // never again after the increment that completes it.
    void process(Batch *b, std::unique_lock<std::mutex> &g)
    {
        while (b->next < b->n) {
            const int i = b->next++;
            if (b->next == b->n) queue_.erase(std::find(queue_.begin(), queue_.end(), b));
            g.unlock();
            b->fn(i);
            g.lock();
            if (++b->finished == b->n) { b->done_cv.notify_all(); return; }
        }
    }
    void loop()
    {
        std::unique_lock<std::mutex> g(mu_);
        for (;;) {
            cv_.wait(g, [&] { return stop_ || !queue_.empty(); });
            if (stop_) return;
            process(queue_.front(), g);
        }
    }
    Workers() {}
    ~Workers()
    {
        { std::lock_guard<std::mutex> g(mu_); stop_ = true; }
        cv_.notify_all();
        for (std::thread &t : threads_) t.join();
    }
};

struct Request {
    const unsigned char *img;
    int width, height, Nx, Ny, improve;
    double *corners;
    unsigned char *found;
    int host_threads;
};

// VG_DETECT_TRACE=1: seconds spent per stage, summed over the passes of all lanes (they overlap, so the sum exceeds the
// wall time), printed at the end of each call
struct StageClock {
    std::atomic<long long> ns[4];
    StageClock() { for (auto &x : ns) x = 0; }
};
inline long long now_ns()
{
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

#define DET_CUDA(call)                                                                                      \
    do {                                                                                                    \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess) return std::string("CUDA error in " #call ": ") + cudaGetErrorString(e__);  \
    } while (0)

// one pass: the images `ids` (indices into the request) at one scale; the ones without a pattern are appended to `still`
std::string run_pass(Lane &B, const Request &q, const int *ids, const int np, const double sigma2, std::vector<int> &still,
                     StageClock &clk)
{
    long long t0 = now_ns();
    auto lap = [&](const int stage) { const long long t = now_ns(); clk.ns[stage] += t - t0; t0 = t; };
    const int W = q.width, H = q.height, P = q.Nx * q.Ny;
    const size_t N = (size_t)W * H;
    const int R = (int)std::round(1.5 * sigma2);                    // INIT_RADIUS (:231)
    cudaStream_t st = B.st;
    // staged through pinned memory: the caller's pageable pages would serialise the lanes inside the driver
    Workers::get().run(np, [&](const int i) { std::memcpy(B.h_img + (size_t)i * N, q.img + (size_t)ids[i] * N, N); });
    lap(0);
    DET_CUDA(cudaMemcpyAsync(B.img, B.h_img, N * np, cudaMemcpyHostToDevice, st));
    if (corner_response_launch(B.img, np, W, H, 0.7, sigma2, B.resp, B.gradx, B.grady, nullptr, B.s1, B.s2, B.avg, nullptr, st,
                               B.work, B.work_bytes))
        return vg_last_error();
    DET_CUDA(cudaMemcpyAsync(B.h_s1, B.s1, N * np, cudaMemcpyDeviceToHost, st));
    DET_CUDA(cudaMemcpyAsync(B.h_s2, B.s2, N * np, cudaMemcpyDeviceToHost, st));
    for (;;) {
        DET_CUDA(cudaMemsetAsync(B.count, 0, sizeof(unsigned int) * np, st));
        const dim3 grid((W + 31) / 32, (H + 7) / 8, np);
        local_maxima_kernel<<<grid, 256, 0, st>>>(B.resp, B.avg, W, H, R, B.maxima, B.count, B.cap);
        count_launch(&launch_counter());
        DET_CUDA(cudaGetLastError());
        DET_CUDA(cudaMemcpyAsync(B.h_count, B.count, sizeof(unsigned int) * np, cudaMemcpyDeviceToHost, st));
        DET_CUDA(cudaStreamSynchronize(st));
        const unsigned int most = *std::max_element(B.h_count, B.h_count + np);
        if (most <= (unsigned)B.cap) break;
        DET_CUDA(B.reserve_maxima(B.images, (int)most));           // a noisy image: room for the longest list, again
    }
    for (int i = 0; i < np; i++)
        if (B.h_count[i])
            DET_CUDA(cudaMemcpyAsync(B.h_maxima + (size_t)i * B.cap, B.maxima + (size_t)i * B.cap,
                                     sizeof(det::Maximum) * B.h_count[i], cudaMemcpyDeviceToHost, st));
    DET_CUDA(cudaStreamSynchronize(st));
    lap(1);
    // host stages, one image per thread
    std::vector<std::vector<det::Pt>> grids(np);
    std::vector<unsigned char> job_ok((size_t)np * P, 0);
    std::vector<RefineJob> jobs((size_t)np * P);
    Workers::get().run(np, [&](const int i) {
        const det::Frame F{B.h_img + (size_t)i * N, B.h_s1 + (size_t)i * N, B.h_s2 + (size_t)i * N, W, H};
        std::vector<det::Maximum> mx(B.h_maxima + (size_t)i * B.cap, B.h_maxima + (size_t)i * B.cap + B.h_count[i]);
        grids[i] = det::detect_at_scale(F, mx, q.Nx, q.Ny, R);
        if ((int)grids[i].size() != P || !q.improve) return;
        std::vector<double> reach(P);
        det::refinement_reach(grids[i], q.Nx, reach.data());
        for (int k = 0; k < P; k++) {
            RefineJob &J = jobs[(size_t)i * P + k];
            J.prior[0] = grids[i][k].u; J.prior[1] = grids[i][k].v;
            J.reach = reach[k];
            J.img = i; J.pad = 0;
            job_ok[(size_t)i * P + k] = det::init_point(F, grids[i][k], R, J.x) ? 1 : 0;
        }
    });
    lap(2);
    // results; refinement of the grids found, while this pass's gradient maps are on the device
    std::vector<size_t> where;                                      // slot in q.corners of each packed job
    int nj = 0;
    for (int i = 0; i < np; i++) {
        if ((int)grids[i].size() != P) { still.push_back(ids[i]); continue; }
        q.found[ids[i]] = 1;
        for (int k = 0; k < P; k++) {
            const size_t slot = ((size_t)ids[i] * P + k) * 2;
            q.corners[slot] = grids[i][k].u; q.corners[slot + 1] = grids[i][k].v;
            if (q.improve && job_ok[(size_t)i * P + k]) { B.h_jobs[nj++] = jobs[(size_t)i * P + k]; where.push_back(slot); }
        }
    }
    if (nj) {
        DET_CUDA(cudaMemcpyAsync(B.jobs, B.h_jobs, sizeof(RefineJob) * nj, cudaMemcpyHostToDevice, st));
        subpixel_refine_kernel<<<(nj + REFINE_WARPS - 1) / REFINE_WARPS, 32 * REFINE_WARPS, 0, st>>>(B.gradx, B.grady, W, H, B.jobs, nj,
                                                                                                     B.refined, nullptr);
        count_launch(&launch_counter());
        DET_CUDA(cudaGetLastError());
        DET_CUDA(cudaMemcpyAsync(B.h_refined, B.refined, sizeof(double) * 2 * nj, cudaMemcpyDeviceToHost, st));
        DET_CUDA(cudaStreamSynchronize(st));
        for (int j = 0; j < nj; j++) { q.corners[where[j]] = B.h_refined[2 * j]; q.corners[where[j] + 1] = B.h_refined[2 * j + 1]; }
    }
    lap(3);
    return std::string();
}

}  // namespace
}  // namespace vg

using namespace vg;

extern "C" {

int vg_detect_pattern(const unsigned char *img, int n_img, int width, int height, int Nx, int Ny, int improve, double *corners,
                      unsigned char *found)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(VG_ERR_CUDA, "no CUDA device: this engine has no CPU path");
    }
    if (n_img < 0 || width < 8 || height < 8 || Nx < 2 || Ny < 2) return fail(VG_ERR_INVALID, "vg_detect_pattern: bad size");
    if (n_img == 0) return VG_OK;
    if (!img || !corners || !found) return fail(VG_ERR_INVALID, "null argument");
    const size_t N = (size_t)width * height;
    const int P = Nx * Ny;
    // images per pass: about 8 MB of pixels (8 images of 1280 x 800); VG_DETECT_CHUNK overrides (developer knob, tests)
    size_t per_pass = std::max<size_t>(1, ((size_t)8 << 20) / N);
    if (const char *e = std::getenv("VG_DETECT_CHUNK")) per_pass = (size_t)std::max(1, std::atoi(e));
    const int chunk = (int)std::min<size_t>((size_t)n_img, per_pass);
    const int cap = (int)std::min<size_t>(N / 4 + 16, 1u << 15);    // maxima kept per image; grown on overflow
    const int hw = std::max(1, (int)std::thread::hardware_concurrency());
    int dev = 0;
    VG_CUDA(cudaGetDevice(&dev));
    Pipeline &pl = pipeline();
    std::lock_guard<std::mutex> lock(pl.mu);
    int want_lanes = DEFAULT_LANES;
    if (const char *e = std::getenv("VG_DETECT_LANES")) want_lanes = std::max(1, std::min(MAX_LANES, std::atoi(e)));
    const int lanes = std::max(1, std::min(want_lanes, (n_img + chunk - 1) / chunk));
    for (int l = 0; l < lanes; l++) VG_CUDA(pl.lane[l].reserve(dev, width, height, chunk, P, cap));
    const Request q{img, width, height, Nx, Ny, improve, corners, found, hw};
    std::vector<int> pending(n_img);
    for (int i = 0; i < n_img; i++) { pending[i] = i; found[i] = 0; }
    static const double SIGMA[3] = {1.4, 2, 1};                     // detectPattern's scales (:225)
    StageClock clk;
    const long long t_call = now_ns();
    for (int scale = 0; scale < 3 && !pending.empty(); scale++) {
        const int np = (int)pending.size(), passes = (np + chunk - 1) / chunk;
        std::atomic<int> next(0);
        std::mutex out_mu;
        std::vector<int> still;
        std::string error;
        auto drive = [&](const int l) {
            cudaSetDevice(dev);
            for (int b = next.fetch_add(1); b < passes; b = next.fetch_add(1)) {
                std::vector<int> mine;
                const std::string err = run_pass(pl.lane[l], q, pending.data() + (size_t)b * chunk, std::min(chunk, np - b * chunk),
                                                 SIGMA[scale], mine, clk);
                std::lock_guard<std::mutex> g(out_mu);
                if (!err.empty() && error.empty()) error = err;
                still.insert(still.end(), mine.begin(), mine.end());
                if (!err.empty()) return;
            }
        };
        const int nl = std::min(lanes, passes);
        if (nl == 1) drive(0);
        else {
            std::vector<std::thread> drivers;
            for (int l = 0; l < nl; l++) drivers.emplace_back(drive, l);
            for (auto &t : drivers) t.join();
        }
        if (!error.empty()) return fail(VG_ERR_CUDA, error);
        std::sort(still.begin(), still.end());
        pending.swap(still);
    }
    if (const char *e = std::getenv("VG_DETECT_TRACE"))
        if (e[0] == '1')
            std::fprintf(stderr, "[vg detect] %d images, %d per pass, %d lanes: wall %.2f ms; summed over passes: staging %.2f, "
                         "upload + response + scan + download %.2f, host stages %.2f, refinement %.2f ms\n", n_img, chunk, lanes,
                         (now_ns() - t_call) * 1e-6, clk.ns[0] * 1e-6, clk.ns[1] * 1e-6, clk.ns[2] * 1e-6, clk.ns[3] * 1e-6);
    return VG_OK;
}

// The host stages of one scale on their own (no GPU involved): from the image, its two blurred copies and the list of
// local maxima of the response to the candidates (in graph order), the grid and the refinement's start values.
int vg_detector_host_stages(const unsigned char *img, const unsigned char *s1, const unsigned char *s2, int width, int height,
                            const float *max_val, const int *max_uv, int n_max, int Nx, int Ny, int init_radius, int *cand,
                            int cand_cap, int *n_cand, int *grid, double *start, double *reach)
{
    if (!img || !s1 || !s2 || width < 8 || height < 8 || n_max < 0 || Nx < 2 || Ny < 2 || init_radius < 1 ||
        init_radius > det::CircleTable::RMAX / 2 - 1)
        return fail(VG_ERR_INVALID, "vg_detector_host_stages: bad argument");
    const det::Frame F{img, s1, s2, width, height};
    std::vector<det::Maximum> mx(n_max);
    for (int i = 0; i < n_max; i++) mx[i] = det::Maximum{max_val[i], max_uv[2 * i], max_uv[2 * i + 1]};
    std::sort(mx.begin(), mx.end(), [](const det::Maximum &a, const det::Maximum &b) { return a.v != b.v ? a.v < b.v : a.u < b.u; });
    const std::vector<det::Pt> c = det::select_candidates(F, mx, Nx, Ny, init_radius);
    if (n_cand) *n_cand = (int)c.size();
    if (cand)
        for (int i = 0; i < (int)c.size() && i < cand_cap; i++) { cand[2 * i] = c[i].u; cand[2 * i + 1] = c[i].v; }
    const std::vector<det::Pt> g = det::detect_at_scale(F, mx, Nx, Ny, init_radius);
    if ((int)g.size() != Nx * Ny) return 0;
    for (int k = 0; k < Nx * Ny; k++) {
        if (grid) { grid[2 * k] = g[k].u; grid[2 * k + 1] = g[k].v; }
        if (start && !det::init_point(F, g[k], init_radius, start + 5 * k)) return fail(VG_ERR_INVALID, "corner without transitions");
    }
    if (reach) det::refinement_reach(g, Nx, reach);
    return 1;
}

// SubpixelCorner(gradu, gradv, prior, 7, length).Evaluate(params) for n parameter vectors on one pair of gradient maps
// (host buffers): cost[n], gradient[5 n]
int vg_subpixel_evaluate(const float *gradx, const float *grady, int width, int height, int n, const double *prior,
                         const double *length, const double *params, double *cost, double *gradient)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(VG_ERR_CUDA, "no CUDA device: this engine has no CPU path");
    }
    if (n < 0 || width < 1 || height < 1) return fail(VG_ERR_INVALID, "vg_subpixel_evaluate: bad size");
    if (n == 0) return VG_OK;
    if (!gradx || !grady || !prior || !length || !params || !cost || !gradient) return fail(VG_ERR_INVALID, "null argument");
    const size_t N = (size_t)width * height;
    std::vector<RefineJob> jobs(n);
    for (int i = 0; i < n; i++) {
        jobs[i].prior[0] = prior[2 * i]; jobs[i].prior[1] = prior[2 * i + 1];
        for (int k = 0; k < 5; k++) jobs[i].x[k] = params[5 * i + k];
        jobs[i].reach = length[i]; jobs[i].img = 0; jobs[i].pad = 0;
    }
    float *d_g = nullptr;
    RefineJob *d_jobs = nullptr;
    double *d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_g, 2 * N * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_jobs, sizeof(RefineJob) * n);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(double) * 6 * n);
    if (e == cudaSuccess) e = cudaMemcpy(d_g, gradx, N * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_g + N, grady, N * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_jobs, jobs.data(), sizeof(RefineJob) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        subpixel_evaluate_kernel<<<(n + REFINE_WARPS - 1) / REFINE_WARPS, 32 * REFINE_WARPS>>>(d_g, d_g + N, width, height, d_jobs, n,
                                                                                              d_out, d_out + n);
        count_launch(&launch_counter());
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(cost, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(gradient, d_out + n, sizeof(double) * 5 * n, cudaMemcpyDeviceToHost);
    cudaFree(d_g); cudaFree(d_jobs); cudaFree(d_out);
    return e == cudaSuccess ? VG_OK : fail_cuda(e, "vg_subpixel_evaluate");
}

// the refinement alone (improveCorners' second loop): n corners on one pair of gradient maps, start values given
int vg_subpixel_refine(const float *gradx, const float *grady, int width, int height, int n, const double *prior,
                       const double *length, const double *start, double *refined, int *iterations)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(VG_ERR_CUDA, "no CUDA device: this engine has no CPU path");
    }
    if (n < 0 || width < 1 || height < 1) return fail(VG_ERR_INVALID, "vg_subpixel_refine: bad size");
    if (n == 0) return VG_OK;
    if (!gradx || !grady || !prior || !length || !start || !refined) return fail(VG_ERR_INVALID, "null argument");
    const size_t N = (size_t)width * height;
    std::vector<RefineJob> jobs(n);
    for (int i = 0; i < n; i++) {
        jobs[i].prior[0] = prior[2 * i]; jobs[i].prior[1] = prior[2 * i + 1];
        for (int k = 0; k < 5; k++) jobs[i].x[k] = start[5 * i + k];
        jobs[i].reach = length[i]; jobs[i].img = 0; jobs[i].pad = 0;
    }
    float *d_g = nullptr;
    RefineJob *d_jobs = nullptr;
    double *d_out = nullptr;
    int *d_it = nullptr;
    cudaError_t e = cudaMalloc(&d_g, 2 * N * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_jobs, sizeof(RefineJob) * n);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(double) * 2 * n);
    if (e == cudaSuccess) e = cudaMalloc(&d_it, sizeof(int) * n);
    if (e == cudaSuccess) e = cudaMemcpy(d_g, gradx, N * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_g + N, grady, N * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_jobs, jobs.data(), sizeof(RefineJob) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        subpixel_refine_kernel<<<(n + REFINE_WARPS - 1) / REFINE_WARPS, 32 * REFINE_WARPS>>>(d_g, d_g + N, width, height, d_jobs, n,
                                                                                            d_out, d_it);
        count_launch(&launch_counter());
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(refined, d_out, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && iterations) e = cudaMemcpy(iterations, d_it, sizeof(int) * n, cudaMemcpyDeviceToHost);
    cudaFree(d_g); cudaFree(d_jobs); cudaFree(d_out); cudaFree(d_it);
    return e == cudaSuccess ? VG_OK : fail_cuda(e, "vg_subpixel_refine");
}

}  // extern "C"
