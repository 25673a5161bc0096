// ref_entry_detector.cpp -- C entry points around the REFERENCE's own checkerboard detector (TEST INFRASTRUCTURE ONLY).
// Compiled by `make -C oracle ref` together with $(REFERENCE)/src/calibration/corner_detector.cpp against oracle/shim
// (OpenCV / Ceres / Eigen stand-ins); outputs oracle/_ref/libvisgeom_refdet.so (-O2) and libvisgeom_refdet_O0.so.
// No reference source is copied into this repository.
//
// Why two builds: CornerDetector::initPoin (corner_detector.cpp:1261-1298) is declared `int` and has no return
// statement; g++ >= 8 at -O1 and above ends such a function in an unreachable trap, so the -O2 build must not call it.
// The -O0 build runs detectPattern exactly as written (improveCorners included).  The -O2 build runs detectPattern
// with IMPROVE_DETECTION off and, when refinement is asked for, walks improveCorners' loop (:162-198) here with the
// reference's own SubpixelCorner / getTransitions and the five assignments of initPoin written out;
// tests/test_detector_oracle.py holds the two builds against each other.
#define private public
#include "calibration/corner_detector.h"
#undef private

#include <cstring>

namespace {
void load(CornerDetector &det, const uint8_t *img, int w, int h)
{
    Mat8u m(h, w);
    std::memcpy(m.data, img, (size_t)w * h);
    det.setImage(m);
}

// initPoin's body (:1265-1287) without the missing return
void init_point(CornerDetector &det, const Vector2i &pt, double *data)
{
    Vector2iVec tr = det.getTransitions(pt);
    const Vector2i &A = tr[0], &C = tr[1], &B = tr[2], &D = tr[3];
    Matrix2d M;
    Vector2d b;
    M(0, 0) = A[1] - C[1];
    M(0, 1) = C[0] - A[0];
    M(1, 0) = B[1] - D[1];
    M(1, 1) = D[0] - B[0];
    b[0] = A[1] * M(0, 1) + A[0] * M(0, 0);
    b[1] = B[1] * M(1, 1) + B[0] * M(1, 0);
    Vector2d E = M.inverse() * b;
    data[0] = E[0];
    data[1] = E[1];
    data[2] = atan2(A[1] - C[1], A[0] - C[0]);
    data[3] = atan2(B[1] - D[1], B[0] - D[0]);
    data[4] = 0;
}

// improveCorners (:162-198) with init_point in place of initPoin; init5 (optional) receives the five start values
void improve(CornerDetector &det, Vector2dVec &pointVec, double *init5, int *iters)
{
    const int Nx = det._Nx;
    vector<double> radVec;
    for (int i = 0; i < (int)pointVec.size(); i++) {
        double radMax = 7;
        if (i > Nx) radMax = min(radMax, (pointVec[i] - pointVec[i - Nx]).norm() * 0.7);
        else radMax = min(radMax, (pointVec[i] - pointVec[i + Nx]).norm() * 0.7);
        if (i > 0) radMax = min(radMax, (pointVec[i] - pointVec[i - 1]).norm() * 0.7);
        else radMax = min(radMax, (pointVec[i] - pointVec[i + 1]).norm() * 0.7);
        radVec.push_back(radMax);
    }
    for (int i = 0; i < (int)pointVec.size(); i++) {
        Vector2d x = pointVec[i];
        double d[5];
        init_point(det, round(x), d);
        if (init5) std::memcpy(init5 + 5 * i, d, sizeof(d));
        ceres::GradientProblem problem(new SubpixelCorner(det._gradx, det._grady, x, 7, radVec[i]));
        ceres::GradientProblemSolver::Options options;
        ceres::GradientProblemSolver::Summary summary;
        ceres::Solve(options, problem, d, &summary);
        if (iters) iters[i] = summary.num_iterations;
        pointVec[i][0] = d[0];
        pointVec[i][1] = d[1];
    }
}
}  // namespace

extern "C" {

// CornerDetector(Nx, Ny, 3, improve).detectPattern on one image (unified_calibration.cpp:995,1031-1033).
// own_improve != 0: detection with IMPROVE_DETECTION off, then the loop above (the only route of the -O2 build).
// corners: Nx * Ny * 2 doubles.  Returns 1 when the pattern was found.
int vgref_detect_pattern(const uint8_t *img, int w, int h, int Nx, int Ny, int do_improve, int own_improve, double *corners,
                         double *init5, int *iters)
{
    CornerDetector det(Nx, Ny, 3, do_improve && !own_improve);
    load(det, img, w, h);
    Vector2dVec pts;
    if (!det.detectPattern(pts)) return 0;
    if (do_improve && own_improve) improve(det, pts, init5, iters);
    for (size_t i = 0; i < pts.size(); i++) { corners[2 * i] = pts[i][0]; corners[2 * i + 1] = pts[i][1]; }
    return 1;
}

// the stages of one scale of detectPattern (:231-240), for staged comparisons: candidates in _ptVec order, the arcs of
// the graph with their signs, the selected pattern (indices into the candidates)
int vgref_detector_stages(const uint8_t *img, int w, int h, int Nx, int Ny, double sigma2, int cap, int *pts, int *n_arcs,
                          int *arcs, int *arc_sign, int arc_cap, int *pattern, int *n_pattern, double *avg)
{
    CornerDetector det(Nx, Ny, 3, false);
    load(det, img, w, h);
    det.INIT_RADIUS = round(1.5 * sigma2);
    det.computeResponse(0.7, sigma2);
    if (avg) *avg = det._avgVal;
    det.selectCandidates();
    const int n_hyp = (int)det._hypHeap.size();
    *n_pattern = 0;
    if (n_hyp < Nx * Ny) {              // detectPattern does not build the graph then (:236)
        auto heap = det._hypHeap;
        int k = 0;
        while (!heap.empty()) {
            pop_heap(heap.begin(), heap.end(), CornerDetector::comp);
            if (k < cap) { pts[2 * k] = heap.back().second[0]; pts[2 * k + 1] = heap.back().second[1]; }
            heap.pop_back(); k++;
        }
        if (n_arcs) for (int i = 0; i < min(k, cap); i++) n_arcs[i] = 0;
        return n_hyp;
    }
    det.constructGraph();
    const int n = (int)det._ptVec.size();
    int a = 0;
    for (int i = 0; i < n && i < cap; i++) {
        pts[2 * i] = det._ptVec[i][0]; pts[2 * i + 1] = det._ptVec[i][1];
        n_arcs[i] = (int)det._arcVec[i].size();
        for (int j : det._arcVec[i]) {
            if (a < arc_cap) { arcs[a] = j; arc_sign[a] = det._arcSign[make_pair(i, j)]; }
            a++;
        }
    }
    vector<int> idx = det.selectPattern();
    *n_pattern = (int)idx.size();
    for (size_t i = 0; i < idx.size(); i++) pattern[i] = idx[i];
    return n;
}

// SubpixelCorner(gradu, gradv, prior, steps, length).Evaluate (:31-100)
int vgref_subpixel_evaluate(const float *gradx, const float *grady, int w, int h, const double *prior, int steps, double length,
                            const double *params, double *cost, double *gradient)
{
    Mat32f gx(h, w), gy(h, w);
    std::memcpy(gx.data, gradx, sizeof(float) * (size_t)w * h);
    std::memcpy(gy.data, grady, sizeof(float) * (size_t)w * h);
    SubpixelCorner f(gx, gy, Vector2d(prior[0], prior[1]), steps, length);
    return f.Evaluate(params, cost, gradient) ? 1 : 0;
}

// ceres::Solve(GradientProblemSolver::Options(), GradientProblem(SubpixelCorner), params) as improveCorners runs it
int vgref_subpixel_solve(const float *gradx, const float *grady, int w, int h, const double *prior, int steps, double length,
                         double *params, double *final_cost)
{
    Mat32f gx(h, w), gy(h, w);
    std::memcpy(gx.data, gradx, sizeof(float) * (size_t)w * h);
    std::memcpy(gy.data, grady, sizeof(float) * (size_t)w * h);
    ceres::GradientProblem problem(new SubpixelCorner(gx, gy, Vector2d(prior[0], prior[1]), steps, length));
    ceres::GradientProblemSolver::Options options;
    ceres::GradientProblemSolver::Summary summary;
    ceres::Solve(options, problem, params, &summary);
    if (final_cost) *final_cost = summary.final_cost;
    return summary.num_iterations;
}

// the maps computeResponse leaves in the detector (for the product's lazily recomputed gradients)
int vgref_detector_maps(const uint8_t *img, int w, int h, double sigma2, float *resp, float *gradx, float *grady, float *imgrad,
                        uint8_t *src1, uint8_t *src2)
{
    CornerDetector det(9, 6, 3, false);
    load(det, img, w, h);
    // cv::Mat::create leaves new memory unset; the reference never writes the border of _gradx / _grady
    det._gradx.setTo(0); det._grady.setTo(0);
    det.computeResponse(0.7, sigma2);
    const size_t N = (size_t)w * h;
    if (resp) std::memcpy(resp, det._resp.data, N * 4);
    if (gradx) std::memcpy(gradx, det._gradx.data, N * 4);
    if (grady) std::memcpy(grady, det._grady.data, N * 4);
    if (imgrad) std::memcpy(imgrad, det._imgrad.data, N * 4);
    if (src1) std::memcpy(src1, det._src1.data, N);
    if (src2) std::memcpy(src2, det._src2.data, N);
    return 0;
}

}  // extern "C"
