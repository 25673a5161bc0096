// vg_project.cu -- the ICamera point API on the GPU (C ABI: vg_project_points*, vg_reconstruct_points*): what a
// visgeom caller does with a camera object outside the calibration functor --
//   ICamera::projectPoint / projectionJacobian / intrinsicJacobian   generic_camera.h:39-50
//   ICamera::projectPointCloud / reconstructPointCloud               generic_camera.h:64-113
//   EnhancedCamera / UnifiedCamera / MeiCamera                       eucm.h:85-226, ucm.h:81-197, mei.h:90-285
// batched over points: one thread per point (a warp per 32 consecutive points, their structs moved through shared-memory
// tiles so that global memory is touched in whole runs), the same Camera<MODEL>::eval the calibration kernel uses (projection,
// dP/dX and dP/dintrinsics share rho, eta and their reciprocals).  HBM-bound: 24 B in, up to 16 + 48 + 16 K + 1 B out
// per point.  There is no CPU path.
#include "vg_common.h"
#include "vg_math.cuh"

#include <cstdint>

namespace vg {
namespace {

// The arrays are arrays of small structs (3 doubles in; 2, 6 and 2 K doubles out per point): a thread that reads and
// writes its own point's struct touches memory 24 ... 160 bytes apart from its neighbour's, a fraction of every sector
// it moves.  So a warp takes 32 consecutive points, whose structs are one contiguous run of each array, and moves each
// run through a shared-memory tile: the lanes read / write the run element by element (coalesced), each lane picks up
// or drops its own struct in the tile (row pitch odd: no bank conflicts).
constexpr int PP_WARPS = 8;

template <int W>      // doubles per point
__device__ __forceinline__ void tile_to_global(const double *tile, double *g, const long long p0, const int np, const int lane)
{
    for (int e = lane; e < np * W; e += 32) g[p0 * W + e] = tile[(e / W) * (W + 1) + e % W];
}

template <int MODEL>
__global__ void __launch_bounds__(32 * PP_WARPS)
project_points_kernel(const double *__restrict__ intr_g, const long long n, const double *__restrict__ X,
                      double *__restrict__ uv, double *__restrict__ dPdX, double *__restrict__ dPdintr,
                      unsigned char *__restrict__ ok_out)
{
    using CAM = Camera<MODEL>;
    constexpr int K = CAM::K;
    __shared__ double tiles[PP_WARPS][32 * (2 * K + 1)];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double *tile = tiles[wib];
    double intr[K];
#pragma unroll
    for (int i = 0; i < K; i++) intr[i] = __ldg(intr_g + i);
    const typename CAM::Consts cc = CAM::prepare(intr);
    const long long warps = (long long)gridDim.x * PP_WARPS, wid = (long long)blockIdx.x * PP_WARPS + wib;
    for (long long p0 = wid * 32; p0 < n; p0 += warps * 32) {
        const int np = (int)min(32LL, n - p0);
        for (int e = lane; e < np * 3; e += 32) tile[(e / 3) * 4 + e % 3] = X[p0 * 3 + e];
        __syncwarp();
        const double x = tile[lane * 4], y = tile[lane * 4 + 1], z = tile[lane * 4 + 2];
        __syncwarp();
        double u = 0, v = 0, Pu[3], Pv[3], Ju[K], Jv[K];
        bool ok = false;
        if (lane < np) ok = CAM::eval(intr, cc, x, y, z, u, v, Pu, Pv, Ju, Jv);
        const unsigned okmask = __ballot_sync(0xffffffffu, ok);
        if (ok_out && lane < np) ok_out[p0 + lane] = ok ? 1 : 0;
        // a failed projection leaves the image point alone and zero Jacobians (eucm.h:46-54,141-150,198-206)
        if (dPdintr) {
#pragma unroll
            for (int q = 0; q < K; q++) { tile[lane * (2 * K + 1) + q] = ok ? Ju[q] : 0.0; tile[lane * (2 * K + 1) + K + q] = ok ? Jv[q] : 0.0; }
            __syncwarp();
            tile_to_global<2 * K>(tile, dPdintr, p0, np, lane);
            __syncwarp();
        }
        if (dPdX) {
#pragma unroll
            for (int q = 0; q < 3; q++) { tile[lane * 7 + q] = ok ? Pu[q] : 0.0; tile[lane * 7 + 3 + q] = ok ? Pv[q] : 0.0; }
            __syncwarp();
            tile_to_global<6>(tile, dPdX, p0, np, lane);
            __syncwarp();
        }
        if (uv) {
            tile[lane * 3] = u; tile[lane * 3 + 1] = v;
            __syncwarp();
            for (int e = lane; e < np * 2; e += 32)
                if (okmask >> (e >> 1) & 1u) uv[p0 * 2 + e] = tile[(e >> 1) * 3 + (e & 1)];
            __syncwarp();
        }
    }
}

// back-projection: EUCM eucm.h:85-106; UCM ucm.h:81-103 and MEI mei.h:90-112 (the reference ignores the distortion
// terms there) share g = sqrt(1 + u2 (1 - xi^2)).  Same tiling as above.
template <int MODEL>
__global__ void __launch_bounds__(32 * PP_WARPS)
reconstruct_points_kernel(const double *__restrict__ intr_g, const long long n, const double *__restrict__ uv,
                          double *__restrict__ X, unsigned char *__restrict__ ok_out)
{
    constexpr int K = Camera<MODEL>::K;
    __shared__ double tiles[PP_WARPS][32 * 4];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double *tile = tiles[wib];
    double p[K];
#pragma unroll
    for (int i = 0; i < K; i++) p[i] = __ldg(intr_g + i);
    const double fu = p[K - 4], fv = p[K - 3], u0 = p[K - 2], v0 = p[K - 1];
    const long long warps = (long long)gridDim.x * PP_WARPS, wid = (long long)blockIdx.x * PP_WARPS + wib;
    for (long long p0 = wid * 32; p0 < n; p0 += warps * 32) {
        const int np = (int)min(32LL, n - p0);
        for (int e = lane; e < np * 2; e += 32) tile[(e >> 1) * 3 + (e & 1)] = uv[p0 * 2 + e];
        __syncwarp();
        const double xn = (tile[lane * 3] - u0) / fu, yn = (tile[lane * 3 + 1] - v0) / fv;
        __syncwarp();
        const double u2 = xn * xn + yn * yn;
        bool ok = lane < np;
        double zz;
        if (MODEL == MODEL_EUCM) {
            const double alpha = p[0], beta = p[1], gamma = 1.0 - alpha;
            const double det = 1.0 - (alpha - gamma) * beta * u2;
            ok = ok && !(det < 0.0);
            zz = (1.0 - u2 * alpha * alpha * beta) / (gamma + alpha * sqrt(det));
        } else {
            const double xi = p[0];
            const double g = sqrt(1.0 + u2 * (1.0 - xi * xi));
            const double en = -g - xi * u2, ed = xi * xi * u2 - 1.0;
            zz = ed / (ed + xi * en);
        }
        const unsigned okmask = __ballot_sync(0xffffffffu, ok);
        if (ok_out && lane < np) ok_out[p0 + lane] = ok ? 1 : 0;
        tile[lane * 4] = xn; tile[lane * 4 + 1] = yn; tile[lane * 4 + 2] = zz;
        __syncwarp();
        for (int e = lane; e < np * 3; e += 32)                        // (a failed point is left alone, eucm.h:100)
            if (okmask >> (e / 3) & 1u) X[p0 * 3 + e] = tile[(e / 3) * 4 + e % 3];
        __syncwarp();
    }
}

int model_K(int model) { return model == VG_MODEL_EUCM ? 6 : model == VG_MODEL_UCM ? 5 : model == VG_MODEL_MEI ? 10 : -1; }

int grid_for(long long n)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long blocks = (n + 255) / 256;
    const long long cap = (long long)sms * 5;           // a multiple of the SM count (5 CTAs of 40 KB tiles fit an SM)
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace
}  // namespace vg

using namespace vg;

extern "C" {

int vg_project_points_dev(int model, const double *intr, long long n, const double *X, double *uv, double *dPdX,
                          double *dPdintr, unsigned char *ok, void *stream)
{
    if (model_K(model) < 0) return fail(VG_ERR_INVALID, "invalid camera model name");
    if (n < 0 || !intr || (n > 0 && !X)) return fail(VG_ERR_INVALID, "vg_project_points: bad arguments");
    if (n == 0) return VG_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = grid_for(n);
    if (model == VG_MODEL_EUCM) project_points_kernel<MODEL_EUCM><<<grid, 256, 0, st>>>(intr, n, X, uv, dPdX, dPdintr, ok);
    else if (model == VG_MODEL_UCM) project_points_kernel<MODEL_UCM><<<grid, 256, 0, st>>>(intr, n, X, uv, dPdX, dPdintr, ok);
    else project_points_kernel<MODEL_MEI><<<grid, 256, 0, st>>>(intr, n, X, uv, dPdX, dPdintr, ok);
    count_launch(&launch_counter());
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? VG_OK : fail_cuda(e, "project_points_kernel launch");
}

int vg_reconstruct_points_dev(int model, const double *intr, long long n, const double *uv, double *X, unsigned char *ok,
                              void *stream)
{
    if (model_K(model) < 0) return fail(VG_ERR_INVALID, "invalid camera model name");
    if (n < 0 || !intr || (n > 0 && (!uv || !X))) return fail(VG_ERR_INVALID, "vg_reconstruct_points: bad arguments");
    if (n == 0) return VG_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = grid_for(n);
    if (model == VG_MODEL_EUCM) reconstruct_points_kernel<MODEL_EUCM><<<grid, 256, 0, st>>>(intr, n, uv, X, ok);
    else if (model == VG_MODEL_UCM) reconstruct_points_kernel<MODEL_UCM><<<grid, 256, 0, st>>>(intr, n, uv, X, ok);
    else reconstruct_points_kernel<MODEL_MEI><<<grid, 256, 0, st>>>(intr, n, uv, X, ok);
    count_launch(&launch_counter());
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? VG_OK : fail_cuda(e, "reconstruct_points_kernel launch");
}

// host buffers: one device allocation per call (these are convenience entry points; bulk users keep their points on
// the device and call the _dev forms)
static int with_device_buffers(int model, const double *intr, long long n, const double *in, int in_per, double *out0, int out0_per,
                               double *out1, int out1_per, double *out2, int out2_per, unsigned char *ok, bool project)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(VG_ERR_CUDA, "no CUDA device: this engine has no CPU path");
    }
    const int K = model_K(model);
    if (K < 0) return fail(VG_ERR_INVALID, "invalid camera model name");
    if (n < 0 || !intr || (n > 0 && !in)) return fail(VG_ERR_INVALID, "null argument");
    if (n == 0) return VG_OK;
    const size_t nn = (size_t)n;
    const size_t total = 8 * (16 + nn * in_per + nn * out0_per + nn * out1_per + nn * out2_per) + nn + 64;
    char *base = nullptr;
    VG_CUDA(cudaMalloc(&base, total));
    double *d_intr = reinterpret_cast<double *>(base), *d_in = d_intr + 16, *d0 = d_in + nn * in_per, *d1 = d0 + nn * out0_per,
           *d2 = d1 + nn * out1_per;
    unsigned char *d_ok = reinterpret_cast<unsigned char *>(d2 + nn * out2_per);
    cudaError_t e = cudaMemcpy(d_intr, intr, 8 * K, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_in, in, 8 * nn * in_per, cudaMemcpyHostToDevice);
    // outputs a failed point leaves alone keep the caller's values
    if (e == cudaSuccess && out0) e = cudaMemcpy(d0, out0, 8 * nn * out0_per, cudaMemcpyHostToDevice);
    int rc = VG_OK;
    if (e == cudaSuccess) {
        rc = project ? vg_project_points_dev(model, d_intr, n, d_in, out0 ? d0 : nullptr, out1 ? d1 : nullptr, out2 ? d2 : nullptr, d_ok, nullptr)
                     : vg_reconstruct_points_dev(model, d_intr, n, d_in, d0, d_ok, nullptr);
    }
    if (rc == VG_OK && e == cudaSuccess && out0) e = cudaMemcpy(out0, d0, 8 * nn * out0_per, cudaMemcpyDeviceToHost);
    if (rc == VG_OK && e == cudaSuccess && out1) e = cudaMemcpy(out1, d1, 8 * nn * out1_per, cudaMemcpyDeviceToHost);
    if (rc == VG_OK && e == cudaSuccess && out2) e = cudaMemcpy(out2, d2, 8 * nn * out2_per, cudaMemcpyDeviceToHost);
    if (rc == VG_OK && e == cudaSuccess && ok) e = cudaMemcpy(ok, d_ok, nn, cudaMemcpyDeviceToHost);
    cudaFree(base);
    if (rc) return rc;
    return e == cudaSuccess ? VG_OK : fail_cuda(e, "vg_project_points / vg_reconstruct_points");
}

int vg_project_points(int model, const double *intr, long long n, const double *X, double *uv, double *dPdX, double *dPdintr,
                      unsigned char *ok)
{
    const int K = model_K(model);
    return with_device_buffers(model, intr, n, X, 3, uv, 2, dPdX, 6, dPdintr, 2 * (K > 0 ? K : 1), ok, true);
}

int vg_reconstruct_points(int model, const double *intr, long long n, const double *uv, double *X, unsigned char *ok)
{
    if (n > 0 && !X) return fail(VG_ERR_INVALID, "null argument");
    return with_device_buffers(model, intr, n, uv, 2, X, 3, nullptr, 0, nullptr, 0, ok, false);
}

}  // extern "C"
