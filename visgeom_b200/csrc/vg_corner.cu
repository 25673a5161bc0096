// vg_corner.cu -- first stage of the checkerboard detector on the GPU:
//   CornerDetector::computeResponse      src/calibration/corner_detector.cpp:262-329
// Two Gaussian blurs of the 8-bit image (cv::GaussianBlur :266, :270 -- OpenCV's bit-exact fixed-point path for CV_8U:
// 8.8 fixed-point taps, BORDER_REFLECT_101, integer passes, one rounding (sum + 2^15) >> 16), the "sharp" gradient
// (:283-290), the saddle response -Iuu Ivv + Iuv^2 - 0.001 |grad|^4 of the wider blur kept where > 0.01 (:297-315) and
// the mean of the kept values (:318), fused in ONE pass over the image: the blurred rows live in sliding windows of
// registers (a warp walks down a strip of 30 columns), runs both separable blurs there, and writes the four float maps.  HBM-bound integer / stencil work
// (1 B read, 16 B written per pixel); no tensor cores, no GEMM.  Arithmetic of the stencil is written with explicit
// round-to-nearest intrinsics (no FMA contraction), so every float written equals the CPU restatement's bit for bit;
// only the mean is summed in a different (fixed) order.  Batched over images (grid.z).  No CPU path.
#include "vg_common.h"
#include "vg_detector.cuh"

#include <cmath>
#include <cstdint>

namespace vg {
namespace {

constexpr int RMAX = 4;                       // blur radius <= 4 (sigma_2 <= 4)
constexpr int COLS = 30;                      // output columns of a warp: 32 lanes minus one halo lane on each side
constexpr int STRIP = 64;                     // output rows of a warp
constexpr int WARPS = 8;
struct Taps { int r; int k[2 * RMAX + 1]; };

__device__ __forceinline__ int reflect101(int i, const int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

// One warp = a strip of COLS columns x STRIP rows; lane l owns column x0 - 1 + l and walks down the image keeping the
// last rows of the horizontally blurred values (integers) in registers: the vertical passes and the 3 x 3 stencil come
// from those windows, the left / right neighbours of the blurred rows through two shuffles per row.  No shared memory,
// ~70 instructions per pixel.  R2 = radius of the wide blur (compile time: the windows are register arrays).
template <int R2>
__global__ void __launch_bounds__(32 * WARPS)
corner_response_kernel(const unsigned char *__restrict__ img, const int W, const int H, const Taps t1, const Taps t2,
                       float *__restrict__ resp, float *__restrict__ gradx, float *__restrict__ grady,
                       float *__restrict__ imgrad, unsigned char *__restrict__ s1out, unsigned char *__restrict__ s2out,
                       double *__restrict__ part_acc, unsigned int *__restrict__ part_cnt, const int warps_x, const int strips_y)
{
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long wid = (long long)blockIdx.x * WARPS + wib;          // warp index inside the image
    const int per_img = warps_x * strips_y;
    if (wid >= per_img) return;
    const int wy = (int)(wid / warps_x), wx = (int)(wid - (long long)wy * warps_x);
    const int x = wx * COLS - 1 + lane, ya = wy * STRIP, yb = min(H, ya + STRIP);
    const size_t base = (size_t)blockIdx.y * W * H;
    const unsigned char *src = img + base;
    // this lane's five (2 R2 + 1) source columns, reflected at the image border (BORDER_REFLECT_101)
    int col[2 * R2 + 1];
#pragma unroll
    for (int j = 0; j <= 2 * R2; j++) col[j] = reflect101(x - R2 + j, W);
    int k1[3], k2[2 * R2 + 1];
#pragma unroll
    for (int j = 0; j < 3; j++) k1[j] = t1.k[j];
#pragma unroll
    for (int j = 0; j <= 2 * R2; j++) k2[j] = t2.k[j];
    // windows: h2 rows yy - 2 R2 .. yy, h1 rows yy - R2 - 1 .. yy - R2 + 1 (kept as the last R2 + 2 rows)
    unsigned int h2w[2 * R2 + 1], h1w[R2 + 2];
    int s1c[3] = {0, 0, 0}, s2l[3] = {0, 0, 0}, s2c[3] = {0, 0, 0}, s2r[3] = {0, 0, 0}, s1l = 0, s1r = 0, s1lm = 0, s1rm = 0;
#pragma unroll
    for (int j = 0; j <= 2 * R2; j++) h2w[j] = 0;
#pragma unroll
    for (int j = 0; j < R2 + 2; j++) h1w[j] = 0;
    double acc = 0.0;
    unsigned int cnt = 0;
    const bool writer = lane >= 1 && lane <= COLS && x < W;
    for (int yy = ya - 1 - R2; yy <= yb + R2; yy++) {
        // horizontal passes of input row yy (8.8 taps x 8-bit pixels: exact)
        const unsigned char *row = src + (size_t)reflect101(yy, H) * W;
        unsigned int px[2 * R2 + 1];
#pragma unroll
        for (int j = 0; j <= 2 * R2; j++) px[j] = __ldg(row + col[j]);
        unsigned int a = 0, b = 0;
#pragma unroll
        for (int j = 0; j < 3; j++) a += (unsigned)k1[j] * px[R2 - 1 + j];
#pragma unroll
        for (int j = 0; j <= 2 * R2; j++) b += (unsigned)k2[j] * px[j];
#pragma unroll
        for (int j = 0; j < 2 * R2; j++) h2w[j] = h2w[j + 1];
        h2w[2 * R2] = b;
#pragma unroll
        for (int j = 0; j < R2 + 1; j++) h1w[j] = h1w[j + 1];
        h1w[R2 + 1] = a;
        // vertical passes centred on row c = yy - R2, one rounding (sum + 2^15) >> 16
        unsigned int v1 = 0, v2 = 0;
#pragma unroll
        for (int j = 0; j < 3; j++) v1 += (unsigned)k1[j] * h1w[j];
#pragma unroll
        for (int j = 0; j <= 2 * R2; j++) v2 += (unsigned)k2[j] * h2w[j];
        v1 = min((v1 + 32768u) >> 16, 255u);
        v2 = min((v2 + 32768u) >> 16, 255u);
        // blurred rows c - 2, c - 1, c of this column and (wide blur) of its neighbours
        s1c[0] = s1c[1]; s1c[1] = s1c[2]; s1c[2] = (int)v1;
        s2c[0] = s2c[1]; s2c[1] = s2c[2]; s2c[2] = (int)v2;
        s2l[0] = s2l[1]; s2l[1] = s2l[2]; s2l[2] = __shfl_up_sync(0xffffffffu, (int)v2, 1);
        s2r[0] = s2r[1]; s2r[1] = s2r[2]; s2r[2] = __shfl_down_sync(0xffffffffu, (int)v2, 1);
        s1lm = s1l; s1l = __shfl_up_sync(0xffffffffu, (int)v1, 1);      // s1 of row c (kept one step: row c - 1 is the stencil's)
        s1rm = s1r; s1r = __shfl_down_sync(0xffffffffu, (int)v1, 1);
        const int o = yy - R2 - 1;                                       // output row: the middle of the three
        if (o < ya || o >= yb || !writer) continue;
        const size_t idx = base + (size_t)o * W + x;
        float f_resp = 0.f, f_gx = 0.f, f_gy = 0.f, f_mag = 0.f;
        if (x >= 1 && x < W - 1 && o >= 1 && o < H - 1) {
            // sharp gradient (:283-290): ((a - b) - 0.3 (c - d)) / 2 -- the differences are integers, then every
            // operation is rounded on its own as in the reference's double arithmetic (no FMA contraction)
            const double gxs = __dmul_rn(__dsub_rn((double)(s1rm - s1lm), __dmul_rn(0.3, (double)(s2r[1] - s2l[1]))), 0.5);
            const double gys = __dmul_rn(__dsub_rn((double)(s1c[2] - s1c[0]), __dmul_rn(0.3, (double)(s2c[2] - s2c[0]))), 0.5);
            f_gx = __double2float_rn(__dmul_rn(gxs, 0.01));
            f_gy = __double2float_rn(__dmul_rn(gys, 0.01));
            if (imgrad)        // (the detector pipeline does not ask for this map: its host stages recompute what they need)
                f_mag = __double2float_rn(__dmul_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(gxs, gxs), __dmul_rn(gys, gys))), 0.01));
            // saddle response of the wider blur (:297-309).  -Iuu Ivv + Iuv^2 is a multiple of 1/16 below 2^21 and
            // |grad|^4 = n^2 with n < 2^15: both exact in integers, hence exact as doubles; what is rounded is
            // 0.001 * q and the final subtraction, as in the reference
            const int iuu = s2l[1] + s2r[1] - 2 * s2c[1], ivv = s2c[0] + s2c[2] - 2 * s2c[1];
            const int iuv4 = s2l[0] + s2r[2] - s2l[2] - s2r[0];
            // the reference divides the pixel differences by the int 2 (:304-305): integer division, towards zero
            const int gx2 = (s2r[1] - s2l[1]) / 2, gy2 = (s2c[2] - s2c[0]) / 2;
            const long long n = (long long)gx2 * gx2 + (long long)gy2 * gy2;
            const double sdet = (double)(iuv4 * iuv4 - 16 * iuu * ivv) * 0.0625;
            const double q = (double)(n * n);
            const double val = __dsub_rn(sdet, __dmul_rn(0.001, q));
            if (val > 0.01) {
                f_resp = __double2float_rn(val);
                acc = __dadd_rn(acc, val);
                cnt++;
            }
        }
        resp[idx] = f_resp; gradx[idx] = f_gx; grady[idx] = f_gy;
        if (imgrad) imgrad[idx] = f_mag;
        // the detector's host stages work from the two blurred images (vg_detector.cu); defined on the border too
        if (s1out) { s1out[idx] = (unsigned char)s1c[1]; s2out[idx] = (unsigned char)s2c[1]; }
    }
    // the strip's sum and count in a fixed order
    for (int off = 16; off > 0; off >>= 1) {
        acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, off));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    }
    if (lane == 0) {
        part_acc[(size_t)blockIdx.y * per_img + wid] = acc;
        part_cnt[(size_t)blockIdx.y * per_img + wid] = cnt;
    }
}

// mean of the kept responses of each image: the tiles' partial sums in a fixed order
__global__ void __launch_bounds__(256)
corner_mean_kernel(const double *__restrict__ part_acc, const unsigned int *__restrict__ part_cnt, const int tiles,
                   double *__restrict__ avg, long long *__restrict__ count)
{
    __shared__ double sa[256];
    __shared__ unsigned long long sc[256];
    const int tid = threadIdx.x;
    double a = 0.0;
    unsigned long long c = 0;
    for (int i = tid; i < tiles; i += 256) {
        a = __dadd_rn(a, part_acc[(size_t)blockIdx.x * tiles + i]);
        c += part_cnt[(size_t)blockIdx.x * tiles + i];
    }
    sa[tid] = a; sc[tid] = c;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (tid < off) { sa[tid] = __dadd_rn(sa[tid], sa[tid + off]); sc[tid] += sc[tid + off]; }
        __syncthreads();
    }
    if (tid == 0) {
        if (avg) avg[blockIdx.x] = sa[0] / (double)sc[0];        // (0 / 0 -> NaN when nothing was kept, as the reference)
        if (count) count[blockIdx.x] = (long long)sc[0];
    }
}

// OpenCV's fixed-point Gaussian taps (getGaussianKernelBitExact + getGaussianKernelFixedPoint_ED, 8 fractional bits)
bool gaussian_taps(int n, double sigma, Taps *t)
{
    if (n < 1 || (n & 1) == 0 || n > 2 * RMAX + 1 || !(sigma > 0.0)) return false;
    double w[2 * RMAX + 1], sum = 0.0, err = 0.0;
    long s = 0;
    for (int i = 0; i < n; i++) { const double x = i - (n - 1) * 0.5; w[i] = std::exp(-0.5 / (sigma * sigma) * x * x); sum += w[i]; }
    for (int i = 0; i < n / 2; i++) {
        const double adj = w[i] / sum * 256.0 + err;
        const long v = std::lrint(adj);
        err = adj - (double)v;
        t->k[i] = t->k[n - 1 - i] = (int)v;
        s += v;
    }
    t->k[n / 2] = (int)(256 - 2 * s);
    t->r = n / 2;
    return true;
}

}  // namespace
}  // namespace vg

using namespace vg;

size_t vg::corner_response_work_bytes(int n_img, int width, int height)
{
    const size_t tiles = (size_t)((width + COLS - 1) / COLS) * ((height + STRIP - 1) / STRIP);
    return tiles * n_img * (sizeof(double) + sizeof(unsigned int));
}

// computeResponse for the detector pipeline (vg_detector.cu): imgrad may be NULL (not written), s1 / s2 receive the
// two blurred 8-bit images when given
int vg::corner_response_launch(const unsigned char *img, int n_img, int width, int height, double sigma1, double sigma2,
                               float *resp, float *gradx, float *grady, float *imgrad, unsigned char *s1, unsigned char *s2,
                               double *avg, long long *count, void *stream, void *work, size_t work_bytes)
{
    if (n_img < 0 || width < 3 || height < 3) return fail(VG_ERR_INVALID, "vg_corner_response: bad image size");
    if (n_img == 0) return VG_OK;
    if (!img || !resp || !gradx || !grady || (!s1) != (!s2)) return fail(VG_ERR_INVALID, "null argument");
    Taps t1, t2;
    // FILTER_SIZE_1 = 3, FILTER_SIZE_2 = 1 + 2 ceil(SIGMA_2)  (corner_detector.cpp:265,269)
    if (!gaussian_taps(3, sigma1, &t1) || !gaussian_taps(1 + 2 * (int)std::ceil(sigma2), sigma2, &t2))
        return fail(VG_ERR_UNSUPPORTED, "vg_corner_response: sigma out of range (0 < sigma_2 <= 4)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int warps_x = (width + COLS - 1) / COLS, strips_y = (height + STRIP - 1) / STRIP;
    const int tiles = warps_x * strips_y;
    const dim3 grid((tiles + WARPS - 1) / WARPS, n_img);
    // per-strip partial sums: the caller's work area, else a grow-only scratch of the calling thread (work queued on one
    // stream at a time per thread)
    static thread_local struct { int dev = -1; char *p = nullptr; size_t bytes = 0; } scratch;
    const size_t n_part = (size_t)tiles * n_img, need = n_part * (sizeof(double) + sizeof(unsigned int));
    char *area = static_cast<char *>(work);
    if (area && work_bytes < need) return fail(VG_ERR_INVALID, "vg_corner_response: work area too small");
    if (!area) {
        int dev = 0;
        VG_CUDA(cudaGetDevice(&dev));
        if (scratch.dev != dev || scratch.bytes < need) {
            if (scratch.p) { cudaDeviceSynchronize(); cudaFree(scratch.p); scratch.p = nullptr; scratch.bytes = 0; }
            VG_CUDA(cudaMalloc(&scratch.p, need));
            scratch.bytes = need; scratch.dev = dev;
        }
        area = scratch.p;
    }
    double *part_acc = reinterpret_cast<double *>(area);
    unsigned int *part_cnt = reinterpret_cast<unsigned int *>(area + n_part * sizeof(double));
#define VG_CORNER_LAUNCH(R) corner_response_kernel<R><<<grid, 32 * WARPS, 0, st>>>(img, width, height, t1, t2, resp, gradx, grady, imgrad, \
                                                                                   s1, s2, part_acc, part_cnt, warps_x, strips_y)
    switch (t2.r) {
    case 1: VG_CORNER_LAUNCH(1); break;
    case 2: VG_CORNER_LAUNCH(2); break;
    case 3: VG_CORNER_LAUNCH(3); break;
    default: VG_CORNER_LAUNCH(4); break;
    }
#undef VG_CORNER_LAUNCH
    count_launch(&launch_counter());
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && (avg || count)) {
        corner_mean_kernel<<<n_img, 256, 0, st>>>(part_acc, part_cnt, tiles, avg, count);
        count_launch(&launch_counter());
        e = cudaGetLastError();
    }
    return e == cudaSuccess ? VG_OK : fail_cuda(e, "corner_response_kernel launch");
}

extern "C" {

int vg_corner_response_dev(const unsigned char *img, int n_img, int width, int height, double sigma1, double sigma2,
                           float *resp, float *gradx, float *grady, float *imgrad, double *avg, long long *count, void *stream)
{
    if (n_img > 0 && !imgrad) return fail(VG_ERR_INVALID, "null argument");
    return corner_response_launch(img, n_img, width, height, sigma1, sigma2, resp, gradx, grady, imgrad, nullptr, nullptr, avg,
                                  count, stream, nullptr, 0);
}

int vg_corner_response(const unsigned char *img, int n_img, int width, int height, double sigma1, double sigma2,
                       float *resp, float *gradx, float *grady, float *imgrad, double *avg, long long *count)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(VG_ERR_CUDA, "no CUDA device: this engine has no CPU path");
    }
    if (n_img < 0 || width < 3 || height < 3) return fail(VG_ERR_INVALID, "vg_corner_response: bad image size");
    if (n_img == 0) return VG_OK;
    if (!img || !resp || !gradx || !grady || !imgrad) return fail(VG_ERR_INVALID, "null argument");
    const size_t N = (size_t)width * height * n_img;
    unsigned char *d_img = nullptr;
    float *d_out = nullptr;
    double *d_avg = nullptr;
    long long *d_cnt = nullptr;
    cudaError_t e = cudaMalloc(&d_img, N);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, 4 * N * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_avg, sizeof(double) * n_img);
    if (e == cudaSuccess) e = cudaMalloc(&d_cnt, sizeof(long long) * n_img);
    if (e == cudaSuccess) e = cudaMemcpy(d_img, img, N, cudaMemcpyHostToDevice);
    int rc = VG_OK;
    if (e == cudaSuccess)
        rc = vg_corner_response_dev(d_img, n_img, width, height, sigma1, sigma2, d_out, d_out + N, d_out + 2 * N, d_out + 3 * N,
                                    d_avg, d_cnt, nullptr);
    if (rc == VG_OK && e == cudaSuccess) e = cudaMemcpy(resp, d_out, N * sizeof(float), cudaMemcpyDeviceToHost);
    if (rc == VG_OK && e == cudaSuccess) e = cudaMemcpy(gradx, d_out + N, N * sizeof(float), cudaMemcpyDeviceToHost);
    if (rc == VG_OK && e == cudaSuccess) e = cudaMemcpy(grady, d_out + 2 * N, N * sizeof(float), cudaMemcpyDeviceToHost);
    if (rc == VG_OK && e == cudaSuccess) e = cudaMemcpy(imgrad, d_out + 3 * N, N * sizeof(float), cudaMemcpyDeviceToHost);
    if (rc == VG_OK && e == cudaSuccess && avg) e = cudaMemcpy(avg, d_avg, sizeof(double) * n_img, cudaMemcpyDeviceToHost);
    if (rc == VG_OK && e == cudaSuccess && count) e = cudaMemcpy(count, d_cnt, sizeof(long long) * n_img, cudaMemcpyDeviceToHost);
    cudaFree(d_img); cudaFree(d_out); cudaFree(d_avg); cudaFree(d_cnt);
    if (rc) return rc;
    return e == cudaSuccess ? VG_OK : fail_cuda(e, "vg_corner_response");
}

}  // extern "C"
