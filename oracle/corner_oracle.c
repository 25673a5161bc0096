/* corner_oracle.c -- CPU restatement of the first stage of visgeom's checkerboard detector (TEST INFRASTRUCTURE ONLY:
 * imported by tests/, smoke() and bench.py's cpu_baseline leg, never by the product).
 *
 *   CornerDetector::computeResponse      src/calibration/corner_detector.cpp:262-329
 *     two Gaussian blurs of the 8-bit image (cv::GaussianBlur, :266, :270), then per interior pixel the "sharp"
 *     gradient (:283-290), the saddle response -Iuu Ivv + Iuv^2 - 0.001 (gx^2 + gy^2)^2 of the wider blur (:297-309),
 *     kept where > 0.01 (:310-315), and the average of the kept values (:318).
 *
 * cv::GaussianBlur is a THIRD-PARTY dependency that is not under /root/reference (OpenCV "2.4.9+", README.md:21,
 * unpinned, not vendored).  For 8-bit images OpenCV >= 3.4.2 / 4.x takes its bit-exact fixed-point path; that published
 * algorithm is restated here: the kernel exp(-x^2 / (2 sigma^2)) normalised in double, quantised to 8 fractional bits
 * with the rounding error carried from tap to tap and the centre tap taking what is left of 256
 * (getGaussianKernelFixedPoint_ED), border BORDER_REFLECT_101, both passes in integers, one rounding at the end:
 * (sum + 2^15) >> 16.  Pinned bit for bit against cv2 4.13 (the python wheel of this container) by
 * tests/golden/make_corner_golden.py -> tests/golden/corner_response.npz.  The stencil part is pinned on the
 * reference's own corner_detector.cpp compiled against an OpenCV stand-in (oracle/_ref/libvisgeom_refdet.so, whose
 * GaussianBlur is the function below): tests/test_detector_oracle.py, fixtures tests/golden/detector.npz. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* 8.8 fixed-point Gaussian taps, n odd */
void vgo_gaussian_kernel_u8(int n, double sigma, int *k)
{
    double t[64], sum = 0.0, err = 0.0;
    long s = 0;
    for (int i = 0; i < n; i++) {
        const double x = i - (n - 1) * 0.5;
        t[i] = exp(-0.5 / (sigma * sigma) * x * x);
        sum += t[i];
    }
    for (int i = 0; i < n / 2; i++) {
        const double adj = t[i] / sum * 256.0 + err;
        const long v = lrint(adj);                   /* cvRound: to nearest, ties to even */
        err = adj - (double)v;
        k[i] = k[n - 1 - i] = (int)v;
        s += v;
    }
    k[n / 2] = (int)(256 - 2 * s);
}

static inline int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

/* cv::GaussianBlur(src, dst, Size(n, n), sigma, sigma) for CV_8UC1 */
void vgo_gaussian_blur_u8(const uint8_t *src, int width, int height, int n, double sigma, uint8_t *dst)
{
    int k[64];
    const int r = n / 2;
    vgo_gaussian_kernel_u8(n, sigma, k);
    uint32_t *h = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)width * height);
    for (int v = 0; v < height; v++)
        for (int u = 0; u < width; u++) {
            uint32_t a = 0;
            for (int j = 0; j < n; j++) a += (uint32_t)k[j] * src[(size_t)v * width + reflect101(u + j - r, width)];
            h[(size_t)v * width + u] = a;
        }
    for (int v = 0; v < height; v++)
        for (int u = 0; u < width; u++) {
            uint32_t a = 0;
            for (int j = 0; j < n; j++) a += (uint32_t)k[j] * h[(size_t)reflect101(v + j - r, height) * width + u];
            a = (a + 32768u) >> 16;
            dst[(size_t)v * width + u] = (uint8_t)(a > 255u ? 255u : a);
        }
    free(h);
}

/* computeResponse (corner_detector.cpp:262-329).  resp / imgrad are zero on the one-pixel border (:272-273); gradx /
 * grady are left unset there by the reference (cv::Mat::create does not clear): zero here.  Returns the number of kept
 * responses; *avg = their mean (NaN if none, as 0 / 0 in the reference). */
long vgo_corner_response(const uint8_t *img, int width, int height, double sigma1, double sigma2, float *resp,
                         float *gradx, float *grady, float *imgrad, double *avg)
{
    const size_t N = (size_t)width * height;
    uint8_t *s1 = (uint8_t *)malloc(N), *s2 = (uint8_t *)malloc(N);
    vgo_gaussian_blur_u8(img, width, height, 3, sigma1, s1);                              /* :265-266 */
    vgo_gaussian_blur_u8(img, width, height, 1 + 2 * (int)ceil(sigma2), sigma2, s2);      /* :269-270 */
    memset(resp, 0, N * sizeof(float));
    memset(imgrad, 0, N * sizeof(float));
    memset(gradx, 0, N * sizeof(float));
    memset(grady, 0, N * sizeof(float));
    double acc = 0.0;
    long count = 0;
#define S1(v, u) ((double)s1[(size_t)(v) * width + (u)])
#define S2(v, u) ((double)s2[(size_t)(v) * width + (u)])
    for (int v = 1; v < height - 1; v++)
        for (int u = 1; u < width - 1; u++) {
            const double gxs = (S1(v, u + 1) - S1(v, u - 1) - 0.3 * (S2(v, u + 1) - S2(v, u - 1))) / 2.;
            const double gys = (S1(v + 1, u) - S1(v - 1, u) - 0.3 * (S2(v + 1, u) - S2(v - 1, u))) / 2.;
            gradx[(size_t)v * width + u] = (float)(gxs * 0.01);
            grady[(size_t)v * width + u] = (float)(gys * 0.01);
            imgrad[(size_t)v * width + u] = (float)(sqrt(gxs * gxs + gys * gys) * 0.01);
            const double iuu = S2(v, u - 1) + S2(v, u + 1) - 2 * S2(v, u);
            const double ivv = S2(v - 1, u) + S2(v + 1, u) - 2 * S2(v, u);
            const double iuv = (S2(v - 1, u - 1) + S2(v + 1, u + 1) - S2(v + 1, u - 1) - S2(v - 1, u + 1)) / 4;
            /* :304-305 divide two 8-bit pixels' difference by the int 2 * HSIZE: an INTEGER division (towards zero) */
            const double gx = (double)(((int)s2[(size_t)v * width + u + 1] - (int)s2[(size_t)v * width + u - 1]) / 2);
            const double gy = (double)(((int)s2[(size_t)(v + 1) * width + u] - (int)s2[(size_t)(v - 1) * width + u]) / 2);
            const double gsq = gx * gx + gy * gy;
            const double val = -iuu * ivv + iuv * iuv - 0.001 * (gsq * gsq);
            if (val > 0.01) {
                resp[(size_t)v * width + u] = (float)val;
                acc += val;
                count++;
            }
        }
#undef S1
#undef S2
    *avg = acc / (double)count;
    free(s1); free(s2);
    return count;
}
