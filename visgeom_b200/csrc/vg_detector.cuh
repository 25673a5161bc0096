// vg_detector.cuh -- what vg_corner.cu and vg_detector.cu share.
#pragma once
#include "vg_common.h"

namespace vg {
// computeResponse (vg_corner.cu) with the outputs the detector pipeline wants: imgrad optional, the blurred 8-bit
// images s1 / s2 optional (both or neither)
int corner_response_launch(const unsigned char *img, int n_img, int width, int height, double sigma1, double sigma2,
                           float *resp, float *gradx, float *grady, float *imgrad, unsigned char *s1, unsigned char *s2,
                           double *avg, long long *count, void *stream, void *work, size_t work_bytes);
// device bytes the launch needs as its work area (per-strip partial sums); work = NULL uses a scratch of the calling thread
size_t corner_response_work_bytes(int n_img, int width, int height);
}  // namespace vg
