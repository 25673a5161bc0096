// vg_common.h -- error plumbing shared by the C-ABI translation units.
#pragma once

#include <cuda_runtime.h>

#include <string>

#include "../../include/visgeom_b200.h"

namespace vg {

void set_error(const std::string &msg);
int fail(int code, const std::string &msg);
int fail_cuda(cudaError_t e, const char *where);
unsigned long long &launch_counter();
#ifndef VG_COUNT_LAUNCH_DEFINED
#define VG_COUNT_LAUNCH_DEFINED
inline void count_launch(unsigned long long *c) { __atomic_fetch_add(c, 1ull, __ATOMIC_RELAXED); }   // several host threads
#endif

#define VG_CUDA(call)                                              \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return vg::fail_cuda(e__, #call);  \
    } while (0)

}  // namespace vg
