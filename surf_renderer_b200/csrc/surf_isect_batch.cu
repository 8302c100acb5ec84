// surf_isect_batch.cu - translation unit of libsurf_b200.so: k_intersect_batch<P, MODE>, the camera-ray intersection
// kernel over a strided batch of scenes (also: single scenes with triangle sets, small dense frames, math_mode 4).
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "surf_view.h"

namespace surf {

#include "surf_runtime.cuh"
#include "surf_ptx.cuh"
#include "surf_batch.cuh"
#include "surf_launch.cuh"
#include "surf_intersect.cuh"

template <int P, int MODE>
static int launch_batch(const IsectParams& prm, const BatchArgs& ba, int grid, size_t smem, cudaStream_t st) {
    auto kern = k_intersect_batch<P, MODE>;
    SURF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    timer_mark(0, 0, st);
    kern<<<grid, kThreads, smem, st>>>(prm, ba);
    timer_mark(0, 1, st);
    SURF_LAUNCHED("k_intersect_batch");
    return SURF_OK;
}

int launch_intersect_batch(const IsectParams& prm, const BatchArgs& ba, int P, int mode, int grid, size_t smem, cudaStream_t st) {
#define SURF_DISPATCH(PP)                                                   \
    if (P == PP) {                                                          \
        if (mode == 0) return launch_batch<PP, 0>(prm, ba, grid, smem, st);  \
        if (mode == 1) return launch_batch<PP, 1>(prm, ba, grid, smem, st);  \
        return launch_batch<PP, 2>(prm, ba, grid, smem, st);                 \
    }
    SURF_DISPATCH(2)
    SURF_DISPATCH(4)
    SURF_DISPATCH(8)
#undef SURF_DISPATCH
    return fail(SURF_ERR_BAD_ARG, "unsupported pixels_per_thread");
}

}  // namespace surf
