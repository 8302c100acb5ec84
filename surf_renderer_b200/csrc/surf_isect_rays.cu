// surf_isect_rays.cu - translation unit of libsurf_b200.so: the intersection kernels other than the camera-ray
// plane-filter pair - k_intersect_screen (math_mode 3), k_intersect_rays / k_intersect_generic (per-ray origins),
// k_intersect_shadow (light-origin filters) - with their prep kernels and host-side launch logic.
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
#include "surf_intersect_rays.cuh"

template <int P>
static int launch_screen(const ScreenParams& prm, int grid, size_t smem, cudaStream_t st) {
    auto kern = k_intersect_screen<P>;
    SURF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    timer_mark(0, 0, st);
    kern<<<grid, kThreads, smem, st>>>(prm);
    timer_mark(0, 1, st);
    SURF_LAUNCHED("k_intersect_screen");
    return SURF_OK;
}

int run_intersect_screen(const Frame& f, const SurfOptions* opt, cudaStream_t st) {
    ScreenParams prm;
    prm.sc = f.sc; prm.cam = f.ws.cam; prm.circ = f.ws.circ; prm.rays = f.ws.rays; prm.zbuf = f.ws.zbuf;
    prm.pix0 = f.pix0; prm.n_pix = f.n; prm.W = f.cam.W; prm.total = f.sc.total;
    int P = opt->pixels_per_thread ? opt->pixels_per_thread : 8;
    if (P != 4 && P != 8 && P != 16) return fail(SURF_ERR_BAD_ARG, "pixels_per_thread must be 4, 8 or 16 for math_mode 3");
    const int row0 = f.pix0 / f.cam.W, row1 = (f.pix0 + f.n - 1) / f.cam.W;
    prm.row0 = row0;
    prm.tiles_x = (f.cam.W + 8 * P - 1) / (8 * P);
    const int tiles_y = (row1 - row0 + 1 + 31) / 32;
    prm.n_tiles = prm.tiles_x * tiles_y;
    const int occ = P <= 8 ? 3 : 2;
    const int grid_max = sm_count() * occ;
    int chunk = opt->chunk_prims ? opt->chunk_prims : 1024;
    if (chunk < 32 || chunk > 2048 || chunk % 32) return fail(SURF_ERR_BAD_ARG, "chunk_prims must be a multiple of 32 in [32, 2048]");
    if (!opt->chunk_prims)
        while (chunk > 64 && (long long)((f.sc.total + chunk - 1) / chunk) * prm.n_tiles < 4LL * grid_max) chunk /= 2;
    prm.chunk = chunk;
    prm.n_chunks = (f.sc.total + chunk - 1) / chunk;
    const long long items = (long long)prm.n_tiles * prm.n_chunks;
    const int grid = (int)std::min<long long>(items, grid_max);
    const size_t smem = (size_t)kStages * chunk * sizeof(float4);
    if (P == 4) return launch_screen<4>(prm, grid, smem, st);
    if (P == 8) return launch_screen<8>(prm, grid, smem, st);
    return launch_screen<16>(prm, grid, smem, st);
}

template <int MODE>
static int run_intersect_rays(const Frame& f, unsigned long long* zbuf, cudaStream_t st, const int* n_live = nullptr,
                              long long n_rays_cap = 0) {
    const long long cap = n_rays_cap > 0 ? n_rays_cap : f.n;        // rays stored (row stride of `gray`)
    if (cap > 0x7fffffffLL) return fail(SURF_ERR_UNSUPPORTED, "more than 2^31 rays in one launch");
    k_prep_rays<<<(f.sc.total + 255) / 256, 256, 0, st>>>(f.sc, f.ws.obound, f.ws.packed);
    SURF_LAUNCHED("k_prep_rays");
    constexpr int P = 4;
    RayParams prm;
    prm.sc = f.sc; prm.cam = f.ws.cam; prm.packed = f.ws.packed; prm.gray = f.ws.gray; prm.zbuf = zbuf; prm.n_pix = (int)cap;
    prm.n_live = n_live;
    const int tile = kThreads * P;
    prm.n_tiles = (int)((cap + tile - 1) / tile);
    const int grid_max = sm_count() * 2;
    int chunk = 1024;
    while (chunk > 64) {
        long long items = 0;
        for (int s = 0; s < f.sc.n_sets; ++s) {
            const int ppc = (chunk * 2) / rec_f4(f.sc.sets[s].kind);
            items += (f.sc.sets[s].count + ppc - 1) / ppc;
        }
        if (items * prm.n_tiles >= 4LL * grid_max) break;
        chunk /= 2;
    }
    prm.stage_f4 = chunk * 2;
    int nchunks = 0;
    for (int s = 0; s < kMaxSets; ++s) {
        prm.chunks_before[s] = nchunks;
        if (s < f.sc.n_sets) {
            const int ppc = prm.stage_f4 / rec_f4(f.sc.sets[s].kind);
            nchunks += (f.sc.sets[s].count + ppc - 1) / ppc;
        }
    }
    prm.chunks_before[kMaxSets] = nchunks;
    prm.n_chunks = nchunks;
    const long long items = (long long)prm.n_tiles * nchunks;
    const int grid = (int)std::min<long long>(items, grid_max);
    const size_t smem = (size_t)kStages * prm.stage_f4 * sizeof(float4);
    auto kern = k_intersect_rays<P, MODE>;
    SURF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreads, smem, st>>>(prm);
    SURF_LAUNCHED("k_intersect_rays");
    return SURF_OK;
}

// orthographic camera (per-pixel origins, utils.py:463-468): generic rays + the per-ray-origin filtered kernel;
// math_mode 1 keeps the exact-only brute force for cross-checking
int run_intersect_ortho(const Frame& f, const SurfOptions* opt, cudaStream_t st) {
    if (opt->math_mode == 1) {
        k_intersect_generic<<<(f.n + 255) / 256, 256, 0, st>>>(f.sc, f.ws.cam, f.pix0, f.n, f.ws.zbuf);
        SURF_LAUNCHED("k_intersect_generic");
        return SURF_OK;
    }
    SURF_CUDA(cudaMemsetAsync(f.ws.obound, 0, 4, st));
    k_rays_ortho<<<(f.n + 255) / 256, 256, 0, st>>>(f.ws.cam, f.pix0, f.n, f.ws.gray, f.ws.obound);
    SURF_LAUNCHED("k_rays_ortho");
    return run_intersect_rays<0>(f, f.ws.zbuf, st);
}

int run_intersect_rays_shadow(const Frame& f, unsigned long long* zbuf, cudaStream_t st, const int* n_live, long long n_rays_cap) {
    return run_intersect_rays<1>(f, zbuf, st, n_live, n_rays_cap);
}

// shadow rays: per-light records + k_intersect_shadow (see surf_intersect.cuh)
int run_intersect_shadow(const Frame& f, const int* n_live, cudaStream_t st) {
    constexpr int P = 8;
    const int L = f.sc.n_lights;
    ShadowIsectParams prm;
    prm.sc = f.sc; prm.packed = f.ws.packed; prm.packed_stride = packed_f4_total(f.sc);
    prm.gray = f.ws.gray; prm.cap = (size_t)f.n * L; prm.zbuf2 = f.ws.zbuf2; prm.n_live = n_live;
    prm.n = f.n; prm.n_lights = L;
    k_prep_lights<<<dim3((f.sc.total + 255) / 256, L), 256, 0, st>>>(f.sc, f.ws.packed, prm.packed_stride);
    SURF_LAUNCHED("k_prep_lights");
    const int tile = kThreads * P;
    prm.tiles_per_light = (f.n + tile - 1) / tile;
    const int grid_max = sm_count() * 2;
    int chunk = 1024;
    while (chunk > 64) {
        long long items = 0;
        for (int s = 0; s < f.sc.n_sets; ++s) {
            const int ppc = (chunk * 2) / rec_f4(f.sc.sets[s].kind);
            items += (f.sc.sets[s].count + ppc - 1) / ppc;
        }
        if (items * prm.tiles_per_light * L >= 4LL * grid_max) break;
        chunk /= 2;
    }
    prm.stage_f4 = chunk * 2;
    int nchunks = 0;
    for (int s = 0; s < kMaxSets; ++s) {
        prm.chunks_before[s] = nchunks;
        if (s < f.sc.n_sets) {
            const int ppc = prm.stage_f4 / rec_f4(f.sc.sets[s].kind);
            nchunks += (f.sc.sets[s].count + ppc - 1) / ppc;
        }
    }
    prm.chunks_before[kMaxSets] = nchunks;
    prm.n_chunks = nchunks;
    const long long items = (long long)prm.tiles_per_light * L * nchunks;
    if (items > 0x7fffffffLL) return fail(SURF_ERR_UNSUPPORTED, "too many shadow work items");
    const int grid = (int)std::min<long long>(items, grid_max);
    const size_t smem = (size_t)kStages * prm.stage_f4 * sizeof(float4);
    auto kern = k_intersect_shadow<P>;
    SURF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreads, smem, st>>>(prm);
    SURF_LAUNCHED("k_intersect_shadow");
    return SURF_OK;
}


}  // namespace surf
