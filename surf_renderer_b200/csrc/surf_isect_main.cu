// surf_isect_main.cu - translation unit of libsurf_b200.so: k_intersect<P, MODE>, the single-scene camera-ray
// intersection + z-buffer kernel (config E's hot kernel), and run_intersect, the host-side chunking / dispatch of every
// perspective plane-filter launch (the strided-batch kernel itself lives in surf_isect_batch.cu).
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
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
static int launch_intersect(const IsectParams& prm, int grid, size_t smem, cudaStream_t st) {
    auto kern = k_intersect<P, MODE>;
    SURF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    timer_mark(0, 0, st);
    kern<<<grid, kThreads, smem, st>>>(prm);
    timer_mark(0, 1, st);
    SURF_LAUNCHED("k_intersect");
    return SURF_OK;
}

// `ba` non-null: strided batch (perspective, plane-filter modes only) - the work grid gets a scene dimension
int run_intersect(const Frame& f, const SurfOptions* opt, cudaStream_t st, const BatchArgs* ba) {
    if (ba && (f.cam.proj != 0 || opt->math_mode == 3)) return fail(SURF_ERR_UNSUPPORTED, "no fused batch for this mode");
    if (f.cam.proj != 0) return run_intersect_ortho(f, opt, st);
    if (opt->math_mode == 3) return run_intersect_screen(f, opt, st);
    // math_mode 4 ("dense"): the batch kernel's body - 2-D pixel tiles, per-disk filter minima - on a single scene,
    // i.e. a batch of one.  For small frames with splats several pixels wide (bunny 256x256: -11 %).
    // Scenes with triangle sets take the same body for its packed triangle filter (see chunk_triangles_packed).
    int mode = opt->math_mode;
    BatchArgs one_scene;
    bool has_triangles = false;
    for (int k = 0; k < f.sc.n_sets; ++k) has_triangles |= f.sc.sets[k].kind == KIND_TRIANGLE;
    // Small frames (<= 256x256 pixels) take it too: a primitive that matters at such a resolution is several pixels
    // wide, which is the regime the dense body is built for (bunny 256x256: -19 %; with nothing to narrow: +7 %).
    const bool small_frame = f.n <= 256 * 256;
    if (mode == 4 || (mode == 0 && (has_triangles || small_frame))) {
        mode = 0;
        if (!ba) {
            std::memset(&one_scene, 0, sizeof(one_scene));
            one_scene.n_scenes = 1;
            ba = &one_scene;
        }
    }
    IsectParams prm;
    prm.sc = f.sc;
    // Disk sets of large single frames go through k_intersect_const (records via the constant bank / uniform registers,
    // surf_isect_const.cu); the staged kernel below then handles the scene's other sets.  math_mode 5 keeps the staged
    // kernel for everything (A/B, and the cross-check of the GPU suite).
    if ((mode == 0 || mode == 6) && !ba && (opt->pixels_per_thread == 0 || opt->pixels_per_thread == 8)) {
        int kept = 0;
        for (int k = 0; k < f.sc.n_sets; ++k) {
            const SetView& sv = f.sc.sets[k];
            if (sv.kind == KIND_DISK && sv.count >= 256 && const_path_fits(f, sv)) {
                const int rc = run_intersect_const(f, sv, mode == 0 ? 1 : 0, st);
                if (rc) return rc;
            } else prm.sc.sets[kept++] = sv;
        }
        if (kept == 0) return SURF_OK;
        prm.sc.n_sets = kept;
    }
    if (mode == 5 || mode == 6) mode = 0;
    const SceneView& scv = prm.sc;
    prm.cam = f.ws.cam; prm.packed = f.ws.packed; prm.rays = f.ws.rays; prm.zbuf = f.ws.zbuf;
    prm.n_pix = f.n;
    int P = opt->pixels_per_thread ? opt->pixels_per_thread : 8;
    if (P != 2 && P != 4 && P != 8) return fail(SURF_ERR_BAD_ARG, "pixels_per_thread must be 2, 4 or 8");
    const int tile = kThreads * P;
    prm.tiles_per_scene = (f.n + tile - 1) / tile;
    prm.tiles_x = 0; prm.W = f.cam.W; prm.pix0 = f.pix0;
    if (ba && P == 8 && f.cam.W % 64 == 0) {       // 2-D tiles for the batch kernel (see IsectParams)
        const int row0 = f.pix0 / f.cam.W, row1 = (f.pix0 + f.n - 1) / f.cam.W;
        prm.tiles_x = f.cam.W / 64;
        prm.tiles_per_scene = prm.tiles_x * ((row1 - row0 + 1 + 31) / 32);
    }
    prm.n_tiles = prm.tiles_per_scene * (ba ? ba->n_scenes : 1);
    // stage capacity: chunk_prims disk records (2 float4 each); keep >= 4x grid items for balance on small frames
    int chunk = opt->chunk_prims ? opt->chunk_prims : 1024;
    if (chunk < 32 || chunk > 2048 || chunk % 32) return fail(SURF_ERR_BAD_ARG, "chunk_prims must be a multiple of 32 in [32, 2048]");
    const int grid_max = sm_count() * ((P == 4 && !ba) ? SURF_ISECT_P4_BLOCKS : 2);
    if (!opt->chunk_prims) {
        // pick the largest chunk whose item count splits over the persistent grid with <= 1.5% quantisation loss
        // (items are dealt as equal contiguous ranges: the slowest CTA runs ceil(items / grid) of them)
        auto items_for = [&](int ch) {
            long long items = 0;
            for (int s = 0; s < scv.n_sets; ++s) {
                const int ppc = (ch * 2) / rec_f4(scv.sets[s].kind);
                items += (scv.sets[s].count + ppc - 1) / ppc;
            }
            return items * prm.n_tiles;
        };
        int best = 64;
        double best_loss = 1e30;
        for (int ch = 1024; ch >= 64; ch /= 2) {
            const long long items = items_for(ch);
            const long long per = (items + grid_max - 1) / grid_max;
            const double loss = (double)per * grid_max / (double)items - 1.0;
            if (loss <= 0.015) { best = ch; best_loss = loss; break; }
            if (loss < best_loss) { best = ch; best_loss = loss; }
        }
        chunk = best;
    }
    prm.stage_f4 = chunk * 2;
    int nchunks = 0;
    for (int s = 0; s < kMaxSets; ++s) {
        prm.chunks_before[s] = nchunks;
        if (s < scv.n_sets) {
            const int ppc = prm.stage_f4 / rec_f4(scv.sets[s].kind);
            nchunks += (scv.sets[s].count + ppc - 1) / ppc;
        }
    }
    prm.chunks_before[kMaxSets] = nchunks;
    prm.n_chunks = nchunks;
    const long long items = (long long)prm.n_tiles * nchunks;
    const int grid = (int)std::min<long long>(items, grid_max);
    const size_t smem = (size_t)kStages * prm.stage_f4 * sizeof(float4);
    if (mode < 0 || mode > 2) return fail(SURF_ERR_BAD_ARG, "math_mode must be 0..5");
    if (ba) {
        // hybrid distribution: 3/4 of the items as contiguous static shares, the rest drawn dynamically in short runs
        // measured (B200): splat batches / dense splat frames are fastest with 12/16 static (config D 2.45 -> 2.28 ms,
        // bunny 256x256 0.189 -> 0.184 ms), triangle meshes with 14/16 (their items are 4x heavier per record)
        static const int s16_env = getenv("SURF_DYN_STATIC16") ? atoi(getenv("SURF_DYN_STATIC16")) : -1;     // tuning knobs
        const int s16 = s16_env >= 0 ? s16_env : (has_triangles ? 14 : 12);
        static const int rpc = getenv("SURF_DYN_RUNS") ? atoi(getenv("SURF_DYN_RUNS")) : 12;
        prm.static_per = (int)((items * s16 / 16) / grid);
        prm.dyn_begin = prm.static_per * grid;
        prm.run_len = (int)std::max<long long>(1, (items - prm.dyn_begin) / ((long long)rpc * grid));
        prm.work_counter = (int*)f.ws.obound + 255;          // last cell of the 1 KB counter block
        SURF_CUDA(cudaMemsetAsync(prm.work_counter, 0, sizeof(int), st));
        return launch_intersect_batch(prm, *ba, P, mode, grid, smem, st);
    }
#define SURF_DISPATCH(PP)                                                  \
    if (P == PP) {                                                         \
        if (mode == 0) return launch_intersect<PP, 0>(prm, grid, smem, st); \
        if (mode == 1) return launch_intersect<PP, 1>(prm, grid, smem, st); \
        return launch_intersect<PP, 2>(prm, grid, smem, st);                \
    }
    SURF_DISPATCH(2)
    SURF_DISPATCH(4)
    SURF_DISPATCH(8)
#undef SURF_DISPATCH
    return fail(SURF_ERR_BAD_ARG, "unsupported pixels_per_thread");
}


}  // namespace surf
