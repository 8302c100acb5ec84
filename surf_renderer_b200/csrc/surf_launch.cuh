// surf_launch.cuh - part of libsurf_b200.so (included by every translation unit inside namespace surf).
// The per-call frame description and the host-side launch functions that cross translation units.  The library is
// built from five translation units so that they compile in parallel and a change to the shading / backward code
// does not recompile the intersection templates:
//   surf_kernels.cu      C ABI, frame / shade / backward / splat kernels, host orchestration
//   surf_isect_main.cu   k_intersect<P, MODE> (single scene) + run_intersect (chunking, dispatch)
//   surf_isect_const.cu  k_filter_const / k_narrow_queue / k_const_fallback (disk sets of large single frames: records
//                        through the constant bank into uniform registers - config E's hot kernel)
//   surf_isect_batch.cu  k_intersect_batch<P, MODE> (strided batches, dense frames, triangle scenes)
//   surf_isect_rays.cu   k_intersect_screen, k_intersect_rays, k_intersect_generic, k_intersect_shadow + prep kernels
#pragma once

inline int g_sm_count = 0;
inline int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

struct Frame {           // everything derived from (scene, camera, options) once per call
    SceneView sc;
    CamArgs cam;
    int pix0, n;
    ShadeFlags fl;
    bool shadow;
    Workspace ws;
};

// surf_isect_main.cu.  `ba` non-null: strided batch (perspective, plane-filter modes only)
int run_intersect(const Frame& f, const SurfOptions* opt, cudaStream_t st, const BatchArgs* ba = nullptr);
// surf_isect_const.cu: one disk set of a perspective frame, records streamed through the constant bank
int run_intersect_const(const Frame& f, const SetView& sv, int filter, cudaStream_t st);     // filter 0: plane, 1: sphere + plane
bool const_path_fits(const Frame& f, const SetView& sv);       // the workspace carries a candidate queue for this frame
// surf_isect_rays.cu
int run_intersect_screen(const Frame& f, const SurfOptions* opt, cudaStream_t st);                 // math_mode 3
int run_intersect_ortho(const Frame& f, const SurfOptions* opt, cudaStream_t st);                  // per-pixel origins
int run_intersect_rays_shadow(const Frame& f, unsigned long long* zbuf, cudaStream_t st, const int* n_live,
                              long long n_rays_cap);                                                // math_mode 2 cross-check
int run_intersect_shadow(const Frame& f, const int* n_live, cudaStream_t st);                      // light-origin filters
