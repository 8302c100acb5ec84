// surf_kernels.cu - sm_100a kernels + the C ABI of libsurf_b200.so (include/surf_b200.h).
//
// Pipeline of one forward frame (all on the caller's stream, no host sync):
//   k_setup      1 thread    camera basis / image-plane scale -> CamState            (utils.py:402-456)
//   k_prep       O(M)        per-primitive exact plane constants + conservative filter records, packed
//                            contiguously per primitive set for TMA bulk staging       (utils.py:288-297)
//   k_raygen     O(N)        unit ray directions [3,N] + z-buffer key init             (utils.py:439-478)
//   k_intersect  O(M*N)      THE hot kernel: persistent CTAs, primitives streamed global->shared with
//                            cp.async.bulk + mbarrier (3-stage ring), P pixels per thread in registers,
//                            packed FFMA2 (fma.rn.f32x2) conservative disk filter, exact reference-order
//                            narrow phase on the rare candidates, per-pixel (depth,index) merged with a
//                            64-bit atomicMin so the [M,N] distance tensor never exists  (utils.py:481-512,
//                            renderer.py:170-189)
//   k_shade      O(N*L)      winner -> depth/nearest/pos/normal + Phong shading + tonemap (renderer.py:82-125,
//                            :266-340)
// Backward: k_backward O(N*L) recomputes the winning hit per pixel (surf_math.cuh backward_pixel), reduces
// light/material/colour gradients across the warp and CTA, primitive gradients with a warp-segmented
// reduction keyed by `nearest` before red.global.add; k_backward_finalize folds the double accumulators.
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
#include "surf_frame_kernels.cuh"
#include "surf_fast.cuh"
#include "surf_shade.cuh"
#include "surf_backward.cuh"
#include "surf_splats.cuh"
#include "surf_scatter.cuh"

// ---------------------------------------------------------------------------------------------------
// FP32 FMA-pipe microbenchmark (roofline denominator check)
// ---------------------------------------------------------------------------------------------------
template <bool PACKED>
__global__ void __launch_bounds__(256) k_fma_peak(int iters, float* out) {
    float a = 1.0f + threadIdx.x * 1e-7f, b = 0.999f;
    if (PACKED) {
        unsigned long long x[8];
        for (int j = 0; j < 8; ++j) x[j] = pack2(a + j, a - j);
        const unsigned long long bb = pack2(b, b), cc = pack2(1e-3f, 2e-3f);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = fma2(x[j], bb, cc);
        }
        float s = 0.f;
        for (int j = 0; j < 8; ++j) { float lo, hi; unpack2(x[j], lo, hi); s += lo + hi; }
        if (s == 12345.678f) out[0] = s;
    } else {
        float x[16];
        for (int j = 0; j < 16; ++j) x[j] = a + j;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = fmaf(x[j], b, 1e-3f);
        }
        float s = 0.f;
        for (int j = 0; j < 16; ++j) s += x[j];
        if (s == 12345.678f) out[0] = s;
    }
}

// ---------------------------------------------------------------------------------------------------
// packed Adam: optimizer.step() of the inverse-rendering loop (test_optimization.py:122, torch.optim.Adam without
// amsgrad / weight decay) for leaves whose gradients sit in ONE flat buffer (MSEStep).  torch's fused multi-tensor Adam
// spends ~80 us on config E's four leaves (two of 300 000 floats, two tiny ones: ten thread blocks); this is one
// grid-wide launch plus a one-thread launch that advances the step counter and the bias corrections on the device
// (graph-capturable: nothing step-dependent is a launch parameter).
// ---------------------------------------------------------------------------------------------------
struct AdamTensors { int count; float* param[SURF_ADAM_MAX_TENSORS]; long long begin[SURF_ADAM_MAX_TENSORS + 1]; };

__global__ void k_adam_advance(float* __restrict__ state, float beta1, float beta2) {
    // state[0] = step count, [1] = 1 - beta1^t, [2] = sqrt(1 - beta2^t)
    const float t = state[0] + 1.f;
    state[0] = t;
    state[1] = 1.f - powf(beta1, t);
    state[2] = sqrtf(1.f - powf(beta2, t));
}

__global__ void __launch_bounds__(256) k_adam_packed(const __grid_constant__ AdamTensors tn, const float* __restrict__ grads,
                                                     float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                     const float* __restrict__ state, float lr, float beta1, float beta2, float eps) {
    const long long total = tn.begin[tn.count];
    const float bc1 = state[1], bc2_sqrt = state[2];
    const float step_size = lr / bc1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int j = 0;
        while (i >= tn.begin[j + 1]) ++j;
        const float g = grads[i];
        const float m = beta1 * exp_avg[i] + (1.f - beta1) * g;
        const float v = beta2 * exp_avg_sq[i] + (1.f - beta2) * g * g;
        exp_avg[i] = m;
        exp_avg_sq[i] = v;
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        float* p = tn.param[j] + (i - tn.begin[j]);
        *p = *p - step_size * (m / denom);
    }
}

// ---------------------------------------------------------------------------------------------------
// host-side orchestration (device-pointer API)
// ---------------------------------------------------------------------------------------------------
static int make_frame(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* opt, void* workspace,
                      size_t workspace_bytes, Frame* f, bool step = false) {
    if (!scene || !camera || !opt) return fail(SURF_ERR_BAD_ARG, "null scene/camera/options");
    std::string err;
    if (!build_scene_view(*scene, &f->sc, &err)) return fail(SURF_ERR_BAD_ARG, err);
    if (!check_camera(*camera, &err)) return fail(SURF_ERR_BAD_ARG, err);
    const int N = camera->width * camera->height;
    f->pix0 = opt->pixel_begin;
    int p1 = opt->pixel_end;
    if (f->pix0 == 0 && p1 == 0) p1 = N;
    if (f->pix0 < 0 || p1 > N || p1 <= f->pix0) return fail(SURF_ERR_BAD_ARG, "bad pixel range");
    f->n = p1 - f->pix0;
    f->cam = CamArgs{camera->eye, camera->at, camera->up, camera->proj, camera->width, camera->height,
                     camera->fovy, camera->focal_length, camera->near_clip, camera->far_clip};
    f->fl = ShadeFlags{opt->double_sided, opt->use_quartic};
    f->shadow = opt->shadow != 0;
    if (f->sc.n_lights > kMaxShadowLights && f->shadow) return fail(SURF_ERR_UNSUPPORTED, "shadow rays support at most 254 lights");
    if (!workspace) return fail(SURF_ERR_WORKSPACE, "null workspace");
    carve(workspace, f->sc.total, f->n, f->sc.n_lights, f->shadow, &f->ws, camera->proj != 0, step);
    if (f->ws.bytes > workspace_bytes)       // no room for the candidate queue: the staged intersection kernel needs none
        carve(workspace, f->sc.total, f->n, f->sc.n_lights, f->shadow, &f->ws, camera->proj != 0, step, false);
    if (f->ws.bytes > workspace_bytes) return fail(SURF_ERR_WORKSPACE, "workspace too small; see surf_workspace_bytes");
    return SURF_OK;
}

static int run_common_prologue(const Frame& f, float* ray_out, cudaStream_t st, bool need_rays_and_zbuf) {
    k_setup<<<1, 32, 0, st>>>(f.cam, f.ws.cam);
    SURF_LAUNCHED("k_setup");
    if (need_rays_and_zbuf) {
        k_raygen<<<(f.n + 255) / 256, 256, 0, st>>>(f.ws.cam, f.pix0, f.n, f.ws.rays, ray_out, f.ws.zbuf);
        SURF_LAUNCHED("k_raygen");
    }
    return SURF_OK;
}

// the MSE loss of an inverse-rendering step, fused into the shading epilogue (surf_step_mse)
struct StepLoss {
    const float* target; float* g_image; double* loss_acc; float scale;
    cudaEvent_t target_ready = nullptr;      // optional: the target arrives on another stream; the shading kernel waits for it
};

// k_shade / k_shade_batch with 4 pixels per thread (all-vector global accesses) on large frames, 1 on small ones
static int launch_shade(ShadeParams& sh, const StepLoss* loss, const BatchArgs* ba, int n_scenes, cudaStream_t st) {
    sh.target = loss ? loss->target : nullptr;
    sh.g_image = loss ? loss->g_image : nullptr;
    sh.loss_acc = loss ? loss->loss_acc : nullptr;
    sh.loss_scale = loss ? loss->scale : 0.f;
    // PX pixels per thread: 2 (vector accesses) on frames large enough to fill the machine that way, else 1
    const int px = ((sh.n % 2 == 0) && (long long)sh.n * n_scenes >= 256 * 1024) ? kShadePX : 1;
    const int per_cta = 256 * px;
    const int tiles = (sh.n + per_cta - 1) / per_cta;
    const int wave = sm_count() * kShadeBlocksPerSM;                 // resident CTAs: persistent, one wave
    const int ns = ba ? n_scenes : 1;
    const dim3 grid(std::max(1, std::min(tiles, (wave + ns - 1) / ns)), ns);
    timer_mark(1, 0, st);
    if (ba) {
        if (px > 1) k_shade_batch<kShadePX><<<grid, 256, 0, st>>>(sh, *ba);
        else k_shade_batch<1><<<grid, 256, 0, st>>>(sh, *ba);
    } else {
        if (px > 1) k_shade<kShadePX><<<grid, 256, 0, st>>>(sh);
        else k_shade<1><<<grid, 256, 0, st>>>(sh);
    }
    timer_mark(1, 1, st);
    SURF_LAUNCHED(ba ? "k_shade_batch" : "k_shade");
    return SURF_OK;
}

static int forward_impl(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* opt, void* workspace,
                        size_t workspace_bytes, const SurfOutputs* out, cudaStream_t st, const StepLoss* loss = nullptr) {
    if (!out) return fail(SURF_ERR_BAD_ARG, "null outputs");
    Frame f;
    int rc = make_frame(scene, camera, opt, workspace, workspace_bytes, &f);
    if (rc) return rc;
    if ((rc = run_common_prologue(f, out->ray_dir, st, true))) return rc;
    // per-primitive records: the filters of k_intersect, the unit normals k_shade reads back, and the index-range
    // check.  (Orthographic frames replace them with origin-independent records inside run_intersect_ortho.)
    k_prep<<<(f.sc.total + 255) / 256, 256, 0, st>>>(f.sc, f.ws.cam, f.ws.packed);
    SURF_LAUNCHED("k_prep");
    if (f.cam.proj == 0 && opt->math_mode == 3) {
        k_prep_screen<<<(f.sc.total + 255) / 256, 256, 0, st>>>(f.sc, f.ws.cam, f.ws.circ);
        SURF_LAUNCHED("k_prep_screen");
    }
    if ((rc = run_intersect(f, opt, st))) return rc;
    if (f.shadow) {
        ShadowParams sp{f.sc, f.ws.cam, f.ws.rays, f.ws.zbuf, f.ws.vis, f.pix0, f.n};
        if (opt->math_mode == 1) {       // exact-only brute force kept for cross-checking
            dim3 grid((f.n + 127) / 128, f.sc.n_lights);
            k_shadow<<<grid, 128, 0, st>>>(sp);
            SURF_LAUNCHED("k_shadow");
        } else {
            // all lights at once: one ray list of (light, hit pixel) pairs, one intersection launch
            const int L = f.sc.n_lights;
            const size_t cap = (size_t)f.n * L;
            int* n_live = (int*)f.ws.obound + 1;                      // [0] = origin bound, [1..] = live-ray counter(s)
            int* slot_of = (int*)(f.ws.gray + 7 * cap);
            const int per_light = opt->math_mode != 2 ? 1 : 0;   // math_mode 2 keeps the per-ray-origin filter (cross-check)
            SURF_CUDA(cudaMemsetAsync(f.ws.obound, 0, 4 * (1 + (size_t)L), st));
            k_rays_shadow<<<dim3((f.n + 255) / 256, L), 256, 0, st>>>(sp, cap, f.ws.gray, f.ws.zbuf2, f.ws.obound, n_live, slot_of,
                                                                     per_light);
            SURF_LAUNCHED("k_rays_shadow");
            if (per_light) {
                if ((rc = run_intersect_shadow(f, n_live, st))) return rc;
            } else if ((rc = run_intersect_rays_shadow(f, f.ws.zbuf2, st, n_live, (long long)cap))) return rc;
            k_shadow_resolve<<<(unsigned)((cap + 255) / 256), 256, 0, st>>>(f.ws.zbuf, f.ws.zbuf2, slot_of, f.n, cap, f.ws.vis);
            SURF_LAUNCHED("k_shadow_resolve");
        }
    }
    ShadeParams sh;
    sh.sc = f.sc; sh.cam = f.ws.cam; sh.packed = f.ws.packed; sh.rays = f.ws.rays; sh.zbuf = f.ws.zbuf;
    sh.vis = f.shadow ? f.ws.vis : nullptr;
    sh.pix0 = f.pix0; sh.n = f.n; sh.fl = f.fl;
    sh.image = out->image; sh.depth = out->depth; sh.normal = out->normal; sh.pos = out->pos;
    sh.nearest = (long long*)out->nearest;
    if (loss && loss->target_ready) SURF_CUDA(cudaStreamWaitEvent(st, loss->target_ready, 0));
    return launch_shade(sh, loss, nullptr, 1, st);
}

static int backward_impl(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* opt, void* workspace,
                         size_t workspace_bytes, const int64_t* nearest, const float* depth, const SurfOutGrads* og,
                         const SurfSceneGrads* sg, cudaStream_t st, bool workspace_is_warm, float* step_loss = nullptr) {
    if (!nearest || !depth || !og || !sg) return fail(SURF_ERR_BAD_ARG, "null nearest/depth/out_grads/scene_grads");
    Frame f;
    int rc = make_frame(scene, camera, opt, workspace, workspace_bytes, &f);
    if (rc) return rc;
    const SlotMap sm = slot_map(f.sc.n_materials, f.sc.n_lights, f.sc.n_colors);     // slots past kMaxAccSlots: see BwdAcc
    if (!workspace_is_warm) {
        // recompute camera state and rays (the forward call may have used a different workspace)
        k_setup<<<1, 32, 0, st>>>(f.cam, f.ws.cam);
        SURF_LAUNCHED("k_setup");
        k_raygen<<<(f.n + 255) / 256, 256, 0, st>>>(f.ws.cam, f.pix0, f.n, f.ws.rays, nullptr, f.ws.zbuf);
        SURF_LAUNCHED("k_raygen");
        if (f.shadow) return fail(SURF_ERR_UNSUPPORTED, "shadow backward needs the forward workspace (visibility)");
    }
    SURF_CUDA(cudaMemsetAsync(f.ws.acc, 0, sizeof(double) * kMaxAccSlots, st));
    SURF_CUDA(cudaMemsetAsync(f.ws.prim_acc, 0, sizeof(double) * 7 * (size_t)f.sc.total, st));
    BackwardParams bp;
    bp.sc = f.sc; bp.cam = f.ws.cam; bp.rays = f.ws.rays; bp.vis = f.shadow ? f.ws.vis : nullptr;
    bp.nearest = (const long long*)nearest; bp.depth = depth;
    bp.g_image = og->image; bp.g_depth = og->depth; bp.g_normal = og->normal; bp.g_pos = og->pos;
    bp.pix0 = f.pix0; bp.n = f.n; bp.fl = f.fl; bp.sm = sm; bp.acc = f.ws.acc; bp.prim_acc = f.ws.prim_acc;
    for (int s = 0; s < kMaxSets; ++s) {
        bp.gp.prim_pos[s] = sg->sets[s].pos; bp.gp.prim_normal[s] = sg->sets[s].normal; bp.gp.prim_radius[s] = sg->sets[s].radius;
    }
    bp.gp.light_pos = sg->light_pos; bp.gp.atten = sg->light_attenuation; bp.gp.ambient = sg->ambient;
    bp.gp.colors = sg->colors; bp.gp.albedo = sg->albedo; bp.gp.coeffs = sg->coeffs; bp.gp.gamma = sg->gamma;
    timer_mark(2, 0, st);
    k_backward<<<std::min((f.n + 127) / 128, sm_count() * 16), 128, 0, st>>>(bp);
    timer_mark(2, 1, st);
    SURF_LAUNCHED("k_backward");
    FinalizeParams fp{bp.gp, sm, f.ws.acc, f.sc.n_materials, f.sc.n_lights, f.sc.n_colors, f.sc.light_pos_stride,
                      f.sc, f.ws.prim_acc, f.ws.loss_acc, step_loss};
    k_backward_finalize<<<(f.sc.total + sm.total + 127) / 128, 128, 0, st>>>(fp);
    SURF_LAUNCHED("k_backward_finalize");
    return SURF_OK;
}

// ---------------------------------------------------------------------------------------------------
// render_splats_along_ray: host orchestration (single scene = a batch of one)
// ---------------------------------------------------------------------------------------------------
static int splat_frame(int n_scenes, const SurfScene* scene, const SurfCamera* camera, const SurfOptions* opt, const SurfSplats* sp,
                       const SurfSplatBatch* batch, void* workspace, size_t ws_stride, SplatParams* p, CamArgs* cam,
                       SplatWorkspace* ws) {
    if (!scene || !camera || !opt || !sp) return fail(SURF_ERR_BAD_ARG, "null scene/camera/options/splats");
    if (n_scenes < 1 || n_scenes > 65535) return fail(SURF_ERR_BAD_ARG, "batch of 1..65535 scenes");
    std::string err;
    SceneView sc;
    if (!build_scene_view(*scene, &sc, &err, true)) return fail(SURF_ERR_BAD_ARG, err);
    if (!check_camera(*camera, &err)) return fail(SURF_ERR_BAD_ARG, err);
    if (camera->proj != 0) return fail(SURF_ERR_UNSUPPORTED, "render_splats_along_ray is defined for the perspective frustum");
    if (scene->light_pos_stride != 4) return fail(SURF_ERR_BAD_ARG, "along-ray lights must be homogeneous [L,4] (torch.mm with the 4x4 view matrix)");
    const int K = sp->samples > 1 ? sp->samples : 1;
    const int n_src = camera->width * camera->height;
    if (sp->pos) {
        if (K != 1 || sp->estimate_normals) return fail(SURF_ERR_BAD_ARG, "explicit fragment positions exclude samples > 1 and normal estimation");
        if (sp->count < 1) return fail(SURF_ERR_BAD_ARG, "no splats");
    } else {
        if (sp->count != n_src) return fail(SURF_ERR_BAD_ARG, "one splat per pixel: count must equal width*height");
        if (!sp->z) return fail(SURF_ERR_BAD_ARG, "splat depths are required");
        if (sp->z_stride != 1 && sp->z_stride != 3) return fail(SURF_ERR_BAD_ARG, "z_stride must be 1 or 3");
    }
    if (sp->estimate_normals < 0 || sp->estimate_normals > 2) return fail(SURF_ERR_BAD_ARG, "estimate_normals: 0 (given), 1 (plane), 2 (avg_normal)");
    if (!sp->estimate_normals) {
        if (!sp->normal) return fail(SURF_ERR_BAD_ARG, "splat normals are required unless they are estimated");
        if (sp->normal_stride != 3 && sp->normal_stride != 4) return fail(SURF_ERR_BAD_ARG, "normal_stride must be 3 or 4");
    }
    if (K > 1 && !sp->material_idx) return fail(SURF_ERR_BAD_ARG, "supersampling needs material_idx (renderer.py:605)");
    if (!workspace) return fail(SURF_ERR_WORKSPACE, "null workspace");
    const int n_in = sp->pos ? sp->count : n_src;
    carve_splats(workspace, n_in, sc.n_lights, ws);
    if (ws->bytes > ws_stride) return fail(SURF_ERR_WORKSPACE, "workspace too small; see surf_splats_workspace_bytes");
    *cam = CamArgs{camera->eye, camera->at, camera->up, 0, camera->width, camera->height, camera->fovy,
                   camera->focal_length, camera->near_clip, camera->far_clip};
    std::memset(p, 0, sizeof(*p));
    p->sc = sc;
    p->sc.light_pos = ws->light_cc; p->sc.light_pos_stride = 3; p->sc.gamma = nullptr;
    p->cam = ws->cam;
    p->z = sp->z; p->z_stride = sp->z_stride ? sp->z_stride : 1;
    p->estimate = sp->estimate_normals;
    p->normal = p->estimate ? ws->nest : sp->normal;
    p->normal_stride = p->estimate ? 3 : sp->normal_stride;
    p->mat = sp->material_idx; p->vis = sp->light_vis;
    p->pos_in = sp->pos;
    p->W = camera->width; p->H = camera->height; p->K = K;
    p->n_src = n_in;
    p->n = sp->pos ? sp->count : n_src * K * K;
    p->fl = ShadeFlags{sp->ndc_stride ? opt->double_sided : 0, opt->use_quartic, sp->ndc_stride ? 1 : 0};
    if (sp->ndc_stride) {
        if (!sp->pos || (sp->ndc_stride != 3 && sp->ndc_stride != 4)) return fail(SURF_ERR_BAD_ARG, "NDC splats: pos [count, 3|4] required");
        // inverse of the right-handed [-1, 1] perspective matrix (diffrend/torch/ops.py:19-68)
        const double th = tan(camera->fovy / 2.0), aspect = (double)camera->width / (double)camera->height;
        const double m22 = ((double)camera->near_clip + camera->far_clip) / ((double)camera->far_clip - camera->near_clip);
        const double m23 = -2.0 * camera->near_clip * camera->far_clip / ((double)camera->far_clip - camera->near_clip);
        p->ndc_stride = sp->ndc_stride;
        p->ndc_a0 = (float)(aspect * th); p->ndc_a1 = (float)th; p->ndc_b = (float)(1.0 / m23); p->ndc_c = (float)(-m22 / m23);
    }
    p->nest = ws->nest; p->gpos_src = ws->gpos_src; p->gn_src = ws->gn_src; p->minmax = ws->minmax;
    p->far_clip = camera->far_clip;
    p->sm = slot_map(sc.n_materials, sc.n_lights, sc.n_colors);
    p->acc = ws->acc;
    if (p->sm.total > kMaxAccSlots) return fail(SURF_ERR_UNSUPPORTED, "too many materials/lights/colours for the along-ray backward accumulators");
    SplatBatchArgs& ba = p->ba;
    std::memset(&ba, 0, sizeof(ba));
    if (batch) {
        ba.z = batch->z; ba.normal = batch->normal; ba.mat = batch->material_idx; ba.vis = batch->light_vis;
        ba.light_pos = batch->light_pos; ba.eye = batch->eye;
    }
    ba.ws_stride = n_scenes > 1 ? (long long)ws_stride : 0;
    ba.out_stride = p->n;
    return SURF_OK;
}

static int splats_forward_impl(int n_scenes, const SurfScene* scene, const SurfCamera* camera, const SurfOptions* opt,
                               const SurfSplats* sp, const SurfSplatBatch* batch, void* workspace, size_t ws_stride,
                               const SurfOutputs* out, cudaStream_t st) {
    if (!out) return fail(SURF_ERR_BAD_ARG, "null outputs");
    SplatParams p; CamArgs cam; SplatWorkspace ws;
    int rc = splat_frame(n_scenes, scene, camera, opt, sp, batch, workspace, ws_stride, &p, &cam, &ws);
    if (rc) return rc;
    if (sp->norm_depth && !out->depth) return fail(SURF_ERR_BAD_ARG, "norm_depth needs out->depth");
    const unsigned B = (unsigned)n_scenes;
    k_splat_setup<<<B, 64, 0, st>>>(cam, p.ba.eye, ws.cam, scene->light_pos, p.ba.light_pos, scene->n_lights, ws.light_cc,
                                    ws.minmax, p.ba.ws_stride);
    SURF_LAUNCHED("k_splat_setup");
    if (p.estimate) {
        k_splat_normals<<<dim3((p.n_src + 255) / 256, B), 256, 0, st>>>(p);
        SURF_LAUNCHED("k_splat_normals");
    }
    p.image = out->image; p.depth = out->depth; p.normal_out = out->normal; p.pos = out->pos;
    p.norm_depth = sp->norm_depth;
    timer_mark(1, 0, st);
    k_splat_forward<<<dim3((p.n + 255) / 256, B), 256, 0, st>>>(p);
    timer_mark(1, 1, st);
    SURF_LAUNCHED("k_splat_forward");
    if (p.norm_depth) {
        k_splat_normdepth<<<dim3((p.n + 255) / 256, B), 256, 0, st>>>(p);
        SURF_LAUNCHED("k_splat_normdepth");
    }
    return SURF_OK;
}

static int splats_backward_impl(int n_scenes, const SurfScene* scene, const SurfCamera* camera, const SurfOptions* opt,
                                const SurfSplats* sp, const SurfSplatBatch* batch, void* workspace, size_t ws_stride,
                                const SurfOutGrads* og, const SurfSceneGrads* sg, const SurfSplatGrads* spg, cudaStream_t st) {
    if (!og || !sg || !spg) return fail(SURF_ERR_BAD_ARG, "null out_grads/scene_grads/splat_grads");
    SplatParams p; CamArgs cam; SplatWorkspace ws;
    int rc = splat_frame(n_scenes, scene, camera, opt, sp, batch, workspace, ws_stride, &p, &cam, &ws);
    if (rc) return rc;
    const unsigned B = (unsigned)n_scenes;
    // the workspace may be a fresh one: recompute camera, lights and (if estimated) the normals
    k_splat_setup<<<B, 64, 0, st>>>(cam, p.ba.eye, ws.cam, scene->light_pos, p.ba.light_pos, scene->n_lights, ws.light_cc,
                                    ws.minmax, p.ba.ws_stride);
    SURF_LAUNCHED("k_splat_setup");
    if (p.estimate) {
        k_splat_normals<<<dim3((p.n_src + 255) / 256, B), 256, 0, st>>>(p);
        SURF_LAUNCHED("k_splat_normals");
    }
    const size_t stride = n_scenes > 1 ? ws_stride : 0;
    SURF_CUDA(cudaMemset2DAsync(ws.acc, stride ? stride : sizeof(double) * kMaxAccSlots, 0, sizeof(double) * kMaxAccSlots, B, st));
    const bool scatter = p.K > 1 || p.estimate != 0;
    if (scatter) {     // gpos_src and gn_src are adjacent: one memset per scene
        const size_t bytes = (size_t)((char*)ws.gn_src - (char*)ws.gpos_src) + (size_t)p.n_src * 12;
        SURF_CUDA(cudaMemset2DAsync(ws.gpos_src, stride ? stride : bytes, 0, bytes, B, st));
    }
    p.g_image = og->image; p.g_depth = og->depth; p.g_normal = og->normal; p.g_pos = og->pos;
    p.gz = spg->z; p.gnormal = spg->normal; p.gpos = spg->pos;
    for (int s = 0; s < kMaxSets; ++s) p.gp.prim_pos[s] = p.gp.prim_normal[s] = p.gp.prim_radius[s] = nullptr;
    p.gp.light_pos = nullptr; p.gp.atten = sg->light_attenuation; p.gp.ambient = sg->ambient;
    p.gp.colors = sg->colors; p.gp.albedo = sg->albedo; p.gp.coeffs = sg->coeffs; p.gp.gamma = nullptr;
    const int gx = std::max(1, std::min((p.n + kBwdThreads - 1) / kBwdThreads, (sm_count() * 8 + n_scenes - 1) / n_scenes));
    timer_mark(2, 0, st);
    k_splat_backward<<<dim3(gx, B), kBwdThreads, 0, st>>>(p);
    timer_mark(2, 1, st);
    SURF_LAUNCHED("k_splat_backward");
    if (p.estimate) {
        k_splat_normals_backward<<<dim3((p.n_src + 255) / 256, B), 256, 0, st>>>(p);
        SURF_LAUNCHED("k_splat_normals_backward");
    }
    if (scatter) {
        k_splat_src_finalize<<<dim3((p.n_src + 255) / 256, B), 256, 0, st>>>(p);
        SURF_LAUNCHED("k_splat_src_finalize");
    }
    SplatFinalizeParams fp;
    fp.gp = p.gp;
    fp.gp.light_pos = sg->light_pos;
    fp.sm = p.sm; fp.acc = ws.acc; fp.cam = ws.cam; fp.L = scene->n_lights;
    fp.ws_stride = p.ba.ws_stride; fp.light_stride = p.ba.light_pos;
    k_splat_finalize<<<dim3((p.sm.total + 127) / 128, B), 128, 0, st>>>(fp);
    SURF_LAUNCHED("k_splat_finalize");
    return SURF_OK;
}

}  // namespace surf

// ===================================================================================================
// C ABI
// ===================================================================================================
using namespace surf;

namespace surf { struct HostState; }
struct SurfContext {
    int device;
    cudaStream_t stream;
    cudaStream_t copy_stream;    // the target image of a step is uploaded here, next to the intersection kernels
    cudaEvent_t target_ready;
    char* stage_up; char* stage_down;     // pinned staging (kStageBytes each): the small arrays of a call travel as ONE copy per direction
    void* arena; size_t arena_bytes;
    uint64_t h2d, d2h;
    surf::HostState* state;      // the call in flight between surf_step_host_begin and surf_step_host_end
};

extern "C" {

int surf_abi_version(void) { return SURF_ABI_VERSION; }
const char* surf_last_error(void) { return g_error.c_str(); }
int surf_last_launch_count(void) { return g_launches.load(); }
void surf_set_kernel_timing(int32_t enabled) {
    g_timers.enabled = enabled != 0;
    for (int k = 0; k < 3; ++k) g_timers.count[k] = 0;
}
double surf_last_kernel_ms(int32_t which) {
    if (which < 0 || which > 2 || !g_timers.created || g_timers.count[which] == 0) return -1.0;
    return timer_ms(which, g_timers.count[which] - 1);
}
double surf_mean_kernel_ms(int32_t which, int32_t* launches) {
    if (launches) *launches = 0;
    if (which < 0 || which > 2 || !g_timers.created || g_timers.count[which] == 0) return -1.0;
    const long long n = g_timers.count[which];
    const long long first = n > kTimerRing ? n - kTimerRing : 0;
    double sum = 0.0;
    int used = 0;
    for (long long i = first; i < n; ++i) {
        const double ms = timer_ms(which, i);
        if (ms >= 0.0) { sum += ms; ++used; }
    }
    if (launches) *launches = used;
    return used ? sum / used : -1.0;
}

size_t surf_workspace_bytes(int32_t total_prims, int32_t n_pixels, int32_t n_lights, int32_t shadow) {
    return std::max(surf_workspace_bytes_ex(total_prims, n_pixels, n_lights, shadow, 1, 0),
                    surf_workspace_bytes_ex(total_prims, n_pixels, n_lights, shadow, 0, 0));
}
size_t surf_workspace_bytes_ex(int32_t total_prims, int32_t n_pixels, int32_t n_lights, int32_t shadow,
                               int32_t orthographic, int32_t step) {
    Workspace ws;
    carve(nullptr, total_prims, n_pixels, n_lights, shadow != 0, &ws, orthographic != 0, step != 0);
    return ws.bytes;
}

int surf_check_indices(const void* workspace, void* cuda_stream) {
    if (!workspace) return fail(SURF_ERR_BAD_ARG, "null workspace");
    Workspace ws;
    carve(const_cast<void*>(workspace), 0, 0, 0, false, &ws, false);
    int flag = 0;
    SURF_CUDA(cudaMemcpyAsync(&flag, &ws.cam->bad_index, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)cuda_stream));
    SURF_CUDA(cudaStreamSynchronize((cudaStream_t)cuda_stream));
    if (flag) return fail(SURF_ERR_BAD_ARG, "material_idx or light color_idx out of range (the reference's index_select raises IndexError)");
    return SURF_OK;
}

int surf_forward(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* options, void* workspace,
                 size_t workspace_bytes, const SurfOutputs* out, void* cuda_stream) {
    g_launches = 0;
    return forward_impl(scene, camera, options, workspace, workspace_bytes, out, (cudaStream_t)cuda_stream);
}

int surf_backward(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* options, void* workspace,
                  size_t workspace_bytes, const int64_t* nearest, const float* depth, const SurfOutGrads* out_grads,
                  const SurfSceneGrads* scene_grads, void* cuda_stream) {
    g_launches = 0;
    // options->forced_nearest == 2: the caller guarantees `workspace` still holds this frame's forward state
    const bool warm = options && options->forced_nearest == 2;
    return backward_impl(scene, camera, options, workspace, workspace_bytes, nearest, depth, out_grads, scene_grads,
                         (cudaStream_t)cuda_stream, warm);
}

// ---- batches of independent scenes (GAN real-sample batches, GAN/gan.py:326-377; BASELINE configs[3]) ----
// One host call launches every scene's kernels back to back, fanned out over a small pool of internal streams that
// fork from / join to the caller's stream with events, so the small per-scene kernels overlap and the host pays one
// call instead of one Python round trip per scene.
}  // extern "C"

namespace {
constexpr int kBatchStreams = 16;
struct StreamPool {
    int device = -1;
    cudaStream_t st[kBatchStreams] = {};
    cudaEvent_t fork = nullptr, join[kBatchStreams] = {};
    std::mutex in_use;       // one fork / launch / join sequence at a time: the events are shared by all callers
};
static std::mutex g_pool_mutex;
static std::vector<StreamPool*> g_pools;

static StreamPool* stream_pool() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    for (StreamPool* p : g_pools)
        if (p->device == dev) return p;
    StreamPool* p = new StreamPool();
    p->device = dev;
    bool ok = cudaEventCreateWithFlags(&p->fork, cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; k < kBatchStreams && ok; ++k)
        ok = cudaStreamCreateWithFlags(&p->st[k], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&p->join[k], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { delete p; return nullptr; }
    g_pools.push_back(p);
    return p;
}

template <class PerScene>
static int run_batch(int32_t n_scenes, cudaStream_t caller, PerScene&& per_scene) {
    if (n_scenes < 1) return fail(SURF_ERR_BAD_ARG, "empty batch");
    StreamPool* pool = stream_pool();
    if (!pool) return fail(SURF_ERR_CUDA, "could not create the batch stream pool");
    const int lanes = std::min<int>(kBatchStreams, n_scenes);
    // Two threads in surf_*_batch on one device (forward on the main thread, backward on autograd's) would otherwise
    // re-record `fork` / `join[k]` between the other's record and wait.
    std::lock_guard<std::mutex> busy(pool->in_use);
    SURF_CUDA(cudaEventRecord(pool->fork, caller));
    for (int k = 0; k < lanes; ++k) SURF_CUDA(cudaStreamWaitEvent(pool->st[k], pool->fork, 0));
    int rc = SURF_OK;
    for (int b = 0; b < n_scenes && rc == SURF_OK; ++b) rc = per_scene(b, pool->st[b % lanes]);
    for (int k = 0; k < lanes; ++k) {       // always re-join, also on error, so the caller's stream stays ordered
        cudaEventRecord(pool->join[k], pool->st[k]);
        cudaStreamWaitEvent(caller, pool->join[k], 0);
    }
    return rc;
}
}  // namespace

extern "C" {

int surf_forward_batch(int32_t n_scenes, const SurfScene* scenes, const SurfCamera* cameras, const SurfOptions* options,
                       void* const* workspaces, const size_t* workspace_bytes, const SurfOutputs* outs, void* cuda_stream) {
    g_launches = 0;
    if (!scenes || !cameras || !options || !workspaces || !workspace_bytes || !outs)
        return fail(SURF_ERR_BAD_ARG, "null batch argument");
    return run_batch(n_scenes, (cudaStream_t)cuda_stream, [&](int b, cudaStream_t st) {
        return forward_impl(&scenes[b], &cameras[b], options, workspaces[b], workspace_bytes[b], &outs[b], st);
    });
}

int surf_backward_batch(int32_t n_scenes, const SurfScene* scenes, const SurfCamera* cameras, const SurfOptions* options,
                        void* const* workspaces, const size_t* workspace_bytes, const int64_t* const* nearest,
                        const float* const* depth, const SurfOutGrads* out_grads, const SurfSceneGrads* scene_grads,
                        void* cuda_stream) {
    g_launches = 0;
    if (!scenes || !cameras || !options || !workspaces || !workspace_bytes || !nearest || !depth || !out_grads || !scene_grads)
        return fail(SURF_ERR_BAD_ARG, "null batch argument");
    const bool warm = options->forced_nearest == 2;
    return run_batch(n_scenes, (cudaStream_t)cuda_stream, [&](int b, cudaStream_t st) {
        return backward_impl(&scenes[b], &cameras[b], options, workspaces[b], workspace_bytes[b], nearest[b], depth[b],
                             &out_grads[b], &scene_grads[b], st, warm);
    });
}

}  // extern "C"
namespace {
template <class T>
inline T* advance(T* base, int64_t elems) { return base ? base + elems : nullptr; }

// scene / camera / gradient structs of element b of a strided batch
void strided_scene(const SurfScene& s0, const SurfBatchLayout& L, int64_t b, SurfScene* s) {
    *s = s0;
    for (int k = 0; k < s0.n_sets && k < SURF_MAX_SETS; ++k) {
        s->sets[k].pos = advance(s0.sets[k].pos, b * L.set_pos[k]);
        s->sets[k].normal = advance(s0.sets[k].normal, b * L.set_normal[k]);
        s->sets[k].radius = advance(s0.sets[k].radius, b * L.set_radius[k]);
        s->sets[k].material_idx = advance(s0.sets[k].material_idx, b * L.set_material_idx[k]);
    }
    s->light_pos = advance(s0.light_pos, b * L.light_pos);
    s->light_color_idx = advance(s0.light_color_idx, b * L.light_color_idx);
    s->light_attenuation = advance(s0.light_attenuation, b * L.light_attenuation);
    s->ambient = advance(s0.ambient, b * L.ambient);
    s->colors = advance(s0.colors, b * L.colors);
    s->albedo = advance(s0.albedo, b * L.albedo);
    s->coeffs = advance(s0.coeffs, b * L.coeffs);
    s->gamma = advance(s0.gamma, b * L.gamma);
}
void strided_camera(const SurfCamera& c0, const SurfBatchLayout& L, int64_t b, SurfCamera* c) {
    *c = c0;
    c->eye = advance(c0.eye, b * L.eye);
    c->at = advance(c0.at, b * L.at);
    c->up = advance(c0.up, b * L.up);
}
void strided_grads(const SurfSceneGrads& g0, int n_sets, const SurfBatchLayout& L, int64_t b, SurfSceneGrads* g) {
    *g = g0;
    for (int k = 0; k < n_sets && k < SURF_MAX_SETS; ++k) {
        g->sets[k].pos = advance(g0.sets[k].pos, b * L.set_pos[k]);
        g->sets[k].normal = advance(g0.sets[k].normal, b * L.set_normal[k]);
        g->sets[k].radius = advance(g0.sets[k].radius, b * L.set_radius[k]);
    }
    g->light_pos = advance(g0.light_pos, b * L.light_pos);
    g->light_attenuation = advance(g0.light_attenuation, b * L.light_attenuation);
    g->ambient = advance(g0.ambient, b * L.ambient);
    g->colors = advance(g0.colors, b * L.colors);
    g->albedo = advance(g0.albedo, b * L.albedo);
    g->coeffs = advance(g0.coeffs, b * L.coeffs);
    g->gamma = advance(g0.gamma, b * L.gamma);
}
int64_t frame_pixels(const SurfCamera& c, const SurfOptions& o) {
    return (o.pixel_begin == 0 && o.pixel_end == 0) ? (int64_t)c.width * c.height : (int64_t)o.pixel_end - o.pixel_begin;
}
BatchArgs batch_args(const SurfBatchLayout& L, int n_scenes, size_t ws_stride) {
    BatchArgs ba;
    ba.n_scenes = n_scenes;
    ba.ws_stride = (long long)ws_stride;
    for (int k = 0; k < kMaxSets; ++k) {
        ba.set_pos[k] = L.set_pos[k]; ba.set_normal[k] = L.set_normal[k];
        ba.set_radius[k] = L.set_radius[k]; ba.set_mat[k] = L.set_material_idx[k];
    }
    ba.light_pos = L.light_pos; ba.light_color_idx = L.light_color_idx; ba.light_atten = L.light_attenuation;
    ba.ambient = L.ambient; ba.colors = L.colors; ba.albedo = L.albedo; ba.coeffs = L.coeffs; ba.gamma = L.gamma;
    ba.eye = L.eye; ba.at = L.at; ba.up = L.up;
    return ba;
}

// the fused batch kernels cover the GAN case: perspective, no shadow rays, ray-plane filter modes
bool fused_batch_ok(const SurfCamera& c, const SurfOptions& o, int n_scenes) {
    return c.proj == 0 && !o.shadow && o.math_mode != 3 && n_scenes <= 65535;
}

// Whole batch in FIVE launches: every kernel gets a scene dimension (blockIdx.y; k_intersect_batch: the work item).
int forward_strided_fused(int n_scenes, const SurfScene* scene0, const SurfCamera* camera0, const SurfBatchLayout* layout,
                          const SurfOptions* opt, void* workspace, size_t ws_stride, const SurfOutputs* out0, cudaStream_t st,
                          const StepLoss* loss = nullptr) {
    Frame f;
    int rc = make_frame(scene0, camera0, opt, workspace, ws_stride, &f);
    if (rc) return rc;
    const BatchArgs ba = batch_args(*layout, n_scenes, ws_stride);
    const unsigned B = (unsigned)n_scenes;
    k_setup_batch<<<B, 32, 0, st>>>(f.cam, ba, f.ws.cam);
    SURF_LAUNCHED("k_setup_batch");
    k_raygen_batch<<<dim3((f.n + 255) / 256, B), 256, 0, st>>>(f.ws.cam, ba, f.pix0, f.n, f.ws.rays, out0->ray_dir,
                                                              3LL * f.n, f.ws.zbuf);
    SURF_LAUNCHED("k_raygen_batch");
    k_prep_batch<<<dim3((f.sc.total + 255) / 256, B), 256, 0, st>>>(f.sc, ba, f.ws.cam, f.ws.packed);
    SURF_LAUNCHED("k_prep_batch");
    if ((rc = run_intersect(f, opt, st, &ba))) return rc;
    ShadeParams sh;
    sh.sc = f.sc; sh.cam = f.ws.cam; sh.packed = f.ws.packed; sh.rays = f.ws.rays; sh.zbuf = f.ws.zbuf; sh.vis = nullptr;
    sh.pix0 = f.pix0; sh.n = f.n; sh.fl = f.fl;
    sh.image = out0->image; sh.depth = out0->depth; sh.normal = out0->normal; sh.pos = out0->pos;
    sh.nearest = (long long*)out0->nearest;
    return launch_shade(sh, loss, &ba, n_scenes, st);
}

int backward_strided_fused(int n_scenes, const SurfScene* scene0, const SurfCamera* camera0, const SurfBatchLayout* layout,
                           const SurfOptions* opt, void* workspace, size_t ws_stride, const int64_t* nearest0,
                           const float* depth0, const SurfOutGrads* og, const SurfSceneGrads* sg, cudaStream_t st, bool warm,
                           float* step_loss = nullptr) {
    Frame f;
    int rc = make_frame(scene0, camera0, opt, workspace, ws_stride, &f);
    if (rc) return rc;
    const SlotMap sm = slot_map(f.sc.n_materials, f.sc.n_lights, f.sc.n_colors);     // slots past kMaxAccSlots: see BwdAcc
    const BatchArgs ba = batch_args(*layout, n_scenes, ws_stride);
    const unsigned B = (unsigned)n_scenes;
    if (!warm) {
        k_setup_batch<<<B, 32, 0, st>>>(f.cam, ba, f.ws.cam);
        SURF_LAUNCHED("k_setup_batch");
        k_raygen_batch<<<dim3((f.n + 255) / 256, B), 256, 0, st>>>(f.ws.cam, ba, f.pix0, f.n, f.ws.rays, nullptr, 0, f.ws.zbuf);
        SURF_LAUNCHED("k_raygen_batch");
    }
    // the accumulators of all scenes sit at the same offsets of their workspaces: two strided memsets
    SURF_CUDA(cudaMemset2DAsync(f.ws.acc, ws_stride, 0, sizeof(double) * kMaxAccSlots, B, st));
    SURF_CUDA(cudaMemset2DAsync(f.ws.prim_acc, ws_stride, 0, sizeof(double) * 7 * (size_t)f.sc.total, B, st));
    BackwardParams bp;
    bp.sc = f.sc; bp.cam = f.ws.cam; bp.rays = f.ws.rays; bp.vis = nullptr;
    bp.nearest = (const long long*)nearest0; bp.depth = depth0;
    bp.g_image = og->image; bp.g_depth = og->depth; bp.g_normal = og->normal; bp.g_pos = og->pos;
    bp.pix0 = f.pix0; bp.n = f.n; bp.fl = f.fl; bp.sm = sm; bp.acc = f.ws.acc; bp.prim_acc = f.ws.prim_acc;
    for (int s = 0; s < kMaxSets; ++s) {
        bp.gp.prim_pos[s] = sg->sets[s].pos; bp.gp.prim_normal[s] = sg->sets[s].normal; bp.gp.prim_radius[s] = sg->sets[s].radius;
    }
    bp.gp.light_pos = sg->light_pos; bp.gp.atten = sg->light_attenuation; bp.gp.ambient = sg->ambient;
    bp.gp.colors = sg->colors; bp.gp.albedo = sg->albedo; bp.gp.coeffs = sg->coeffs; bp.gp.gamma = sg->gamma;
    const int gx = std::max(1, std::min((f.n + 127) / 128, (sm_count() * 16 + n_scenes - 1) / n_scenes));
    timer_mark(2, 0, st);
    k_backward_batch<<<dim3(gx, B), 128, 0, st>>>(bp, ba);
    timer_mark(2, 1, st);
    SURF_LAUNCHED("k_backward_batch");
    FinalizeParams fp{bp.gp, sm, f.ws.acc, f.sc.n_materials, f.sc.n_lights, f.sc.n_colors, f.sc.light_pos_stride,
                      f.sc, f.ws.prim_acc, f.ws.loss_acc, step_loss};
    k_backward_finalize_batch<<<dim3((f.sc.total + sm.total + 127) / 128, B), 128, 0, st>>>(fp, ba);
    SURF_LAUNCHED("k_backward_finalize_batch");
    return SURF_OK;
}
}  // namespace
extern "C" {

int surf_forward_strided(int32_t n_scenes, const SurfScene* scene0, const SurfCamera* camera0,
                         const SurfBatchLayout* layout, const SurfOptions* options, void* workspace,
                         size_t workspace_bytes_per_scene, const SurfOutputs* out0, void* cuda_stream) {
    g_launches = 0;
    if (!scene0 || !camera0 || !layout || !options || !workspace || !out0) return fail(SURF_ERR_BAD_ARG, "null batch argument");
    if (workspace_bytes_per_scene % 256) return fail(SURF_ERR_BAD_ARG, "workspace_bytes_per_scene must be a multiple of 256");
    if (n_scenes < 1) return fail(SURF_ERR_BAD_ARG, "empty batch");
    if (fused_batch_ok(*camera0, *options, n_scenes))
        return forward_strided_fused(n_scenes, scene0, camera0, layout, options, workspace, workspace_bytes_per_scene, out0,
                                     (cudaStream_t)cuda_stream);
    const int64_t n = frame_pixels(*camera0, *options);
    const int64_t rd = camera0->proj == 0 ? 3 * n : 3;
    return run_batch(n_scenes, (cudaStream_t)cuda_stream, [&](int b, cudaStream_t st) {
        SurfScene s; SurfCamera c; SurfOutputs o;
        strided_scene(*scene0, *layout, b, &s);
        strided_camera(*camera0, *layout, b, &c);
        o.image = advance(out0->image, b * n * 3); o.depth = advance(out0->depth, b * n);
        o.normal = advance(out0->normal, b * n * 3); o.pos = advance(out0->pos, b * n * 3);
        o.nearest = advance(out0->nearest, b * n); o.ray_dir = advance(out0->ray_dir, b * rd);
        return forward_impl(&s, &c, options, (char*)workspace + (size_t)b * workspace_bytes_per_scene,
                            workspace_bytes_per_scene, &o, st);
    });
}

int surf_backward_strided(int32_t n_scenes, const SurfScene* scene0, const SurfCamera* camera0,
                          const SurfBatchLayout* layout, const SurfOptions* options, void* workspace,
                          size_t workspace_bytes_per_scene, const int64_t* nearest0, const float* depth0,
                          const SurfOutGrads* out_grads0, const SurfSceneGrads* grads0, void* cuda_stream) {
    g_launches = 0;
    if (!scene0 || !camera0 || !layout || !options || !workspace || !nearest0 || !depth0 || !out_grads0 || !grads0)
        return fail(SURF_ERR_BAD_ARG, "null batch argument");
    if (workspace_bytes_per_scene % 256) return fail(SURF_ERR_BAD_ARG, "workspace_bytes_per_scene must be a multiple of 256");
    if (n_scenes < 1) return fail(SURF_ERR_BAD_ARG, "empty batch");
    const bool warm = options->forced_nearest == 2;
    if (fused_batch_ok(*camera0, *options, n_scenes))
        return backward_strided_fused(n_scenes, scene0, camera0, layout, options, workspace, workspace_bytes_per_scene,
                                      nearest0, depth0, out_grads0, grads0, (cudaStream_t)cuda_stream, warm);
    const int64_t n = frame_pixels(*camera0, *options);
    return run_batch(n_scenes, (cudaStream_t)cuda_stream, [&](int b, cudaStream_t st) {
        SurfScene s; SurfCamera c; SurfOutGrads og; SurfSceneGrads sg;
        strided_scene(*scene0, *layout, b, &s);
        strided_camera(*camera0, *layout, b, &c);
        strided_grads(*grads0, scene0->n_sets, *layout, b, &sg);
        og.image = advance(out_grads0->image, b * n * 3); og.depth = advance(out_grads0->depth, b * n);
        og.normal = advance(out_grads0->normal, b * n * 3); og.pos = advance(out_grads0->pos, b * n * 3);
        return backward_impl(&s, &c, options, (char*)workspace + (size_t)b * workspace_bytes_per_scene,
                             workspace_bytes_per_scene, nearest0 + b * n, depth0 + b * n, &og, &sg, st, warm);
    });
}

// ---- fused inverse-rendering step: forward -> MSE loss + d(loss)/d(image) in the shading epilogue -> backward ----
// One host call, one enqueue (test_optimization.py:100-125).  d(loss)/d(image) and the double loss accumulator live in
// the (step-sized) workspace; depth / nearest are needed by the backward, so when the caller does not want them as
// outputs they must still be provided - the Python wrapper always does.
static int step_frame_check(const SurfOutputs* out, const SurfStepMSE* step, const SurfSceneGrads* grads) {
    if (!out || !step || !grads) return fail(SURF_ERR_BAD_ARG, "null outputs/step/scene_grads");
    if (!step->target_image) return fail(SURF_ERR_BAD_ARG, "surf_step_mse needs target_image");
    if (!out->depth || !out->nearest) return fail(SURF_ERR_BAD_ARG, "surf_step_mse needs out->depth and out->nearest (the backward reads them)");
    return SURF_OK;
}

int surf_step_mse(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* options, void* workspace,
                  size_t workspace_bytes, const SurfOutputs* out, const SurfStepMSE* step, const SurfSceneGrads* scene_grads,
                  void* cuda_stream) {
    g_launches = 0;
    int rc = step_frame_check(out, step, scene_grads);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    Frame f;
    if ((rc = make_frame(scene, camera, options, workspace, workspace_bytes, &f, step->grad_image == nullptr))) return rc;
    SURF_CUDA(cudaMemsetAsync(f.ws.loss_acc, 0, 8, st));
    float* gimg = step->grad_image ? step->grad_image : f.ws.gimg;
    const StepLoss loss{step->target_image, gimg, f.ws.loss_acc, step->loss_scale};
    if ((rc = forward_impl(scene, camera, options, workspace, workspace_bytes, out, st, &loss))) return rc;
    SurfOutGrads og;
    og.image = gimg; og.depth = nullptr; og.normal = nullptr; og.pos = nullptr;
    return backward_impl(scene, camera, options, workspace, workspace_bytes, out->nearest, out->depth, &og, scene_grads, st,
                         true, step->loss);
}

int surf_step_mse_strided(int32_t n_scenes, const SurfScene* scene0, const SurfCamera* camera0, const SurfBatchLayout* layout,
                          const SurfOptions* options, void* workspace, size_t workspace_bytes_per_scene,
                          const SurfOutputs* out0, const SurfStepMSE* step, const SurfSceneGrads* grads0, void* cuda_stream) {
    g_launches = 0;
    if (!scene0 || !camera0 || !layout || !options || !workspace) return fail(SURF_ERR_BAD_ARG, "null batch argument");
    int rc = step_frame_check(out0, step, grads0);
    if (rc) return rc;
    if (workspace_bytes_per_scene % 256) return fail(SURF_ERR_BAD_ARG, "workspace_bytes_per_scene must be a multiple of 256");
    if (n_scenes < 1) return fail(SURF_ERR_BAD_ARG, "empty batch");
    if (!fused_batch_ok(*camera0, *options, n_scenes))
        return fail(SURF_ERR_UNSUPPORTED, "surf_step_mse_strided: perspective frames without shadow rays and math_mode != 3 only");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    Frame f;
    if ((rc = make_frame(scene0, camera0, options, workspace, workspace_bytes_per_scene, &f))) return rc;
    SURF_CUDA(cudaMemset2DAsync(f.ws.loss_acc, workspace_bytes_per_scene, 0, 8, (size_t)n_scenes, st));
    if (!step->grad_image) return fail(SURF_ERR_BAD_ARG, "surf_step_mse_strided needs step->grad_image [B, n, 3]");
    const StepLoss loss{step->target_image, step->grad_image, f.ws.loss_acc, step->loss_scale};
    if ((rc = forward_strided_fused(n_scenes, scene0, camera0, layout, options, workspace, workspace_bytes_per_scene, out0, st, &loss)))
        return rc;
    SurfOutGrads og;
    og.image = step->grad_image; og.depth = nullptr; og.normal = nullptr; og.pos = nullptr;
    return backward_strided_fused(n_scenes, scene0, camera0, layout, options, workspace, workspace_bytes_per_scene,
                                  out0->nearest, out0->depth, &og, grads0, st, true, step->loss);
}

int surf_adam_step(const SurfAdamTensors* tensors, const float* grads_packed, float* exp_avg, float* exp_avg_sq, float* state,
                   float lr, float beta1, float beta2, float eps, void* cuda_stream) {
    g_launches = 0;
    if (!tensors || !grads_packed || !exp_avg || !exp_avg_sq || !state) return fail(SURF_ERR_BAD_ARG, "null adam argument");
    if (tensors->count < 1 || tensors->count > SURF_ADAM_MAX_TENSORS) return fail(SURF_ERR_BAD_ARG, "surf_adam_step: 1..16 tensors");
    AdamTensors tn;
    tn.count = tensors->count;
    long long off = 0;
    for (int j = 0; j < tensors->count; ++j) {
        if (!tensors->param[j] || tensors->size[j] < 0) return fail(SURF_ERR_BAD_ARG, "surf_adam_step: bad tensor");
        tn.param[j] = tensors->param[j];
        tn.begin[j] = off;
        off += tensors->size[j];
    }
    tn.begin[tensors->count] = off;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    k_adam_advance<<<1, 1, 0, st>>>(state, beta1, beta2);
    SURF_LAUNCHED("k_adam_advance");
    const int grid = (int)std::max<long long>(1, std::min<long long>((off + 255) / 256, (long long)sm_count() * 8));
    k_adam_packed<<<grid, 256, 0, st>>>(tn, grads_packed, exp_avg, exp_avg_sq, state, lr, beta1, beta2, eps);
    SURF_LAUNCHED("k_adam_packed");
    return SURF_OK;
}

// ---- projection layer: surfel projection + scatter renderers (projection_layer.py:20-106, utils.py:146-215) ----
static int project_params(const SurfProjection* pr, ProjectParams* p) {
    if (!pr || !pr->eye || !pr->at || !pr->up) return fail(SURF_ERR_BAD_ARG, "null projection / camera vectors");
    if (pr->batch < 1 || pr->batch > 65535 || pr->n_surfels < 1) return fail(SURF_ERR_BAD_ARG, "empty projection batch");
    if (pr->pos_stride != 3 && pr->pos_stride != 4) return fail(SURF_ERR_BAD_ARG, "pos_stride must be 3 or 4");
    if (pr->width < 1 || pr->height < 1) return fail(SURF_ERR_BAD_ARG, "empty viewport");
    const double h = tan(pr->fovy / 2) * 2 * pr->focal_length;
    const double w = h * ((double)pr->width / (double)pr->height);
    std::memset(p, 0, sizeof(*p));
    p->batch = pr->batch; p->n = pr->n_surfels; p->pos_stride = pr->pos_stride; p->W = pr->width; p->H = pr->height;
    p->f = (float)pr->focal_length;
    p->sx_px = (float)(-(pr->width - 1) / w); p->sy_px = (float)((pr->height - 1) / h);
    p->cx = (float)(pr->width / 2.0); p->cy = (float)(pr->height / 2.0);
    p->eye = pr->eye; p->at = pr->at; p->up = pr->up;
    p->eye_stride = pr->eye_stride; p->at_stride = pr->at_stride; p->up_stride = pr->up_stride;
    return SURF_OK;
}

int surf_project_surfels(const SurfProjection* proj, const float* pos_wc, float* px_coord, int64_t* px_idx, void* cuda_stream) {
    g_launches = 0;
    ProjectParams p;
    int rc = project_params(proj, &p);
    if (rc) return rc;
    if (!pos_wc || !px_coord) return fail(SURF_ERR_BAD_ARG, "null positions / px_coord");
    p.pos = pos_wc; p.px_coord = px_coord; p.px_idx = (long long*)px_idx;
    k_project_surfels<<<dim3((p.n + 255) / 256, p.batch), 256, 0, (cudaStream_t)cuda_stream>>>(p);
    SURF_LAUNCHED("k_project_surfels");
    return SURF_OK;
}

int surf_project_surfels_backward(const SurfProjection* proj, const float* pos_wc, const float* g_px_coord, float* g_pos,
                                  void* cuda_stream) {
    g_launches = 0;
    ProjectParams p;
    int rc = project_params(proj, &p);
    if (rc) return rc;
    if (!pos_wc || !g_px_coord || !g_pos) return fail(SURF_ERR_BAD_ARG, "null positions / gradients");
    p.pos = pos_wc; p.g_px = g_px_coord; p.g_pos = g_pos;
    k_project_surfels_backward<<<dim3((p.n + 255) / 256, p.batch), 256, 0, (cudaStream_t)cuda_stream>>>(p);
    SURF_LAUNCHED("k_project_surfels_backward");
    return SURF_OK;
}

static int scatter_params(const SurfScatter* sc, ScatterParams* p) {
    if (!sc) return fail(SURF_ERR_BAD_ARG, "null scatter description");
    if (sc->batch < 1 || sc->batch > 65535 || sc->n < 1 || sc->channels < 1 || sc->n_dst < 1) return fail(SURF_ERR_BAD_ARG, "empty scatter");
    if (sc->mode != 0 && sc->mode != 1) return fail(SURF_ERR_BAD_ARG, "scatter mode: 0 (mean) or 1 (weighted blended OIT)");
    std::memset(p, 0, sizeof(*p));
    p->batch = sc->batch; p->n = sc->n; p->channels = sc->channels; p->n_dst = sc->n_dst; p->mode = sc->mode;
    p->use_depth = sc->use_depth; p->use_center_dist = sc->use_center_dist;
    const double s2 = (double)sc->sigma * sc->sigma;
    p->alpha0 = (float)(1.0 / (2.0 * 3.14159265358979323846 * s2));
    p->inv_2s2 = (float)(1.0 / (2.0 * s2));
    p->z_scale = sc->z_scale;
    p->eps = sc->mode == 1 ? 1e-8f : 0.f;
    return SURF_OK;
}

int surf_scatter_forward(const SurfScatter* scatter, const float* x, const int64_t* idx, const float* z, const float* center_dist_2,
                         float* out, float* denom, uint8_t* mask, void* cuda_stream) {
    g_launches = 0;
    ScatterParams p;
    int rc = scatter_params(scatter, &p);
    if (rc) return rc;
    if (!x || !idx || !out || !denom) return fail(SURF_ERR_BAD_ARG, "null scatter operand");
    if (p.mode == 1 && ((p.use_depth && !z) || (p.use_center_dist && !center_dist_2))) return fail(SURF_ERR_BAD_ARG, "OIT scatter needs z / center_dist_2");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    p.x = x; p.idx = (const long long*)idx; p.z = z; p.cd2 = center_dist_2; p.out = out; p.denom = denom; p.mask = mask;
    SURF_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)p.batch * p.n_dst * p.channels, st));
    SURF_CUDA(cudaMemsetAsync(denom, 0, sizeof(float) * (size_t)p.batch * p.n_dst, st));
    k_scatter_accum<<<dim3((p.n + 255) / 256, p.batch), 256, 0, st>>>(p);
    SURF_LAUNCHED("k_scatter_accum");
    const size_t total = (size_t)p.batch * p.n_dst;
    k_scatter_normalize<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p);
    SURF_LAUNCHED("k_scatter_normalize");
    return SURF_OK;
}

int surf_scatter_backward(const SurfScatter* scatter, const float* x, const int64_t* idx, const float* z, const float* center_dist_2,
                          const float* out, const float* denom, const float* g_out, float* g_x, float* g_z, float* g_center_dist_2,
                          void* cuda_stream) {
    g_launches = 0;
    ScatterParams p;
    int rc = scatter_params(scatter, &p);
    if (rc) return rc;
    if (!x || !idx || !out || !denom || !g_out) return fail(SURF_ERR_BAD_ARG, "null scatter operand");
    p.x = x; p.idx = (const long long*)idx; p.z = z; p.cd2 = center_dist_2; p.out = (float*)out; p.denom = (float*)denom;
    p.g_out = g_out; p.g_x = g_x; p.g_z = g_z; p.g_cd2 = g_center_dist_2;
    k_scatter_backward<<<dim3((p.n + 255) / 256, p.batch), 256, 0, (cudaStream_t)cuda_stream>>>(p);
    SURF_LAUNCHED("k_scatter_backward");
    return SURF_OK;
}

static int bilinear_params(const SurfBilinear* bl, BilinearParams* p) {
    if (!bl) return fail(SURF_ERR_BAD_ARG, "null bilinear description");
    if (bl->batch < 1 || bl->batch > 65535 || bl->n < 1 || bl->channels < 1 || bl->width < 1 || bl->height < 1)
        return fail(SURF_ERR_BAD_ARG, "empty bilinear scatter");
    std::memset(p, 0, sizeof(*p));
    p->batch = bl->batch; p->n = bl->n; p->channels = bl->channels; p->W = bl->width; p->H = bl->height;
    p->use_depth = bl->use_depth; p->use_center_dist = bl->use_center_dist; p->want_depth = bl->compute_depth;
    const double s2 = (double)bl->sigma * bl->sigma;
    p->alpha0 = (float)(1.0 / (2.0 * 3.14159265358979323846 * s2));
    p->inv_2s2 = (float)(1.0 / (2.0 * s2));
    p->z_scale = bl->z_scale;
    p->eps = 1e-8f;
    return SURF_OK;
}

size_t surf_bilinear_acc_floats(const SurfBilinear* bl) {
    return bl ? (size_t)bl->batch * 4 * (size_t)bl->width * bl->height * (size_t)(bl->channels + 3) : 0;
}

int surf_bilinear_oit_forward(const SurfBilinear* bl, const float* px_coord, const float* x, float* acc, float* out, float* mask,
                              float* depth, void* cuda_stream) {
    g_launches = 0;
    BilinearParams p;
    int rc = bilinear_params(bl, &p);
    if (rc) return rc;
    if (!px_coord || !x || !acc || !out || !mask) return fail(SURF_ERR_BAD_ARG, "null bilinear operand");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    p.px = px_coord; p.x = x; p.acc = acc; p.out = out; p.mask = mask; p.depth = bl->compute_depth ? depth : nullptr;
    SURF_CUDA(cudaMemsetAsync(acc, 0, sizeof(float) * surf_bilinear_acc_floats(bl), st));
    k_bilinear_accum<<<dim3((p.n + 255) / 256, p.batch), 256, 0, st>>>(p);
    SURF_LAUNCHED("k_bilinear_accum");
    const size_t total = (size_t)p.batch * p.W * p.H;
    k_bilinear_normalize<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p);
    SURF_LAUNCHED("k_bilinear_normalize");
    return SURF_OK;
}

int surf_bilinear_oit_backward(const SurfBilinear* bl, const float* px_coord, const float* x, const float* acc, const float* g_out,
                               const float* g_mask, const float* g_depth, float* g_x, float* g_px_coord, void* cuda_stream) {
    g_launches = 0;
    BilinearParams p;
    int rc = bilinear_params(bl, &p);
    if (rc) return rc;
    if (!px_coord || !x || !acc) return fail(SURF_ERR_BAD_ARG, "null bilinear operand");
    p.px = px_coord; p.x = x; p.acc = (float*)acc;
    p.g_out = g_out; p.g_mask = g_mask; p.g_depth = g_depth; p.g_x = g_x; p.g_px = g_px_coord;
    k_bilinear_backward<<<dim3((p.n + 255) / 256, p.batch), 256, 0, (cudaStream_t)cuda_stream>>>(p);
    SURF_LAUNCHED("k_bilinear_backward");
    return SURF_OK;
}

int surf_gaussian_blur(const float* image, float* scratch, float* out, int32_t batch, int32_t height, int32_t width, int32_t channels,
                       float sigma, void* cuda_stream) {
    g_launches = 0;
    if (!image || !scratch || !out || batch < 1 || height < 1 || width < 1 || channels < 1 || !(sigma > 0.f))
        return fail(SURF_ERR_BAD_ARG, "bad blur argument");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int half = (int)floor((double)sigma * 3.0);
    double sum = 0.0;
    for (int d = -half; d <= half; ++d) sum += exp(-(double)(d * d) / (2.0 * (double)sigma * sigma));
    const float inv_2s2 = (float)(1.0 / (2.0 * (double)sigma * sigma)), norm = (float)(1.0 / sum);
    const size_t total = (size_t)batch * height * width * channels;
    const unsigned grid = (unsigned)((total + 255) / 256);
    k_blur_pass<<<grid, 256, 0, st>>>(image, scratch, batch, height, width, channels, 0, half, inv_2s2, norm);
    SURF_LAUNCHED("k_blur_pass");
    k_blur_pass<<<grid, 256, 0, st>>>(scratch, out, batch, height, width, channels, 1, half, inv_2s2, norm);
    SURF_LAUNCHED("k_blur_pass");
    return SURF_OK;
}

size_t surf_splats_workspace_bytes(int32_t n_splats, int32_t n_lights) {
    SplatWorkspace ws;
    carve_splats(nullptr, n_splats, n_lights, &ws);
    return align_up(ws.bytes, 256);
}

int surf_splats_forward(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* options,
                        const SurfSplats* splats, void* workspace, size_t workspace_bytes, const SurfOutputs* out,
                        void* cuda_stream) {
    g_launches = 0;
    return splats_forward_impl(1, scene, camera, options, splats, nullptr, workspace, workspace_bytes, out, (cudaStream_t)cuda_stream);
}

int surf_splats_backward(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* options,
                         const SurfSplats* splats, void* workspace, size_t workspace_bytes, const SurfOutGrads* out_grads,
                         const SurfSceneGrads* scene_grads, const SurfSplatGrads* splat_grads, void* cuda_stream) {
    g_launches = 0;
    return splats_backward_impl(1, scene, camera, options, splats, nullptr, workspace, workspace_bytes, out_grads, scene_grads,
                                splat_grads, (cudaStream_t)cuda_stream);
}

int surf_splats_forward_strided(int32_t n_scenes, const SurfScene* scene0, const SurfCamera* camera0, const SurfOptions* options,
                                const SurfSplats* splats0, const SurfSplatBatch* batch, void* workspace,
                                size_t workspace_bytes_per_scene, const SurfOutputs* out0, void* cuda_stream) {
    g_launches = 0;
    if (!batch) return fail(SURF_ERR_BAD_ARG, "null batch layout");
    if (workspace_bytes_per_scene % 256) return fail(SURF_ERR_BAD_ARG, "workspace_bytes_per_scene must be a multiple of 256");
    return splats_forward_impl(n_scenes, scene0, camera0, options, splats0, batch, workspace, workspace_bytes_per_scene, out0,
                               (cudaStream_t)cuda_stream);
}

int surf_splats_backward_strided(int32_t n_scenes, const SurfScene* scene0, const SurfCamera* camera0, const SurfOptions* options,
                                 const SurfSplats* splats0, const SurfSplatBatch* batch, void* workspace,
                                 size_t workspace_bytes_per_scene, const SurfOutGrads* out_grads0, const SurfSceneGrads* scene_grads,
                                 const SurfSplatGrads* splat_grads0, void* cuda_stream) {
    g_launches = 0;
    if (!batch) return fail(SURF_ERR_BAD_ARG, "null batch layout");
    if (workspace_bytes_per_scene % 256) return fail(SURF_ERR_BAD_ARG, "workspace_bytes_per_scene must be a multiple of 256");
    return splats_backward_impl(n_scenes, scene0, camera0, options, splats0, batch, workspace, workspace_bytes_per_scene,
                                out_grads0, scene_grads, splat_grads0, (cudaStream_t)cuda_stream);
}

double surf_fma_peak(int32_t mode, int32_t iters, void* cuda_stream) {
    cudaStream_t st = (cudaStream_t)cuda_stream;
    float* out = nullptr;
    if (cudaMalloc(&out, 4) != cudaSuccess) return -1.0;
    const int grid = sm_count() * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {   // first pass warms up
        cudaEventRecord(e0, st);
        if (mode == 1) k_fma_peak<true><<<grid, 256, 0, st>>>(iters, out);
        else k_fma_peak<false><<<grid, 256, 0, st>>>(iters, out);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    if (cudaGetLastError() != cudaSuccess || ms <= 0.f) return -1.0;
    // lane-FMAs: both variants execute 16 lane-FMAs per thread per iteration
    const double lane_fma = (double)grid * 256.0 * (double)iters * 16.0;
    return lane_fma / (ms * 1e-3);
}

}  // extern "C"

namespace surf {

// bump allocator over the context arena
constexpr size_t kStageBytes = 64 << 10;

struct Bump {
    char* base; size_t off, cap;
    void* take(size_t bytes) {
        off = align_up(off, 256);
        void* p = base ? base + off : nullptr;
        off += bytes;
        return p;
    }
};

struct HostPlan {        // device mirrors of every host array of one call
    SurfScene dscene;
    SurfCamera dcam;
    SurfOutputs dout;
    SurfOutGrads dgout;
    SurfSceneGrads dgrads;
    void* workspace; size_t workspace_bytes;
    float* d_target; double* d_loss;
    size_t grads_begin, grads_end;     // arena range holding the gradient accumulators (zeroed per call); its last
    float* d_loss_f;                   // float is the loss of a step (so one collective reduces gradients and loss)
};
struct HostState {                     // a host call between its begin (H2D + kernels) and end (D2H + synchronise)
    HostPlan pl;
    SurfScene hscene;                  // counts / strides of the caller's arrays
    int n;
    bool want_bwd, has_target, active;
};

static size_t set_pos_floats(const SurfPrimSet& s) {
    return (size_t)s.count * (s.kind == SURF_TRIANGLE ? 3 : 1) * s.pos_stride;
}

// Lays out the arena; with b.base == nullptr it only measures.
static void plan_host(const SurfScene& hs, const SurfCamera& hc, const SurfOptions& opt, int n, bool want_bwd,
                      bool has_target, Bump& b, HostPlan* pl) {
    pl->dscene = hs;
    pl->dcam = hc;
    int total = 0;
    for (int k = 0; k < hs.n_sets; ++k) {
        const SurfPrimSet& s = hs.sets[k];
        SurfPrimSet& d = pl->dscene.sets[k];
        total += s.count;
        d.pos = (const float*)b.take(set_pos_floats(s) * 4);
        d.normal = s.normal ? (const float*)b.take((size_t)s.count * s.normal_stride * 4) : nullptr;
        d.radius = s.radius ? (const float*)b.take((size_t)s.count * 4) : nullptr;
        d.material_idx = (const int32_t*)b.take((size_t)s.count * 4);
    }
    pl->dscene.light_pos = (const float*)b.take((size_t)hs.n_lights * hs.light_pos_stride * 4);
    pl->dscene.light_color_idx = (const int32_t*)b.take((size_t)hs.n_lights * 4);
    pl->dscene.light_attenuation = (const float*)b.take((size_t)hs.n_lights * 12);
    pl->dscene.ambient = (const float*)b.take(12);
    pl->dscene.colors = (const float*)b.take((size_t)hs.n_colors * 12);
    pl->dscene.albedo = (const float*)b.take((size_t)hs.n_materials * 12);
    pl->dscene.coeffs = (const float*)b.take((size_t)hs.n_materials * 12);
    pl->dscene.gamma = hs.gamma ? (const float*)b.take(4) : nullptr;
    pl->dcam.eye = (const float*)b.take(12);
    pl->dcam.at = (const float*)b.take(12);
    pl->dcam.up = (const float*)b.take(12);
    pl->dout.image = (float*)b.take((size_t)n * 12);
    pl->dout.depth = (float*)b.take((size_t)n * 4);
    pl->dout.normal = (float*)b.take((size_t)n * 12);
    pl->dout.pos = (float*)b.take((size_t)n * 12);
    pl->dout.nearest = (int64_t*)b.take((size_t)n * 8);
    pl->dout.ray_dir = (float*)b.take((size_t)(hc.proj == 0 ? n : 1) * 12);
    pl->workspace_bytes = surf_workspace_bytes(total, n, hs.n_lights, opt.shadow);
    pl->workspace = b.take(pl->workspace_bytes);
    pl->d_target = nullptr; pl->d_loss = nullptr; pl->d_loss_f = nullptr;
    pl->grads_begin = pl->grads_end = 0;
    memset(&pl->dgout, 0, sizeof(pl->dgout));
    memset(&pl->dgrads, 0, sizeof(pl->dgrads));
    if (!want_bwd) return;
    pl->dgout.image = (const float*)b.take((size_t)n * 12);
    pl->dgout.depth = (const float*)b.take((size_t)n * 4);
    pl->dgout.normal = (const float*)b.take((size_t)n * 12);
    pl->dgout.pos = (const float*)b.take((size_t)n * 12);
    if (has_target) {
        pl->d_target = (float*)b.take((size_t)n * 12);
        pl->d_loss = (double*)b.take(8);
    }
    b.off = align_up(b.off, 256);
    pl->grads_begin = b.off;
    for (int k = 0; k < hs.n_sets; ++k) {
        const SurfPrimSet& s = hs.sets[k];
        pl->dgrads.sets[k].pos = (float*)b.take(set_pos_floats(s) * 4);
        pl->dgrads.sets[k].normal = s.normal ? (float*)b.take((size_t)s.count * s.normal_stride * 4) : nullptr;
        pl->dgrads.sets[k].radius = s.kind == SURF_SPHERE ? (float*)b.take((size_t)s.count * 4) : nullptr;
    }
    pl->dgrads.light_pos = (float*)b.take((size_t)hs.n_lights * hs.light_pos_stride * 4);
    pl->dgrads.light_attenuation = (float*)b.take((size_t)hs.n_lights * 12);
    pl->dgrads.ambient = (float*)b.take(12);
    pl->dgrads.colors = (float*)b.take((size_t)hs.n_colors * 12);
    pl->dgrads.albedo = (float*)b.take((size_t)hs.n_materials * 12);
    pl->dgrads.coeffs = (float*)b.take((size_t)hs.n_materials * 12);
    pl->dgrads.gamma = (float*)b.take(4);
    pl->d_loss_f = (float*)b.take(4);
    pl->grads_end = align_up(b.off, 256);
    b.off = pl->grads_end;
}

__global__ void k_loss_to_float(const double* __restrict__ acc, float* __restrict__ out) { out[0] = (float)acc[0]; }

static int h2d(SurfContext* c, const void* dst, const void* src, size_t bytes) {
    if (!bytes) return SURF_OK;
    SURF_CUDA(cudaMemcpyAsync((void*)dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    c->h2d += bytes;
    return SURF_OK;
}
static int d2h(SurfContext* c, void* dst, const void* src, size_t bytes) {
    if (!bytes || !dst) return SURF_OK;
    SURF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    c->d2h += bytes;
    return SURF_OK;
}

// first half of a host call: H2D of the inputs, every kernel of the forward (and backward); asynchronous on c->stream.
// `loss_scale` <= 0: the mean over this call's pixels.
static int host_begin(SurfContext* c, const SurfScene* hs, const SurfCamera* hc, const SurfOptions* opt,
                      const SurfOutputs* hout, const SurfOutGrads* hgout, const float* target, float loss_scale, bool want_bwd) {
    if (!c || !hs || !hc || !opt) return fail(SURF_ERR_BAD_ARG, "null context/scene/camera/options");
    SURF_CUDA(cudaSetDevice(c->device));
    if (!c->state) c->state = new HostState();
    c->state->active = false;
    g_launches = 0;
    c->h2d = c->d2h = 0;
    SceneView probe;
    std::string err;
    if (!build_scene_view(*hs, &probe, &err)) return fail(SURF_ERR_BAD_ARG, err);
    if (!check_camera(*hc, &err)) return fail(SURF_ERR_BAD_ARG, err);
    // the index arrays are host memory here: range-check them like the reference's index_select would
    for (int k = 0; k < hs->n_sets; ++k)
        for (int i = 0; i < hs->sets[k].count; ++i) {
            const int m = hs->sets[k].material_idx[i];
            if (m < 0 || m >= hs->n_materials) return fail(SURF_ERR_BAD_ARG, "material_idx out of range");
        }
    for (int l = 0; l < hs->n_lights; ++l)
        if (hs->light_color_idx[l] < 0 || hs->light_color_idx[l] >= hs->n_colors) return fail(SURF_ERR_BAD_ARG, "light color_idx out of range");
    const int N = hc->width * hc->height;
    int p0 = opt->pixel_begin, p1 = opt->pixel_end;
    if (p0 == 0 && p1 == 0) p1 = N;
    if (p0 < 0 || p1 > N || p1 <= p0) return fail(SURF_ERR_BAD_ARG, "bad pixel range");
    const int n = p1 - p0;
    const bool has_target = want_bwd && target != nullptr;

    HostPlan& pl = c->state->pl;
    Bump measure{nullptr, 0, 0};
    plan_host(*hs, *hc, *opt, n, want_bwd, has_target, measure, &pl);
    const size_t need = align_up(measure.off, 256) + 256;
    if (need > c->arena_bytes) {
        if (c->arena) SURF_CUDA(cudaFree(c->arena));
        c->arena = nullptr; c->arena_bytes = 0;
        SURF_CUDA(cudaMalloc(&c->arena, need));
        c->arena_bytes = need;
    }
    Bump b{(char*)c->arena, 0, c->arena_bytes};
    plan_host(*hs, *hc, *opt, n, want_bwd, has_target, b, &pl);

    int rc;
    for (int k = 0; k < hs->n_sets; ++k) {
        const SurfPrimSet& s = hs->sets[k];
        const SurfPrimSet& d = pl.dscene.sets[k];
        if ((rc = h2d(c, d.pos, s.pos, set_pos_floats(s) * 4))) return rc;
        if (s.normal && (rc = h2d(c, d.normal, s.normal, (size_t)s.count * s.normal_stride * 4))) return rc;
        if (s.radius && (rc = h2d(c, d.radius, s.radius, (size_t)s.count * 4))) return rc;
        if ((rc = h2d(c, d.material_idx, s.material_idx, (size_t)s.count * 4))) return rc;
    }
    {   // lights, colours, materials, tonemap, camera vectors: a dozen arrays of a few bytes each.  Their device mirrors are
        // neighbours in the arena, so they are packed into pinned staging at the same offsets and uploaded as ONE copy
        // (each separate copy costs 5 - 10 us of DMA latency on the stream).
        struct Item { const void* dst; const void* src; size_t bytes; };
        const Item items[] = {
            {pl.dscene.light_pos, hs->light_pos, (size_t)hs->n_lights * hs->light_pos_stride * 4},
            {pl.dscene.light_color_idx, hs->light_color_idx, (size_t)hs->n_lights * 4},
            {pl.dscene.light_attenuation, hs->light_attenuation, (size_t)hs->n_lights * 12},
            {pl.dscene.ambient, hs->ambient, 12},
            {pl.dscene.colors, hs->colors, (size_t)hs->n_colors * 12},
            {pl.dscene.albedo, hs->albedo, (size_t)hs->n_materials * 12},
            {pl.dscene.coeffs, hs->coeffs, (size_t)hs->n_materials * 12},
            {pl.dscene.gamma, hs->gamma, hs->gamma ? (size_t)4 : (size_t)0},
            {pl.dcam.eye, hc->eye, 12}, {pl.dcam.at, hc->at, 12}, {pl.dcam.up, hc->up, 12}};
        const char* lo = (const char*)pl.dscene.light_pos;
        const char* hi = (const char*)pl.dcam.up + 12;
        if ((size_t)(hi - lo) <= kStageBytes) {
            for (const Item& it : items)
                if (it.bytes) std::memcpy(c->stage_up + ((const char*)it.dst - lo), it.src, it.bytes);
            if ((rc = h2d(c, lo, c->stage_up, (size_t)(hi - lo)))) return rc;
        } else {
            for (const Item& it : items)
                if ((rc = h2d(c, it.dst, it.src, it.bytes))) return rc;
        }
    }

    // with a target image the loss and d(loss)/d(image) are fused into the shading epilogue (mean over this call's pixels)
    StepLoss sl{pl.d_target, (float*)pl.dgout.image, pl.d_loss, loss_scale > 0.f ? loss_scale : 1.0f / (3.0f * (float)n)};
    if (has_target) {
        // the target is only read by the shading kernel: upload it on the copy stream, under the intersection stage
        // (the previous call on this context ended with a stream synchronisation, so d_target is free)
        SURF_CUDA(cudaMemcpyAsync((void*)pl.d_target, target, (size_t)n * 12, cudaMemcpyHostToDevice, c->copy_stream));
        SURF_CUDA(cudaEventRecord(c->target_ready, c->copy_stream));
        c->h2d += (size_t)n * 12;
        sl.target_ready = c->target_ready;
        SURF_CUDA(cudaMemsetAsync(pl.d_loss, 0, 8, c->stream));
    }
    if ((rc = forward_impl(&pl.dscene, &pl.dcam, opt, pl.workspace, pl.workspace_bytes, &pl.dout, c->stream,
                           has_target ? &sl : nullptr)))
        return rc;

    if (hout) {
        if ((rc = d2h(c, hout->image, pl.dout.image, (size_t)n * 12))) return rc;
        if ((rc = d2h(c, hout->depth, pl.dout.depth, (size_t)n * 4))) return rc;
        if ((rc = d2h(c, hout->normal, pl.dout.normal, (size_t)n * 12))) return rc;
        if ((rc = d2h(c, hout->pos, pl.dout.pos, (size_t)n * 12))) return rc;
        if ((rc = d2h(c, hout->nearest, pl.dout.nearest, (size_t)n * 8))) return rc;
        if ((rc = d2h(c, hout->ray_dir, pl.dout.ray_dir, (size_t)(hc->proj == 0 ? n : 1) * 12))) return rc;
    }
    if (want_bwd) {
        SurfOutGrads og = pl.dgout;
        if (has_target) {
            og.depth = nullptr; og.normal = nullptr; og.pos = nullptr;
        } else {
            if (!hgout) return fail(SURF_ERR_BAD_ARG, "need out_grads or target_image");
            if (hgout->image) { if ((rc = h2d(c, og.image, hgout->image, (size_t)n * 12))) return rc; } else og.image = nullptr;
            if (hgout->depth) { if ((rc = h2d(c, og.depth, hgout->depth, (size_t)n * 4))) return rc; } else og.depth = nullptr;
            if (hgout->normal) { if ((rc = h2d(c, og.normal, hgout->normal, (size_t)n * 12))) return rc; } else og.normal = nullptr;
            if (hgout->pos) { if ((rc = h2d(c, og.pos, hgout->pos, (size_t)n * 12))) return rc; } else og.pos = nullptr;
        }
        SURF_CUDA(cudaMemsetAsync((char*)c->arena + pl.grads_begin, 0, pl.grads_end - pl.grads_begin, c->stream));
        if ((rc = backward_impl(&pl.dscene, &pl.dcam, opt, pl.workspace, pl.workspace_bytes, pl.dout.nearest,
                                pl.dout.depth, &og, &pl.dgrads, c->stream, true)))
            return rc;
        if (has_target) {
            k_loss_to_float<<<1, 1, 0, c->stream>>>(pl.d_loss, pl.d_loss_f);
            SURF_LAUNCHED("k_loss_to_float");
        }
    }
    c->state->hscene = *hs;
    c->state->n = n;
    c->state->want_bwd = want_bwd;
    c->state->has_target = has_target;
    c->state->active = true;
    return SURF_OK;
}

// second half: D2H of the gradients and the loss, then synchronise
static int host_end(SurfContext* c, const SurfSceneGrads* hgrads, float* loss) {
    if (!c || !c->state || !c->state->active) return fail(SURF_ERR_BAD_ARG, "no host call in flight on this context");
    SURF_CUDA(cudaSetDevice(c->device));
    HostState& hsx = *c->state;
    hsx.active = false;
    const HostPlan& pl = hsx.pl;
    const SurfScene* hs = &hsx.hscene;
    const bool has_target = hsx.has_target;
    int rc;
    if (hsx.want_bwd) {
        if (!hgrads) return fail(SURF_ERR_BAD_ARG, "null scene_grads");
        for (int k = 0; k < hs->n_sets; ++k) {
            const SurfPrimSet& s = hs->sets[k];
            if ((rc = d2h(c, hgrads->sets[k].pos, pl.dgrads.sets[k].pos, set_pos_floats(s) * 4))) return rc;
            if (s.normal && (rc = d2h(c, hgrads->sets[k].normal, pl.dgrads.sets[k].normal, (size_t)s.count * s.normal_stride * 4))) return rc;
            if (s.kind == SURF_SPHERE && (rc = d2h(c, hgrads->sets[k].radius, pl.dgrads.sets[k].radius, (size_t)s.count * 4))) return rc;
        }
    }
    // the small gradient arrays and the loss: neighbours in the arena, ONE copy into pinned staging, handed out after the wait
    struct Item { void* dst; const void* src; size_t bytes; };
    float loss_f = 0.f;
    Item items[8];
    int n_items = 0;
    if (hsx.want_bwd) {
        items[n_items++] = {hgrads->light_pos, pl.dgrads.light_pos, (size_t)hs->n_lights * hs->light_pos_stride * 4};
        items[n_items++] = {hgrads->light_attenuation, pl.dgrads.light_attenuation, (size_t)hs->n_lights * 12};
        items[n_items++] = {hgrads->ambient, pl.dgrads.ambient, 12};
        items[n_items++] = {hgrads->colors, pl.dgrads.colors, (size_t)hs->n_colors * 12};
        items[n_items++] = {hgrads->albedo, pl.dgrads.albedo, (size_t)hs->n_materials * 12};
        items[n_items++] = {hgrads->coeffs, pl.dgrads.coeffs, (size_t)hs->n_materials * 12};
        if (hs->gamma) items[n_items++] = {hgrads->gamma, pl.dgrads.gamma, 4};
        if (has_target && loss) items[n_items++] = {&loss_f, pl.d_loss_f, 4};
    }
    const char* lo = hsx.want_bwd ? (const char*)pl.dgrads.light_pos : nullptr;
    const char* hi = hsx.want_bwd ? (const char*)pl.d_loss_f + 4 : nullptr;
    const bool staged = hsx.want_bwd && (size_t)(hi - lo) <= kStageBytes;
    if (staged) {
        SURF_CUDA(cudaMemcpyAsync(c->stage_down, lo, (size_t)(hi - lo), cudaMemcpyDeviceToHost, c->stream));
        for (int k = 0; k < n_items; ++k)
            if (items[k].dst) c->d2h += items[k].bytes;
    } else {
        for (int k = 0; k < n_items; ++k)
            if ((rc = d2h(c, items[k].dst, items[k].src, items[k].bytes))) return rc;
    }
    SURF_CUDA(cudaStreamSynchronize(c->stream));
    if (staged)
        for (int k = 0; k < n_items; ++k)
            if (items[k].dst && items[k].bytes) std::memcpy(items[k].dst, c->stage_down + ((const char*)items[k].src - lo), items[k].bytes);
    if (has_target && loss) *loss = loss_f;
    return SURF_OK;
}

static int host_call(SurfContext* c, const SurfScene* hs, const SurfCamera* hc, const SurfOptions* opt,
                     const SurfOutputs* hout, const SurfOutGrads* hgout, const float* target, float* loss,
                     const SurfSceneGrads* hgrads, bool want_bwd) {
    if (want_bwd && !hgrads) return fail(SURF_ERR_BAD_ARG, "null scene_grads");
    int rc = host_begin(c, hs, hc, opt, hout, hgout, target, 0.f, want_bwd);
    if (rc) return rc;
    return host_end(c, hgrads, loss);
}

}  // namespace surf

extern "C" {

SurfContext* surf_context_create(int32_t device) {
    if (cudaSetDevice(device) != cudaSuccess) { g_error = "cudaSetDevice failed"; return nullptr; }
    SurfContext* c = new SurfContext();
    c->device = device; c->arena = nullptr; c->arena_bytes = 0; c->h2d = c->d2h = 0; c->state = nullptr;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->target_ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaHostAlloc((void**)&c->stage_up, surf::kStageBytes, cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc((void**)&c->stage_down, surf::kStageBytes, cudaHostAllocDefault) != cudaSuccess) {
        g_error = "cudaStreamCreate failed";
        delete c;
        return nullptr;
    }
    return c;
}

void surf_context_destroy(SurfContext* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->arena) cudaFree(c->arena);
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->copy_stream);
    cudaEventDestroy(c->target_ready);
    cudaFreeHost(c->stage_up);
    cudaFreeHost(c->stage_down);
    delete c->state;
    delete c;
}

int surf_context_last_transfer(const SurfContext* c, uint64_t* h2d, uint64_t* d2h) {
    if (!c) return fail(SURF_ERR_BAD_ARG, "null context");
    if (h2d) *h2d = c->h2d;
    if (d2h) *d2h = c->d2h;
    return SURF_OK;
}

int surf_step_host_begin(SurfContext* ctx, const SurfScene* scene, const SurfCamera* camera, const SurfOptions* options,
                         const float* target_image, float loss_scale) {
    if (!target_image) return fail(SURF_ERR_BAD_ARG, "surf_step_host_begin needs target_image");
    return host_begin(ctx, scene, camera, options, nullptr, nullptr, target_image, loss_scale, true);
}
int surf_step_host_end(SurfContext* ctx, const SurfSceneGrads* scene_grads, float* loss) { return host_end(ctx, scene_grads, loss); }
int surf_context_device_grads(SurfContext* ctx, float** block, size_t* n_floats) {
    if (!ctx || !ctx->state || !ctx->state->active || !ctx->state->want_bwd) return fail(SURF_ERR_BAD_ARG, "no step in flight on this context");
    const HostPlan& pl = ctx->state->pl;
    if (block) *block = (float*)((char*)ctx->arena + pl.grads_begin);
    if (n_floats) *n_floats = (pl.grads_end - pl.grads_begin) / 4;
    return SURF_OK;
}
void* surf_context_stream(SurfContext* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int surf_render_host(SurfContext* ctx, const SurfScene* scene, const SurfCamera* camera, const SurfOptions* options,
                     const SurfOutputs* out) {
    return host_call(ctx, scene, camera, options, out, nullptr, nullptr, nullptr, nullptr, false);
}

int surf_render_backward_host(SurfContext* ctx, const SurfScene* scene, const SurfCamera* camera,
                              const SurfOptions* options, const SurfOutputs* out, const SurfOutGrads* out_grads,
                              const float* target_image, float* loss, const SurfSceneGrads* scene_grads) {
    return host_call(ctx, scene, camera, options, out, out_grads, target_image, loss, scene_grads, true);
}

}  // extern "C"

