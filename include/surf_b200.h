/*
 * surf_b200.h - C ABI of libsurf_b200.so: the B200 (sm_100a) implementation of DiffRend's
 * ray-cast render path.
 *
 * Every entry point replaces (part of) one reference interface; the citations are into the
 * reference tree (fmannan/surf_renderer):
 *
 *   surf_forward / surf_render_host      <- diffrend/torch/renderer.py:136-355  render(scene, **params)
 *        ray generation                  <- diffrend/torch/utils.py:439-478     generate_rays
 *        intersections + z-buffer        <- diffrend/torch/utils.py:238-366,481-512 ; renderer.py:170-202
 *        shading + tonemap               <- renderer.py:82-125 fragment_shader ; :318-340 ; utils.py:430-432
 *        shadow rays (opt.shadow)        <- renderer.py:291-314
 *   surf_backward / surf_render_backward_host
 *                                        <- torch autograd through the same graph (loss.backward(),
 *                                           e.g. diffrend/torch/test_optimization.py:100-125)
 *
 * The reference has no FFI for this path (it is a Python function); the Python binding a
 * maintainer would add is ctypes - see INTEGRATION.md and surf_renderer_b200/_lib.py.
 *
 * Conventions
 *   - plain C, POD structs, raw pointers + sizes, no torch / C++ types.
 *   - "device" entry points take DEVICE pointers and a CUDA stream (as void*); they are asynchronous
 *     and stream-ordered, hold no device memory and no global state (thread-local error string only).
 *   - "host" entry points take HOST pointers with the same structs; a SurfContext owns the staging
 *     buffers, the device workspace and a stream.  They return after the results are in host memory.
 *   - all entry points return 0 on success or a negative SurfStatus; surf_last_error() describes it.
 *   - all float data is IEEE fp32; index outputs are int64 like torch's LongTensor.
 *   - row strides are given in floats so that the reference's homogeneous [M,4] arrays and plain
 *     [M,3] arrays are both accepted without a copy (the reference slices [:, :3], utils.py:245,288).
 */
#ifndef SURF_B200_H_
#define SURF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SURF_ABI_VERSION 2
#define SURF_MAX_SETS 8

typedef enum SurfStatus {
    SURF_OK = 0,
    SURF_ERR_BAD_ARG = -1,
    SURF_ERR_UNSUPPORTED = -2,
    SURF_ERR_CUDA = -3,
    SURF_ERR_WORKSPACE = -4
} SurfStatus;

/* primitive kinds; the order of SurfScene.sets is the insertion order of scene['objects']
 * (utils.py:486) and defines the global primitive index reported in `nearest`. */
typedef enum SurfPrimKind { SURF_DISK = 0, SURF_PLANE = 1, SURF_SPHERE = 2, SURF_TRIANGLE = 3 } SurfPrimKind;

typedef struct SurfPrimSet {
    int32_t kind;            /* SurfPrimKind */
    int32_t count;           /* M_k */
    const float* pos;        /* disk/plane/sphere: [M_k, pos_stride]; triangle: face [M_k, 3, pos_stride] */
    int32_t pos_stride;      /* 3 or 4 */
    const float* normal;     /* disk/plane/triangle: [M_k, normal_stride] (un-normalised ok); sphere: NULL */
    int32_t normal_stride;   /* 3 or 4 */
    const float* radius;     /* disk/sphere: [M_k]; else NULL */
    const int32_t* material_idx; /* [M_k] */
} SurfPrimSet;

typedef struct SurfScene {
    int32_t n_sets;
    SurfPrimSet sets[SURF_MAX_SETS];
    int32_t n_lights;
    const float* light_pos;      /* [L, light_pos_stride] */
    int32_t light_pos_stride;    /* 3 or 4 */
    const int32_t* light_color_idx; /* [L] rows of `colors` */
    const float* light_attenuation; /* [L,3] (kc, kl, kq) */
    const float* ambient;        /* [3] */
    int32_t n_colors;
    const float* colors;         /* [C,3] */
    int32_t n_materials;
    const float* albedo;         /* [K,3] */
    const float* coeffs;         /* [K,3] (kd, ks, shininess) */
    const float* gamma;          /* [1] tonemap gamma, or NULL for no tonemap (renderer.py:339) */
} SurfScene;

typedef struct SurfCamera {
    int32_t proj;            /* 0 perspective, 1 orthographic (utils.py:462-469) */
    int32_t width, height;   /* viewport[2]-viewport[0], viewport[3]-viewport[1] */
    double fovy;             /* radians */
    double focal_length;
    const float* eye;        /* [3] (first three of the reference's 4-vector) */
    const float* at;         /* [3] */
    const float* up;         /* [3] */
    float near_clip, far_clip;
} SurfCamera;

typedef struct SurfOptions {
    int32_t double_sided;    /* renderer.py:107-112 */
    int32_t use_quartic;     /* renderer.py:90 */
    int32_t shadow;          /* renderer.py:291-314 */
    int32_t pixel_begin;     /* render flat pixels [pixel_begin, pixel_end) of the H*W row-major grid;     */
    int32_t pixel_end;       /* 0,0 = whole frame.  Output arrays are sized for the range (row bands / GPU) */
    int32_t forced_nearest;  /* backward only: 2 = `workspace` still holds this frame's forward state (camera,
                                rays, shadow visibility), skip recomputing it; 0/1 = recompute               */
    int32_t pixels_per_thread; /* 0 = library default; tuning knob (modes 0-2: 2,4,8; mode 3: 4,8,16)        */
    int32_t chunk_prims;     /* 0 = library default; primitives staged per TMA bulk copy (multiple of 32)   */
    int32_t math_mode;       /* intersection kernel.  All modes run the same exact narrow phase and give bit-identical
                                results.
                                0 = default.  Disk sets of at least 256 primitives in single frames above 256x256
                                    pixels: filter records streamed through the constant bank into uniform registers
                                    (k_filter_const) - bounding-sphere test, 3 FMA-pipe lane-instr per ray-disk test;
                                    the candidates go through the plane filter and the exact test in k_narrow_queue.
                                    Everything else: ray-plane filter, packed FFMA2, records staged in shared memory
                                    by TMA (k_intersect / k_intersect_batch), 10 lane-instr per ray-disk test - the
                                    SURVEY 8(d) formulation; frames of at most 256x256 pixels and scenes with triangle
                                    sets take the dense body (mode 4).
                                1 = staged kernel, scalar FFMA; 2 = staged kernel, packed, one branch per primitive;
                                3 = per-pair screen-space bounding-circle test (2.25 lane-instr per test);
                                4 = dense: the staged filter with 2-D pixel tiles and per-disk minima (the strided
                                    batch kernel on one scene) for frames with splats several pixels wide;
                                5 = the staged kernel for everything (mode 0 without the constant-bank path);
                                6 = mode 0 with the plane filter (10 lane-instr per test) in k_filter_const.        */
} SurfOptions;

/* outputs for n = pixel_end - pixel_begin pixels (row-major).  Any pointer may be NULL to skip it. */
typedef struct SurfOutputs {
    float* image;            /* [n,3] */
    float* depth;            /* [n]   miss = far+1 (renderer.py:180) */
    float* normal;           /* [n,3] */
    float* pos;              /* [n,3] */
    int64_t* nearest;        /* [n]   miss = 0 (argmin of an all-(far+1) column) */
    float* ray_dir;          /* perspective: [3,n]; orthographic: [3,1] */
} SurfOutputs;

/* incoming gradients of the outputs (NULL = zero) */
typedef struct SurfOutGrads {
    const float* image;      /* [n,3] */
    const float* depth;      /* [n]   */
    const float* normal;     /* [n,3] */
    const float* pos;        /* [n,3] */
} SurfOutGrads;

/* gradient accumulators, same layout/strides as the inputs they belong to; caller zero-initialises,
 * the library adds.  NULL = not wanted.  (Disk radius has no gradient in the reference: SURVEY A.5.) */
typedef struct SurfPrimSetGrads {
    float* pos;              /* disk/plane/sphere centre; triangle: face (only vertex 0 receives gradient) */
    float* normal;
    float* radius;           /* sphere only */
} SurfPrimSetGrads;

typedef struct SurfSceneGrads {
    SurfPrimSetGrads sets[SURF_MAX_SETS];
    float* light_pos;        /* [L, light_pos_stride] */
    float* light_attenuation;/* [L,3] */
    float* ambient;          /* [3] */
    float* colors;           /* [C,3] */
    float* albedo;           /* [K,3] */
    float* coeffs;           /* [K,3] */
    float* gamma;            /* [1] */
} SurfSceneGrads;

int surf_abi_version(void);
const char* surf_last_error(void);

/* bytes of device scratch surf_forward / surf_backward need for this problem size
 * (n_lights and shadow only matter when options.shadow is set: + 4*L*n bytes of visibility) */
size_t surf_workspace_bytes(int32_t total_prims, int32_t n_pixels, int32_t n_lights, int32_t shadow);
/* the exact need of one call: `orthographic` = camera proj 1 (per-pixel ray origins: + 40 B per pixel; shadow frames
 * carry that list anyway), `step` = the workspace also holds d(loss)/d(image) of surf_step_mse (+ 12 B per pixel).
 * Perspective frames above 256x256 pixels include the candidate queue of the constant-bank intersection path
 * (80 B per pixel); a workspace without room for it is accepted and renders through the staged kernel.
 * surf_workspace_bytes() = the larger of orthographic = 0 / 1, step = 0. */
size_t surf_workspace_bytes_ex(int32_t total_prims, int32_t n_pixels, int32_t n_lights, int32_t shadow,
                               int32_t orthographic, int32_t step);

/* Index arrays (material_idx, light_color_idx) in DEVICE memory cannot be range-checked by the host without a
 * synchronisation, so the kernels clamp them to the valid range and record the fact in the workspace.  This call
 * synchronises `cuda_stream` and returns SURF_ERR_BAD_ARG if the last forward that used `workspace` saw an index out of
 * range - where the reference's index_select (renderer.py:284-286) raises IndexError.  The host-pointer entry points
 * check their (host) index arrays up front instead. */
int surf_check_indices(const void* workspace, void* cuda_stream);

/* ---- device-pointer API (what the torch autograd.Function calls) ---- */
int surf_forward(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* options,
                 void* workspace, size_t workspace_bytes, const SurfOutputs* out, void* cuda_stream);

/* `nearest` [n] int64 and `depth` [n] are the forward outputs (hit <=> depth <= far). */
int surf_backward(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* options,
                  void* workspace, size_t workspace_bytes,
                  const int64_t* nearest, const float* depth,
                  const SurfOutGrads* out_grads, const SurfSceneGrads* scene_grads, void* cuda_stream);

/* ---- one inverse-rendering step in one call: render -> mean((image - target)^2) -> backward ----
 * Replaces the loop body of diffrend/torch/test_optimization.py:100-125 (`res = render(scene)`, `loss = mean((im -
 * target)^2)`, `loss.backward()`): forward of the pixel range, the loss and d(loss)/d(image) fused into the shading
 * epilogue, then the backward into `scene_grads` - all enqueued by this call.  out->depth and out->nearest are
 * required (the backward reads them); the other outputs may be NULL.  The workspace must be sized with
 * surf_workspace_bytes_ex(..., step = 1) unless `grad_image` is provided. */
typedef struct SurfStepMSE {
    const float* target_image;   /* [n,3] target of this call's pixel range */
    float loss_scale;            /* weight of every squared error: 1/(3*W*H) = the mean over the whole frame, also when
                                    this call renders one band of it (the bands' partial losses then add up) */
    float* loss;                 /* device [1] or NULL; the library ADDS loss_scale * sum((image-target)^2) to it
                                    (caller zero-initialises, like the gradient accumulators) */
    float* grad_image;           /* optional device [n,3] (strided: [B,n,3], required): receives d(loss)/d(image) */
} SurfStepMSE;
int surf_step_mse(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* options, void* workspace,
                  size_t workspace_bytes, const SurfOutputs* out, const SurfStepMSE* step,
                  const SurfSceneGrads* scene_grads, void* cuda_stream);

/* ---- the optimizer of that loop: `optimizer.step()` (test_optimization.py:122; torch.optim.Adam, no amsgrad, no weight
 * decay) for parameter tensors whose gradients are packed back to back in ONE flat buffer (the layout MSEStep gives
 * them for the single all-reduce).  exp_avg / exp_avg_sq have the gradients' packed layout; `state` is three device
 * floats (step count, 1 - beta1^t, sqrt(1 - beta2^t)), all zero-initialised by the caller and advanced on the device,
 * so the call replays from a CUDA graph. */
#define SURF_ADAM_MAX_TENSORS 16
typedef struct SurfAdamTensors {
    int32_t count;
    float* param[SURF_ADAM_MAX_TENSORS];     /* device pointers of the parameter tensors, in packed order */
    int64_t size[SURF_ADAM_MAX_TENSORS];     /* their element counts */
} SurfAdamTensors;
int surf_adam_step(const SurfAdamTensors* tensors, const float* grads_packed, float* exp_avg, float* exp_avg_sq,
                   float* state, float lr, float beta1, float beta2, float eps, void* cuda_stream);

/* ---- batches of independent scenes: the per-element loop of GAN.get_real_samples (GAN/gan.py:326-377) ----
 * Arrays of n_scenes scenes / cameras / workspaces / outputs (one shared SurfOptions).  All kernels of all scenes are
 * launched by this one call, fanned out over internal streams that fork from and join to `cuda_stream`; gradients of
 * parameters shared between scenes may alias (the accumulation is atomic). */
int surf_forward_batch(int32_t n_scenes, const SurfScene* scenes, const SurfCamera* cameras, const SurfOptions* options,
                       void* const* workspaces, const size_t* workspace_bytes, const SurfOutputs* outs, void* cuda_stream);
int surf_backward_batch(int32_t n_scenes, const SurfScene* scenes, const SurfCamera* cameras, const SurfOptions* options,
                        void* const* workspaces, const size_t* workspace_bytes, const int64_t* const* nearest,
                        const float* const* depth, const SurfOutGrads* out_grads, const SurfSceneGrads* scene_grads,
                        void* cuda_stream);

/* ---- strided batches: one scene description + element strides between consecutive scenes ----
 * Replaces the same per-element loop (GAN.get_real_samples, diffrend/torch/GAN/gan.py:326-377: one render() call per
 * batch element over scenes that differ only in splat positions / normals and camera eye) for callers that hold
 * the batch as stacked tensors (splat positions [B,M,3], camera eyes [B,3], shared lights and materials).  Scene b reads every pointer of `scene0` / `camera0` advanced by b * stride ELEMENTS (floats / int32);
 * a stride of 0 means all scenes share that array (its gradient then receives the sum over the batch).  All scenes
 * have the same primitive counts, light / colour / material counts, viewport and scalar camera parameters.
 * Outputs, `nearest`, `depth`, incoming gradients are contiguous per scene: [B, n, ...] (ray_dir [B,3,n] or [B,3,1]);
 * `workspace` holds B slices of `workspace_bytes_per_scene` (>= surf_workspace_bytes, rounded up to 256).
 * Gradient accumulators in `grads0` have the layout (and strides) of the inputs they belong to. */
typedef struct SurfBatchLayout {
    int64_t set_pos[SURF_MAX_SETS], set_normal[SURF_MAX_SETS], set_radius[SURF_MAX_SETS], set_material_idx[SURF_MAX_SETS];
    int64_t light_pos, light_color_idx, light_attenuation, ambient, colors, albedo, coeffs, gamma;
    int64_t eye, at, up;
} SurfBatchLayout;
int surf_forward_strided(int32_t n_scenes, const SurfScene* scene0, const SurfCamera* camera0,
                         const SurfBatchLayout* layout, const SurfOptions* options, void* workspace,
                         size_t workspace_bytes_per_scene, const SurfOutputs* out0, void* cuda_stream);
int surf_backward_strided(int32_t n_scenes, const SurfScene* scene0, const SurfCamera* camera0,
                          const SurfBatchLayout* layout, const SurfOptions* options, void* workspace,
                          size_t workspace_bytes_per_scene, const int64_t* nearest0, const float* depth0,
                          const SurfOutGrads* out_grads0, const SurfSceneGrads* grads0, void* cuda_stream);

/* fused step over a strided batch: target_image / grad_image are [B,n,3]; *loss receives the sum over the scenes
 * (perspective frames without shadow rays, math_mode != 3 - otherwise SURF_ERR_UNSUPPORTED) */
int surf_step_mse_strided(int32_t n_scenes, const SurfScene* scene0, const SurfCamera* camera0,
                          const SurfBatchLayout* layout, const SurfOptions* options, void* workspace,
                          size_t workspace_bytes_per_scene, const SurfOutputs* out0, const SurfStepMSE* step,
                          const SurfSceneGrads* grads0, void* cuda_stream);

/* ---- one-splat-per-pixel renderer: diffrend/torch/renderer.py:537-751 render_splats_along_ray ----
 * Splat k sits on the ray of flat pixel k at camera-space depth z[k] (negative = in front of the camera,
 * clamped with -relu(-z)); it is shaded in camera coordinates with the same fragment shader as render()
 * (no double_sided, no tonemap; use_quartic honoured).  `scene` supplies lights / colours / materials only
 * (its primitive sets are ignored; light_pos must be homogeneous [L,4], it is multiplied by the view matrix).
 * Outputs: image [n,3], depth [n] = |pos|, pos [n,3] (camera coordinates), normal [n,3] (copy of the input). */
typedef struct SurfSplats {
    int32_t count;            /* must equal width * height */
    const float* z;           /* [count] depths, or column 2 of a [count,3] position array (z_stride = 3) */
    int32_t z_stride;         /* 1 or 3 */
    const float* normal;      /* [count, normal_stride] camera-space normals (used as given, not normalised); ignored
                                 when estimate_normals != 0 */
    int32_t normal_stride;    /* 3 or 4 */
    const int32_t* material_idx; /* [count], or NULL = material 0 */
    const float* light_vis;   /* [L, count] per-light visibility (constant), or NULL */
    const float* pos;         /* optional [count,3] explicit camera-space fragment positions; when non-NULL `z` is ignored,
                                 count need not equal W*H, and samples / estimate_normals must be 1 / 0 */
    int32_t samples;          /* supersampling K (renderer.py:603-673): every pixel's splat plane is intersected with its
                                 K x K sub-pixel rays; outputs are [(H K) (W K), ...].  0 or 1 = off */
    int32_t estimate_normals; /* 0 = `normal` is given; 1 = 3x3 constrained plane fit (utils.py:886-923, the GAN's
                                 normal_estimation_method='plane'); 2 = average neighbour cross product (utils.py:854-883) */
    int32_t ndc_stride;       /* 0, or 3 | 4: `pos` holds NDC coordinates [count, ndc_stride] - render_splats_NDC
                                 (renderer.py:358-474): unprojected with the inverse perspective of the camera's
                                 fovy / near / far (ops.py:61-68), shaded with the UNNORMALISED view vector -pos like the
                                 reference (:437) and with options.double_sided honoured; gradients go to
                                 SurfSplatGrads.pos in the same layout */
    float* norm_depth;        /* optional [n] output: norm_depth_image_only (renderer.py:677-686): depth normalised to
                                 [0, 1] over the frame, fragments at or beyond the camera's far plane mapped to 0 */
} SurfSplats;
typedef struct SurfSplatGrads {
    float* z;                 /* same stride as SurfSplats.z; caller zero-initialises, the library adds */
    float* normal;            /* same stride as SurfSplats.normal (given normals only) */
    float* pos;               /* [count,3], only with explicit positions */
} SurfSplatGrads;
/* bytes of device scratch per scene (camera, camera-space lights, accumulators, estimated normals, per-splat gradient
 * accumulators); n_splats = width * height (or `count` with explicit positions) */
size_t surf_splats_workspace_bytes(int32_t n_splats, int32_t n_lights);
int surf_splats_forward(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* options,
                        const SurfSplats* splats, void* workspace, size_t workspace_bytes,
                        const SurfOutputs* out, void* cuda_stream);
int surf_splats_backward(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* options,
                         const SurfSplats* splats, void* workspace, size_t workspace_bytes,
                         const SurfOutGrads* out_grads, const SurfSceneGrads* scene_grads,
                         const SurfSplatGrads* splat_grads, void* cuda_stream);

/* ---- a batch of along-ray frames in one call: the per-element loop of the GAN generator step, GAN/gan.py:563-597
 * (each element: its own depths / normals, camera eye and light positions; shared materials and camera scalars).
 * Scene b reads the pointers of `splats0` / `scene0->light_pos` / `camera0->eye` advanced by b * stride ELEMENTS
 * (0 = shared); outputs and incoming gradients are [B, n, ...]; `workspace` holds B slices of
 * workspace_bytes_per_scene (>= surf_splats_workspace_bytes, a multiple of 256).  Gradients of shared arrays receive
 * the sum over the batch. */
typedef struct SurfSplatBatch {
    int64_t z, normal, material_idx, light_vis;   /* strides of the SurfSplats arrays */
    int64_t light_pos;                            /* stride of scene0->light_pos ([B, L, 4] -> 4 L) */
    int64_t eye;                                  /* stride of camera0->eye */
} SurfSplatBatch;
int surf_splats_forward_strided(int32_t n_scenes, const SurfScene* scene0, const SurfCamera* camera0,
                                const SurfOptions* options, const SurfSplats* splats0, const SurfSplatBatch* batch,
                                void* workspace, size_t workspace_bytes_per_scene, const SurfOutputs* out0,
                                void* cuda_stream);
int surf_splats_backward_strided(int32_t n_scenes, const SurfScene* scene0, const SurfCamera* camera0,
                                 const SurfOptions* options, const SurfSplats* splats0, const SurfSplatBatch* batch,
                                 void* workspace, size_t workspace_bytes_per_scene, const SurfOutGrads* out_grads0,
                                 const SurfSceneGrads* scene_grads, const SurfSplatGrads* splat_grads0, void* cuda_stream);

/* ---- projection layer: surfels of one view projected into another camera and scattered onto its pixel grid ----
 * surf_project_surfels       <- diffrend/torch/projection_layer.py:20-86  project_surfels + project_image_coordinates:
 *                               world -> camera (lookat) -> image plane (x = f X / Z, nonzero_divide) -> pixel
 *                               coordinates px_coord [B,N,3] = (pixel x, pixel y, depth -Z) and the destination index
 *                               px_idx [B,N] = round(py - 0.5) * W + round(px - 0.5), or W*H ("dump") outside the frame
 * surf_scatter_forward       <- diffrend/torch/utils.py:146-175 scatter_mean_dim0 (mode 0: out = mean of the surfels
 *                               that land on a pixel, mask = pixel received nothing) and utils.py:178-215
 *                               scatter_weighted_blended_oit (mode 1: weights alpha(center_dist_2) * exp(-z_scale z),
 *                               out = sum(x a w) / (sum(a w) + 1e-8))
 * Camera vectors are [B,3] with element strides (0 = shared).  `out` [B,n_dst,C] and `denom` [B,n_dst] are written by
 * the library (zeroed first); gradients are ADDED to caller-zeroed arrays like everywhere else in this ABI. */
typedef struct SurfProjection {
    int32_t batch, n_surfels, pos_stride;   /* positions [B, N, 3|4] */
    int32_t width, height;
    double fovy, focal_length;
    const float* eye; const float* at; const float* up;
    int64_t eye_stride, at_stride, up_stride;
} SurfProjection;
int surf_project_surfels(const SurfProjection* proj, const float* pos_wc, float* px_coord, int64_t* px_idx, void* cuda_stream);
int surf_project_surfels_backward(const SurfProjection* proj, const float* pos_wc, const float* g_px_coord, float* g_pos,
                                  void* cuda_stream);
typedef struct SurfScatter {
    int32_t batch, n, channels;   /* x [B, n, C], idx [B, n] */
    int32_t n_dst;                /* destinations per batch element; an index outside [0, n_dst) is dropped */
    int32_t mode;                 /* 0 = scatter_mean_dim0, 1 = scatter_weighted_blended_oit */
    float sigma, z_scale;         /* OIT: Gaussian sigma of the pixel-centre weight, Beer-Lambert extinction */
    int32_t use_depth, use_center_dist;
} SurfScatter;
int surf_scatter_forward(const SurfScatter* scatter, const float* x, const int64_t* idx, const float* z,
                         const float* center_dist_2, float* out, float* denom, uint8_t* mask, void* cuda_stream);
int surf_scatter_backward(const SurfScatter* scatter, const float* x, const int64_t* idx, const float* z,
                          const float* center_dist_2, const float* out, const float* denom, const float* g_out,
                          float* g_x, float* g_z, float* g_center_dist_2, void* cuda_stream);

/* surf_bilinear_oit_*      <- projection_layer.py:170-240, the scatter stage of projection_renderer_differentiable_fast:
 *                            every surfel goes to the four pixels around px_coord with bilinear weights, each corner
 *                            scatter blended like scatter_weighted_blended_oit (sigma 0.5, z_scale 2 in the reference's
 *                            call), the four normalised results summed: out [B,P,C], soft mask [B,P], optionally the
 *                            new depth [B,P].  `acc` (surf_bilinear_acc_floats floats, written by the forward) carries
 *                            the per-corner accumulators to the backward, which ADDS d/dx and d/d(px_coord).
 * surf_gaussian_blur      <- projection_layer.py:156-168 blur(): separable, zero-padded, taps within 3 sigma; the
 *                            operation is its own adjoint, so the backward is the same call on the gradient. */
typedef struct SurfBilinear {
    int32_t batch, n, channels, width, height;
    float sigma, z_scale;
    int32_t use_depth, use_center_dist, compute_depth;
} SurfBilinear;
size_t surf_bilinear_acc_floats(const SurfBilinear* bl);
int surf_bilinear_oit_forward(const SurfBilinear* bl, const float* px_coord, const float* x, float* acc, float* out,
                              float* mask, float* depth, void* cuda_stream);
int surf_bilinear_oit_backward(const SurfBilinear* bl, const float* px_coord, const float* x, const float* acc,
                               const float* g_out, const float* g_mask, const float* g_depth, float* g_x,
                               float* g_px_coord, void* cuda_stream);
int surf_gaussian_blur(const float* image, float* scratch, float* out, int32_t batch, int32_t height, int32_t width,
                       int32_t channels, float sigma, void* cuda_stream);

/* ---- host-pointer API (self-contained: H2D, kernels, D2H) ---- */
typedef struct SurfContext SurfContext;
SurfContext* surf_context_create(int32_t device);
void surf_context_destroy(SurfContext* ctx);
int surf_render_host(SurfContext* ctx, const SurfScene* scene, const SurfCamera* camera,
                     const SurfOptions* options, const SurfOutputs* out);
/* forward + backward in one call: renders, copies outputs out (if non-NULL), then back-propagates the
 * host `out_grads` and copies the scene gradients back.  If `target_image` is non-NULL the image
 * gradient is instead d/d(image) of mean((image-target)^2) and *loss receives that loss (the
 * inverse-rendering step of test_optimization.py:100-125). */
int surf_render_backward_host(SurfContext* ctx, const SurfScene* scene, const SurfCamera* camera,
                              const SurfOptions* options, const SurfOutputs* out,
                              const SurfOutGrads* out_grads, const float* target_image, float* loss,
                              const SurfSceneGrads* scene_grads);
/* The same step in two halves, for callers that reduce the gradients across devices before reading them (one process
 * per GPU, each rendering one band of the frame):
 *   surf_step_host_begin   H2D of the scene and the target band, forward, fused MSE loss (weight `loss_scale` per
 *                          squared error; <= 0: the mean over this call's pixels), backward - asynchronous on the
 *                          context's stream;
 *   surf_context_device_grads  the device block that holds every gradient array of the call, contiguous, with the loss
 *                          as its last float - run the collective (e.g. ncclAllReduce) over it on surf_context_stream();
 *   surf_step_host_end     D2H of the gradients and the loss into the caller's host arrays, then synchronise. */
int surf_step_host_begin(SurfContext* ctx, const SurfScene* scene, const SurfCamera* camera, const SurfOptions* options,
                         const float* target_image, float loss_scale);
int surf_context_device_grads(SurfContext* ctx, float** block, size_t* n_floats);
void* surf_context_stream(SurfContext* ctx);
int surf_step_host_end(SurfContext* ctx, const SurfSceneGrads* scene_grads, float* loss);
/* bytes moved by the last host call */
int surf_context_last_transfer(const SurfContext* ctx, uint64_t* h2d_bytes, uint64_t* d2h_bytes);

/* ---- measurement helpers ---- */
/* FP32 FMA-pipe microbenchmark on the current device: lane-instructions/s for scalar FFMA (mode 0)
 * and packed FFMA2 (mode 1, counted as two lane-FMAs per lane).  Returns <0 on error. */
double surf_fma_peak(int32_t mode, int32_t iters, void* cuda_stream);
/* number of kernels the last device/host call on this thread launched */
int surf_last_launch_count(void);
/* per-kernel device timing with CUDA events recorded on the launching stream (off by default).
 * which: 0 = intersection (k_intersect), 1 = shading (k_shade), 2 = backward (k_backward).
 * surf_last_kernel_ms synchronises on that kernel's end event; returns <0 when nothing was recorded. */
void surf_set_kernel_timing(int32_t enabled);          /* (re-)enabling resets the recorded launches */
double surf_last_kernel_ms(int32_t which);
/* mean duration over the launches recorded since timing was enabled (ring of the last 256); the events are
 * recorded asynchronously, only this call synchronises.  *launches receives how many were averaged. */
double surf_mean_kernel_ms(int32_t which, int32_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* SURF_B200_H_ */
