// surf_frame_kernels.cuh - part of libsurf_b200.so (single translation unit: included by surf_kernels.cu inside namespace surf).
// k_setup / k_prep / k_raygen: per-frame camera state, primitive records, rays
#pragma once

// ---------------------------------------------------------------------------------------------------
// k_setup / k_prep / k_raygen
// ---------------------------------------------------------------------------------------------------
struct CamArgs {
    const float* eye; const float* at; const float* up;
    int proj, W, H;
    double fovy, focal;
    float near_clip, far_clip;
};

// ---------------------------------------------------------------------------------------------------
// strided batches (surf_forward_strided / surf_backward_strided): scene b is scene 0 with every input pointer
// advanced by b * stride elements and its workspace / outputs by fixed byte / element strides.  The *_batch
// kernels below and in the other headers take the scene index from blockIdx.y (k_intersect: from the work item),
// build the parameter block of their scene in shared memory and run the same body as the single-scene kernels.
// ---------------------------------------------------------------------------------------------------
struct BatchArgs {
    int n_scenes;
    long long ws_stride;          // bytes between consecutive per-scene workspaces
    long long set_pos[kMaxSets], set_normal[kMaxSets], set_radius[kMaxSets], set_mat[kMaxSets];   // elements
    long long light_pos, light_color_idx, light_atten, ambient, colors, albedo, coeffs, gamma;
    long long eye, at, up;
};

template <class T>
__host__ __device__ __forceinline__ T* adv(T* p, long long elems) { return p ? p + elems : nullptr; }
template <class T>
__host__ __device__ __forceinline__ T* ws_at(T* p, const BatchArgs& ba, int b) {
    return p ? (T*)((char*)p + ba.ws_stride * b) : nullptr;
}
template <class T>
__host__ __device__ __forceinline__ const T* ws_at(const T* p, const BatchArgs& ba, int b) {
    return p ? (const T*)((const char*)p + ba.ws_stride * b) : nullptr;
}

__host__ __device__ inline void scene_at(SceneView* sc, const BatchArgs& ba, int b) {
    for (int k = 0; k < kMaxSets; ++k) {
        SetView& sv = sc->sets[k];
        sv.pos = adv(sv.pos, b * ba.set_pos[k]);
        sv.normal = adv(sv.normal, b * ba.set_normal[k]);
        sv.radius = adv(sv.radius, b * ba.set_radius[k]);
        sv.mat = adv(sv.mat, b * ba.set_mat[k]);
    }
    sc->light_pos = adv(sc->light_pos, b * ba.light_pos);
    sc->light_color_idx = adv(sc->light_color_idx, b * ba.light_color_idx);
    sc->light_atten = adv(sc->light_atten, b * ba.light_atten);
    sc->ambient = adv(sc->ambient, b * ba.ambient);
    sc->colors = adv(sc->colors, b * ba.colors);
    sc->albedo = adv(sc->albedo, b * ba.albedo);
    sc->coeffs = adv(sc->coeffs, b * ba.coeffs);
    sc->gamma = adv(sc->gamma, b * ba.gamma);
}

__global__ void k_setup(CamArgs a, CamState* cs) {
    if (threadIdx.x == 0 && blockIdx.x == 0)
        camera_setup(a.eye, a.at, a.up, a.proj, a.W, a.H, a.fovy, a.focal, a.near_clip, a.far_clip, cs);
}

__global__ void k_setup_batch(CamArgs a, const __grid_constant__ BatchArgs ba, CamState* cs0) {
    const int b = blockIdx.x;
    if (threadIdx.x == 0)
        camera_setup(a.eye + b * ba.eye, a.at + b * ba.at, a.up + b * ba.up, a.proj, a.W, a.H, a.fovy, a.focal, a.near_clip,
                     a.far_clip, ws_at(cs0, ba, b));
}

// filter records of primitive g for rays from the common origin o (the eye; a light for k_prep_lights)
__device__ __forceinline__ void prep_body(const SceneView& sc, Vec3 o, float4* __restrict__ packed, int g) {
    if (g >= sc.total) return;
    const int s = find_set(sc, g);
    const SetView& sv = sc.sets[s];
    const int i = g - sv.first;
    F4 r[4];
    if (sv.kind == KIND_DISK) {
        prep_disk(ld3(sv.pos + (size_t)i * sv.pos_stride), ld3(sv.normal + (size_t)i * sv.normal_stride), sv.radius[i], o,
                  &r[0], &r[1]);
    } else if (sv.kind == KIND_PLANE) {
        prep_plane(ld3(sv.pos + (size_t)i * sv.pos_stride), ld3(sv.normal + (size_t)i * sv.normal_stride), o, &r[0]);
    } else if (sv.kind == KIND_SPHERE) {
        prep_sphere(ld3(sv.pos + (size_t)i * sv.pos_stride), sv.radius[i], o, &r[0]);
    } else {
        const float* f = sv.pos + (size_t)i * 3 * sv.pos_stride;
        prep_triangle(ld3(f), ld3(f + sv.pos_stride), ld3(f + 2 * sv.pos_stride),
                      ld3(sv.normal + (size_t)i * sv.normal_stride), o, &r[0], &r[1], &r[2], &r[3]);
    }
    const int nf4 = rec_f4(sv.kind);
    float4* dst = packed + sv.rec_off + (size_t)i * nf4;
    for (int k = 0; k < nf4; ++k) dst[k] = make_float4(r[k].x, r[k].y, r[k].z, r[k].w);
}

__global__ void __launch_bounds__(256) k_prep(const __grid_constant__ SceneView sc, const CamState* __restrict__ cs,
                                              float4* __restrict__ packed) {
    prep_body(sc, v3(cs->eye[0], cs->eye[1], cs->eye[2]), packed, blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void __launch_bounds__(256) k_prep_batch(const __grid_constant__ SceneView sc0, const __grid_constant__ BatchArgs ba,
                                                    const CamState* cs0, float4* packed0) {
    __shared__ SceneView sc;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) { sc = sc0; scene_at(&sc, ba, b); }
    __syncthreads();
    const CamState* cs = ws_at(cs0, ba, b);
    prep_body(sc, v3(cs->eye[0], cs->eye[1], cs->eye[2]), ws_at(packed0, ba, b), blockIdx.x * blockDim.x + threadIdx.x);
}

__device__ __forceinline__ void raygen_body(const CamState* __restrict__ cs, int pix0, int n, float* __restrict__ rays,
                                            float* __restrict__ ray_out, unsigned long long* __restrict__ zbuf, int k);

__global__ void __launch_bounds__(256) k_raygen(const CamState* __restrict__ cs, int pix0, int n,
                                                float* __restrict__ rays, float* __restrict__ ray_out,
                                                unsigned long long* __restrict__ zbuf) {
    raygen_body(cs, pix0, n, rays, ray_out, zbuf, blockIdx.x * blockDim.x + threadIdx.x);
}

// ray_out0: [B, 3, n] (perspective) or [B, 3, 1] (orthographic)
__global__ void __launch_bounds__(256) k_raygen_batch(const CamState* cs0, const __grid_constant__ BatchArgs ba, int pix0, int n,
                                                      float* rays0, float* ray_out0, long long ray_out_stride,
                                                      unsigned long long* zbuf0) {
    const int b = blockIdx.y;
    raygen_body(ws_at(cs0, ba, b), pix0, n, ws_at(rays0, ba, b), adv(ray_out0, b * ray_out_stride), ws_at(zbuf0, ba, b),
                blockIdx.x * blockDim.x + threadIdx.x);
}

__device__ __forceinline__ void raygen_body(const CamState* __restrict__ cs, int pix0, int n, float* __restrict__ rays,
                                            float* __restrict__ ray_out, unsigned long long* __restrict__ zbuf, int k) {
    if (k >= n) return;
    zbuf[k] = kMissKey;
    if (cs->proj == 0) {
        Vec3 d = pixel_ray_dir(*cs, pix0 + k);
        rays[k] = d.x; rays[(size_t)n + k] = d.y; rays[2 * (size_t)n + k] = d.z;
        if (ray_out) { ray_out[k] = d.x; ray_out[(size_t)n + k] = d.y; ray_out[2 * (size_t)n + k] = d.z; }
    } else if (k == 0 && ray_out) {
        ray_out[0] = cs->odir[0]; ray_out[1] = cs->odir[1]; ray_out[2] = cs->odir[2];
    }
}

__global__ void __launch_bounds__(256) k_prep_screen(const __grid_constant__ SceneView sc, const CamState* __restrict__ cs,
                                                     float4* __restrict__ circ) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= sc.total) return;
    const int s = find_set(sc, g);
    const F4 r = prep_screen(*cs, sc.sets[s], g - sc.sets[s].first);
    circ[g] = make_float4(r.x, r.y, r.z, r.w);
}

