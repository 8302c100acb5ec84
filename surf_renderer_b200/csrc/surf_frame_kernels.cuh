// surf_frame_kernels.cuh - part of libsurf_b200.so (included by surf_kernels.cu inside namespace surf, after surf_batch.cuh).
// k_setup / k_prep / k_raygen: per-frame camera state, primitive records, rays
#pragma once

// ---------------------------------------------------------------------------------------------------
// k_setup / k_prep / k_raygen
// ---------------------------------------------------------------------------------------------------
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

__global__ void __launch_bounds__(256) k_prep(const __grid_constant__ SceneView sc, CamState* __restrict__ cs,
                                              float4* __restrict__ packed) {
    // camera rays: the edge-function triangle filter assumes hits at t >= 0, i.e. a non-negative near plane
    prep_body(sc, v3(cs->eye[0], cs->eye[1], cs->eye[2]), packed, blockIdx.x * blockDim.x + threadIdx.x, &cs->bad_index,
              cs->near_clip >= 0.f ? 1 : 2);
}

__global__ void __launch_bounds__(256) k_prep_batch(const __grid_constant__ SceneView sc0, const __grid_constant__ BatchArgs ba,
                                                    CamState* cs0, float4* packed0) {
    __shared__ SceneView sc;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) { sc = sc0; scene_at(&sc, ba, b); }
    __syncthreads();
    CamState* cs = ws_at(cs0, ba, b);
    prep_body(sc, v3(cs->eye[0], cs->eye[1], cs->eye[2]), ws_at(packed0, ba, b), blockIdx.x * blockDim.x + threadIdx.x,
              &cs->bad_index, cs->near_clip >= 0.f ? 1 : 2);
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

