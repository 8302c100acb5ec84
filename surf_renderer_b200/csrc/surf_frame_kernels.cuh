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

__global__ void k_setup(CamArgs a, CamState* cs) {
    if (threadIdx.x == 0 && blockIdx.x == 0)
        camera_setup(a.eye, a.at, a.up, a.proj, a.W, a.H, a.fovy, a.focal, a.near_clip, a.far_clip, cs);
}

__global__ void __launch_bounds__(256) k_prep(const __grid_constant__ SceneView sc, const CamState* __restrict__ cs,
                                              float4* __restrict__ packed) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= sc.total) return;
    const int s = find_set(sc, g);
    const SetView& sv = sc.sets[s];
    const int i = g - sv.first;
    const Vec3 o = v3(cs->eye[0], cs->eye[1], cs->eye[2]);
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

__global__ void __launch_bounds__(256) k_raygen(const CamState* __restrict__ cs, int pix0, int n,
                                                float* __restrict__ rays, float* __restrict__ ray_out,
                                                unsigned long long* __restrict__ zbuf) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
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

