// surf_batch.cuh - part of libsurf_b200.so (included by every translation unit inside namespace surf).
// Camera arguments, strided-batch addressing and the per-primitive record preparation shared by the frame kernels
// (surf_kernels.cu) and the intersection translation units.  Inline / forceinline code only: no kernels here.
#pragma once

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

// filter records of primitive g for rays from the common origin o (the eye; a light for k_prep_lights)
// `bad_index` (may be null): set when a material index (every thread checks its primitive) or a light's colour index
// (thread 0) is out of range - the kernels clamp such indices, surf_check_indices() reports them
__device__ __forceinline__ void prep_body(const SceneView& sc, Vec3 o, float4* __restrict__ packed, int g, int* bad_index = nullptr,
                                          int tri_form = 0 /* 0: line form (shadow rays); 1: edge functions; 2: always pass */) {
    if (g >= sc.total) return;
    if (bad_index) {
        bool bad = false;
        if (g == 0)
            for (int l = 0; l < sc.n_lights; ++l) bad |= sc.light_color_idx[l] < 0 || sc.light_color_idx[l] >= sc.n_colors;
        const int s0 = find_set(sc, g);
        const int m = sc.sets[s0].mat[g - sc.sets[s0].first];
        bad |= m < 0 || m >= sc.n_materials;
        if (bad) atomicOr(bad_index, 1);
    }
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
        if (tri_form == 0)
            prep_triangle_line(ld3(f), ld3(f + sv.pos_stride), ld3(f + 2 * sv.pos_stride),
                               ld3(sv.normal + (size_t)i * sv.normal_stride), o, &r[0], &r[1], &r[2], &r[3]);
        else
            prep_triangle(ld3(f), ld3(f + sv.pos_stride), ld3(f + 2 * sv.pos_stride),
                          ld3(sv.normal + (size_t)i * sv.normal_stride), o, &r[0], &r[1], &r[2], &r[3], tri_form == 1);
    }
    const int nf4 = rec_f4(sv.kind);
    float4* dst = packed + sv.rec_off + (size_t)i * nf4;
    for (int k = 0; k < nf4; ++k) dst[k] = make_float4(r[k].x, r[k].y, r[k].z, r[k].w);
}

