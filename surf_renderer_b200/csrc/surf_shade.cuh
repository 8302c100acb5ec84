// surf_shade.cuh - part of libsurf_b200.so (single translation unit: included by surf_kernels.cu inside namespace surf).
// k_shade (resolve + Phong epilogue) and the shadow-ray kernels
#pragma once

// ---------------------------------------------------------------------------------------------------
// k_shade: resolve + Phong shading epilogue
// ---------------------------------------------------------------------------------------------------
struct ShadeParams {
    SceneView sc;
    const CamState* cam;
    const float* rays;
    const unsigned long long* zbuf;
    const float* vis;     // [L, n] or null
    int pix0, n;
    ShadeFlags fl;
    float* image; float* depth; float* normal; float* pos; long long* nearest;
};

__device__ __forceinline__ void pixel_ray(const CamState& cs, const float* rays, int n, int pix0, int k, Vec3* o, Vec3* d) {
    if (cs.proj == 0) {
        *o = v3(cs.eye[0], cs.eye[1], cs.eye[2]);
        *d = v3(rays[k], rays[(size_t)n + k], rays[2 * (size_t)n + k]);
    } else {
        *o = pixel_ray_origin_ortho(cs, pix0 + k);
        *d = v3(cs.odir[0], cs.odir[1], cs.odir[2]);
    }
}

// cooperative [256,3] -> coalesced store through shared memory
__device__ __forceinline__ void store3(float* __restrict__ dst, float (*sm)[3], int base, int n, const float v[3]) {
    const int tid = threadIdx.x;
    __syncthreads();
    sm[tid][0] = v[0]; sm[tid][1] = v[1]; sm[tid][2] = v[2];
    __syncthreads();
    const float* flat = &sm[0][0];
    const int lim = min(256, n - base) * 3;
    for (int j = tid; j < lim; j += 256) dst[(size_t)base * 3 + j] = flat[j];
}

__device__ __forceinline__ void shade_body(const ShadeParams& p, float (*sm)[3], int block);

__global__ void __launch_bounds__(256) k_shade(const __grid_constant__ ShadeParams p) {
    __shared__ float sm[256][3];
    shade_body(p, sm, blockIdx.x);
}

// strided batch: blockIdx.y = scene; outputs are [B, n, ...]
__global__ void __launch_bounds__(256) k_shade_batch(const __grid_constant__ ShadeParams p0, const __grid_constant__ BatchArgs ba) {
    __shared__ float sm[256][3];
    __shared__ ShadeParams p;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) {
        p = p0;
        scene_at(&p.sc, ba, b);
        p.cam = ws_at(p0.cam, ba, b); p.rays = ws_at(p0.rays, ba, b); p.zbuf = ws_at(p0.zbuf, ba, b);
        const long long n = p0.n;
        p.image = adv(p0.image, b * n * 3); p.depth = adv(p0.depth, b * n); p.normal = adv(p0.normal, b * n * 3);
        p.pos = adv(p0.pos, b * n * 3); p.nearest = adv(p0.nearest, b * n);
    }
    __syncthreads();
    shade_body(p, sm, blockIdx.x);
}

__device__ __forceinline__ void shade_body(const ShadeParams& p, float (*sm)[3], int block) {
    const int base = block * 256;
    const int k = base + threadIdx.x;
    const bool live = k < p.n;
    PixelOut po;
    if (live) {
        Vec3 o, d;
        pixel_ray(*p.cam, p.rays, p.n, p.pix0, k, &o, &d);
        float vis_l[16];
        const float* vis = nullptr;
        if (p.vis) {
            for (int l = 0; l < p.sc.n_lights && l < 16; ++l) vis_l[l] = p.vis[(size_t)l * p.n + k];
            vis = vis_l;
        }
        po = resolve_pixel(p.sc, *p.cam, o, d, p.zbuf[k], p.fl, vis);
        if (p.depth) p.depth[k] = po.depth;
        if (p.nearest) p.nearest[k] = po.nearest;
    } else {
        po = PixelOut();
    }
    if (p.image) store3(p.image, sm, base, p.n, po.image);
    if (p.normal) store3(p.normal, sm, base, p.n, po.normal);
    if (p.pos) store3(p.pos, sm, base, p.n, po.pos);
}

// ---------------------------------------------------------------------------------------------------
// shadow rays (renderer.py:291-314): per light, a ray from frag_pos + 0.1 L toward the light against all
// primitives; the light is visible iff nothing is hit strictly between 0 and |L|, or the nearest such hit
// is the fragment's own primitive.  Per-pixel origins -> exact tests over the raw arrays.
// ---------------------------------------------------------------------------------------------------
struct ShadowParams {
    SceneView sc;
    const CamState* cam;
    const float* rays;
    const unsigned long long* zbuf;
    float* vis;      // [L, n]
    int pix0, n;
};

__global__ void __launch_bounds__(128) k_shadow(const __grid_constant__ ShadowParams p) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int l = blockIdx.y;
    if (k >= p.n) return;
    Vec3 o, d;
    pixel_ray(*p.cam, p.rays, p.n, p.pix0, k, &o, &d);
    const unsigned long long key = p.zbuf[k];
    const int self = key == kMissKey ? 0 : (int)(key & 0xFFFFFFFFull);
    Fragment f = fragment_at(p.sc, self, o, d);
    p.vis[(size_t)l * p.n + k] = shadow_visibility(p.sc, f.P, self, l);
}

// Shadow rays (renderer.py:293-299) of ALL lights in one launch (blockIdx.y = light): origin frag_pos + 0.1 L,
// direction L, t_max = |light - frag_pos|.  Only HIT pixels cast shadow rays (a miss pixel's visibility never reaches
// an output: its image is masked), and the rays are COMPACTED: ray (light l, pixel k) is stored in slot
// j = slot_of[l*n + k] of `gray` / `zbuf2` (warp-aggregated atomic counter), so ONE k_intersect_rays launch walks the
// *n_live rays of every light - half the work on config E, a quarter on the bunny frame, and L times more parallel
// work per launch on small frames.  The order of the slots varies from run to run; every ray's result is independent
// of its slot.  `cap` = n * n_lights is the row stride of `gray`.
__global__ void __launch_bounds__(256) k_rays_shadow(const __grid_constant__ ShadowParams p, size_t cap, float* __restrict__ gray,
                                                     unsigned long long* __restrict__ zbuf2, float* __restrict__ obound,
                                                     int* __restrict__ n_live, int* __restrict__ slot_of, int per_light) {
    // per_light: light l's rays are compacted into their own range [l*n, l*n + n_live[l]) (k_intersect_shadow walks
    // the lights one after the other with per-light records); otherwise one list [0, n_live[0]) for all lights
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int l = blockIdx.y;
    float len = 0.f;
    Vec3 so = v3(0.f, 0.f, 0.f), L = v3(0.f, 0.f, 0.f);
    float dist = 0.f;
    bool live = false;
    if (k < p.n) {
        const unsigned long long key = p.zbuf[k];
        if (key != kMissKey) {
            live = true;
            Vec3 o, d;
            pixel_ray(*p.cam, p.rays, p.n, p.pix0, k, &o, &d);
            Fragment f = fragment_at(p.sc, (int)(key & 0xFFFFFFFFull), o, d);
            Vec3 Lv = vsub(ld3(p.sc.light_pos + (size_t)l * p.sc.light_pos_stride), f.P);
            dist = xsqrt(sq3_seq(Lv));
            L = v3(xdiv(Lv.x, dist), xdiv(Lv.y, dist), xdiv(Lv.z, dist));
            so = vadd(f.P, vscale(0.1f, L));
            len = sqrtf(so.x * so.x + so.y * so.y + so.z * so.z);
            if (!(len == len) || !(dist == dist) || isinf(len)) { L = v3(0.f, 0.f, 0.f); len = 0.f; }
        }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, live);
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && ballot) base = per_light ? l * p.n + atomicAdd(n_live + l, __popc(ballot)) : atomicAdd(n_live, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (k < p.n) {
        const int j = live ? base + __popc(ballot & ((1u << lane) - 1u)) : -1;
        slot_of[(size_t)l * p.n + k] = j;
        if (live) {
            zbuf2[j] = kMissKey;
            gray[j] = so.x; gray[cap + j] = so.y; gray[2 * cap + j] = so.z;
            gray[3 * cap + j] = L.x; gray[4 * cap + j] = L.y; gray[5 * cap + j] = L.z;
            gray[6 * cap + j] = dist;
        }
    }
    for (int off = 16; off > 0; off >>= 1) len = fmaxf(len, __shfl_xor_sync(0xffffffffu, len, off));
    if (lane == 0 && len > 0.f) atomicMax((int*)obound, __float_as_int(len));
}

// visible iff nothing was hit inside (0, |L|), or the nearest such hit is the fragment's own primitive (:306-309);
// one thread per (light, pixel), vis is [L, n]
__global__ void __launch_bounds__(256) k_shadow_resolve(const unsigned long long* __restrict__ zbuf,
                                                        const unsigned long long* __restrict__ zbuf2,
                                                        const int* __restrict__ slot_of, int n, size_t total, float* __restrict__ vis) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int k = (int)(idx % (size_t)n);
    const int j = slot_of[idx];
    const unsigned long long self = zbuf[k], hit = j >= 0 ? zbuf2[j] : kMissKey;
    const bool visible = hit == kMissKey || self == kMissKey || (unsigned)(hit & 0xFFFFFFFFull) == (unsigned)(self & 0xFFFFFFFFull);
    vis[idx] = visible ? 1.f : 0.f;
}
