// surf_shade.cuh - part of libsurf_b200.so (included by surf_kernels.cu inside namespace surf).
// k_shade (resolve + Phong epilogue, optionally fused with the MSE loss of an inverse-rendering step) and the
// shadow-ray list kernels
#pragma once

// ---------------------------------------------------------------------------------------------------
// k_shade: resolve + Phong shading epilogue (renderer.py:183-189, 266-340)
// ---------------------------------------------------------------------------------------------------
#ifndef SURF_SHADE_BLOCKS
#define SURF_SHADE_BLOCKS 4        // resident CTAs per SM k_shade is compiled for (4 -> 64 registers)
#endif
#ifndef SURF_SHADE_PX
#define SURF_SHADE_PX 2            // pixels per thread on large frames
#endif
constexpr int kShadeBlocksPerSM = SURF_SHADE_BLOCKS;
constexpr int kShadePX = SURF_SHADE_PX;

struct ShadeParams {
    SceneView sc;
    const CamState* cam;
    const float4* packed;   // per-primitive records of k_prep*: .xyz of a planar primitive's first float4 = unit normal
    const float* rays;
    const unsigned long long* zbuf;
    const float* vis;     // [L, n] or null
    int pix0, n;
    ShadeFlags fl;
    float* image; float* depth; float* normal; float* pos; long long* nearest;
    // fused loss of an inverse-rendering step (surf_step_mse; test_optimization.py:104): with `target` [n,3] set, the
    // kernel also writes g_image = 2 loss_scale (image - target) and adds loss_scale * sum((image - target)^2) to *loss_acc
    const float* target; float* g_image; double* loss_acc; float loss_scale;
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

// cooperative [256,3] -> coalesced store through shared memory (k_splat_forward)
__device__ __forceinline__ void store3(float* __restrict__ dst, float (*sm)[3], int base, int n, const float v[3]) {
    const int tid = threadIdx.x;
    __syncthreads();
    sm[tid][0] = v[0]; sm[tid][1] = v[1]; sm[tid][2] = v[2];
    __syncthreads();
    const float* flat = &sm[0][0];
    const int lim = min(256, n - base) * 3;
    for (int j = tid; j < lim; j += 256) dst[(size_t)base * 3 + j] = flat[j];
}

// winner of pixel k -> outputs.  Hit pixels: depth is the t the exact narrow phase stored in the z-buffer key, the
// unit normal of a planar primitive comes from its k_prep record (both bit-identical to the reference-order
// recomputation), shading uses the fast forms of surf_fast.cuh.  Miss pixels report primitive 0 at its unmasked
// distance (argmin of an all-miss column) through the reference-order path.
__device__ __forceinline__ PixelOut resolve_fast(const ShadeParams& p, const CamState& cs, const LightS* lights, Vec3 eye,
                                                 Vec3 o, Vec3 d, unsigned long long key, int k) {
    PixelOut po;
    if (key == kMissKey) {
        const Fragment f = fragment_at(p.sc, 0, o, d);
        po.nearest = 0;
        po.depth = cs.far_plus1;
        po.normal[0] = f.n.x; po.normal[1] = f.n.y; po.normal[2] = f.n.z;
        po.pos[0] = f.P.x; po.pos[1] = f.P.y; po.pos[2] = f.P.z;
        po.image[0] = po.image[1] = po.image[2] = 0.f;
        return po;
    }
    const int idx = (int)(key & 0xFFFFFFFFull);
    const float t = float_from_order_key((uint32_t)(key >> 32));
    const int set = find_set(p.sc, idx);
    const SetView& sv = p.sc.sets[set];
    const int local = idx - sv.first;
    const Vec3 P = ray_point(o, t, d);
    Vec3 n;
    if (sv.kind == KIND_SPHERE) {
        n = unit_eps(vsub(P, ld3(sv.pos + (size_t)local * sv.pos_stride)), nullptr);          // utils.py:275
    } else {
        const float4 A = p.packed[sv.rec_off + (size_t)local * rec_f4(sv.kind)];
        n = v3(A.x, A.y, A.z);
    }
    const MatF mt = load_material(p.sc, clampi(sv.mat[local], 0, p.sc.n_materials - 1));
    float lit[3];
    shade_fast(p.sc, lights, eye, P, n, mt, p.fl, p.vis ? p.vis + k : nullptr, (size_t)p.n, lit);
    composite_fast(lit, true, p.sc.gamma, po.image);
    po.nearest = idx;
    po.depth = t;
    po.normal[0] = n.x; po.normal[1] = n.y; po.normal[2] = n.z;
    po.pos[0] = P.x; po.pos[1] = P.y; po.pos[2] = P.z;
    return po;
}

// PX consecutive pixels per thread.  With n % PX == 0 every global access of the kernel is a vector: PX = 2 -> rays
// 3 x LDG.64, keys 1 x LDG.128, image / normal / pos / d(image) 3 x STG.64 each, depth STG.64, nearest STG.128.
// (PX = 4 doubles the widths but needs 4 x 11 output registers live at once; it spills at the 80 registers three
// resident CTAs allow, and this kernel is latency-bound - dependent key -> record -> material gathers - so
// occupancy wins.)
// Persistent CTAs: the grid is sized to one wave (launch_shade) and every CTA strides over the 256*PX-pixel tiles, so the
// light table is staged (and its barrier paid) once per CTA and no partial last wave is left idle.
template <int PX> struct VecF;
template <> struct VecF<1> { typedef float T; };
template <> struct VecF<2> { typedef float2 T; };
template <> struct VecF<4> { typedef float4 T; };

template <int PX>
__device__ __forceinline__ void load_px(const float* __restrict__ src, float out[PX]) {
    const typename VecF<PX>::T v = *reinterpret_cast<const typename VecF<PX>::T*>(src);
    const float* f = reinterpret_cast<const float*>(&v);
#pragma unroll
    for (int j = 0; j < PX; ++j) out[j] = f[j];
}
// 3*PX contiguous floats at dst (12*PX-byte aligned): three vector stores
template <int PX>
__device__ __forceinline__ void store_3px(float* __restrict__ dst, const float f[3 * PX]) {
    typedef typename VecF<PX>::T V;
    V* q = reinterpret_cast<V*>(dst);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        V v;
        float* w = reinterpret_cast<float*>(&v);
#pragma unroll
        for (int j = 0; j < PX; ++j) w[j] = f[i * PX + j];
        q[i] = v;
    }
}

template <int PX>
__device__ __forceinline__ void shade_body(const ShadeParams& p, LightS* lights, float* red, int block, int n_blocks) {
    stage_lights(p.sc, lights);
    const CamState& cs = *p.cam;
    const Vec3 eye = v3(cs.eye[0], cs.eye[1], cs.eye[2]);
    const bool vec = PX > 1 && (p.n % PX) == 0;
    const bool persp = cs.proj == 0;
    float err = 0.f;
    const int n_tiles = (p.n + 256 * PX - 1) / (256 * PX);
    for (int tile = block; tile < n_tiles; tile += n_blocks) {
        const int k0 = (tile * 256 + threadIdx.x) * PX;
        if (k0 >= p.n) continue;
        unsigned long long key[PX];
        float dx[PX], dy[PX], dz[PX];
        if (vec) {
            if (PX == 1) key[0] = p.zbuf[k0];
            else {
#pragma unroll
                for (int j = 0; j < PX; j += 2) {
                    const ulonglong2 kk = *reinterpret_cast<const ulonglong2*>(p.zbuf + k0 + j);
                    key[j] = kk.x; key[j + (PX > 1 ? 1 : 0)] = kk.y;
                }
            }
            if (persp) {
                load_px<PX>(p.rays + k0, dx);
                load_px<PX>(p.rays + (size_t)p.n + k0, dy);
                load_px<PX>(p.rays + 2 * (size_t)p.n + k0, dz);
            }
        } else {
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                const int k = min(k0 + j, p.n - 1);
                key[j] = p.zbuf[k];
                if (persp) { dx[j] = p.rays[k]; dy[j] = p.rays[(size_t)p.n + k]; dz[j] = p.rays[2 * (size_t)p.n + k]; }
            }
        }
        float img[3 * PX], nrm[3 * PX], pos[3 * PX], gi[3 * PX], dep[PX];
        long long nea[PX];
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            const int k = min(k0 + j, p.n - 1);
            Vec3 o = eye, d;
            if (persp) d = v3(dx[j], dy[j], dz[j]);
            else { o = pixel_ray_origin_ortho(cs, p.pix0 + k); d = v3(cs.odir[0], cs.odir[1], cs.odir[2]); }
            const PixelOut po = resolve_fast(p, cs, lights, eye, o, d, key[j], k);
#pragma unroll
            for (int c = 0; c < 3; ++c) { img[3 * j + c] = po.image[c]; nrm[3 * j + c] = po.normal[c]; pos[3 * j + c] = po.pos[c]; }
            dep[j] = po.depth;
            nea[j] = po.nearest;
            if (p.target) {
                const bool live = k0 + j < p.n;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float diff = po.image[c] - p.target[(size_t)k * 3 + c];
                    gi[3 * j + c] = 2.f * p.loss_scale * diff;
                    if (live) err = fmaf(diff, diff, err);
                }
            }
        }
        if (vec) {
            if (p.image) store_3px<PX>(p.image + (size_t)k0 * 3, img);
            if (p.normal) store_3px<PX>(p.normal + (size_t)k0 * 3, nrm);
            if (p.pos) store_3px<PX>(p.pos + (size_t)k0 * 3, pos);
            if (p.target && p.g_image) store_3px<PX>(p.g_image + (size_t)k0 * 3, gi);
            if (p.depth) {
                typename VecF<PX>::T v;
                float* w = reinterpret_cast<float*>(&v);
#pragma unroll
                for (int j = 0; j < PX; ++j) w[j] = dep[j];
                *reinterpret_cast<typename VecF<PX>::T*>(p.depth + k0) = v;
            }
            if (p.nearest) {
                if (PX == 1) p.nearest[k0] = nea[0];
                else {
#pragma unroll
                    for (int j = 0; j < PX; j += 2)
                        *reinterpret_cast<longlong2*>(p.nearest + k0 + j) = make_longlong2(nea[j], nea[j + (PX > 1 ? 1 : 0)]);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                const int k = k0 + j;
                if (k >= p.n) break;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (p.image) p.image[(size_t)k * 3 + c] = img[3 * j + c];
                    if (p.normal) p.normal[(size_t)k * 3 + c] = nrm[3 * j + c];
                    if (p.pos) p.pos[(size_t)k * 3 + c] = pos[3 * j + c];
                    if (p.target && p.g_image) p.g_image[(size_t)k * 3 + c] = gi[3 * j + c];
                }
                if (p.depth) p.depth[k] = dep[j];
                if (p.nearest) p.nearest[k] = nea[j];
            }
        }
    }
    if (p.target) {          // CTA-uniform
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = err;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (int w = 0; w < 8; ++w) s += red[w];
            if (s != 0.f) atomicAdd(p.loss_acc, (double)s * (double)p.loss_scale);
        }
    }
}

template <int PX>
__global__ void __launch_bounds__(256, kShadeBlocksPerSM) k_shade(const __grid_constant__ ShadeParams p) {
    __shared__ LightS lights[kLightTable];
    __shared__ float red[8];
    shade_body<PX>(p, lights, red, blockIdx.x, gridDim.x);
}

// strided batch: blockIdx.y = scene; outputs are [B, n, ...]
template <int PX>
__global__ void __launch_bounds__(256, kShadeBlocksPerSM) k_shade_batch(const __grid_constant__ ShadeParams p0, const __grid_constant__ BatchArgs ba) {
    __shared__ LightS lights[kLightTable];
    __shared__ float red[8];
    __shared__ ShadeParams p;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) {
        p = p0;
        scene_at(&p.sc, ba, b);
        p.cam = ws_at(p0.cam, ba, b); p.packed = ws_at(p0.packed, ba, b); p.rays = ws_at(p0.rays, ba, b);
        p.zbuf = ws_at(p0.zbuf, ba, b);
        const long long n = p0.n;
        p.image = adv(p0.image, b * n * 3); p.depth = adv(p0.depth, b * n); p.normal = adv(p0.normal, b * n * 3);
        p.pos = adv(p0.pos, b * n * 3); p.nearest = adv(p0.nearest, b * n);
        p.target = adv(p0.target, b * n * 3); p.g_image = adv(p0.g_image, b * n * 3);
    }
    __syncthreads();
    shade_body<PX>(p, lights, red, blockIdx.x, gridDim.x);
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
