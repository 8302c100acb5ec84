// surf_backward.cuh - part of libsurf_b200.so (single translation unit: included by surf_kernels.cu inside namespace surf).
// k_backward / k_backward_finalize / k_mse_grad and the gradient sink
#pragma once

// ---------------------------------------------------------------------------------------------------
// k_backward
// ---------------------------------------------------------------------------------------------------
struct GradPtrs {
    float* prim_pos[kMaxSets]; float* prim_normal[kMaxSets]; float* prim_radius[kMaxSets];
    float* light_pos; float* atten; float* ambient; float* colors; float* albedo; float* coeffs; float* gamma;
};
// accumulator slot map (doubles): [albedo K*3][coeffs K*3][light_pos L*3][atten L*3][colors C*3][ambient 3][gamma 1]
struct SlotMap { int albedo, coeffs, light_pos, atten, colors, ambient, gamma, total; };

__host__ __device__ inline SlotMap slot_map(int K, int L, int Cn) {
    SlotMap m;
    m.albedo = 0; m.coeffs = K * 3; m.light_pos = m.coeffs + K * 3; m.atten = m.light_pos + L * 3;
    m.colors = m.atten + L * 3; m.ambient = m.colors + Cn * 3; m.gamma = m.ambient + 3; m.total = m.gamma + 1;
    return m;
}

struct BackwardParams {
    SceneView sc;
    const CamState* cam;
    const float* rays;
    const float* vis;
    const long long* nearest;
    const float* depth;
    const float* g_image; const float* g_depth; const float* g_normal; const float* g_pos;
    int pix0, n;
    ShadeFlags fl;
    GradPtrs gp;
    SlotMap sm;
    double* acc;
    double* prim_acc;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct DeviceSink {
    const BackwardParams& p;
    double* cta_acc;           // shared, [sm.total]
    float alb[3], cf[3], amb[3], gam;
    float lp[3], at[3], col[3];
    __device__ DeviceSink(const BackwardParams& prm, double* shared_acc) : p(prm), cta_acc(shared_acc) {
        for (int c = 0; c < 3; ++c) alb[c] = cf[c] = amb[c] = lp[c] = at[c] = col[c] = 0.f;
        gam = 0.f;
    }
    __device__ void albedo(int, int c, float v) { alb[c] += v; }
    __device__ void coeff(int, int c, float v) { cf[c] += v; }
    __device__ void ambient(int c, float v) { amb[c] += v; }
    __device__ void gamma(float v) { gam += v; }
    __device__ void light_pos(int, int c, float v) { lp[c] += v; }
    __device__ void atten(int, int c, float v) { at[c] += v; }
    __device__ void color(int, int c, float v) { col[c] += v; }
    __device__ void add_cta(int slot, float warp_total) { atomicAdd(&cta_acc[slot], (double)warp_total); }
    __device__ void end_light(int l, int crow) {
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float a = warp_sum(lp[c]), b = warp_sum(at[c]), e = warp_sum(col[c]);
            if (lane == 0) {
                if (a != 0.f) add_cta(p.sm.light_pos + l * 3 + c, a);
                if (b != 0.f) add_cta(p.sm.atten + l * 3 + c, b);
                if (e != 0.f) add_cta(p.sm.colors + crow * 3 + c, e);
            }
            lp[c] = at[c] = col[c] = 0.f;
        }
    }
    __device__ void end_splat(int m) { flush_scalars(m); }
    __device__ void end_pixel(int set, int local, int idx, int m, const float* g7) {
        flush_scalars(m);
        flush_primitive(set, local, idx, g7);
    }
    __device__ void flush_scalars(int m) {
        const int lane = threadIdx.x & 31;
        // global scalars
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float a = warp_sum(amb[c]);
            if (lane == 0 && a != 0.f) add_cta(p.sm.ambient + c, a);
        }
        float gsum = warp_sum(gam);
        if (lane == 0 && gsum != 0.f) add_cta(p.sm.gamma, gsum);
        // per-material rows: warp-uniform material is the common case (splat scenes use one material)
        const int m0 = __shfl_sync(0xffffffffu, m, 0);
        if (__all_sync(0xffffffffu, m == m0)) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float a = warp_sum(alb[c]), b = warp_sum(cf[c]);
                if (lane == 0) {
                    if (a != 0.f) add_cta(p.sm.albedo + m0 * 3 + c, a);
                    if (b != 0.f) add_cta(p.sm.coeffs + m0 * 3 + c, b);
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (alb[c] != 0.f) atomicAdd(&cta_acc[p.sm.albedo + m * 3 + c], (double)alb[c]);
                if (cf[c] != 0.f) atomicAdd(&cta_acc[p.sm.coeffs + m * 3 + c], (double)cf[c]);
            }
        }
    }
    __device__ void flush_primitive(int set, int local, int idx, const float* g7) {
        const int lane = threadIdx.x & 31;
        // per-primitive gradients: warp-segmented reduction keyed by the winner index, then one
        // red.global.add per component from the segment leader
        const unsigned peers = __match_any_sync(0xffffffffu, idx);
        const int leader = __ffs(peers) - 1;
        float v[7];
#pragma unroll
        for (int c = 0; c < 7; ++c) v[c] = g7[c];
        unsigned rest = peers & ~(1u << leader);
        // every lane walks the union of peer sets in lock-step (max 31 steps, usually 0-3)
        const unsigned any_rest = __reduce_or_sync(0xffffffffu, rest);
        if (any_rest) {
            for (int src = 0; src < 32; ++src) {
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    float o = __shfl_sync(0xffffffffu, g7[c], src);
                    if (lane == leader && ((rest >> src) & 1u)) v[c] += o;
                }
            }
        }
        if (lane == leader) {
            // double accumulation: per-pixel contributions of a grazing primitive cancel heavily, and a
            // sequential fp32 atomic sum would carry ~1e-4 relative noise (the reference sums pairwise)
            double* dst = p.prim_acc + (size_t)idx * 7;
#pragma unroll
            for (int c = 0; c < 7; ++c)
                if (v[c] != 0.f) atomicAdd(dst + c, (double)v[c]);
        }
        (void)set; (void)local;
    }
};

// Persistent grid-stride CTAs: the light / material / colour accumulators live in shared memory for the whole life
// of the CTA and are flushed with one double atomicAdd per slot per CTA.  (One flush per 128 pixels put ~8000
// same-address double atomics per slot on the L2 and made the kernel 5x slower than its instruction count.)
__device__ __forceinline__ void backward_body(const BackwardParams& p, double* cta_acc);

__global__ void __launch_bounds__(128, 6) k_backward(const __grid_constant__ BackwardParams p) {
    __shared__ double cta_acc[kMaxAccSlots];
    backward_body(p, cta_acc);
}

__device__ __forceinline__ void grads_at(GradPtrs* gp, const BatchArgs& ba, int b) {
    for (int k = 0; k < kMaxSets; ++k) {
        gp->prim_pos[k] = adv(gp->prim_pos[k], b * ba.set_pos[k]);
        gp->prim_normal[k] = adv(gp->prim_normal[k], b * ba.set_normal[k]);
        gp->prim_radius[k] = adv(gp->prim_radius[k], b * ba.set_radius[k]);
    }
    gp->light_pos = adv(gp->light_pos, b * ba.light_pos);
    gp->atten = adv(gp->atten, b * ba.light_atten);
    gp->ambient = adv(gp->ambient, b * ba.ambient);
    gp->colors = adv(gp->colors, b * ba.colors);
    gp->albedo = adv(gp->albedo, b * ba.albedo);
    gp->coeffs = adv(gp->coeffs, b * ba.coeffs);
    gp->gamma = adv(gp->gamma, b * ba.gamma);
}

// strided batch: blockIdx.y = scene; nearest / depth / incoming gradients are [B, n, ...]
__global__ void __launch_bounds__(128, 6) k_backward_batch(const __grid_constant__ BackwardParams p0,
                                                        const __grid_constant__ BatchArgs ba) {
    __shared__ double cta_acc[kMaxAccSlots];
    __shared__ BackwardParams p;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) {
        p = p0;
        scene_at(&p.sc, ba, b);
        grads_at(&p.gp, ba, b);
        p.cam = ws_at(p0.cam, ba, b); p.rays = ws_at(p0.rays, ba, b);
        p.acc = ws_at(p0.acc, ba, b); p.prim_acc = ws_at(p0.prim_acc, ba, b);
        const long long n = p0.n;
        p.nearest = adv(p0.nearest, b * n); p.depth = adv(p0.depth, b * n);
        p.g_image = adv(p0.g_image, b * n * 3); p.g_depth = adv(p0.g_depth, b * n);
        p.g_normal = adv(p0.g_normal, b * n * 3); p.g_pos = adv(p0.g_pos, b * n * 3);
    }
    __syncthreads();
    backward_body(p, cta_acc);
}

__device__ __forceinline__ void backward_body(const BackwardParams& p, double* cta_acc) {
    for (int j = threadIdx.x; j < p.sm.total; j += blockDim.x) cta_acc[j] = 0.0;
    __syncthreads();
    const Vec3 eye = v3(p.cam->eye[0], p.cam->eye[1], p.cam->eye[2]);
    for (int base = blockIdx.x * blockDim.x; base < p.n; base += gridDim.x * blockDim.x) {   // uniform trip count per CTA
        const int k = base + threadIdx.x;
        const bool live = k < p.n;
        const int kk = live ? k : p.n - 1;      // dead lanes shadow the last pixel with zero incoming gradients
        const float dep = p.depth[kk];
        const bool hit = live && dep <= p.cam->far_clip && dep >= p.cam->near_clip;
        PixelGrads g;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            g.image[c] = (live && p.g_image) ? p.g_image[(size_t)kk * 3 + c] : 0.f;
            g.pos[c] = (live && p.g_pos) ? p.g_pos[(size_t)kk * 3 + c] : 0.f;
            g.normal[c] = (live && p.g_normal) ? p.g_normal[(size_t)kk * 3 + c] : 0.f;
        }
        g.depth = (live && p.g_depth) ? p.g_depth[kk] : 0.f;
        // a warp whose pixels are all misses with no incoming geometry gradient has nothing to contribute
        const bool needed = hit || g.pos[0] != 0.f || g.pos[1] != 0.f || g.pos[2] != 0.f ||
                            g.normal[0] != 0.f || g.normal[1] != 0.f || g.normal[2] != 0.f;
        if (!__any_sync(0xffffffffu, needed)) continue;
        Vec3 o, d;
        pixel_ray(*p.cam, p.rays, p.n, p.pix0, kk, &o, &d);
        float vis_l[16];
        const float* vis = nullptr;
        if (p.vis) {
            for (int l = 0; l < p.sc.n_lights && l < 16; ++l) vis_l[l] = p.vis[(size_t)l * p.n + kk];
            vis = vis_l;
        }
        DeviceSink sink(p, cta_acc);
        backward_pixel(p.sc, eye, o, d, (int)p.nearest[kk], hit, p.fl, vis, g, sink);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < p.sm.total; j += blockDim.x)
        if (cta_acc[j] != 0.0) atomicAdd(p.acc + j, cta_acc[j]);
}

struct FinalizeParams {
    GradPtrs gp; SlotMap sm; const double* acc; int K, L, Cn, light_pos_stride;
    SceneView sc; const double* prim_acc;
};
// The leaves are updated with atomics: in a batch several scenes (concurrent streams, or the scene dimension of
// k_backward_finalize_batch) may share one gradient array (SurfBatchLayout stride 0, or aliased pointers).
__device__ __forceinline__ void finalize_body(const FinalizeParams& p, int j) {
    if (j < p.sc.total) {      // per-primitive accumulators -> fp32 leaves (caller's strides)
        const int s = find_set(p.sc, j);
        const SetView& sv = p.sc.sets[s];
        const int local = j - sv.first;
        const double* a = p.prim_acc + (size_t)j * 7;
        float* gpos = p.gp.prim_pos[s];
        if (gpos) {
            const size_t row = sv.kind == KIND_TRIANGLE ? (size_t)local * 3 * sv.pos_stride : (size_t)local * sv.pos_stride;
            for (int c = 0; c < 3; ++c) atomicAdd(gpos + row + c, (float)a[c]);
        }
        float* gnr = p.gp.prim_normal[s];
        if (gnr && sv.kind != KIND_SPHERE)
            for (int c = 0; c < 3; ++c) atomicAdd(gnr + (size_t)local * sv.normal_stride + c, (float)a[3 + c]);
        float* grd = p.gp.prim_radius[s];
        if (grd && sv.kind == KIND_SPHERE) atomicAdd(grd + local, (float)a[6]);
        return;
    }
    j -= p.sc.total;
    if (j >= p.sm.total) return;
    const float v = (float)p.acc[j];
    if (j < p.sm.coeffs) { if (p.gp.albedo) atomicAdd(p.gp.albedo + (j - p.sm.albedo), v); }
    else if (j < p.sm.light_pos) { if (p.gp.coeffs) atomicAdd(p.gp.coeffs + (j - p.sm.coeffs), v); }
    else if (j < p.sm.atten) {
        const int q = j - p.sm.light_pos;
        if (p.gp.light_pos) atomicAdd(p.gp.light_pos + (size_t)(q / 3) * p.light_pos_stride + q % 3, v);
    }
    else if (j < p.sm.colors) { if (p.gp.atten) atomicAdd(p.gp.atten + (j - p.sm.atten), v); }
    else if (j < p.sm.ambient) { if (p.gp.colors) atomicAdd(p.gp.colors + (j - p.sm.colors), v); }
    else if (j < p.sm.gamma) { if (p.gp.ambient) atomicAdd(p.gp.ambient + (j - p.sm.ambient), v); }
    else { if (p.gp.gamma) atomicAdd(p.gp.gamma, v); }
}

__global__ void __launch_bounds__(128) k_backward_finalize(const __grid_constant__ FinalizeParams p) {
    finalize_body(p, blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void __launch_bounds__(128) k_backward_finalize_batch(const __grid_constant__ FinalizeParams p0,
                                                                 const __grid_constant__ BatchArgs ba) {
    __shared__ FinalizeParams p;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) {
        p = p0;
        scene_at(&p.sc, ba, b);
        grads_at(&p.gp, ba, b);
        p.acc = ws_at(p0.acc, ba, b); p.prim_acc = ws_at(p0.prim_acc, ba, b);
    }
    __syncthreads();
    finalize_body(p, blockIdx.x * blockDim.x + threadIdx.x);
}
