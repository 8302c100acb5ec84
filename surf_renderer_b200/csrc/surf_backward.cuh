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

// ---------------------------------------------------------------------------------------------------
// k_backward (fast forms, surf_fast.cuh): thread per pixel, persistent grid-stride CTAs.
//   * one pass over the lights for the composite (dL/dI needs the sum over lights), one for the gradients; both use
//     the SFU approximations - gradients are held to 1e-4, not to bit-exactness;
//   * hit pixels take t from the saved depth, so no division sits on the geometry path;
//   * light / material / colour / ambient / gamma gradients: per light the nine lane values (light position 3,
//     attenuation 3, colour row 3) are summed over the warp with warp_sum8 (9 shuffles for 8 values) + one plain
//     butterfly, then added by the owning lanes to this WARP's private float accumulators in shared memory (no
//     atomics); the CTA folds them into the double accumulators in the workspace once, at the end of its life;
//   * per-primitive gradients: pixels of one warp are consecutive in a row, so lanes that hit the same primitive
//     are (almost always) adjacent - a segmented shuffle-down scan (5 steps x 6 values) feeds one set of double
//     atomics per run of equal winners.  Equal winners that are not adjacent simply issue two sets of atomics.
// Accumulator slots beyond kMaxAccSlots (scenes with hundreds of materials / lights) skip shared memory and the
// double buffer: their contributions go straight into the fp32 leaf gradients with float atomics.
// ---------------------------------------------------------------------------------------------------
struct BwdAcc {
    const GradPtrs& gp;
    const SlotMap& sm;
    int light_pos_stride;
    float* wacc;                   // this warp's [kMaxAccSlots] float accumulators (shared memory)
    // called by at most one lane per slot per instruction
    __device__ __forceinline__ void add(int slot, float v) const {
        if (v == 0.f) return;
        if (slot < kMaxAccSlots) wacc[slot] += v;
        else spill(slot, v);
    }
    // several lanes of the warp may name the same slot (pixels with different materials in one warp)
    __device__ __forceinline__ void add_shared(int slot, float v) const {
        if (v == 0.f) return;
        if (slot < kMaxAccSlots) atomicAdd(&wacc[slot], v);
        else spill(slot, v);
    }
    __device__ __noinline__ void spill(int j, float v) const {
        if (j < sm.coeffs) { if (gp.albedo) atomicAdd(gp.albedo + (j - sm.albedo), v); }
        else if (j < sm.light_pos) { if (gp.coeffs) atomicAdd(gp.coeffs + (j - sm.coeffs), v); }
        else if (j < sm.atten) {
            const int q = j - sm.light_pos;
            if (gp.light_pos) atomicAdd(gp.light_pos + (size_t)(q / 3) * light_pos_stride + q % 3, v);
        }
        else if (j < sm.colors) { if (gp.atten) atomicAdd(gp.atten + (j - sm.atten), v); }
        else if (j < sm.ambient) { if (gp.colors) atomicAdd(gp.colors + (j - sm.colors), v); }
        else if (j < sm.gamma) { if (gp.ambient) atomicAdd(gp.ambient + (j - sm.ambient), v); }
        else { if (gp.gamma) atomicAdd(gp.gamma, v); }
    }
};

// sums v[0..N) over runs of adjacent lanes with equal `idx`; the first lane of each run (return value true) ends up
// with the run's totals
template <int N>
__device__ __forceinline__ bool segmented_sum(int idx, float v[N], int lane) {
    const int prev = __shfl_up_sync(0xffffffffu, idx, 1);
    const bool head = lane == 0 || idx != prev;
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const unsigned above = lane == 31 ? 0u : (heads & ~((2u << lane) - 1u));
    const int seg_end = above ? __ffs(above) - 1 : 32;
    if (__popc(heads) == 32) return head;          // every lane its own run: nothing to add
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
#pragma unroll
        for (int c = 0; c < N; ++c) {
            const float o = __shfl_down_sync(0xffffffffu, v[c], off);
            if (lane + off < seg_end) v[c] += o;
        }
    }
    return head;
}

// shading backward of one fragment (renderer.py:82-125, 330-340 differentiated): accumulates the
// light / material / colour / ambient / gamma gradients and ADDS d/dP, d/dn to *gP, *gn.  Called by all 32 lanes.
__device__ __forceinline__ void backward_shading_fast(const SceneView& sc, const LightS* lights, Vec3 eye, Vec3 P, Vec3 n, int m,
                                                      bool hit, ShadeFlags fl, const float* __restrict__ vis, size_t vis_stride,
                                                      const float g_image[3], const BwdAcc& acc, const SlotMap& sm, int lane,
                                                      Vec3* gP_io, Vec3* gn_io) {
    const MatF mt = load_material(sc, m);
    float inv_len = 1.f;
    const Vec3 V = fl.raw_view ? v3(eye.x - P.x, eye.y - P.y, eye.z - P.z) : view_vector(eye, P, &inv_len);
    const float Vn = f_dot(V, n);
    const float sg = fl.double_sided ? sign_or_zero(Vn) : 1.f;
    const float amb[3] = {sc.ambient[0], sc.ambient[1], sc.ambient[2]};
    // pass 1: the composite, for dLoss/d(lit)
    float gI[3] = {0.f, 0.f, 0.f};
    float g_gamma = 0.f;
    bool active = false;
    if (hit) {
        float lit[3];
        shade_fast(sc, lights, eye, P, n, mt, fl, vis, vis_stride, lit);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float Ic = lit[c];
            if (Ic > 0.f && g_image[c] != 0.f) {
                if (sc.gamma) {
                    const float gm = sc.gamma[0];
                    const float lg = f_lg2(Ic);
                    const float pw = gm == 0.f ? 1.f : f_ex2(gm * lg);       // Ic^gm
                    gI[c] = g_image[c] * gm * pw * f_rcp(Ic);
                    g_gamma = fmaf(g_image[c] * pw, lg * kLn2, g_gamma);
                } else {
                    gI[c] = g_image[c];
                }
            }
            active |= gI[c] != 0.f;
        }
    }
    Vec3 gP = *gP_io, gn = *gn_io, gV = v3(0.f, 0.f, 0.f);
    float alb[3] = {0.f, 0.f, 0.f}, cf[3] = {0.f, 0.f, 0.f};
    const int vidx = warp_sum8_index(lane);
    const bool owner = (lane & 3) == 0;
    for (int l = 0; l < sc.n_lights; ++l) {
        const LightS ls = light_at(sc, lights, l);
        const LightF e = eval_light_fast(ls, P, n, V, Vn, fl.use_quartic);
        const float Dsg = sg * e.D, Ssg = sg * e.S;
        const float Dp = fmaxf(Dsg, 0.f), Sp = fmaxf(Ssg, 0.f);
        const float lgS = f_lg2(Sp);
        const float spec = mt.sh == 0.f ? 1.f : f_ex2(mt.sh * lgS);
        const float scal = fmaf(mt.ks, spec, mt.kd * Dp);
        const float v = vis ? vis[(size_t)l * vis_stride] : 1.f;
        const float col[3] = {ls.cr, ls.cg, ls.cb};
        float red[8];          // light_pos 3, atten 3, colour 0..1   (+ colour 2 separately)
        float col2;
        float g_scal = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float g_tint = gI[c] * scal * v;
            g_scal = fmaf(gI[c], col[c] * mt.A[c] * v, g_scal);
            alb[c] = fmaf(g_tint, col[c], fmaf(gI[c], amb[c], alb[c]));
            const float gc = g_tint * mt.A[c];
            if (c < 2) red[6 + c] = active ? gc : 0.f; else col2 = active ? gc : 0.f;
        }
        cf[0] = fmaf(g_scal, Dp, cf[0]);
        cf[1] = fmaf(g_scal, spec, cf[1]);
        const float g_spec = g_scal * mt.ks;
        float g_Sp = 0.f;
        if (Sp > 0.f) {
            cf[2] = fmaf(g_spec * spec, lgS * kLn2, cf[2]);
            if (mt.sh != 0.f) g_Sp = g_spec * mt.sh * spec * f_rcp(Sp);
        }
        const float g_D = (active && Dsg > 0.f) ? g_scal * mt.kd * sg : 0.f;
        const float g_S = (active && Ssg > 0.f) ? g_Sp * sg : 0.f;
        // D = att (n.L):   S = V.R,  R = 2 (n.L) n - L
        const Vec3 R = f_axpy(2.f * e.nL, n, vneg(e.L));
        Vec3 gL = vscale(g_D * e.att, n);
        const float g_att = g_D * e.nL;
        const Vec3 gR = vscale(g_S, V);
        const float g_s = -2.f * f_dot(gR, n);          // s = -(n.L)
        const Vec3 g_inc = f_axpy(g_s, n, gR);          // inc = -L
        gL = v3(gL.x - g_inc.x, gL.y - g_inc.y, gL.z - g_inc.z);
        const float g_den = e.den_nz ? -g_att * e.att * e.att : 0.f;
        const float dpw = fl.use_quartic ? 4.f * e.d2 * e.dl : 2.f * e.dl;
        const float g_dl = g_den * fmaf(ls.a2, dpw, ls.a1);
        Vec3 gLv = gL;
        if (e.dl_nz) {
            const float gl_dot = f_dot(gL, e.L);
            const float rdl = f_rcp(e.dl);
            gLv = v3(fmaf(gL.x - gl_dot * e.L.x, rdl, g_dl * e.L.x), fmaf(gL.y - gl_dot * e.L.y, rdl, g_dl * e.L.y),
                     fmaf(gL.z - gl_dot * e.L.z, rdl, g_dl * e.L.z));
        }
        if (active) {
            gn = f_axpy(g_D * e.att, e.L, gn);
            gV = f_axpy(g_S, R, gV);
            gn = f_axpy(2.f * e.nL, gR, gn);            // -2 s gR
            gn = f_axpy(-g_s, e.L, gn);                 // g_s inc
            gP = v3(gP.x - gLv.x, gP.y - gLv.y, gP.z - gLv.z);
        }
        red[0] = active ? gLv.x : 0.f; red[1] = active ? gLv.y : 0.f; red[2] = active ? gLv.z : 0.f;
        red[3] = active ? g_den : 0.f; red[4] = active ? g_den * e.dl : 0.f; red[5] = active ? g_den * e.pw : 0.f;
        const float tot = warp_sum8(red, lane);
        const float tot2 = warp_sum(col2);
        const int crow = clampi(sc.light_color_idx[l], 0, sc.n_colors - 1);
        if (owner) {
            const int slot = vidx < 3 ? sm.light_pos + l * 3 + vidx
                                      : (vidx < 6 ? sm.atten + l * 3 + (vidx - 3) : sm.colors + crow * 3 + (vidx - 6));
            acc.add(slot, tot);
        }
        if (lane == 1) acc.add(sm.colors + crow * 3 + 2, tot2);
    }
    if (active && fl.raw_view) gP = v3(gP.x - gV.x, gP.y - gV.y, gP.z - gV.z);       // V = eye - P
    else if (active) {   // V = Vv / |Vv|
        const float gv_dot = f_dot(gV, V);
        gP = v3(gP.x - (gV.x - gv_dot * V.x) * inv_len, gP.y - (gV.y - gv_dot * V.y) * inv_len,
                gP.z - (gV.z - gv_dot * V.z) * inv_len);
    }
    // material rows, ambient, gamma.  warp-uniform material is the common case (splat scenes use one material)
    float red[8];
    float amb2, gam;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float ga = active ? gI[c] * mt.A[c] * (float)sc.n_lights : 0.f;
        if (c < 2) red[6 + c] = ga; else amb2 = ga;
    }
    gam = g_gamma;
    const int m0 = __shfl_sync(0xffffffffu, m, 0);
    const bool uniform = __all_sync(0xffffffffu, m == m0);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        red[c] = (uniform && active) ? alb[c] : 0.f;
        red[3 + c] = (uniform && active) ? cf[c] : 0.f;
    }
    const float tot = warp_sum8(red, lane);
    const float tot_a2 = warp_sum(amb2);
    const float tot_g = warp_sum(gam);
    if (owner) {
        const int slot = vidx < 3 ? sm.albedo + m0 * 3 + vidx : (vidx < 6 ? sm.coeffs + m0 * 3 + (vidx - 3) : sm.ambient + (vidx - 6));
        acc.add(slot, tot);
    }
    if (lane == 1) acc.add(sm.ambient + 2, tot_a2);
    if (lane == 2) acc.add(sm.gamma, tot_g);
    if (!uniform && active) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            acc.add_shared(sm.albedo + m * 3 + c, alb[c]);
            acc.add_shared(sm.coeffs + m * 3 + c, cf[c]);
        }
    }
    *gP_io = gP;
    *gn_io = gn;
}

__device__ __forceinline__ void backward_body(const BackwardParams& p, float* warp_acc, LightS* lights);

constexpr int kBwdThreads = 128;
__global__ void __launch_bounds__(kBwdThreads, 6) k_backward(const __grid_constant__ BackwardParams p) {
    __shared__ float warp_acc[(kBwdThreads / 32) * kMaxAccSlots];
    __shared__ LightS lights[kLightTable];
    backward_body(p, warp_acc, lights);
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
__global__ void __launch_bounds__(kBwdThreads, 6) k_backward_batch(const __grid_constant__ BackwardParams p0,
                                                                 const __grid_constant__ BatchArgs ba) {
    __shared__ float warp_acc[(kBwdThreads / 32) * kMaxAccSlots];
    __shared__ LightS lights[kLightTable];
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
    backward_body(p, warp_acc, lights);
}

__device__ __forceinline__ void backward_body(const BackwardParams& p, float* warp_acc, LightS* lights) {
    const int n_slots = min(p.sm.total, kMaxAccSlots);
    for (int j = threadIdx.x; j < (kBwdThreads / 32) * kMaxAccSlots; j += blockDim.x) warp_acc[j] = 0.f;
    stage_lights(p.sc, lights);        // ends with __syncthreads
    const int lane = threadIdx.x & 31;
    const BwdAcc acc{p.gp, p.sm, p.sc.light_pos_stride, warp_acc + (threadIdx.x >> 5) * kMaxAccSlots};
    const CamState& cs = *p.cam;
    const Vec3 eye = v3(cs.eye[0], cs.eye[1], cs.eye[2]);
    bool any_sphere = false;
    for (int s = 0; s < p.sc.n_sets; ++s) any_sphere |= p.sc.sets[s].kind == KIND_SPHERE;
    for (int base = blockIdx.x * blockDim.x; base < p.n; base += gridDim.x * blockDim.x) {   // uniform trip count per CTA
        const int k = base + threadIdx.x;
        const bool live = k < p.n;
        const int kk = live ? k : p.n - 1;      // dead lanes shadow the last pixel with zero incoming gradients
        const float dep = p.depth[kk];
        const bool hit = live && dep <= cs.far_clip && dep >= cs.near_clip;
        float g_image[3], g_pos[3], g_normal[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            g_image[c] = (live && p.g_image) ? p.g_image[(size_t)kk * 3 + c] : 0.f;
            g_pos[c] = (live && p.g_pos) ? p.g_pos[(size_t)kk * 3 + c] : 0.f;
            g_normal[c] = (live && p.g_normal) ? p.g_normal[(size_t)kk * 3 + c] : 0.f;
        }
        const float g_depth = (hit && p.g_depth) ? p.g_depth[kk] : 0.f;
        // a warp whose pixels are all misses with no incoming geometry gradient has nothing to contribute
        const bool needed = hit || g_pos[0] != 0.f || g_pos[1] != 0.f || g_pos[2] != 0.f ||
                            g_normal[0] != 0.f || g_normal[1] != 0.f || g_normal[2] != 0.f;
        if (!__any_sync(0xffffffffu, needed)) continue;
        Vec3 o, d;
        pixel_ray(cs, p.rays, p.n, p.pix0, kk, &o, &d);
        const int idx = (int)p.nearest[kk];
        const int set = find_set(p.sc, idx);
        const SetView& sv = p.sc.sets[set];
        const int local = idx - sv.first;
        const int m = clampi(sv.mat[local], 0, p.sc.n_materials - 1);
        // geometry of the winner (fast forms; t of a hit is the saved depth)
        Vec3 P, n, c0 = v3(0.f, 0.f, 0.f), pmo = v3(0.f, 0.f, 0.f);
        float t, inv_nlen = 0.f, b = 1.f, radius = 0.f;
        if (sv.kind == KIND_SPHERE) {
            c0 = ld3(sv.pos + (size_t)local * sv.pos_stride);
            radius = sv.radius[local];
            t = hit ? dep : kMissSentinel;
            P = f_axpy(t, d, o);
            const Vec3 G = v3(P.x - c0.x, P.y - c0.y, P.z - c0.z);
            inv_nlen = f_rsqrt(fmaf(G.z, G.z, kEps) + (fmaf(G.y, G.y, kEps) + fmaf(G.x, G.x, kEps)));
            n = v3(G.x * inv_nlen, G.y * inv_nlen, G.z * inv_nlen);
        } else {
            const size_t prow = (sv.kind == KIND_TRIANGLE) ? (size_t)local * 3 * sv.pos_stride : (size_t)local * sv.pos_stride;
            const Vec3 p0 = ld3(sv.pos + prow);
            const Vec3 nr = ld3(sv.normal + (size_t)local * sv.normal_stride);
            inv_nlen = f_rsqrt(fmaf(nr.z, nr.z, kEps) + (fmaf(nr.y, nr.y, kEps) + fmaf(nr.x, nr.x, kEps)));
            n = v3(nr.x * inv_nlen, nr.y * inv_nlen, nr.z * inv_nlen);
            pmo = v3(p0.x - o.x, p0.y - o.y, p0.z - o.z);
            b = f_dot(n, d);
            t = hit ? dep : f_dot(pmo, n) * f_rcp(b);
            P = f_axpy(t, d, o);
        }
        Vec3 gP = v3(g_pos[0], g_pos[1], g_pos[2]);
        Vec3 gn = v3(g_normal[0], g_normal[1], g_normal[2]);
        backward_shading_fast(p.sc, lights, eye, P, n, m, hit, p.fl, p.vis ? p.vis + kk : nullptr, (size_t)p.n, g_image, acc,
                              p.sm, lane, &gP, &gn);
        float out[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (needed) {
            if (sv.kind == KIND_SPHERE) {
                const Vec3 G = v3(P.x - c0.x, P.y - c0.y, P.z - c0.z);
                const float gdot = f_dot(gn, n);
                const Vec3 gG = v3((gn.x - gdot * n.x) * inv_nlen, (gn.y - gdot * n.y) * inv_nlen, (gn.z - gdot * n.z) * inv_nlen);
                gP = v3(gP.x + gG.x, gP.y + gG.y, gP.z + gG.z);
                Vec3 gc = vneg(gG);
                float gr = 0.f;
                if (hit) {   // t depends on (c, r) only on a real hit; a masked miss has the constant t = 1001
                    const float gt = g_depth + f_dot(gP, d);
                    const float Gd = f_dot(G, d);
                    if (Gd != 0.f && gt != 0.f) {
                        const float q = gt * f_rcp(Gd);
                        gc = f_axpy(q, G, gc);
                        gr = q * radius;
                    }
                }
                out[0] = gc.x; out[1] = gc.y; out[2] = gc.z; out[6] = gr;
            } else {
                const float gt = g_depth + f_dot(gP, d);
                const float rb = f_rcp(b);
                const float ga = gt != 0.f ? gt * rb : 0.f;
                const float gb = gt != 0.f ? -gt * t * rb : 0.f;
                gn = f_axpy(ga, pmo, gn);
                gn = f_axpy(gb, d, gn);
                const float gdot = f_dot(gn, n);
                out[0] = ga * n.x; out[1] = ga * n.y; out[2] = ga * n.z;
                out[3] = (gn.x - gdot * n.x) * inv_nlen;
                out[4] = (gn.y - gdot * n.y) * inv_nlen;
                out[5] = (gn.z - gdot * n.z) * inv_nlen;
            }
        }
        // per-primitive gradients: runs of adjacent lanes with the same winner -> one set of double atomics
        const int key = needed ? idx : -1 - lane;
        bool head;
        if (any_sphere) head = segmented_sum<7>(key, out, lane);
        else head = segmented_sum<6>(key, out, lane);
        if (head && needed) {
            // double accumulation: per-pixel contributions of a grazing primitive cancel heavily, and a
            // sequential fp32 atomic sum would carry ~1e-4 relative noise (the reference sums pairwise)
            double* dst = p.prim_acc + (size_t)idx * 7;
#pragma unroll
            for (int c = 0; c < 7; ++c)
                if (out[c] != 0.f) atomicAdd(dst + c, (double)out[c]);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n_slots; j += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kBwdThreads / 32; ++w) s += warp_acc[w * kMaxAccSlots + j];
        if (s != 0.f) atomicAdd(p.acc + j, (double)s);
    }
}

struct FinalizeParams {
    GradPtrs gp; SlotMap sm; const double* acc; int K, L, Cn, light_pos_stride;
    SceneView sc; const double* prim_acc;
    const double* loss_acc; float* loss_out;      // fused step: *loss_out += *loss_acc (null otherwise)
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
    if (j == 0 && p.loss_out) atomicAdd(p.loss_out, (float)p.loss_acc[0]);
    if (j >= p.sm.total || j >= kMaxAccSlots) return;      // slots past the buffer went straight to the leaves
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
        p.loss_acc = ws_at(p0.loss_acc, ba, b);
    }
    __syncthreads();
    finalize_body(p, blockIdx.x * blockDim.x + threadIdx.x);
}
