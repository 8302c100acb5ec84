// surf_fast.cuh - part of libsurf_b200.so (included by surf_kernels.cu inside namespace surf).  Device-only.
//
// The per-pixel O(N*L) stages - shading (renderer.py:82-125, :330-340) and its backward - are HBM-bound by bytes but
// were issue-bound in practice: evaluated in the reference's operation order with individually rounded IEEE
// div / sqrt / powf they cost ~1500 (forward) and ~4800 (backward) instructions per pixel.  Parity for image and
// gradients is a tolerance (1e-4 relative / 1e-5 absolute, BASELINE.json north_star), not bit-exactness, so these
// stages use the SFU approximations (MUFU.RCP / RSQ / LG2 / EX2, <= 2 ulp each) and explicit FMAs here.  What stays
// bit-exact: the winner index and its depth (taken from the z-buffer key the exact narrow phase wrote), the hit point
// o + t d and the unit normal of planar primitives (read back from the records k_prep computed in reference order).
// surf_math.cuh keeps the reference-order forms: the CPU emulation (tests/emul) and the miss-pixel / sphere corner
// cases run those.
#pragma once

__device__ __forceinline__ float f_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float f_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float f_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float f_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// base >= 0.  pow(0, e>0) = 0, pow(x, 0) = 1, pow(0, e<0) = inf like powf
__device__ __forceinline__ float f_pow(float base, float e) { return e == 0.f ? 1.f : f_ex2(e * f_lg2(base)); }
constexpr float kLn2 = 0.69314718055994530942f;
__device__ __forceinline__ float f_dot(Vec3 a, Vec3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ Vec3 f_axpy(float s, Vec3 a, Vec3 b) { return v3(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)); }
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// ---------------------------------------------------------------------------------------------------
// light table: position, attenuation and colour row of every light, staged once per CTA in shared memory so the
// per-pixel light loop reads three warp-broadcast LDS.128 instead of ten dependent global loads per light
// ---------------------------------------------------------------------------------------------------
struct __align__(16) LightS {
    float px, py, pz, a0;
    float a1, a2, cr, cg;
    float cb, pad0, pad1, pad2;
};
constexpr int kLightTable = 32;        // lights held in shared memory; further lights are read from global memory

__device__ __forceinline__ LightS load_light(const SceneView& sc, int l) {
    LightS s;
    const float* lp = sc.light_pos + (size_t)l * sc.light_pos_stride;
    const float* at = sc.light_atten + 3 * (size_t)l;
    const float* col = sc.colors + 3 * (size_t)clampi(sc.light_color_idx[l], 0, sc.n_colors - 1);
    s.px = lp[0]; s.py = lp[1]; s.pz = lp[2];
    s.a0 = at[0]; s.a1 = at[1]; s.a2 = at[2];
    s.cr = col[0]; s.cg = col[1]; s.cb = col[2];
    s.pad0 = s.pad1 = s.pad2 = 0.f;
    return s;
}
// every thread of the CTA calls this; ends with a __syncthreads
__device__ __forceinline__ void stage_lights(const SceneView& sc, LightS* table) {
    for (int l = threadIdx.x; l < sc.n_lights && l < kLightTable; l += blockDim.x) table[l] = load_light(sc, l);
    __syncthreads();
}
__device__ __forceinline__ LightS light_at(const SceneView& sc, const LightS* table, int l) {
    if (l < kLightTable) return table[l];
    return load_light(sc, l);
}

struct MatF { float A[3]; float kd, ks, sh; };
__device__ __forceinline__ MatF load_material(const SceneView& sc, int m) {
    MatF mt;
    const float* a = sc.albedo + 3 * (size_t)m;
    const float* c = sc.coeffs + 3 * (size_t)m;
    mt.A[0] = a[0]; mt.A[1] = a[1]; mt.A[2] = a[2];
    mt.kd = c[0]; mt.ks = c[1]; mt.sh = c[2];
    return mt;
}

// everything the fragment shader derives for one light at one fragment (renderer.py:89-115)
struct LightF {
    Vec3 L;                      // unit light direction
    float dl, d2, pw, att;       // distance, its square, distance^2 or ^4, attenuation 1 / den
    float nL, VL;                // n.L, V.L
    float D, S;                  // diffuse n.(att L) and specular V.R before sign flip / relu
    bool dl_nz, den_nz;
};
__device__ __forceinline__ LightF eval_light_fast(const LightS& ls, Vec3 P, Vec3 n, Vec3 V, float Vn, int use_quartic) {
    LightF e;
    const Vec3 Lv = v3(ls.px - P.x, ls.py - P.y, ls.pz - P.z);
    e.d2 = f_dot(Lv, Lv);
    e.dl_nz = e.d2 > 0.f;
    const float rinv = e.dl_nz ? f_rsqrt(e.d2) : 1.f;           // nonzero_divide: a zero length divides by one
    e.dl = e.dl_nz ? e.d2 * rinv : 0.f;
    e.L = v3(Lv.x * rinv, Lv.y * rinv, Lv.z * rinv);
    e.pw = use_quartic ? e.d2 * e.d2 : e.d2;
    const float den = fmaf(e.pw, ls.a2, fmaf(e.dl, ls.a1, ls.a0));
    e.den_nz = den != 0.f;
    e.att = f_rcp(e.den_nz ? den : 1.f);
    e.nL = f_dot(n, e.L);
    e.VL = f_dot(V, e.L);
    e.D = e.att * e.nL;
    e.S = fmaf(2.f * e.nL, Vn, -e.VL);                          // V.R with R = 2 (n.L) n - L
    return e;
}

// view vector with the reference's eps-regularised normalisation (utils.py:135-139); *inv_len = 1 / length
__device__ __forceinline__ Vec3 view_vector(Vec3 eye, Vec3 P, float* inv_len) {
    const Vec3 Vv = v3(eye.x - P.x, eye.y - P.y, eye.z - P.z);
    const float s = fmaf(Vv.z, Vv.z, kEps) + (fmaf(Vv.y, Vv.y, kEps) + fmaf(Vv.x, Vv.x, kEps));
    const float inv = f_rsqrt(s);
    *inv_len = inv;
    return v3(Vv.x * inv, Vv.y * inv, Vv.z * inv);
}
__device__ __forceinline__ float sign_or_zero(float dp) { return dp > 0.f ? 1.f : (dp < 0.f ? -1.f : 0.f); }

// sum over lights of the per-light colour (ambient included per light) before mask / relu / tonemap.
// `vis` points at this pixel's entry of light 0's visibility row, `vis_stride` floats to the next light's (null: none)
__device__ __forceinline__ void shade_fast(const SceneView& sc, const LightS* table, Vec3 eye, Vec3 P, Vec3 n, const MatF& mt,
                                           ShadeFlags fl, const float* __restrict__ vis, size_t vis_stride, float rgb[3]) {
    float inv_len;
    const Vec3 V = fl.raw_view ? v3(eye.x - P.x, eye.y - P.y, eye.z - P.z) : view_vector(eye, P, &inv_len);
    const float Vn = f_dot(V, n);
    const float sg = fl.double_sided ? sign_or_zero(Vn) : 1.f;
    float acc[3] = {0.f, 0.f, 0.f};
    for (int l = 0; l < sc.n_lights; ++l) {
        const LightS ls = light_at(sc, table, l);
        const LightF e = eval_light_fast(ls, P, n, V, Vn, fl.use_quartic);
        const float D = fmaxf(sg * e.D, 0.f), S = fmaxf(sg * e.S, 0.f);
        float scal = fmaf(mt.ks, f_pow(S, mt.sh), mt.kd * D);
        if (vis) scal *= vis[(size_t)l * vis_stride];
        acc[0] = fmaf(scal, ls.cr * mt.A[0], acc[0]);
        acc[1] = fmaf(scal, ls.cg * mt.A[1], acc[1]);
        acc[2] = fmaf(scal, ls.cb * mt.A[2], acc[2]);
    }
    const float nl = (float)sc.n_lights;
    rgb[0] = fmaf(nl * sc.ambient[0], mt.A[0], acc[0]);
    rgb[1] = fmaf(nl * sc.ambient[1], mt.A[1], acc[1]);
    rgb[2] = fmaf(nl * sc.ambient[2], mt.A[2], acc[2]);
}

// compositing (renderer.py:330-340): mask by hit, relu, gamma
__device__ __forceinline__ void composite_fast(const float lit[3], bool hit, const float* gamma, float out[3]) {
    const float gm = gamma ? gamma[0] : 1.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = hit ? fmaxf(lit[c], 0.f) : 0.f;
        out[c] = gamma ? f_pow(v, gm) : v;
    }
}

// ---------------------------------------------------------------------------------------------------
// warp reductions for the backward kernels
// ---------------------------------------------------------------------------------------------------
// Sums 8 per-lane values over the warp with 9 shuffles (a plain butterfly needs 40): three halving exchanges leave
// every lane with ONE of the eight sums over 8 lanes, two more xor steps complete it.  On return value j's total sits
// in the four lanes with (lane >> 2) == j; the return value is this lane's total.
__device__ __forceinline__ float warp_sum8(const float v[8], int lane) {
    const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
    float w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = h4 ? v[i] : v[i + 4];
        const float keep = h4 ? v[i + 4] : v[i];
        w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    float x[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = h3 ? w[i] : w[i + 2];
        const float keep = h3 ? w[i + 2] : w[i];
        x[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    float y = (h2 ? x[1] : x[0]) + __shfl_xor_sync(0xffffffffu, h2 ? x[0] : x[1], 4);
    y += __shfl_xor_sync(0xffffffffu, y, 2);
    y += __shfl_xor_sync(0xffffffffu, y, 1);
    return y;                    // total of value index ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)
}
__device__ __forceinline__ int warp_sum8_index(int lane) { return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); }
