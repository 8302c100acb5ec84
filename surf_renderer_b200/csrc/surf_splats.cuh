// surf_splats.cuh - part of libsurf_b200.so (included by surf_kernels.cu inside namespace surf).
// render_splats_along_ray (renderer.py:537-751): one splat per pixel at camera-space depth z on the pixel's ray - the
// GAN generator's differentiable path (GAN/gan.py:563-597).  Everything between the caller's z / normals and the
// images is in these kernels:
//   k_splat_setup             camera basis + lights into camera coordinates              (renderer.py:553, :709)
//   k_splat_normals           3x3 reflect-padded stencil normal estimation: constrained plane fit (utils.py:886-923)
//                             or average of the eight neighbour cross products (utils.py:854-883)
//   k_splat_forward           z -> position (renderer.py:566-580), supersampling: the K x K sub-pixel rays of a pixel
//                             intersect its splat's plane (renderer.py:603-673), shading, depth min / max
//   k_splat_normdepth         norm_depth_image_only (renderer.py:677-686)
//   k_splat_backward          shading backward per fragment; fragment -> splat gradients (atomics when K > 1)
//   k_splat_normals_backward  analytic backward of the stencil, scattered to the nine positions it read
//   k_splat_src_finalize      position gradients -> d/dz, normal gradients -> the caller's normal array
//   k_splat_finalize          light / material accumulators -> leaves (lights back to world coordinates)
// All kernels take the scene of a batch from blockIdx.y (surf_splats_forward_strided: the per-element loop of
// gan.py:563-597 in one call).
#pragma once

struct SplatBatchArgs {       // element strides between consecutive scenes of a batch (0 = shared)
    long long z, normal, mat, vis, light_pos, eye;
    long long ws_stride;      // bytes between per-scene workspaces
    long long out_stride;     // output pixels per scene (n)
};

struct SplatWorkspace {
    CamState* cam;
    float* light_cc;          // [L, 3] lights in camera coordinates
    double* acc;              // [kMaxAccSlots] light / material accumulators
    float* nest;              // [n_src, 3] estimated normals
    float* gpos_src;          // [n_src, 3] d/d(position of the source splat)
    float* gn_src;            // [n_src, 3] d/d(normal of the source splat)
    int* minmax;              // [2] min / max depth as ordered ints (norm_depth_image_only)
    size_t bytes;
};
inline void carve_splats(void* base, int n_src, int n_lights, SplatWorkspace* ws) {
    char* p = (char*)base;
    size_t off = 0;
    ws->cam = (CamState*)(p + off); off += align_up(sizeof(CamState), 256);
    ws->light_cc = (float*)(p + off); off += align_up((size_t)(n_lights > 0 ? n_lights : 1) * 12, 256);
    ws->acc = (double*)(p + off); off += align_up((size_t)kMaxAccSlots * 8, 256);
    ws->minmax = (int*)(p + off); off += 256;
    ws->nest = (float*)(p + off); off += align_up((size_t)n_src * 12, 256);
    ws->gpos_src = (float*)(p + off); off += align_up((size_t)n_src * 12, 256);
    ws->gn_src = (float*)(p + off); off += align_up((size_t)n_src * 12, 256);
    ws->bytes = off;
}

struct SplatParams {
    SceneView sc;                 // lights (camera space, stride 3, in the workspace) / colours / materials
    const CamState* cam;
    const float* z; int z_stride;
    const float* normal; int normal_stride;   // the caller's normals, or the estimated ones (workspace, stride 3)
    const int* mat;
    const float* vis;             // [L, n_src] or null
    const float* pos_in;          // [n, 3] explicit fragment positions (K = 1 only) or null
    int ndc_stride;               // 3 | 4: pos_in holds NDC coordinates [n, ndc_stride] (render_splats_NDC); 0: camera space
    float ndc_a0, ndc_a1, ndc_b, ndc_c;    // inverse perspective (ops.py:61-68): X = a0 x / w', Y = a1 y / w', Z = -w / w', w' = b z + c w
    int W, H, K;                  // source grid and samples per pixel edge; output grid is (H K) x (W K)
    int n_src, n;                 // W H source splats, n output fragments
    int estimate;                 // 0: `normal` is the caller's; 1 / 2: estimated (plane fit / average normal)
    ShadeFlags fl;
    float* image; float* depth; float* normal_out; float* pos;                       // forward outputs [n, ...]
    const float* g_image; const float* g_depth; const float* g_normal; const float* g_pos;   // backward inputs
    float* gz; float* gnormal; float* gpos;    // backward outputs (caller's strides); gpos: explicit positions
    float* nest; float* gpos_src; float* gn_src; int* minmax;
    float* norm_depth; float far_clip;
    SlotMap sm; double* acc;
    GradPtrs gp;
    SplatBatchArgs ba;
};

// per-scene view of the parameter block (blockIdx.y = scene)
__device__ __forceinline__ void splat_scene(SplatParams& p, int b) {
    const SplatBatchArgs& ba = p.ba;
    auto ws = [&](auto* q) { return q ? (decltype(q))((char*)q + ba.ws_stride * b) : q; };
    p.cam = (const CamState*)((const char*)p.cam + ba.ws_stride * b);
    p.sc.light_pos = (const float*)((const char*)p.sc.light_pos + ba.ws_stride * b);
    p.acc = ws(p.acc); p.nest = ws(p.nest); p.gpos_src = ws(p.gpos_src); p.gn_src = ws(p.gn_src); p.minmax = ws(p.minmax);
    p.z = adv(p.z, b * ba.z);
    if (p.estimate) p.normal = p.nest;          // (already advanced)
    else p.normal = adv(p.normal, b * ba.normal);
    p.mat = adv(p.mat, b * ba.mat);
    p.vis = adv(p.vis, b * ba.vis);
    p.pos_in = adv(p.pos_in, (long long)b * ba.out_stride * (p.ndc_stride ? p.ndc_stride : 3));
    p.image = adv(p.image, (long long)b * ba.out_stride * 3); p.depth = adv(p.depth, (long long)b * ba.out_stride);
    p.normal_out = adv(p.normal_out, (long long)b * ba.out_stride * 3); p.pos = adv(p.pos, (long long)b * ba.out_stride * 3);
    p.norm_depth = adv(p.norm_depth, (long long)b * ba.out_stride);
    p.g_image = adv(p.g_image, (long long)b * ba.out_stride * 3); p.g_depth = adv(p.g_depth, (long long)b * ba.out_stride);
    p.g_normal = adv(p.g_normal, (long long)b * ba.out_stride * 3); p.g_pos = adv(p.g_pos, (long long)b * ba.out_stride * 3);
    p.gz = adv(p.gz, b * ba.z); p.gnormal = adv(p.gnormal, b * ba.normal);
    p.gpos = adv(p.gpos, (long long)b * ba.out_stride * (p.ndc_stride ? p.ndc_stride : 3));
    p.gp.light_pos = adv(p.gp.light_pos, b * ba.light_pos);
}

// scene b: camera basis from its eye, its lights into camera coordinates; zero the per-scene reduction cells
__global__ void k_splat_setup(CamArgs a, long long eye_stride, CamState* cs0, const float* light_pos4, long long light_stride,
                              int n_lights, float* light_cc0, int* minmax0, long long ws_stride) {
    const int b = blockIdx.x;
    CamState* cs = (CamState*)((char*)cs0 + ws_stride * b);
    float* light_cc = (float*)((char*)light_cc0 + ws_stride * b);
    int* minmax = (int*)((char*)minmax0 + ws_stride * b);
    if (threadIdx.x == 0) {
        camera_setup(a.eye + b * eye_stride, a.at, a.up, 0, a.W, a.H, a.fovy, a.focal, a.near_clip, a.far_clip, cs);
        minmax[0] = 0x7f800000;       // +inf
        minmax[1] = 0;
    }
    __syncthreads();
    const float* lp = light_pos4 + b * light_stride;
    for (int l = threadIdx.x; l < n_lights; l += blockDim.x) {
        Vec3 v = light_to_camera(*cs, lp + 4 * (size_t)l);
        light_cc[3 * l] = v.x; light_cc[3 * l + 1] = v.y; light_cc[3 * l + 2] = v.z;
    }
}

// ---------------------------------------------------------------------------------------------------
// geometry helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect_index(int i, int n) {        // pad mode 'reflect' of one pixel (utils.py:745-769)
    if (n == 1) return 0;
    if (i < 0) return -i;
    if (i >= n) return 2 * n - 2 - i;
    return i;
}
// camera-space position of source splat (row, col) from its depth (renderer.py:566-580)
__device__ __forceinline__ Vec3 splat_position(const SplatParams& p, const CamState& cs, int row, int col, float* x_out = nullptr,
                                               float* y_out = nullptr) {
    const int s = row * p.W + col;
    const float zin = p.z[(size_t)s * p.z_stride];
    const float Z = zin < 0.f ? zin : -0.f;
    float x, y;
    pixel_xy(cs, s, &x, &y);
    const float inv_f = f_rcp(-cs.neg_focal);
    if (x_out) { *x_out = x; *y_out = y; }
    return v3(-Z * x * inv_f, -Z * y * inv_f, Z);
}
__device__ __forceinline__ Vec3 unit_fast(Vec3 u, float* inv_len) {          // utils.py:135-139, eps inside the sum
    const float s = fmaf(u.z, u.z, kEps) + (fmaf(u.y, u.y, kEps) + fmaf(u.x, u.x, kEps));
    const float inv = f_rsqrt(s);
    if (inv_len) *inv_len = inv;
    return v3(u.x * inv, u.y * inv, u.z * inv);
}
__device__ __forceinline__ Vec3 f_cross(Vec3 a, Vec3 b) {
    return v3(fmaf(a.y, b.z, -a.z * b.y), fmaf(a.z, b.x, -a.x * b.z), fmaf(a.x, b.y, -a.y * b.x));
}
__device__ __forceinline__ Vec3 unit_backward(Vec3 g_u, Vec3 u, float inv_len) {     // u = d * inv_len  ->  d/dd
    const float gd = f_dot(g_u, u);
    return v3((g_u.x - gd * u.x) * inv_len, (g_u.y - gd * u.y) * inv_len, (g_u.z - gd * u.z) * inv_len);
}

// neighbour order of grad_spatial2d (utils.py:772-792): dy outer, dx inner, centre skipped
__device__ __constant__ int kNbDy[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
__device__ __constant__ int kNbDx[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
// ring of find_average_normal (utils.py:869-876)
__device__ __constant__ int kRingA[8] = {4, 2, 1, 0, 3, 5, 6, 7};
__device__ __constant__ int kRingB[8] = {2, 1, 0, 3, 5, 6, 7, 4};

// Normal of splat (row, col) from its 3x3 neighbourhood; with g_n != null also scatters d/d(position) of the nine
// splats it read into gpos_src (atomics: neighbouring stencils overlap).
__device__ __forceinline__ Vec3 stencil_normal(const SplatParams& p, const CamState& cs, int row, int col, const Vec3* g_n) {
    const Vec3 p0 = splat_position(p, cs, row, col);
    Vec3 u[8];
    float inv[8];
    int src[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int r = reflect_index(row + kNbDy[k], p.H), c = reflect_index(col + kNbDx[k], p.W);
        src[k] = r * p.W + c;
        const Vec3 q = splat_position(p, cs, r, c);
        u[k] = unit_fast(v3(q.x - p0.x, q.y - p0.y, q.z - p0.z), &inv[k]);
    }
    Vec3 n;
    Vec3 g_u[8];
    if (p.estimate == 1) {
        // constrained plane fit: [nx, ny] = (M^T M)^-1 M^T (-uz), M rows = (ux, uy); n = unit(nx, ny, 1)
        float a = 0.f, b = 0.f, d = 0.f, r0 = 0.f, r1 = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a = fmaf(u[k].x, u[k].x, a); b = fmaf(u[k].x, u[k].y, b); d = fmaf(u[k].y, u[k].y, d);
            r0 = fmaf(-u[k].x, u[k].z, r0); r1 = fmaf(-u[k].y, u[k].z, r1);
        }
        const float det = fmaf(a, d, -b * b) + 1e-12f;
        const float rdet = 1.f / det;
        const float nx = (d * r0 - b * r1) * rdet, ny = (a * r1 - b * r0) * rdet;
        float inv_v;
        n = unit_fast(v3(nx, ny, 1.f), &inv_v);
        if (g_n) {
            const Vec3 gv = unit_backward(*g_n, n, inv_v);
            const float gnx = gv.x, gny = gv.y;
            const float g_r0 = (gnx * d - gny * b) * rdet, g_r1 = (gny * a - gnx * b) * rdet;
            const float gdet = -(gnx * nx + gny * ny) * rdet;
            const float g_a = gny * r1 * rdet + gdet * d;
            const float g_d = gnx * r0 * rdet + gdet * a;
            const float g_b = -(gnx * r1 + gny * r0) * rdet - 2.f * b * gdet;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                g_u[k] = v3(2.f * u[k].x * g_a + u[k].y * g_b - u[k].z * g_r0, u[k].x * g_b + 2.f * u[k].y * g_d - u[k].z * g_r1,
                            -u[k].x * g_r0 - u[k].y * g_r1);
        }
    } else {
        // mean of the eight cross products of neighbouring difference vectors, normalised, clamped to [0, 1]
        Vec3 m = v3(0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const Vec3 c = f_cross(u[kRingA[k]], u[kRingB[k]]);
            m = v3(m.x + c.x, m.y + c.y, m.z + c.z);
        }
        m = v3(m.x * 0.125f, m.y * 0.125f, m.z * 0.125f);
        float inv_m;
        const Vec3 nu = unit_fast(m, &inv_m);
        n = v3(fminf(fmaxf(nu.x, 0.f), 1.f), fminf(fmaxf(nu.y, 0.f), 1.f), fminf(fmaxf(nu.z, 0.f), 1.f));
        if (g_n) {
            // clamp passes the gradient inside [0, 1] (torch.clamp: inclusive bounds)
            const Vec3 gc = v3((nu.x >= 0.f && nu.x <= 1.f) ? g_n->x : 0.f, (nu.y >= 0.f && nu.y <= 1.f) ? g_n->y : 0.f,
                               (nu.z >= 0.f && nu.z <= 1.f) ? g_n->z : 0.f);
            Vec3 gm = unit_backward(gc, nu, inv_m);
            gm = v3(gm.x * 0.125f, gm.y * 0.125f, gm.z * 0.125f);
#pragma unroll
            for (int k = 0; k < 8; ++k) g_u[k] = v3(0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 8; ++k) {          // c = a x b:  g_a = b x g_c,  g_b = g_c x a
                const int ia = kRingA[k], ib = kRingB[k];
                const Vec3 ga = f_cross(u[ib], gm), gb = f_cross(gm, u[ia]);
                g_u[ia] = v3(g_u[ia].x + ga.x, g_u[ia].y + ga.y, g_u[ia].z + ga.z);
                g_u[ib] = v3(g_u[ib].x + gb.x, g_u[ib].y + gb.y, g_u[ib].z + gb.z);
            }
        }
    }
    if (g_n) {
        Vec3 g0 = v3(0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const Vec3 gd = unit_backward(g_u[k], u[k], inv[k]);
            if (gd.x != 0.f) atomicAdd(p.gpos_src + 3 * (size_t)src[k], gd.x);
            if (gd.y != 0.f) atomicAdd(p.gpos_src + 3 * (size_t)src[k] + 1, gd.y);
            if (gd.z != 0.f) atomicAdd(p.gpos_src + 3 * (size_t)src[k] + 2, gd.z);
            g0 = v3(g0.x - gd.x, g0.y - gd.y, g0.z - gd.z);
        }
        const size_t s0 = (size_t)row * p.W + col;
        atomicAdd(p.gpos_src + 3 * s0, g0.x); atomicAdd(p.gpos_src + 3 * s0 + 1, g0.y); atomicAdd(p.gpos_src + 3 * s0 + 2, g0.z);
    }
    return n;
}

__global__ void __launch_bounds__(256) k_splat_normals(const __grid_constant__ SplatParams p0) {
    SplatParams p = p0;
    splat_scene(p, blockIdx.y);
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= p.n_src) return;
    const Vec3 n = stencil_normal(p, *p.cam, s / p.W, s % p.W, nullptr);
    p.nest[3 * (size_t)s] = n.x; p.nest[3 * (size_t)s + 1] = n.y; p.nest[3 * (size_t)s + 2] = n.z;
}

__global__ void __launch_bounds__(256) k_splat_normals_backward(const __grid_constant__ SplatParams p0) {
    SplatParams p = p0;
    splat_scene(p, blockIdx.y);
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= p.n_src) return;
    const Vec3 g = v3(p.gn_src[3 * (size_t)s], p.gn_src[3 * (size_t)s + 1], p.gn_src[3 * (size_t)s + 2]);
    if (g.x == 0.f && g.y == 0.f && g.z == 0.f) return;
    stencil_normal(p, *p.cam, s / p.W, s % p.W, &g);
}

// fragment k of the output grid: its source splat and, for K > 1, the unit ray of its sub-pixel (renderer.py:646-657;
// reshape_upsampled_data :476-481 puts the x-shift index on the ROW and the y-shift index on the COLUMN sub-position)
struct Fragment2 {
    int src;              // source splat
    Vec3 P, n;            // position, normal
    Vec3 P0, ray;         // source splat position; unit sub-pixel ray (K > 1)
    float x, y;           // image-plane coordinates of the source pixel
    float inv_rn, d, inv_ray_len;   // 1 / (ray . n), plane offset P0 . n, 1 / |ray before normalisation|
    float depth;
};
__device__ __forceinline__ Fragment2 splat_fragment2(const SplatParams& p, const CamState& cs, int k) {
    Fragment2 f;
    const int K = p.K;
    int row, col, a = 0, b = 0;
    if (K == 1) { row = k / p.W; col = k - row * p.W; }
    else {
        const int WK = p.W * K;
        const int R = k / WK, Cc = k - R * WK;
        row = R / K; a = R - row * K;
        col = Cc / K; b = Cc - col * K;
    }
    f.src = row * p.W + col;
    const float* np_ = p.normal + (size_t)f.src * p.normal_stride;
    f.n = v3(np_[0], np_[1], np_[2]);
    if (p.pos_in && p.ndc_stride) {
        // renderer.py:384-387: pos_CC = Minv [x y z w]^T, divided by its w
        const float* q = p.pos_in + (size_t)p.ndc_stride * k;
        const float w = p.ndc_stride == 4 ? q[3] : 1.f;
        const float iw = 1.f / fmaf(p.ndc_b, q[2], p.ndc_c * w);
        f.P = v3(p.ndc_a0 * q[0] * iw, p.ndc_a1 * q[1] * iw, -w * iw);
        f.P0 = f.P; f.x = f.y = 0.f;
        f.inv_ray_len = iw;           // (reused as 1 / w' by the backward)
    } else if (p.pos_in) {
        f.P = ld3(p.pos_in + 3 * (size_t)k);
        f.P0 = f.P; f.x = f.y = 0.f;
    } else {
        f.P0 = splat_position(p, cs, row, col, &f.x, &f.y);
        f.P = f.P0;
    }
    f.inv_rn = 0.f; f.d = 0.f; f.ray = v3(0.f, 0.f, 0.f);
    if (!(p.pos_in && p.ndc_stride)) f.inv_ray_len = 0.f;
    if (K > 1) {
        // sub-pixel ray through (x + deltax dx / 2, y + deltay dy / 2, -focal); deltax = linspace(-1, 1, K)[a] etc.
        const float h_img = 2.f * cs.sy, w_img = 2.f * cs.sx;
        const float sub_w = w_img / (float)(K * p.W - 1), sub_h = h_img / (float)(K * p.H - 1);
        const float deltax = -1.f + 2.f * (float)a / (float)(K - 1), deltay = 1.f - 2.f * (float)b / (float)(K - 1);
        const Vec3 r = v3(fmaf(deltax, 0.5f * sub_w, f.x), fmaf(deltay, 0.5f * sub_h, f.y), cs.neg_focal);
        f.ray = unit_fast(r, &f.inv_ray_len);
        f.d = f_dot(f.P0, f.n);
        f.inv_rn = 1.f / f_dot(f.ray, f.n);
        const float t = f.d * f.inv_rn;
        f.P = v3(t * f.ray.x, t * f.ray.y, t * f.ray.z);
    }
    f.depth = sqrtf(f_dot(f.P, f.P));
    return f;
}

__device__ __forceinline__ int float_as_ordered_int(float v) { return __float_as_int(v); }      // depths are >= 0

__global__ void __launch_bounds__(256) k_splat_forward(const __grid_constant__ SplatParams p0) {
    __shared__ LightS lights[kLightTable];
    __shared__ SplatParams p;
    if (threadIdx.x == 0) { p = p0; splat_scene(p, blockIdx.y); }
    __syncthreads();
    stage_lights(p.sc, lights);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    float dmin = INFINITY, dmax = 0.f;
    if (k < p.n) {
        const Fragment2 f = splat_fragment2(p, *p.cam, k);
        const int m = p.mat ? clampi(p.mat[f.src], 0, p.sc.n_materials - 1) : 0;
        const MatF mt = load_material(p.sc, m);
        float lit[3];
        shade_fast(p.sc, lights, v3(0.f, 0.f, 0.f), f.P, f.n, mt, p.fl, p.vis ? p.vis + f.src : nullptr, (size_t)p.n_src, lit);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (p.image) p.image[3 * (size_t)k + c] = fmaxf(lit[c], 0.f);          // relu, no tonemap (:741)
            if (p.pos) p.pos[3 * (size_t)k + c] = c == 0 ? f.P.x : (c == 1 ? f.P.y : f.P.z);
            if (p.normal_out) p.normal_out[3 * (size_t)k + c] = c == 0 ? f.n.x : (c == 1 ? f.n.y : f.n.z);
        }
        if (p.depth) p.depth[k] = f.depth;
        dmin = dmax = f.depth;
    }
    if (p.norm_depth) {       // CTA-uniform
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
            dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(p.minmax, float_as_ordered_int(dmin));
            atomicMax(p.minmax + 1, float_as_ordered_int(dmax));
        }
    }
}

// renderer.py:677-686: where(depth >= far, min, depth), then (. - min) / (max - min)
__global__ void __launch_bounds__(256) k_splat_normdepth(const __grid_constant__ SplatParams p0) {
    SplatParams p = p0;
    splat_scene(p, blockIdx.y);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= p.n) return;
    const float lo = __int_as_float(p.minmax[0]), hi = __int_as_float(p.minmax[1]);
    const float d = p.depth[k];
    const float v = d >= p.far_clip ? lo : d;
    p.norm_depth[k] = (v - lo) / (hi - lo);
}

__global__ void __launch_bounds__(kBwdThreads) k_splat_backward(const __grid_constant__ SplatParams p0) {
    __shared__ float warp_acc[(kBwdThreads / 32) * kMaxAccSlots];
    __shared__ LightS lights[kLightTable];
    __shared__ SplatParams p;
    if (threadIdx.x == 0) { p = p0; splat_scene(p, blockIdx.y); }
    for (int j = threadIdx.x; j < (kBwdThreads / 32) * kMaxAccSlots; j += blockDim.x) warp_acc[j] = 0.f;
    __syncthreads();
    stage_lights(p.sc, lights);
    const int lane = threadIdx.x & 31;
    const BwdAcc acc{p.gp, p.sm, 3, warp_acc + (threadIdx.x >> 5) * kMaxAccSlots};
    const CamState& cs = *p.cam;
    const bool scatter = p.K > 1 || p.estimate != 0;     // fragment gradients go through the per-splat accumulators
    for (int base = blockIdx.x * blockDim.x; base < p.n; base += gridDim.x * blockDim.x) {     // uniform trip count per CTA
        const int k = base + threadIdx.x;
        const bool live = k < p.n;
        const int kk = live ? k : p.n - 1;
        float g_image[3];
        Vec3 gP, gn;
        g_image[0] = (live && p.g_image) ? p.g_image[3 * (size_t)kk] : 0.f;
        g_image[1] = (live && p.g_image) ? p.g_image[3 * (size_t)kk + 1] : 0.f;
        g_image[2] = (live && p.g_image) ? p.g_image[3 * (size_t)kk + 2] : 0.f;
        gP = (live && p.g_pos) ? ld3(p.g_pos + 3 * (size_t)kk) : v3(0.f, 0.f, 0.f);
        gn = (live && p.g_normal) ? ld3(p.g_normal + 3 * (size_t)kk) : v3(0.f, 0.f, 0.f);
        const float g_depth = (live && p.g_depth) ? p.g_depth[kk] : 0.f;
        const Fragment2 f = splat_fragment2(p, cs, kk);
        const int m = p.mat ? clampi(p.mat[f.src], 0, p.sc.n_materials - 1) : 0;
        backward_shading_fast(p.sc, lights, v3(0.f, 0.f, 0.f), f.P, f.n, m, live, p.fl, p.vis ? p.vis + f.src : nullptr,
                              (size_t)p.n_src, g_image, acc, p.sm, lane, &gP, &gn);
        if (!live) continue;
        if (f.depth > 0.f) gP = f_axpy(g_depth / f.depth, f.P, gP);            // depth = |P|
        Vec3 gP0 = gP;                 // d/d(source splat position)
        if (p.K > 1) {
            // P = t ray, t = d / (ray . n), d = P0 . n
            const float g_t = f_dot(gP, f.ray);
            const float g_d = g_t * f.inv_rn;
            const float g_rn = -g_t * f.d * f.inv_rn * f.inv_rn;
            gn = f_axpy(g_d, f.P0, gn);
            gn = f_axpy(g_rn, f.ray, gn);
            gP0 = v3(g_d * f.n.x, g_d * f.n.y, g_d * f.n.z);
        }
        if (p.pos_in && p.ndc_stride) {
            if (p.gpos) {       // P = (a0 x, a1 y, -w) / w',  w' = b z + c w
                const float* q = p.pos_in + (size_t)p.ndc_stride * k;
                const float iw = f.inv_ray_len;
                const float g_wp = -f_dot(gP0, f.P) * iw;
                float* g = p.gpos + (size_t)p.ndc_stride * k;
                g[0] += gP0.x * p.ndc_a0 * iw;
                g[1] += gP0.y * p.ndc_a1 * iw;
                g[2] += g_wp * p.ndc_b;
                if (p.ndc_stride == 4) g[3] += -gP0.z * iw + g_wp * p.ndc_c;
                (void)q;
            }
            if (p.gnormal) {
                float* dst = p.gnormal + (size_t)f.src * p.normal_stride;
                dst[0] += gn.x; dst[1] += gn.y; dst[2] += gn.z;
            }
        } else if (p.pos_in) {
            if (p.gpos) { p.gpos[3 * (size_t)k] += gP0.x; p.gpos[3 * (size_t)k + 1] += gP0.y; p.gpos[3 * (size_t)k + 2] += gP0.z; }
            if (p.gnormal) {
                float* dst = p.gnormal + (size_t)f.src * p.normal_stride;
                dst[0] += gn.x; dst[1] += gn.y; dst[2] += gn.z;
            }
        } else if (!scatter) {
            // one fragment per splat, caller's normals: write straight into the leaves
            const float zin = p.z[(size_t)f.src * p.z_stride];
            const float inv_f = f_rcp(-cs.neg_focal);
            const float gZ = gP0.z - (gP0.x * f.x + gP0.y * f.y) * inv_f;
            if (p.gz && zin < 0.f) p.gz[(size_t)f.src * p.z_stride] += gZ;
            if (p.gnormal) {
                float* dst = p.gnormal + (size_t)f.src * p.normal_stride;
                dst[0] += gn.x; dst[1] += gn.y; dst[2] += gn.z;
            }
        } else {
            float* gp_ = p.gpos_src + 3 * (size_t)f.src;
            float* gn_ = p.gn_src + 3 * (size_t)f.src;
            atomicAdd(gp_, gP0.x); atomicAdd(gp_ + 1, gP0.y); atomicAdd(gp_ + 2, gP0.z);
            atomicAdd(gn_, gn.x); atomicAdd(gn_ + 1, gn.y); atomicAdd(gn_ + 2, gn.z);
        }
    }
    __syncthreads();
    const int n_slots = min(p.sm.total, kMaxAccSlots);
    for (int j = threadIdx.x; j < n_slots; j += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kBwdThreads / 32; ++w) s += warp_acc[w * kMaxAccSlots + j];
        if (s != 0.f) atomicAdd(p.acc + j, (double)s);
    }
}

// per source splat: accumulated position gradient -> d/dz, accumulated normal gradient -> the caller's normals
__global__ void __launch_bounds__(256) k_splat_src_finalize(const __grid_constant__ SplatParams p0) {
    SplatParams p = p0;
    splat_scene(p, blockIdx.y);
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= p.n_src) return;
    const CamState& cs = *p.cam;
    if (p.gz) {
        const float zin = p.z[(size_t)s * p.z_stride];
        if (zin < 0.f) {
            float x, y;
            pixel_xy(cs, s, &x, &y);
            const float inv_f = f_rcp(-cs.neg_focal);
            const float* g = p.gpos_src + 3 * (size_t)s;
            p.gz[(size_t)s * p.z_stride] += g[2] - (g[0] * x + g[1] * y) * inv_f;
        }
    }
    if (p.gnormal && !p.estimate) {
        float* dst = p.gnormal + (size_t)s * p.normal_stride;
        const float* g = p.gn_src + 3 * (size_t)s;
        dst[0] += g[0]; dst[1] += g[1]; dst[2] += g[2];
    }
}

// light / material accumulators -> leaves; several scenes of a batch may share a leaf, hence atomics
struct SplatFinalizeParams { GradPtrs gp; SlotMap sm; const double* acc; const CamState* cam; int L; long long ws_stride, light_stride; };
__global__ void __launch_bounds__(128) k_splat_finalize(const __grid_constant__ SplatFinalizeParams p) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (j >= p.sm.total) return;
    const double* acc = (const double*)((const char*)p.acc + p.ws_stride * b);
    const float v = (float)acc[j];
    if (j < p.sm.coeffs) { if (p.gp.albedo) atomicAdd(p.gp.albedo + (j - p.sm.albedo), v); }
    else if (j < p.sm.light_pos) { if (p.gp.coeffs) atomicAdd(p.gp.coeffs + (j - p.sm.coeffs), v); }
    else if (j < p.sm.atten) {
        // camera -> world: l_cc = R^T l_xyz - l_w R^T eye  =>  d/dl_xyz = R g,  d/dl_w = -(R^T eye) . g
        const int q = j - p.sm.light_pos;
        const int l = q / 3, c = q % 3;
        if (p.gp.light_pos && c == 0) {
            const CamState& cs = *(const CamState*)((const char*)p.cam + p.ws_stride * b);
            const double g0 = acc[j], g1 = acc[j + 1], g2 = acc[j + 2];
            float* dst = p.gp.light_pos + b * p.light_stride + 4 * (size_t)l;
            for (int r = 0; r < 3; ++r) atomicAdd(dst + r, (float)(cs.R[3 * r] * g0 + cs.R[3 * r + 1] * g1 + cs.R[3 * r + 2] * g2));
            double gw = 0.0;
            const double gi[3] = {g0, g1, g2};
            for (int i = 0; i < 3; ++i)
                gw -= ((double)cs.R[i] * cs.eye[0] + (double)cs.R[3 + i] * cs.eye[1] + (double)cs.R[6 + i] * cs.eye[2]) * gi[i];
            atomicAdd(dst + 3, (float)gw);
        }
    }
    else if (j < p.sm.colors) { if (p.gp.atten) atomicAdd(p.gp.atten + (j - p.sm.atten), v); }
    else if (j < p.sm.ambient) { if (p.gp.colors) atomicAdd(p.gp.colors + (j - p.sm.colors), v); }
    else if (j < p.sm.gamma) { if (p.gp.ambient) atomicAdd(p.gp.ambient + (j - p.sm.ambient), v); }
}
