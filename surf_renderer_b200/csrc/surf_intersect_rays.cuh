// surf_intersect_rays.cuh - part of libsurf_b200.so (included by surf_isect_rays.cu inside namespace surf, after
// surf_intersect.cuh): k_intersect_screen (math_mode 3), k_intersect_rays / k_intersect_generic (per-ray origins:
// orthographic camera, shadow-ray cross-check), k_intersect_shadow (light-origin filters) and their prep kernels.
#pragma once

// ---------------------------------------------------------------------------------------------------
// k_intersect_screen: opt-in fast intersection kernel (math_mode 3, perspective).  Level 1 classifies EVERY (pixel,
// primitive) pair in registers against the primitive's screen-space bounding circle.  The test is separable: the
// row term (y - v)^2 - rho^2 is shared by the P pixels of a thread (they sit in one image row) and rules all of
// them out when positive; otherwise the column terms are evaluated two pixels per instruction (FADD2 + FFMA2 +
// FMNMX3 per pixel pair).  The rare flagged pairs run the exact reference-order test.  Same persistent-CTA / TMA-ring / atomicMin z-buffer structure as k_intersect.
//   thread -> P consecutive columns of one row; warp -> 4P x 8 pixels; CTA -> 8P x 32 pixels.
// ---------------------------------------------------------------------------------------------------
struct ScreenParams {
    SceneView sc;
    const CamState* cam;
    const float4* circ;              // [total] level-1 records
    const float* rays;               // [3, n]
    unsigned long long* zbuf;        // [n]
    int pix0, n_pix, W, row0;
    int tiles_x, n_tiles, n_chunks, chunk, total;
};

__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// exact test of one (pixel, primitive) pair; returns true and *t on a valid hit.  Out of line: rare path.
__device__ __noinline__ bool exact_pair(const ScreenParams& prm, int idx, int k, float* t_out) {
    const int set = find_set(prm.sc, idx);
    const SetView& sv = prm.sc.sets[set];
    const int local = idx - sv.first;
    const Vec3 o = v3(prm.cam->eye[0], prm.cam->eye[1], prm.cam->eye[2]);
    const Vec3 d = v3(prm.rays[k], prm.rays[(size_t)prm.n_pix + k], prm.rays[2 * (size_t)prm.n_pix + k]);
    Vec3 nn;
    float numer;
    plane_consts_for_origin(sv, local, o, &nn, &numer);
    return exact_hit(sv, local, nn, numer, o, d, prm.cam->near_clip, prm.cam->far_clip, t_out);
}

template <int P>
__global__ void __launch_bounds__(kThreads, (P <= 8 ? 3 : 2)) k_intersect_screen(const __grid_constant__ ScreenParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* stage_buf = reinterpret_cast<float4*>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[kStages];

    const int tid = threadIdx.x;
    const long long n_items = (long long)prm.n_tiles * prm.n_chunks;
    const int lo = (int)(n_items * blockIdx.x / gridDim.x);
    const int hi = (int)(n_items * (blockIdx.x + 1) / gridDim.x);
    if (lo >= hi) return;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int item, int stage) {
        const int c = item % prm.n_chunks;
        const int first = c * prm.chunk;
        const int count = min(prm.chunk, prm.total - first);
        const uint32_t bytes = (uint32_t)count * 16u;
        mbar_expect_tx(&full_bar[stage], bytes);
        tma_bulk_g2s(stage_buf + (size_t)stage * prm.chunk, prm.circ + first, bytes, &full_bar[stage]);
    };
    if (tid == 0)
        for (int k = 0; k < kStages - 1 && lo + k < hi; ++k) issue(lo + k, k);

    // thread geometry inside the CTA tile
    const int warp = tid >> 5, lane = tid & 31;
    const int tcol = (warp & 1) * 4 * P + (lane & 3) * P;     // first column of this thread inside the tile
    const int trow = (warp >> 1) * 8 + (lane >> 2);
    constexpr int TW = 8 * P, TH = 32;

    unsigned long long x2[P / 2];       // image-plane x of the thread's pixels, packed pairs
    float y = 0.f;
    float best_t[P];
    int best_i[P];
    int kbase = 0;                      // output index of the thread's first pixel
    unsigned valid = 0;                 // bit p: pixel p is inside the frame and the launch's pixel range
    int cur_tile = -1;

    auto flush = [&]() {
        if (cur_tile < 0) return;
#pragma unroll
        for (int p = 0; p < P; ++p)
            if (best_i[p] >= 0 && ((valid >> p) & 1u)) {
                unsigned long long key = ((unsigned long long)float_order_key(best_t[p]) << 32) | (unsigned)best_i[p];
                atomicMin(prm.zbuf + kbase + p, key);
            }
    };

    for (int it = lo; it < hi; ++it) {
        const int kk = it - lo;
        const int stage = kk % kStages;
        const uint32_t parity = (uint32_t)((kk / kStages) & 1);
        __syncthreads();
        if (tid == 0 && it + kStages - 1 < hi) issue(it + kStages - 1, (kk + kStages - 1) % kStages);

        const int tile = it / prm.n_chunks;
        if (tile != cur_tile) {
            flush();
            cur_tile = tile;
            const int ty = tile / prm.tiles_x, tx = tile - ty * prm.tiles_x;
            const int row = prm.row0 + ty * TH + trow;
            const int col0 = tx * TW + tcol;
            kbase = row * prm.W + col0 - prm.pix0;
            valid = 0;
            float xs[P];
            float yy = 0.f;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const int col = col0 + p;
                const int k = kbase + p;
                const bool ok = col < prm.W && row < prm.cam->H && k >= 0 && k < prm.n_pix;
                float xv = 3.0e18f;                   // far outside any circle: never flagged
                if (ok) {
                    pixel_xy(*prm.cam, row * prm.W + col, &xv, &yy);
                    valid |= 1u << p;
                }
                xs[p] = xv;
                best_t[p] = INFINITY;
                best_i[p] = -1;
            }
            y = valid ? yy : 3.0e18f;
#pragma unroll
            for (int q = 0; q < P / 2; ++q) x2[q] = pack2(xs[2 * q], xs[2 * q + 1]);
        }

        const int first = (it % prm.n_chunks) * prm.chunk;
        const int count = min(prm.chunk, prm.total - first);
        mbar_wait(&full_bar[stage], parity);
        const float4* __restrict__ s = stage_buf + (size_t)stage * prm.chunk;

        constexpr int G = 4;
        int i = 0;
        for (; i + G <= count; i += G) {
            // row term first: the P pixels of a thread share one image row, so sy = (y - v)^2 - rho^2 > 0 rules out
            // all of them at once.  Only when some lane's row crosses one of the G circles are the column terms
            // evaluated (packed, two pixels per instruction).
            float4 C[G];
            float sy[G];
            float my = INFINITY;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                C[g] = s[i + g];                            // -u, -v, -rho^2, 0   (LDS.128, warp broadcast)
                const float dy = y + C[g].y;
                sy[g] = fmaf(dy, dy, C[g].z);
                my = fminf(my, sy[g]);
            }
            if (!__any_sync(0xffffffffu, my <= 0.f)) continue;
            float m = INFINITY;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const unsigned long long mu = pack2(C[g].x, C[g].x), sy2 = pack2(sy[g], sy[g]);
#pragma unroll
                for (int q = 0; q < P / 2; ++q) {
                    const unsigned long long dx = add2(x2[q], mu);
                    float e0, e1;
                    unpack2(fma2(dx, dx, sy2), e0, e1);
                    m = fminf(m, fminf(e0, e1));
                }
            }
            if (m <= 0.f) {          // rare: some pair of this group lies inside its screen circle
#pragma unroll 1
                for (int g = 0; g < G; ++g) {
                    const float4 C = s[i + g];
                    const float dy = y + C.y;
                    const float sy = fmaf(dy, dy, C.z);
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        float lo_, hi_;
                        unpack2(x2[p >> 1], lo_, hi_);
                        const float dx = ((p & 1) ? hi_ : lo_) + C.x;
                        // slots outside the frame are skipped explicitly: an always-flag record (-rho^2 = -inf: planes,
                        // primitives reaching the camera plane) also "contains" their far-away stand-in coordinate
                        if (((valid >> p) & 1u) && fmaf(dx, dx, sy) <= 0.f) {
                            float t;
                            if (exact_pair(prm, first + i + g, kbase + p, &t) && t < best_t[p]) {
                                best_t[p] = t;
                                best_i[p] = first + i + g;
                            }
                        }
                    }
                }
            }
        }
        for (; i < count; ++i) {     // chunk tail
            const float4 C = s[i];
            const float dy = y + C.y;
            const float sy = fmaf(dy, dy, C.z);
#pragma unroll
            for (int p = 0; p < P; ++p) {
                float lo_, hi_;
                unpack2(x2[p >> 1], lo_, hi_);
                const float dx = ((p & 1) ? hi_ : lo_) + C.x;
                if (((valid >> p) & 1u) && fmaf(dx, dx, sy) <= 0.f) {
                    float t;
                    if (exact_pair(prm, first + i, kbase + p, &t) && t < best_t[p]) {
                        best_t[p] = t;
                        best_i[p] = first + i;
                    }
                }
            }
        }
    }
    flush();
}

// ---------------------------------------------------------------------------------------------------
// k_intersect_rays: the same fused intersection + z-buffer structure for rays with PER-RAY origins
// (orthographic camera pixels, shadow rays).  Disk filter per ray pair: 17 packed FMA-pipe instructions
// (n.o 3, numer 1, n.d 3, t 1, P = o + t d 3, rel = P - c 3, |rel|^2 - r^2 3) + 2 MUFU.RCP.
// MODE 0: camera rays, near <= t <= far;  MODE 1: shadow rays, 0 < t < tmax[ray] (renderer.py:306).
// ---------------------------------------------------------------------------------------------------
struct RayParams {
    SceneView sc;
    const CamState* cam;
    const float4* packed;            // origin-independent records (k_prep_rays)
    const float* gray;               // [7, n]: ox oy oz dx dy dz tmax
    unsigned long long* zbuf;        // [n]
    int n_pix, n_tiles, n_chunks, stage_f4;
    int chunks_before[kMaxSets + 1];
    const int* n_live;               // device count of rays actually stored (compacted shadow rays), or null = n_pix.
                                     // n_pix stays the row stride of `gray`; the work grid shrinks to the live rays.
};

__global__ void __launch_bounds__(256) k_prep_rays(const __grid_constant__ SceneView sc, const float* __restrict__ obound,
                                                   float4* __restrict__ packed) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= sc.total) return;
    const int s = find_set(sc, g);
    const SetView& sv = sc.sets[s];
    const int i = g - sv.first;
    const float ob = obound[0];
    F4 r[4];
    if (sv.kind == KIND_DISK) {
        prep_disk_rays(ld3(sv.pos + (size_t)i * sv.pos_stride), ld3(sv.normal + (size_t)i * sv.normal_stride), sv.radius[i], ob,
                       &r[0], &r[1]);
    } else if (sv.kind == KIND_PLANE) {
        prep_plane_rays(ld3(sv.pos + (size_t)i * sv.pos_stride), ld3(sv.normal + (size_t)i * sv.normal_stride), &r[0]);
    } else if (sv.kind == KIND_SPHERE) {
        prep_sphere_rays(ld3(sv.pos + (size_t)i * sv.pos_stride), sv.radius[i], ob, &r[0]);
    } else {
        const float* f = sv.pos + (size_t)i * 3 * sv.pos_stride;
        prep_triangle_rays(ld3(f), ld3(f + sv.pos_stride), ld3(f + 2 * sv.pos_stride),
                           ld3(sv.normal + (size_t)i * sv.normal_stride), ob, &r[0], &r[1], &r[2], &r[3]);
    }
    const int nf4 = rec_f4(sv.kind);
    float4* dst = packed + sv.rec_off + (size_t)i * nf4;
    for (int k = 0; k < nf4; ++k) dst[k] = make_float4(r[k].x, r[k].y, r[k].z, r[k].w);
}

template <int P>
struct RayRegs {
    unsigned long long ox[P / 2], oy[P / 2], oz[P / 2], dx[P / 2], dy[P / 2], dz[P / 2];
    float tmax[P];
    float best_t[P];
    int best_i[P];
};
template <int P>
__device__ __forceinline__ void ray_of(const RayRegs<P>& r, int p, Vec3* o, Vec3* d) {
    float lo, hi;
    unpack2(r.ox[p >> 1], lo, hi); o->x = (p & 1) ? hi : lo;
    unpack2(r.oy[p >> 1], lo, hi); o->y = (p & 1) ? hi : lo;
    unpack2(r.oz[p >> 1], lo, hi); o->z = (p & 1) ? hi : lo;
    unpack2(r.dx[p >> 1], lo, hi); d->x = (p & 1) ? hi : lo;
    unpack2(r.dy[p >> 1], lo, hi); d->y = (p & 1) ? hi : lo;
    unpack2(r.dz[p >> 1], lo, hi); d->z = (p & 1) ? hi : lo;
}

template <int P, int MODE>
__device__ __forceinline__ void narrow_ray(const RayParams& prm, const SetView& sv, int local, RayRegs<P>& r, int p) {
    Vec3 o, d, nn;
    float numer, t;
    ray_of<P>(r, p, &o, &d);
    plane_consts_for_origin(sv, local, o, &nn, &numer);
    bool hit;
    if (MODE == 0) {
        hit = exact_hit(sv, local, nn, numer, o, d, prm.cam->near_clip, prm.cam->far_clip, &t);
    } else {
        hit = exact_hit(sv, local, nn, numer, o, d, -INFINITY, INFINITY, &t) && t > 0.f && t < r.tmax[p];
    }
    if (hit && t < r.best_t[p]) { r.best_t[p] = t; r.best_i[p] = sv.first + local; }
}

template <int P, int MODE>
__global__ void __launch_bounds__(kThreads, 2) k_intersect_rays(const __grid_constant__ RayParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* stage_buf = reinterpret_cast<float4*>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[kStages];
    const int tid = threadIdx.x;
    constexpr int TILE = kThreads * P;
    const int n_rays = prm.n_live ? *prm.n_live : prm.n_pix;
    const long long n_items = (long long)((n_rays + TILE - 1) / TILE) * prm.n_chunks;
    const int lo = (int)(n_items * blockIdx.x / gridDim.x);
    const int hi = (int)(n_items * (blockIdx.x + 1) / gridDim.x);
    if (lo >= hi) return;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto decode = [&](int c, int* set, int* local0, int* count) {
        int s = 0;
#pragma unroll
        for (int k = 1; k < kMaxSets; ++k)
            if (k < prm.sc.n_sets && c >= prm.chunks_before[k]) s = k;
        const SetView& sv = prm.sc.sets[s];
        const int ppc = prm.stage_f4 / rec_f4(sv.kind);
        const int j = c - prm.chunks_before[s];
        *set = s; *local0 = j * ppc; *count = min(ppc, sv.count - j * ppc);
    };
    auto issue = [&](int item, int stage) {
        int set, local0, count;
        decode(item % prm.n_chunks, &set, &local0, &count);
        const SetView& sv = prm.sc.sets[set];
        const int nf4 = rec_f4(sv.kind);
        const uint32_t bytes = (uint32_t)(count * nf4) * 16u;
        mbar_expect_tx(&full_bar[stage], bytes);
        tma_bulk_g2s(stage_buf + (size_t)stage * prm.stage_f4, prm.packed + sv.rec_off + (size_t)local0 * nf4, bytes, &full_bar[stage]);
    };
    if (tid == 0)
        for (int k = 0; k < kStages - 1 && lo + k < hi; ++k) issue(lo + k, k);

    RayRegs<P> r;
    int cur_tile = -1;
    auto flush = [&]() {
        if (cur_tile < 0) return;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int pix = cur_tile * TILE + p * kThreads + tid;
            if (r.best_i[p] >= 0 && pix < n_rays)
                atomicMin(prm.zbuf + pix, ((unsigned long long)float_order_key(r.best_t[p]) << 32) | (unsigned)r.best_i[p]);
        }
    };

    for (int it = lo; it < hi; ++it) {
        const int kk = it - lo;
        const int stage = kk % kStages;
        const uint32_t parity = (uint32_t)((kk / kStages) & 1);
        __syncthreads();
        if (tid == 0 && it + kStages - 1 < hi) issue(it + kStages - 1, (kk + kStages - 1) % kStages);
        const int tile = it / prm.n_chunks;
        if (tile != cur_tile) {
            flush();
            cur_tile = tile;
            float v[6][P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const int pix = tile * TILE + p * kThreads + tid;
                const bool ok = pix < n_rays;
#pragma unroll
                for (int c = 0; c < 6; ++c) v[c][p] = ok ? prm.gray[(size_t)c * prm.n_pix + pix] : 0.f;
                r.tmax[p] = ok ? prm.gray[(size_t)6 * prm.n_pix + pix] : 0.f;
                r.best_t[p] = INFINITY;
                r.best_i[p] = -1;
            }
#pragma unroll
            for (int q = 0; q < P / 2; ++q) {
                r.ox[q] = pack2(v[0][2 * q], v[0][2 * q + 1]); r.oy[q] = pack2(v[1][2 * q], v[1][2 * q + 1]);
                r.oz[q] = pack2(v[2][2 * q], v[2][2 * q + 1]); r.dx[q] = pack2(v[3][2 * q], v[3][2 * q + 1]);
                r.dy[q] = pack2(v[4][2 * q], v[4][2 * q + 1]); r.dz[q] = pack2(v[5][2 * q], v[5][2 * q + 1]);
            }
        }
        int set, local0, count;
        decode(it % prm.n_chunks, &set, &local0, &count);
        const SetView& sv = prm.sc.sets[set];
        mbar_wait(&full_bar[stage], parity);
        const float4* __restrict__ s = stage_buf + (size_t)stage * prm.stage_f4;
        if (sv.kind == KIND_DISK) {
            for (int i = 0; i < count; ++i) {
                const float4 A = s[2 * i], B = s[2 * i + 1];
                const unsigned long long nx = pack2(A.x, A.x), ny = pack2(A.y, A.y), nz = pack2(A.z, A.z);
                float m = INFINITY;
                float e[P];
#pragma unroll
                for (int q = 0; q < P / 2; ++q) {
                    const unsigned long long no = fma2(nz, r.oz[q], fma2(ny, r.oy[q], mul2(nx, r.ox[q])));
                    const unsigned long long numer = fma2(no, pack2(-1.f, -1.f), pack2(A.w, A.w));
                    const unsigned long long b2 = fma2(nz, r.dz[q], fma2(ny, r.dy[q], mul2(nx, r.dx[q])));
                    float b0, b1;
                    unpack2(b2, b0, b1);
                    const unsigned long long t2 = mul2(numer, pack2(rcp_approx(b0), rcp_approx(b1)));
                    const unsigned long long rx = add2(fma2(t2, r.dx[q], r.ox[q]), pack2(B.x, B.x));
                    const unsigned long long ry = add2(fma2(t2, r.dy[q], r.oy[q]), pack2(B.y, B.y));
                    const unsigned long long rz = add2(fma2(t2, r.dz[q], r.oz[q]), pack2(B.z, B.z));
                    unpack2(fma2(rz, rz, fma2(ry, ry, fma2(rx, rx, pack2(B.w, B.w)))), e[2 * q], e[2 * q + 1]);
                    m = fminf(m, fminf(e[2 * q], e[2 * q + 1]));
                }
                if (m <= 0.f) {
#pragma unroll
                    for (int p = 0; p < P; ++p)
                        if (e[p] <= 0.f) narrow_ray<P, MODE>(prm, sv, local0 + i, r, p);
                }
            }
        } else {
            const int nf4 = rec_f4(sv.kind);
            for (int i = 0; i < count; ++i) {
                const float4* rec = s + (size_t)i * nf4;
                const F4 A = f4(rec[0].x, rec[0].y, rec[0].z, rec[0].w);
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    Vec3 o, d;
                    ray_of<P>(r, p, &o, &d);
                    bool pass = true;
                    if (sv.kind == KIND_SPHERE) pass = sphere_filter_rays(A, o, d);
                    else if (sv.kind == KIND_TRIANGLE)
                        pass = triangle_filter_rays(A, f4(rec[1].x, rec[1].y, rec[1].z, rec[1].w), f4(rec[2].x, rec[2].y, rec[2].z, rec[2].w),
                                                    f4(rec[3].x, rec[3].y, rec[3].z, rec[3].w), o, d);
                    if (pass) narrow_ray<P, MODE>(prm, sv, local0 + i, r, p);
                }
            }
        }
    }
    flush();
}

// generic rays of an orthographic frame: per-pixel origins, one direction
__global__ void __launch_bounds__(256) k_rays_ortho(const CamState* __restrict__ cs, int pix0, int n, float* __restrict__ gray,
                                                    float* __restrict__ obound) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    float len = 0.f;
    if (k < n) {
        const Vec3 o = pixel_ray_origin_ortho(*cs, pix0 + k);
        gray[k] = o.x; gray[(size_t)n + k] = o.y; gray[2 * (size_t)n + k] = o.z;
        gray[3 * (size_t)n + k] = cs->odir[0]; gray[4 * (size_t)n + k] = cs->odir[1]; gray[5 * (size_t)n + k] = cs->odir[2];
        gray[6 * (size_t)n + k] = INFINITY;
        len = sqrtf(o.x * o.x + o.y * o.y + o.z * o.z);
    }
    for (int off = 16; off > 0; off >>= 1) len = fmaxf(len, __shfl_xor_sync(0xffffffffu, len, off));
    if ((threadIdx.x & 31) == 0 && len > 0.f) atomicMax((int*)obound, __float_as_int(len));
}

// generic-origin variant (orthographic camera: per-pixel origins, one direction).  Exact tests only; the
// reference itself only supports this projection up to one tile of pixels.
__global__ void __launch_bounds__(256) k_intersect_generic(const __grid_constant__ SceneView sc,
                                                           const CamState* __restrict__ cs, int pix0, int n,
                                                           unsigned long long* __restrict__ zbuf) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const Vec3 o = pixel_ray_origin_ortho(*cs, pix0 + k);
    const Vec3 d = v3(cs->odir[0], cs->odir[1], cs->odir[2]);
    float best_t = INFINITY;
    int best = -1;
    for (int s = 0; s < sc.n_sets; ++s) {
        const SetView& sv = sc.sets[s];
        for (int i = 0; i < sv.count; ++i) {
            Vec3 nn; float numer, t;
            plane_consts_for_origin(sv, i, o, &nn, &numer);
            if (exact_hit(sv, i, nn, numer, o, d, cs->near_clip, cs->far_clip, &t) && t < best_t) {
                best_t = t; best = sv.first + i;
            }
        }
    }
    if (best >= 0) zbuf[k] = ((unsigned long long)float_order_key(best_t) << 32) | (unsigned)best;
}

// ---------------------------------------------------------------------------------------------------
// k_intersect_shadow: the shadow rays of all lights.  All shadow rays of one light lie on lines through that light, so
// their conservative filters are the CAMERA filters with the light as the common origin: records prepared per light
// by k_prep_lights (disk: n, n.(c - light) | light - c, -(r+slack)^2), ray directions -L in registers, 10 packed
// FMA-pipe instr per ray-disk test in the dense disk loop (instead of 17 for the per-ray-origin filter of
// k_intersect_rays), the packed triangle filter, the sphere discriminant filter.
// The filter tests the whole LINE (no t window), so it also covers the reference's quirk that a hit may lie up to 0.1
// beyond the light (renderer.py:296-306: t is measured from frag_pos + 0.1 L but compared with |light - frag_pos|).
// Candidates run the same exact test as k_intersect_rays - forward ray from frag_pos + 0.1 L, window 0 < t < t_max,
// reference operation order - with the ray read back from `gray`.  Work items: [light][ray tile][primitive chunk];
// light l's compacted rays sit in slots [l*n, l*n + n_live[l]).
// ---------------------------------------------------------------------------------------------------
struct ShadowIsectParams {
    SceneView sc;
    const float4* packed;            // [n_lights][packed_stride] records per light origin
    long long packed_stride;         // float4 units
    const float* gray;               // [7, cap]: origin xyz, direction xyz, t_max
    size_t cap;                      // = n * n_lights
    unsigned long long* zbuf2;       // [cap]
    const int* n_live;               // [n_lights]
    int n, n_lights;
    int tiles_per_light, n_chunks, stage_f4;
    int chunks_before[kMaxSets + 1];
};

__global__ void __launch_bounds__(256) k_prep_lights(const __grid_constant__ SceneView sc, float4* __restrict__ packed,
                                                     long long packed_stride) {
    const int l = blockIdx.y;
    prep_body(sc, ld3(sc.light_pos + (size_t)l * sc.light_pos_stride), packed + (size_t)l * packed_stride,
              blockIdx.x * blockDim.x + threadIdx.x);
}

template <int P>
__global__ void __launch_bounds__(kThreads, 2) k_intersect_shadow(const __grid_constant__ ShadowIsectParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* stage_buf = reinterpret_cast<float4*>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[kStages];
    const int tid = threadIdx.x;
    constexpr int TILE = kThreads * P;
    // Work grid = the LIVE ray tiles of every light (device-side counts): tile_end[l] = live tiles of lights 0..l.
    // Every CTA derives the same table, so the contiguous item ranges balance over what actually has to be traced.
    __shared__ int tile_end[kMaxShadowLights + 1];
    if (tid == 0) {
        int acc = 0;
        for (int l = 0; l < prm.n_lights; ++l) { acc += (prm.n_live[l] + TILE - 1) / TILE; tile_end[l] = acc; }
        tile_end[kMaxShadowLights] = acc;
    }
    __syncthreads();
    const long long n_items = (long long)tile_end[kMaxShadowLights] * prm.n_chunks;
    const int lo = (int)(n_items * blockIdx.x / gridDim.x);
    const int hi = (int)(n_items * (blockIdx.x + 1) / gridDim.x);
    if (lo >= hi) return;
    auto light_of = [&](int tile, int* lt) {
        int l = 0;
        while (tile >= tile_end[l]) ++l;
        *lt = tile - (l ? tile_end[l - 1] : 0);
        return l;
    };
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto decode = [&](int c, int* set, int* local0, int* count) {
        int s = 0;
#pragma unroll
        for (int k = 1; k < kMaxSets; ++k)
            if (k < prm.sc.n_sets && c >= prm.chunks_before[k]) s = k;
        const SetView& sv = prm.sc.sets[s];
        const int ppc = prm.stage_f4 / rec_f4(sv.kind);
        const int j = c - prm.chunks_before[s];
        *set = s; *local0 = j * ppc; *count = min(ppc, sv.count - j * ppc);
    };
    auto issue = [&](int item, int stage) {
        int set, local0, count;
        decode(item % prm.n_chunks, &set, &local0, &count);
        int lt_unused;
        const int l = light_of(item / prm.n_chunks, &lt_unused);
        const SetView& sv = prm.sc.sets[set];
        const int nf4 = rec_f4(sv.kind);
        const uint32_t bytes = (uint32_t)(count * nf4) * 16u;
        mbar_expect_tx(&full_bar[stage], bytes);
        tma_bulk_g2s(stage_buf + (size_t)stage * prm.stage_f4,
                     prm.packed + (size_t)l * prm.packed_stride + sv.rec_off + (size_t)local0 * nf4, bytes, &full_bar[stage]);
    };
    if (tid == 0)
        for (int k = 0; k < kStages - 1 && lo + k < hi; ++k) issue(lo + k, k);

    PixelRegs<P> r;
    int cur_tile = -1;
    size_t slot0 = 0;            // slot of this thread's pixel slot 0 in the current tile
    int live_in_tile = 0;        // rays of the current tile that exist (0: the tile is past the light's ray count)
    auto flush = [&]() {
        if (cur_tile < 0) return;
#pragma unroll
        for (int p = 0; p < P; ++p)
            if (r.best_i[p] >= 0)
                atomicMin(prm.zbuf2 + slot0 + (size_t)p * kThreads,
                          ((unsigned long long)float_order_key(r.best_t[p]) << 32) | (unsigned)r.best_i[p]);
    };

    for (int it = lo; it < hi; ++it) {
        const int kk = it - lo;
        const int stage = kk % kStages;
        const uint32_t parity = (uint32_t)((kk / kStages) & 1);
        __syncthreads();
        if (tid == 0 && it + kStages - 1 < hi) issue(it + kStages - 1, (kk + kStages - 1) % kStages);
        const int tile = it / prm.n_chunks;
        if (tile != cur_tile) {
            flush();
            cur_tile = tile;
            int lt;
            const int l = light_of(tile, &lt);
            const int n_l = prm.n_live[l];
            live_in_tile = max(0, min(TILE, n_l - lt * TILE));
            slot0 = (size_t)l * prm.n + (size_t)lt * TILE + tid;
            float d[3][P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const bool ok = p * kThreads + tid < live_in_tile;
                const size_t sl = slot0 + (size_t)p * kThreads;
                // direction from the light toward the fragment: -L (a null direction gives NaN margins = no candidate)
                d[0][p] = ok ? -prm.gray[3 * prm.cap + sl] : 0.f;
                d[1][p] = ok ? -prm.gray[4 * prm.cap + sl] : 0.f;
                d[2][p] = ok ? -prm.gray[5 * prm.cap + sl] : 0.f;
                r.best_t[p] = INFINITY;
                r.best_i[p] = -1;
            }
#pragma unroll
            for (int q = 0; q < P / 2; ++q) {
                r.dx[q] = pack2(d[0][2 * q], d[0][2 * q + 1]);
                r.dy[q] = pack2(d[1][2 * q], d[1][2 * q + 1]);
                r.dz[q] = pack2(d[2][2 * q], d[2][2 * q + 1]);
            }
        }
        int set, local0, count;
        decode(it % prm.n_chunks, &set, &local0, &count);
        const SetView& sv = prm.sc.sets[set];
        mbar_wait(&full_bar[stage], parity);
        if (live_in_tile == 0) continue;                 // CTA-uniform: nothing to trace in this tile
        const float4* __restrict__ s = stage_buf + (size_t)stage * prm.stage_f4;
        auto nf = [&](int local, const float4&, int p) {     // the exact shadow-ray test of k_intersect_rays<., 1>
            if (p * kThreads + tid >= live_in_tile) return;
            const size_t sl = slot0 + (size_t)p * kThreads;
            const Vec3 o = v3(prm.gray[sl], prm.gray[prm.cap + sl], prm.gray[2 * prm.cap + sl]);
            const Vec3 dir = v3(prm.gray[3 * prm.cap + sl], prm.gray[4 * prm.cap + sl], prm.gray[5 * prm.cap + sl]);
            const float tmax = prm.gray[6 * prm.cap + sl];
            Vec3 nn;
            float numer, t;
            plane_consts_for_origin(sv, local, o, &nn, &numer);
            const bool hit = exact_hit(sv, local, nn, numer, o, dir, -INFINITY, INFINITY, &t) && t > 0.f && t < tmax;
            if (hit && t < r.best_t[p]) { r.best_t[p] = t; r.best_i[p] = sv.first + local; }
        };
        // (-L is a unit vector and the filter tests the LINE through the light: the bounding-sphere form applies)
        if (sv.kind == KIND_DISK) chunk_disks_dense<P, true>(s, local0, count, r, nf);
        else if (sv.kind == KIND_TRIANGLE) chunk_triangles_packed<P, false>(s, local0, count, r, nf);
        else if (sv.kind == KIND_SPHERE) chunk_spheres_fn<P>(s, local0, count, r, nf);
        else chunk_planes_fn<P>(s, local0, count, nf);
    }
    flush();
}
