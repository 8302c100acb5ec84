// surf_intersect.cuh - part of libsurf_b200.so (included by the surf_isect_*.cu translation units inside namespace surf).
// the camera-ray intersection + z-buffer kernels k_intersect / k_intersect_batch and the chunk loops they share with
// k_intersect_shadow (surf_intersect_rays.cuh).  Templates and forceinline device code only.
#pragma once

// ---------------------------------------------------------------------------------------------------
// k_intersect: the fused intersection + z-buffer kernel (perspective: one common ray origin)
// ---------------------------------------------------------------------------------------------------
constexpr int kThreads = 256;
constexpr int kStages = 3;

struct IsectParams {
    SceneView sc;
    const CamState* cam;
    const float4* packed;
    const float* rays;               // [3, n]
    unsigned long long* zbuf;        // [n]
    int n_pix;                       // pixels in this launch's range
    int n_tiles, n_chunks;           // work grid: items = n_tiles * n_chunks
    int stage_f4;                    // float4 capacity of one smem stage
    int chunks_before[kMaxSets + 1]; // prefix sum of chunks per set
    int tiles_per_scene;             // k_intersect_batch: n_tiles = n_scenes * tiles_per_scene (else = n_tiles)
    // k_intersect_batch, 2-D pixel tiles (tiles_x > 0): a CTA owns 64 x 32 pixels, a warp a compact 16 x 16 block
    // (lane -> 16 x 2, slot p -> 2 rows further down), so that a splat a few pixels wide puts its narrow-phase
    // candidates into one or two warps instead of every warp that owns one of its rows.  tiles_x = 0: flat tiles
    // of kThreads * P consecutive pixels (the single-scene kernel's mapping; any width, no masked-out lanes).
    int tiles_x, W, pix0;
    // k_intersect_batch: hybrid work distribution.  Every CTA first walks a contiguous static share of `static_per`
    // items (consecutive items share the pixel tile: rays and running best stay in registers), then the CTAs draw
    // runs of `run_len` items of the remaining pool [dyn_begin, n_items) from *work_counter (zeroed by the host before
    // the launch).  Dense frames have a data-dependent narrow phase, so purely static equal shares finish unevenly
    // (ncu on bunny 256x256: the SMs were busy 78 % of the kernel's duration).
    int* work_counter;
    int static_per, dyn_begin, run_len;
};

// surf_isect_batch.cu: launches k_intersect_batch<P, mode> (run_intersect in surf_isect_main.cu picks P, mode and the grid)
struct BatchArgs;
int launch_intersect_batch(const IsectParams& prm, const BatchArgs& ba, int P, int mode, int grid, size_t smem, cudaStream_t st);

__device__ __forceinline__ int prims_per_chunk(int stage_f4, int kind) { return stage_f4 / rec_f4(kind); }

// decode a global chunk id -> (set, first local primitive, count)
__device__ __forceinline__ void decode_chunk(const IsectParams& p, int c, int* set, int* local0, int* count) {
    int s = 0;
#pragma unroll
    for (int k = 1; k < kMaxSets; ++k)
        if (k < p.sc.n_sets && c >= p.chunks_before[k]) s = k;
    const SetView& sv = p.sc.sets[s];
    const int ppc = prims_per_chunk(p.stage_f4, sv.kind);
    const int j = c - p.chunks_before[s];
    *set = s;
    *local0 = j * ppc;
    *count = min(ppc, sv.count - j * ppc);
}

template <int P>
struct PixelRegs {
    // ray directions of the P pixels this thread owns, stored as packed pairs (pixel 2q, 2q+1)
    unsigned long long dx[P / 2], dy[P / 2], dz[P / 2];
    float best_t[P];
    int best_i[P];
};

template <int P>
__device__ __forceinline__ Vec3 ray_of(const PixelRegs<P>& r, int p) {
    float lo, hi;
    Vec3 d;
    unpack2(r.dx[p >> 1], lo, hi); d.x = (p & 1) ? hi : lo;
    unpack2(r.dy[p >> 1], lo, hi); d.y = (p & 1) ? hi : lo;
    unpack2(r.dz[p >> 1], lo, hi); d.z = (p & 1) ? hi : lo;
    return d;
}

// exact narrow phase for all P pixels of this thread against one candidate primitive
template <int P>
__device__ __forceinline__ void narrow(const IsectParams& prm, const SetView& sv, int local, float4 A, Vec3 eye,
                                       float near_clip, float far_clip, PixelRegs<P>& r) {
#pragma unroll
    for (int p = 0; p < P; ++p) {
        float t;
        Vec3 d = ray_of<P>(r, p);
        bool hit = exact_hit(sv, local, v3(A.x, A.y, A.z), A.w, eye, d, near_clip, far_clip, &t);
        if (hit && t < r.best_t[p]) { r.best_t[p] = t; r.best_i[p] = sv.first + local; }
    }
}

// filter margin e = |(o-c) + t d|^2 - (r+slack)^2 of one disk for pixel pair q (two pixels per instruction)
template <int P>
__device__ __forceinline__ unsigned long long disk_margin2(const float4& A, const float4& B, const PixelRegs<P>& r, int q) {
    const unsigned long long nx = pack2(A.x, A.x), ny = pack2(A.y, A.y), nz = pack2(A.z, A.z);
    unsigned long long b2 = fma2(nz, r.dz[q], fma2(ny, r.dy[q], mul2(nx, r.dx[q])));
    float b0, b1;
    unpack2(b2, b0, b1);
    unsigned long long t2 = mul2(pack2(A.w, A.w), pack2(rcp_approx(b0), rcp_approx(b1)));
    unsigned long long rx = fma2(t2, r.dx[q], pack2(B.x, B.x));
    unsigned long long ry = fma2(t2, r.dy[q], pack2(B.y, B.y));
    unsigned long long rz = fma2(t2, r.dz[q], pack2(B.z, B.z));
    return fma2(rz, rz, fma2(ry, ry, fma2(rx, rx, pack2(B.w, B.w))));
}

// exact narrow phase of one pixel against one disk
template <int P>
__device__ __forceinline__ void narrow_one(const SetView& sv, int local, const float4& A, Vec3 eye, float near_clip,
                                           float far_clip, PixelRegs<P>& r, int p) {
    float t;
    Vec3 d = ray_of<P>(r, p);
    bool hit = exact_hit(sv, local, v3(A.x, A.y, A.z), A.w, eye, d, near_clip, far_clip, &t);
    if (hit && t < r.best_t[p]) { r.best_t[p] = t; r.best_i[p] = sv.first + local; }
}

// MODE 0: packed FFMA2 filter, G disks per branch (no per-primitive control dependency, ILP across disks)
// MODE 1: scalar FFMA filter, one branch per disk          MODE 2: packed FFMA2 filter, one branch per disk
template <int P, int MODE>
__device__ __forceinline__ void chunk_disks(const IsectParams& prm, const SetView& sv, const float4* __restrict__ s,
                                            int local0, int count, Vec3 eye, float near_clip, float far_clip,
                                            PixelRegs<P>& r) {
    int i = 0;
    if (MODE == 0) {
        // Software-pipelined: the records of group k+1 are fetched from shared memory while group k computes,
        // and the (rare) branch taken in iteration k tests the filter minimum of group k-1, which finished long
        // ago - so neither the LDS latency nor the FMNMX3 chain + branch resolution sits on the critical path.
        constexpr int G = (P >= 8) ? 2 : 4;
        const int ngroups = count / G;
        if (ngroups > 0) {
            float4 A[G], B[G];
#pragma unroll
            for (int g = 0; g < G; ++g) { A[g] = s[2 * g]; B[g] = s[2 * g + 1]; }
            float m_prev = INFINITY;
            for (int k = 0; k < ngroups; ++k) {
                float4 An[G], Bn[G];
                const int nxt = (k + 1 < ngroups ? k + 1 : k) * G;     // last iteration re-reads its own group
#pragma unroll
                for (int g = 0; g < G; ++g) { An[g] = s[2 * (nxt + g)]; Bn[g] = s[2 * (nxt + g) + 1]; }
                // Stage-major evaluation: each warp-uniform scalar of a disk record is consumed by the Q = P/2 pixel
                // pairs back to back in the same operand slot, so after the first read it comes from the operand
                // reuse cache.  (An FFMA2 reading two register pairs PLUS a fresh scalar needs 3 register-file
                // cycles instead of 2 - measured, tools/ubench/pipes.cu.)
                constexpr int Q = P / 2;
                float m = INFINITY;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const unsigned long long nx = pack2(A[g].x, A[g].x), ny = pack2(A[g].y, A[g].y), nz = pack2(A[g].z, A[g].z);
                    const unsigned long long nm = pack2(A[g].w, A[g].w);
                    const unsigned long long ox = pack2(B[g].x, B[g].x), oy = pack2(B[g].y, B[g].y), oz = pack2(B[g].z, B[g].z);
                    const unsigned long long nr = pack2(B[g].w, B[g].w);
                    unsigned long long b2[Q], t2[Q], rx[Q], ry[Q], rz[Q], e2[Q];
#pragma unroll
                    for (int q = 0; q < Q; ++q) b2[q] = mul2(nx, r.dx[q]);
#pragma unroll
                    for (int q = 0; q < Q; ++q) b2[q] = fma2(ny, r.dy[q], b2[q]);
#pragma unroll
                    for (int q = 0; q < Q; ++q) b2[q] = fma2(nz, r.dz[q], b2[q]);
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        float b0, b1;
                        unpack2(b2[q], b0, b1);
                        t2[q] = mul2(nm, pack2(rcp_approx(b0), rcp_approx(b1)));
                    }
#pragma unroll
                    for (int q = 0; q < Q; ++q) rx[q] = fma2(t2[q], r.dx[q], ox);
#pragma unroll
                    for (int q = 0; q < Q; ++q) ry[q] = fma2(t2[q], r.dy[q], oy);
#pragma unroll
                    for (int q = 0; q < Q; ++q) rz[q] = fma2(t2[q], r.dz[q], oz);
#pragma unroll
                    for (int q = 0; q < Q; ++q) e2[q] = fma2(rx[q], rx[q], nr);
#pragma unroll
                    for (int q = 0; q < Q; ++q) e2[q] = fma2(ry[q], ry[q], e2[q]);
#pragma unroll
                    for (int q = 0; q < Q; ++q) e2[q] = fma2(rz[q], rz[q], e2[q]);
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        float e0, e1;
                        unpack2(e2[q], e0, e1);
                        m = fminf(m, fminf(e0, e1));     // NaN-ignoring min: NaN margins are misses
                    }
                }
                if (m_prev <= 0.f) {       // rare: a pair of the PREVIOUS group passed the conservative filter
                    const int base = (k - 1) * G;
#pragma unroll 1
                    for (int g = 0; g < G; ++g) {
                        const float4 Ag = s[2 * (base + g)], Bg = s[2 * (base + g) + 1];
#pragma unroll
                        for (int q = 0; q < P / 2; ++q) {
                            float e0, e1;
                            unpack2(disk_margin2<P>(Ag, Bg, r, q), e0, e1);
                            if (e0 <= 0.f) narrow_one<P>(sv, local0 + base + g, Ag, eye, near_clip, far_clip, r, 2 * q);
                            if (e1 <= 0.f) narrow_one<P>(sv, local0 + base + g, Ag, eye, near_clip, far_clip, r, 2 * q + 1);
                        }
                    }
                }
                m_prev = m;
#pragma unroll
                for (int g = 0; g < G; ++g) { A[g] = An[g]; B[g] = Bn[g]; }
            }
            if (m_prev <= 0.f) {
                const int base = (ngroups - 1) * G;
#pragma unroll 1
                for (int g = 0; g < G; ++g) {
                    const float4 Ag = s[2 * (base + g)], Bg = s[2 * (base + g) + 1];
#pragma unroll
                    for (int q = 0; q < P / 2; ++q) {
                        float e0, e1;
                        unpack2(disk_margin2<P>(Ag, Bg, r, q), e0, e1);
                        if (e0 <= 0.f) narrow_one<P>(sv, local0 + base + g, Ag, eye, near_clip, far_clip, r, 2 * q);
                        if (e1 <= 0.f) narrow_one<P>(sv, local0 + base + g, Ag, eye, near_clip, far_clip, r, 2 * q + 1);
                    }
                }
            }
            i = ngroups * G;
        }
    }
    for (; i < count; ++i) {
        const float4 A = s[2 * i];       // n.x n.y n.z numer      (LDS.128, warp-broadcast)
        const float4 B = s[2 * i + 1];   // oc.x oc.y oc.z -(r+slack)^2
        bool any = false;
        if (MODE != 1) {
#pragma unroll
            for (int q = 0; q < P / 2; ++q) {
                float e0, e1;
                unpack2(disk_margin2<P>(A, B, r, q), e0, e1);
                any |= (e0 <= 0.f) | (e1 <= 0.f);
            }
        } else {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                Vec3 d = ray_of<P>(r, p);
                float b = fmaf(A.z, d.z, fmaf(A.y, d.y, A.x * d.x));
                float t = A.w * rcp_approx(b);
                float rx = fmaf(t, d.x, B.x), ry = fmaf(t, d.y, B.y), rz = fmaf(t, d.z, B.z);
                any |= fmaf(rz, rz, fmaf(ry, ry, fmaf(rx, rx, B.w))) <= 0.f;
            }
        }
        if (any) narrow<P>(prm, sv, local0 + i, A, eye, near_clip, far_clip, r);
    }
}

// chunk_disks for dense frames (k_intersect_batch: GAN batches of small frames where a splat is several pixels wide
// and 3e-4 .. 1e-3 of all pairs pass the filter): MODE 0's loop with PER-DISK filter minima, so that the rare path
// re-evaluates only the disk that flagged instead of the whole group.  Measured -7.5 % on bunny 256x256 / config D,
// but -3 % on config E when k_intersect itself is built this way (four more registers, one more FMNMX per group,
// another ptxas schedule of the packed-FMA loop) - hence a separate copy used by the batch kernel only.
// `nf(local, A, p)` runs the exact test of pixel slot p against disk `local` (A = its record's first float4): the
// camera-ray narrow phase for k_intersect_batch, the shadow-ray narrow phase for k_intersect_shadow.
// SPHERE (camera rays only: unit directions): the per-disk minima come from the bounding-sphere test of the disk instead
// of the plane filter - the ray passes within rs of the centre iff (oc . d)^2 >= |oc|^2 - rs^2, FOUR packed FMAs per pixel
// pair and no reciprocal instead of ten + two MUFU.RCP; a disk that passes is re-filtered per pixel by the plane filter in
// the rare path, exactly as before.  c0 = |oc|^2 (1 - 32 u) - rs^2 is formed per disk from the record (rs^2 = -B.w is the
// plane filter's inflated radius): 3 u for the fp32 sum of squares, 16 u for the evaluation of the test (see
// k_sphere_records in surf_isect_const.cu), the rest margin; c0 <= 0 (the eye inside the sphere) always passes.
template <int P, bool SPHERE = false, class NarrowFn>
__device__ __forceinline__ void chunk_disks_dense(const float4* __restrict__ s, int local0, int count, PixelRegs<P>& r,
                                                  NarrowFn&& nf) {
    int i = 0;
    {
        // Software-pipelined: the records of group k+1 are fetched from shared memory while group k computes,
        // and the (rare) branch taken in iteration k tests the filter minimum of group k-1, which finished long
        // ago - so neither the LDS latency nor the FMNMX3 chain + branch resolution sits on the critical path.
        constexpr int G = (P >= 8) ? 2 : 4;
        const int ngroups = count / G;
        if (ngroups > 0) {
            float4 A[G], B[G];
#pragma unroll
            for (int g = 0; g < G; ++g) { A[g] = s[2 * g]; B[g] = s[2 * g + 1]; }
            float m_prev[G];                 // per-disk filter minima of the previous group
#pragma unroll
            for (int g = 0; g < G; ++g) m_prev[g] = INFINITY;
            for (int k = 0; k < ngroups; ++k) {
                float4 An[G], Bn[G];
                const int nxt = (k + 1 < ngroups ? k + 1 : k) * G;     // last iteration re-reads its own group
#pragma unroll
                for (int g = 0; g < G; ++g) { An[g] = s[2 * (nxt + g)]; Bn[g] = s[2 * (nxt + g) + 1]; }
                // Stage-major evaluation: each warp-uniform scalar of a disk record is consumed by the Q = P/2 pixel
                // pairs back to back in the same operand slot, so after the first read it comes from the operand
                // reuse cache.  (An FFMA2 reading two register pairs PLUS a fresh scalar needs 3 register-file
                // cycles instead of 2 - measured, tools/ubench/pipes.cu.)
                constexpr int Q = P / 2;
                float m[G];
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    if (SPHERE) {
                        const float oc2 = __fmaf_rn(B[g].z, B[g].z, __fmaf_rn(B[g].y, B[g].y, __fmul_rn(B[g].x, B[g].x)));
                        const float nc0 = -__fmaf_rn(oc2, 1.f - 1.9073486328125e-6f, B[g].w);        // -(|oc|^2 (1 - 2^-19) - rs^2)
                        const unsigned long long ox = pack2(B[g].x, B[g].x), oy = pack2(B[g].y, B[g].y), oz = pack2(B[g].z, B[g].z);
                        const unsigned long long nc = pack2(nc0, nc0);
                        unsigned long long s2[Q];
#pragma unroll
                        for (int q = 0; q < Q; ++q) s2[q] = mul2(ox, r.dx[q]);
#pragma unroll
                        for (int q = 0; q < Q; ++q) s2[q] = fma2(oy, r.dy[q], s2[q]);
#pragma unroll
                        for (int q = 0; q < Q; ++q) s2[q] = fma2(oz, r.dz[q], s2[q]);
#pragma unroll
                        for (int q = 0; q < Q; ++q) s2[q] = fma2(s2[q], s2[q], nc);
                        float mx = -INFINITY;
#pragma unroll
                        for (int q = 0; q < Q; ++q) {
                            float e0, e1;
                            unpack2(s2[q], e0, e1);
                            mx = fmaxf(mx, fmaxf(e0, e1));          // NaN-ignoring max
                        }
                        m[g] = -mx;
                        continue;
                    }
                    m[g] = INFINITY;
                    const unsigned long long nx = pack2(A[g].x, A[g].x), ny = pack2(A[g].y, A[g].y), nz = pack2(A[g].z, A[g].z);
                    const unsigned long long nm = pack2(A[g].w, A[g].w);
                    const unsigned long long ox = pack2(B[g].x, B[g].x), oy = pack2(B[g].y, B[g].y), oz = pack2(B[g].z, B[g].z);
                    const unsigned long long nr = pack2(B[g].w, B[g].w);
                    unsigned long long b2[Q], t2[Q], rx[Q], ry[Q], rz[Q], e2[Q];
#pragma unroll
                    for (int q = 0; q < Q; ++q) b2[q] = mul2(nx, r.dx[q]);
#pragma unroll
                    for (int q = 0; q < Q; ++q) b2[q] = fma2(ny, r.dy[q], b2[q]);
#pragma unroll
                    for (int q = 0; q < Q; ++q) b2[q] = fma2(nz, r.dz[q], b2[q]);
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        float b0, b1;
                        unpack2(b2[q], b0, b1);
                        t2[q] = mul2(nm, pack2(rcp_approx(b0), rcp_approx(b1)));
                    }
#pragma unroll
                    for (int q = 0; q < Q; ++q) rx[q] = fma2(t2[q], r.dx[q], ox);
#pragma unroll
                    for (int q = 0; q < Q; ++q) ry[q] = fma2(t2[q], r.dy[q], oy);
#pragma unroll
                    for (int q = 0; q < Q; ++q) rz[q] = fma2(t2[q], r.dz[q], oz);
#pragma unroll
                    for (int q = 0; q < Q; ++q) e2[q] = fma2(rx[q], rx[q], nr);
#pragma unroll
                    for (int q = 0; q < Q; ++q) e2[q] = fma2(ry[q], ry[q], e2[q]);
#pragma unroll
                    for (int q = 0; q < Q; ++q) e2[q] = fma2(rz[q], rz[q], e2[q]);
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        float e0, e1;
                        unpack2(e2[q], e0, e1);
                        m[g] = fminf(m[g], fminf(e0, e1));     // NaN-ignoring min: NaN margins are misses
                    }
                }
                float m_any = m_prev[0];
#pragma unroll
                for (int g = 1; g < G; ++g) m_any = fminf(m_any, m_prev[g]);
                if (m_any <= 0.f) {        // rare: a pair of the PREVIOUS group passed the conservative filter
                    const int base = (k - 1) * G;
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        if (!(m_prev[g] <= 0.f)) continue;          // only the disk(s) that flagged
                        const float4 Ag = s[2 * (base + g)], Bg = s[2 * (base + g) + 1];
#pragma unroll
                        for (int q = 0; q < P / 2; ++q) {
                            float e0, e1;
                            unpack2(disk_margin2<P>(Ag, Bg, r, q), e0, e1);
                            if (e0 <= 0.f) nf(local0 + base + g, Ag, 2 * q);
                            if (e1 <= 0.f) nf(local0 + base + g, Ag, 2 * q + 1);
                        }
                    }
                }
#pragma unroll
                for (int g = 0; g < G; ++g) { m_prev[g] = m[g]; A[g] = An[g]; B[g] = Bn[g]; }
            }
            {
                const int base = (ngroups - 1) * G;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    if (!(m_prev[g] <= 0.f)) continue;
                    const float4 Ag = s[2 * (base + g)], Bg = s[2 * (base + g) + 1];
#pragma unroll
                    for (int q = 0; q < P / 2; ++q) {
                        float e0, e1;
                        unpack2(disk_margin2<P>(Ag, Bg, r, q), e0, e1);
                        if (e0 <= 0.f) nf(local0 + base + g, Ag, 2 * q);
                        if (e1 <= 0.f) nf(local0 + base + g, Ag, 2 * q + 1);
                    }
                }
            }
            i = ngroups * G;
        }
    }
    for (; i < count; ++i) {             // group remainder
        const float4 A = s[2 * i], B = s[2 * i + 1];
#pragma unroll
        for (int q = 0; q < P / 2; ++q) {
            float e0, e1;
            unpack2(disk_margin2<P>(A, B, r, q), e0, e1);
            if (e0 <= 0.f) nf(local0 + i, A, 2 * q);
            if (e1 <= 0.f) nf(local0 + i, A, 2 * q + 1);
        }
    }
}

template <int P>
__device__ __forceinline__ void chunk_planes(const IsectParams& prm, const SetView& sv, const float4* __restrict__ s,
                                             int local0, int count, Vec3 eye, float near_clip, float far_clip,
                                             PixelRegs<P>& r) {
    for (int i = 0; i < count; ++i) narrow<P>(prm, sv, local0 + i, s[i], eye, near_clip, far_clip, r);
}

template <int P>
__device__ __forceinline__ void chunk_spheres(const IsectParams& prm, const SetView& sv, const float4* __restrict__ s,
                                              int local0, int count, Vec3 eye, float near_clip, float far_clip,
                                              PixelRegs<P>& r) {
    for (int i = 0; i < count; ++i) {
        const float4 S = s[i];
        bool any = false;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            Vec3 d = ray_of<P>(r, p);
            float hb = fmaf(S.z, d.z, fmaf(S.y, d.y, S.x * d.x));
            any |= fmaf(hb, hb, -S.w) >= 0.f;
        }
        if (any) narrow<P>(prm, sv, local0 + i, S, eye, near_clip, far_clip, r);
    }
}

template <int P>
__device__ __forceinline__ void chunk_triangles(const IsectParams& prm, const SetView& sv, const float4* __restrict__ s,
                                                int local0, int count, Vec3 eye, float near_clip, float far_clip,
                                                PixelRegs<P>& r) {
    for (int i = 0; i < count; ++i) {
        const float4 A = s[4 * i], W0 = s[4 * i + 1], W1 = s[4 * i + 2], W2 = s[4 * i + 3];
        bool any = false;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            Vec3 d = ray_of<P>(r, p);          // edge-function form (prep_triangle): three dot products
            float c0 = fmaf(W0.z, d.z, fmaf(W0.y, d.y, fmaf(W0.x, d.x, W0.w)));
            float c1 = fmaf(W1.z, d.z, fmaf(W1.y, d.y, fmaf(W1.x, d.x, W1.w)));
            float c2 = fmaf(W2.z, d.z, fmaf(W2.y, d.y, fmaf(W2.x, d.x, W2.w)));
            any |= (c0 >= 0.f) & (c1 >= 0.f) & (c2 >= 0.f);
        }
        if (any) narrow<P>(prm, sv, local0 + i, A, eye, near_clip, far_clip, r);
    }
}

// chunk_triangles for the k_intersect_batch body (strided batches, math_mode 4, and single scenes that contain triangle
// sets): the conservative filter packed two pixels per instruction like the disk filter (16 FFMA2/FMUL2 + 2 MUFU.RCP
// per triangle and pixel pair, stage-major so the warp-uniform record scalars stay in the operand-reuse cache), one
// FMNMX3 per pixel for min(c0, c1, c2), a max over the thread's pixels, and the (rare) branch on that maximum taken
// one triangle late so that neither the LDS latency of the next record nor the min/max chain is on the critical path.
// torus 512x512: 0.276 -> 0.241 ms.  (In k_intersect itself the same code moved the register allocation and with it
// the disk loop's schedule, -1.5 % on config E, so the single-scene splat kernel keeps the scalar version above.)
// rare path of chunk_triangles_packed: re-evaluate the flagged triangle's filter per pixel pair and run the exact
// test only on the pixels that pass (instead of on all P pixels of the thread)
// EDGE = true: records of prep_triangle (edge functions G_i . d + slack >= 0: 9 packed FMAs per triangle and pixel
// pair, no reciprocal) - camera rays.  EDGE = false: records of prep_triangle_line (t (d.W_i) + w_i >= 0 with
// t ~ numer rcp(n.d): 16 packed FMAs + 2 MUFU) - the shadow rays of a light, where t may have either sign.
template <int P, bool EDGE, class NarrowFn>
__device__ __forceinline__ void triangle_candidates(const float4* __restrict__ rec, int local, const PixelRegs<P>& r, NarrowFn&& nf) {
    const float4 A = rec[0], W0 = rec[1], W1 = rec[2], W2 = rec[3];
#pragma unroll
    for (int q = 0; q < P / 2; ++q) {
        unsigned long long t2 = 0;
        if (!EDGE) {
            const unsigned long long b2 = fma2(pack2(A.z, A.z), r.dz[q], fma2(pack2(A.y, A.y), r.dy[q], mul2(pack2(A.x, A.x), r.dx[q])));
            float b0, b1;
            unpack2(b2, b0, b1);
            t2 = mul2(pack2(A.w, A.w), pack2(rcp_approx(b0), rcp_approx(b1)));
        }
        float m0 = INFINITY, m1 = INFINITY;
#define SURF_EDGE1(W)                                                                                                   \
        {                                                                                                               \
            float c0, c1;                                                                                               \
            if (EDGE) {                                                                                                 \
                unpack2(fma2(pack2(W.z, W.z), r.dz[q], fma2(pack2(W.y, W.y), r.dy[q], fma2(pack2(W.x, W.x), r.dx[q], pack2(W.w, W.w)))), c0, c1); \
            } else {                                                                                                    \
                const unsigned long long u2 = fma2(pack2(W.z, W.z), r.dz[q], fma2(pack2(W.y, W.y), r.dy[q], mul2(pack2(W.x, W.x), r.dx[q]))); \
                unpack2(fma2(t2, u2, pack2(W.w, W.w)), c0, c1);                                                         \
            }                                                                                                           \
            m0 = fminf(m0, c0); m1 = fminf(m1, c1);                                                                     \
        }
        SURF_EDGE1(W0)
        SURF_EDGE1(W1)
        SURF_EDGE1(W2)
#undef SURF_EDGE1
        if (m0 >= 0.f) nf(local, A, 2 * q);
        if (m1 >= 0.f) nf(local, A, 2 * q + 1);
    }
}

template <int P, bool EDGE, class NarrowFn>
__device__ __forceinline__ void chunk_triangles_packed(const float4* __restrict__ s, int local0, int count, PixelRegs<P>& r,
                                                       NarrowFn&& nf) {
    if (count <= 0) return;
    constexpr int Q = P / 2;
    float4 A = s[0], W0 = s[1], W1 = s[2], W2 = s[3];
    float mx_prev = -INFINITY;
    for (int i = 0; i < count; ++i) {
        const int nxt = (i + 1 < count ? i + 1 : i) * 4;
        const float4 An = s[nxt], W0n = s[nxt + 1], W1n = s[nxt + 2], W2n = s[nxt + 3];
        unsigned long long c0[Q], c1[Q], c2[Q];
        if (EDGE) {
            // stage-major over the Q pixel pairs: each warp-uniform record scalar feeds Q consecutive instructions
#define SURF_EDGE(W, c)                                                                                  \
            _Pragma("unroll") for (int q = 0; q < Q; ++q) c[q] = fma2(pack2(W.x, W.x), r.dx[q], pack2(W.w, W.w)); \
            _Pragma("unroll") for (int q = 0; q < Q; ++q) c[q] = fma2(pack2(W.y, W.y), r.dy[q], c[q]);            \
            _Pragma("unroll") for (int q = 0; q < Q; ++q) c[q] = fma2(pack2(W.z, W.z), r.dz[q], c[q]);
            SURF_EDGE(W0, c0)
            SURF_EDGE(W1, c1)
            SURF_EDGE(W2, c2)
#undef SURF_EDGE
        } else {
            unsigned long long b2[Q], t2[Q], u2[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) b2[q] = mul2(pack2(A.x, A.x), r.dx[q]);
#pragma unroll
            for (int q = 0; q < Q; ++q) b2[q] = fma2(pack2(A.y, A.y), r.dy[q], b2[q]);
#pragma unroll
            for (int q = 0; q < Q; ++q) b2[q] = fma2(pack2(A.z, A.z), r.dz[q], b2[q]);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                float b0, b1;
                unpack2(b2[q], b0, b1);
                t2[q] = mul2(pack2(A.w, A.w), pack2(rcp_approx(b0), rcp_approx(b1)));
            }
#define SURF_EDGE(W, c)                                                                                  \
            _Pragma("unroll") for (int q = 0; q < Q; ++q) u2[q] = mul2(pack2(W.x, W.x), r.dx[q]);            \
            _Pragma("unroll") for (int q = 0; q < Q; ++q) u2[q] = fma2(pack2(W.y, W.y), r.dy[q], u2[q]);     \
            _Pragma("unroll") for (int q = 0; q < Q; ++q) u2[q] = fma2(pack2(W.z, W.z), r.dz[q], u2[q]);     \
            _Pragma("unroll") for (int q = 0; q < Q; ++q) c[q] = fma2(t2[q], u2[q], pack2(W.w, W.w));
            SURF_EDGE(W0, c0)
            SURF_EDGE(W1, c1)
            SURF_EDGE(W2, c2)
#undef SURF_EDGE
        }
        float mx = -INFINITY;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            float a0, a1, b0, b1, e0, e1;
            unpack2(c0[q], a0, a1);
            unpack2(c1[q], b0, b1);
            unpack2(c2[q], e0, e1);
            mx = fmaxf(mx, fmaxf(fminf(a0, fminf(b0, e0)), fminf(a1, fminf(b1, e1))));   // NaN-ignoring: conservative
        }
        if (mx_prev >= 0.f) triangle_candidates<P, EDGE>(s + 4 * (i - 1), local0 + i - 1, r, nf);
        mx_prev = mx;
        A = An; W0 = W0n; W1 = W1n; W2 = W2n;
    }
    if (mx_prev >= 0.f) triangle_candidates<P, EDGE>(s + 4 * (count - 1), local0 + count - 1, r, nf);
}

// spheres / planes with a caller-supplied narrow phase (k_intersect_shadow); the filters are chunk_spheres' / none
template <int P, class NarrowFn>
__device__ __forceinline__ void chunk_spheres_fn(const float4* __restrict__ s, int local0, int count, PixelRegs<P>& r, NarrowFn&& nf) {
    for (int i = 0; i < count; ++i) {
        const float4 S = s[i];
        bool any = false;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            Vec3 d = ray_of<P>(r, p);
            float hb = fmaf(S.z, d.z, fmaf(S.y, d.y, S.x * d.x));
            any |= fmaf(hb, hb, -S.w) >= 0.f;
        }
        if (any) {
#pragma unroll
            for (int p = 0; p < P; ++p) nf(local0 + i, S, p);
        }
    }
}
template <int P, class NarrowFn>
__device__ __forceinline__ void chunk_planes_fn(const float4* __restrict__ s, int local0, int count, NarrowFn&& nf) {
    for (int i = 0; i < count; ++i) {
#pragma unroll
        for (int p = 0; p < P; ++p) nf(local0 + i, s[i], p);
    }
}

// resident CTAs per SM the single-scene kernel is compiled for with P = 4 pixels per thread (2 = the shipped
// configuration; 3 caps the kernel at 84 registers - an A/B knob, see profiles/)
#ifndef SURF_ISECT_P4_BLOCKS
#define SURF_ISECT_P4_BLOCKS 2
#endif
template <int P, int MODE>
__global__ void __launch_bounds__(kThreads, (P == 4 ? SURF_ISECT_P4_BLOCKS : 2)) k_intersect(const __grid_constant__ IsectParams prm) {
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
        int set, local0, count;
        decode_chunk(prm, item % prm.n_chunks, &set, &local0, &count);
        const SetView& sv = prm.sc.sets[set];
        const int nf4 = rec_f4(sv.kind);
        const float4* src = prm.packed + sv.rec_off + (size_t)local0 * nf4;
        const uint32_t bytes = (uint32_t)(count * nf4) * 16u;
        mbar_expect_tx(&full_bar[stage], bytes);
        tma_bulk_g2s(stage_buf + (size_t)stage * prm.stage_f4, src, bytes, &full_bar[stage]);
    };
    if (tid == 0)
        for (int k = 0; k < kStages - 1 && lo + k < hi; ++k) issue(lo + k, k);

    const Vec3 eye = v3(prm.cam->eye[0], prm.cam->eye[1], prm.cam->eye[2]);
    const float near_clip = prm.cam->near_clip, far_clip = prm.cam->far_clip;
    constexpr int TILE = kThreads * P;

    PixelRegs<P> r;
    int cur_tile = -1;

    auto flush = [&]() {
        if (cur_tile < 0) return;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int pix = cur_tile * TILE + p * kThreads + tid;
            if (r.best_i[p] >= 0 && pix < prm.n_pix) {
                unsigned long long key = ((unsigned long long)float_order_key(r.best_t[p]) << 32) | (unsigned)r.best_i[p];
                atomicMin(prm.zbuf + pix, key);
            }
        }
    };

    for (int it = lo; it < hi; ++it) {
        const int k = it - lo;
        const int stage = k % kStages;
        const uint32_t parity = (uint32_t)((k / kStages) & 1);
        __syncthreads();   // every thread is done with item it-1, whose stage is the one refilled below
        if (tid == 0 && it + kStages - 1 < hi) issue(it + kStages - 1, (k + kStages - 1) % kStages);

        const int tile = it / prm.n_chunks;
        if (tile != cur_tile) {
            flush();
            cur_tile = tile;
            float d[3][P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const int pix = tile * TILE + p * kThreads + tid;
                const bool ok = pix < prm.n_pix;
                d[0][p] = ok ? prm.rays[pix] : 0.f;
                d[1][p] = ok ? prm.rays[(size_t)prm.n_pix + pix] : 0.f;
                d[2][p] = ok ? prm.rays[2 * (size_t)prm.n_pix + pix] : 0.f;
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
        decode_chunk(prm, it % prm.n_chunks, &set, &local0, &count);
        const SetView& sv = prm.sc.sets[set];
        mbar_wait(&full_bar[stage], parity);
        const float4* s = stage_buf + (size_t)stage * prm.stage_f4;
        if (sv.kind == KIND_DISK) chunk_disks<P, MODE>(prm, sv, s, local0, count, eye, near_clip, far_clip, r);
        else if (sv.kind == KIND_TRIANGLE) chunk_triangles<P>(prm, sv, s, local0, count, eye, near_clip, far_clip, r);
        else if (sv.kind == KIND_SPHERE) chunk_spheres<P>(prm, sv, s, local0, count, eye, near_clip, far_clip, r);
        else chunk_planes<P>(prm, sv, s, local0, count, eye, near_clip, far_clip, r);
    }
    flush();
}

// k_intersect_batch: the same kernel over a strided batch of scenes.  The work grid is [scene][tile][chunk]; all
// scenes share the chunking of scene 0, and the per-scene pointers - primitive arrays for the exact narrow phase,
// camera, rays, z-buffer - live in a shared-memory copy `sprm` of the parameter block that is re-pointed whenever the
// CTA's contiguous item range crosses into the next scene.  (Kept as a separate body: ptxas schedules the packed-FMA
// loop of k_intersect differently - 3 % slower - when the single-scene kernel is instantiated from a shared template.)
template <int P, int MODE, bool BATCH>
__device__ __forceinline__ void intersect_body(const IsectParams& prm0, const BatchArgs* ba, IsectParams* sprm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* stage_buf = reinterpret_cast<float4*>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[kStages];
    __shared__ int item_of[kStages];          // the item whose records fill each stage; -1 = no more work
    const IsectParams& prm = BATCH ? *sprm : prm0;

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // producer state (thread 0 only): the run it is walking through - first the static share, then runs of the pool
    const int n_items = prm0.n_tiles * prm0.n_chunks;
    int run_next = (int)blockIdx.x * prm0.static_per, run_end = run_next + prm0.static_per;
    auto next_item = [&]() -> int {
        if (run_next >= run_end) {
            const int u = atomicAdd(prm0.work_counter, 1);
            const long long a = (long long)prm0.dyn_begin + (long long)u * prm0.run_len;
            if (a >= n_items) return -1;
            run_next = (int)a;
            run_end = (int)min((long long)n_items, a + prm0.run_len);
        }
        return run_next++;
    };
    auto issue = [&](int stage) {      // chunking (prm0) is the same for every scene of a batch
        const int item = next_item();
        item_of[stage] = item;
        if (item < 0) {                // sentinel: complete the phase without a copy
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[stage])) : "memory");
            return;
        }
        int set, local0, count;
        decode_chunk(prm0, item % prm0.n_chunks, &set, &local0, &count);
        const SetView& sv = prm0.sc.sets[set];
        const int nf4 = rec_f4(sv.kind);
        const float4* packed = BATCH ? ws_at(prm0.packed, *ba, (item / prm0.n_chunks) / prm0.tiles_per_scene) : prm0.packed;
        const float4* src = packed + sv.rec_off + (size_t)local0 * nf4;
        const uint32_t bytes = (uint32_t)(count * nf4) * 16u;
        mbar_expect_tx(&full_bar[stage], bytes);
        tma_bulk_g2s(stage_buf + (size_t)stage * prm0.stage_f4, src, bytes, &full_bar[stage]);
    };
    if (tid == 0)
        for (int k = 0; k < kStages - 1; ++k) issue(k);

    Vec3 eye = v3(prm0.cam->eye[0], prm0.cam->eye[1], prm0.cam->eye[2]);
    const float near_clip = prm0.cam->near_clip, far_clip = prm0.cam->far_clip;    // camera scalars are batch-wide
    constexpr int TILE = kThreads * P;

    PixelRegs<P> r;
    int cur_tile = -1;       // global tile id ([scene][tile] in a batch)
    int cur_lt = 0;          // the same tile counted inside its scene
    int cur_b = -1;
    // flat pixel (relative to pix0) of this thread's slot p in tile cur_lt, or -1 if the slot is outside the frame
    const bool tiles2d = prm0.tiles_x > 0;
    const int lane = tid & 31, warp = tid >> 5;
    const int tcol = (warp & 3) * 16 + (lane & 15), trow = (warp >> 2) * 16 + (lane >> 4);
    auto pix_of = [&](int p) -> int {
        if (!tiles2d) {
            const int pix = cur_lt * TILE + p * kThreads + tid;
            return pix < prm0.n_pix ? pix : -1;
        }
        const int ty = cur_lt / prm0.tiles_x, tx = cur_lt - ty * prm0.tiles_x;
        const int col = tx * 64 + tcol;
        const int row = prm0.pix0 / prm0.W + ty * 32 + trow + 2 * p;
        const int pix = row * prm0.W + col - prm0.pix0;
        return (col < prm0.W && pix >= 0 && pix < prm0.n_pix) ? pix : -1;
    };

    auto flush = [&]() {
        if (cur_tile < 0) return;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int pix = pix_of(p);
            if (r.best_i[p] >= 0 && pix >= 0) {
                unsigned long long key = ((unsigned long long)float_order_key(r.best_t[p]) << 32) | (unsigned)r.best_i[p];
                atomicMin(prm.zbuf + pix, key);
            }
        }
    };

    for (int k = 0;; ++k) {
        const int stage = k % kStages;
        const uint32_t parity = (uint32_t)((k / kStages) & 1);
        __syncthreads();   // every thread is done with step k-1, whose stage is the one refilled below
        if (tid == 0) issue((k + kStages - 1) % kStages);
        mbar_wait(&full_bar[stage], parity);
        const int it = item_of[stage];
        if (it < 0) break;             // CTA-uniform: the sentinel reaches every thread through the same stage

        const int tile = it / prm0.n_chunks;
        if (tile != cur_tile) {
            flush();
            cur_tile = tile;
            cur_lt = tile;
            if (BATCH) {
                const int b = tile / prm0.tiles_per_scene;
                cur_lt = tile - b * prm0.tiles_per_scene;
                if (b != cur_b) {        // CTA-uniform: `tile` depends on the item index only
                    __syncthreads();     // flush() above still used the previous scene's pointers
                    if (tid == 0) {
                        *sprm = prm0;
                        scene_at(&sprm->sc, *ba, b);
                        sprm->cam = ws_at(prm0.cam, *ba, b);
                        sprm->rays = ws_at(prm0.rays, *ba, b);
                        sprm->zbuf = ws_at(prm0.zbuf, *ba, b);
                    }
                    __syncthreads();
                    cur_b = b;
                    eye = v3(prm.cam->eye[0], prm.cam->eye[1], prm.cam->eye[2]);
                }
            }
            float d[3][P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const int pix = pix_of(p);
                const bool ok = pix >= 0;
                d[0][p] = ok ? prm.rays[pix] : 0.f;
                d[1][p] = ok ? prm.rays[(size_t)prm.n_pix + pix] : 0.f;
                d[2][p] = ok ? prm.rays[2 * (size_t)prm.n_pix + pix] : 0.f;
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
        decode_chunk(prm0, it % prm0.n_chunks, &set, &local0, &count);
        const SetView& sv = prm.sc.sets[set];
        const float4* s = stage_buf + (size_t)stage * prm0.stage_f4;
        if (sv.kind == KIND_DISK) {
            if (MODE == 0)
                chunk_disks_dense<P, true>(s, local0, count, r, [&](int local, const float4& A, int p) {
                    narrow_one<P>(sv, local, A, eye, near_clip, far_clip, r, p);
                });
            else chunk_disks<P, MODE>(prm, sv, s, local0, count, eye, near_clip, far_clip, r);
        }
        else if (sv.kind == KIND_TRIANGLE) {
            if (MODE != 1)
                chunk_triangles_packed<P, true>(s, local0, count, r, [&](int local, const float4& A, int p) {
                    narrow_one<P>(sv, local, A, eye, near_clip, far_clip, r, p);
                });
            else chunk_triangles<P>(prm, sv, s, local0, count, eye, near_clip, far_clip, r);
        }
        else if (sv.kind == KIND_SPHERE) chunk_spheres<P>(prm, sv, s, local0, count, eye, near_clip, far_clip, r);
        else chunk_planes<P>(prm, sv, s, local0, count, eye, near_clip, far_clip, r);
    }
    flush();
}

template <int P, int MODE>
__global__ void __launch_bounds__(kThreads, 2) k_intersect_batch(const __grid_constant__ IsectParams prm0,
                                                                 const __grid_constant__ BatchArgs ba) {
    __shared__ IsectParams sprm;
    intersect_body<P, MODE, true>(prm0, &ba, &sprm);
}

