// surf_isect_const.cu - translation unit of libsurf_b200.so: the camera-ray x disk intersection + z-buffer stage whose
// filter records reach the FMA pipe through the UNIFORM datapath (k_filter_const and the kernels around it).
//
// Why: the packed filter of k_intersect (surf_intersect.cuh, chunk_disks) multiplies two pixel-pair registers by a
// warp-uniform scalar of the disk record.  With the record staged in shared memory that scalar arrives in a vector
// register (LDS.128), and an FFMA2 reading that register PLUS two register pairs needs three register-file cycles
// instead of two: the staged kernel tops out at 70 % of the FP32-FMA peak (tools/ubench/pipes.cu).  Blackwell's FFMA2
// also takes the scalar from a uniform register (SASS `FFMA2 R, R.F32x2.HI_LO, UR.F32, R.F32x2.HI_LO`), which costs
// no vector register-file port - but uniform registers can only be loaded from the constant bank (LDCU).  So the
// records are streamed through the 64 KB constant bank by device-to-device copies between launches: the same plane
// filter goes from 114 to 96 cycles per warp and disk (tools/ubench/uniform.cu: 70.0 % -> 83.7 % of peak).
//
// The compiler only keeps the records in uniform registers while the kernel is very plain: a subroutine call (the slow
// path of an IEEE division), a narrow phase of any size in or behind the filter loop, a loop bound that is not used
// exactly as loaded from the parameter block, a hand-written warp-aggregated append - each turned every packed FMA of
// the kernel back to vector-register scalars (bisected on the SASS; tests/test_const_path.py checks the built library).
// So the stage is split:
//   k_sphere_records sphere filter only: the 12-byte records oc / sqrt(|oc|^2 - rs^2 - slack) of the disk set.
//   k_filter_const   the conservative filter only - FILTER 1: bounding sphere, |oc' . d| >= 1, 3 FMA-pipe lane-instructions
//                    per ray-disk test; FILTER 0: the plane filter of k_intersect, 10 + a reciprocal - for all pixels of the
//                    frame x the records resident in one half of the bank.  The work grid [pixel tile][record group] is cut
//                    into equal contiguous ranges over the persistent CTAs by the host (nothing is staged, so a range may
//                    start and stop at any group).  A (half thread, group) whose filter passes is appended to a candidate
//                    queue in the workspace - about six 8-byte entries per pixel and frame on config E.
//   k_narrow_queue   one thread per candidate and pixel pair: the sphere test again, the plane filter, the exact
//                    reference-order hit test for the pairs that pass, 64-bit atomicMin into the z-buffer keys.
//   k_const_fallback tiles whose candidates did not fit the queue (dense close-ups) are flagged per launch and redone
//                    here with filter + narrow phase inline, records read from global memory; exits at once otherwise.
//   k_inside_disks   disks whose bounding sphere holds the eye (no sphere record) against every pixel; exits at once otherwise.
// Consecutive launches alternate between the halves of the bank and between two streams: the copy into one half, the cold
// constant cache and the tail of one kernel overlap the kernel on the other half.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "surf_view.h"

namespace surf {

#include "surf_runtime.cuh"
#include "surf_ptx.cuh"
#include "surf_batch.cuh"
#include "surf_launch.cuh"
#include "surf_intersect.cuh"

// One group of the bank: FILTER 0 - two 32-byte plane-filter records (16 floats); FILTER 1 - four 12-byte sphere records
// (12 floats).  The 64 KB bank holds 1024 / 1365 groups.
constexpr int group_floats(int filter) { return filter == 1 ? 12 : 16; }
constexpr int group_disks(int filter) { return filter == 1 ? 4 : 2; }
constexpr int bank_groups(int filter) { return 16384 / group_floats(filter); }
constexpr int kSlackGroups = 3;                       // the filter loop may read (and ignore) this many groups behind a segment's last
// groups per launch: the whole bank, or one half of it - then consecutive launches alternate between the halves and
// between two streams, so that the copy into one half and the ramp of the next kernel overlap the kernel on the other
constexpr int const_groups(int filter, int halves) { return bank_groups(filter) / halves - kSlackGroups; }
constexpr int kConstGrid = 3 * 160;                   // persistent grid: at most 3 CTAs per SM, 160 SMs
#ifndef SURF_CONST_SPHERE_CTAS
#define SURF_CONST_SPHERE_CTAS 2        // resident CTAs per SM of the sphere-filter kernel (3 fit at 8 pixels per thread: 72 registers)
#endif
#ifndef SURF_CONST_P
#define SURF_CONST_P 16                   // pixels per thread of the constant-bank kernels (tile = 256 x P pixels).  Measured on
                                         // config E, sphere filter: 8 -> 11.58 ms (3 CTAs per SM), 12 -> 11.36 ms, 16 -> 11.21 ms
#endif
constexpr int kConstP = SURF_CONST_P;                            // pixels per thread
__constant__ float c_recs[16384];
// records that never pass, to fill the last group of a launch.  Plane filter: n = (0, 0, 1), numer = 0, o - c = 0,
// -(r + slack)^2 = +inf (the margin is +inf or NaN for every ray); sphere filter: three times oc' = 0
__device__ float g_pad_plane[8] = {0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f, __builtin_huge_valf()};
__device__ float g_pad_sphere[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

struct ConstParams {
    const float* rays;               // [3, n]
    int n_pix, n_tiles;
    int group0;                      // set-local group index of bank group 0 (the segments count bank groups)
    uint2* queue;                    // candidates: x = 2 * thread slot (tile * kThreads + tid) + half of its pixels, y = set-local group index
    int* ctl;                        // [0] entries appended (may exceed capacity), [2] length of the `inside` list
    int capacity;
    unsigned char* flags;            // this launch's row of the [launch][tile] overflow map
    // Work of the CTAs, computed by the host: CTA b runs three segments seg[3 b .. 3 b + 2], each = tiles [x, y) against
    // the record groups [z, w) - a partial first tile, whole tiles, a partial last tile (any of them may be empty).  The
    // table lives in the parameter block (constant bank 0) and the loop bounds are used exactly as loaded: only then
    // does the compiler treat the record addresses as warp-uniform and fetch the records with LDCU into uniform
    // registers.  (Bisected on the SASS: deriving the range on the device - an integer division, or even a select between
    // a loaded bound and 0 - sends the record loads back to the vector datapath.)
    int4 seg[3 * kConstGrid];
};

// filter minima of one disk over the lower and the upper half of the thread's P pixels (a queue entry names a half, which
// halves the pixel pairs k_narrow_queue has to look at again); A, B are warp-uniform (uniform registers)
template <int P>
__device__ __forceinline__ float2 const_margin_min(const float4 A, const float4 B, const PixelRegs<P>& r, float2 m) {
    constexpr int Q = P / 2;
    const unsigned long long nx = pack2(A.x, A.x), ny = pack2(A.y, A.y), nz = pack2(A.z, A.z), nm = pack2(A.w, A.w);
    const unsigned long long ox = pack2(B.x, B.x), oy = pack2(B.y, B.y), oz = pack2(B.z, B.z), nr = pack2(B.w, B.w);
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
        if (q < Q / 2) m.x = fminf(m.x, fminf(e0, e1));     // NaN-ignoring min: NaN margins are misses
        else m.y = fminf(m.y, fminf(e0, e1));
    }
    return m;
}

// FILTER 1 - bounding sphere of the disk (centre c, radius r): a unit ray d passes within r of c iff (oc . d)^2 >= |oc|^2 - r^2,
// i.e. |oc' . d| >= 1 with oc' = oc / sqrt(|oc|^2 - r^2).  THREE FMA-pipe lane-instructions per ray-disk test (the dot product)
// and half an FMNMX3 (running maximum of |oc' . d|, NaN-ignoring) instead of the plane filter's 10 + a reciprocal; what
// passes is re-filtered per pixel by the plane filter in k_narrow_queue before the exact test.  oc' comes from
// k_sphere_records, with the slack that covers every rounding of this evaluation and of the reference-order hit test.
template <int P>
__device__ __forceinline__ float2 sphere_margin_max(float sx, float sy, float sz, const PixelRegs<P>& r, float2 m) {
    constexpr int Q = P / 2;
    const unsigned long long ox = pack2(sx, sx), oy = pack2(sy, sy), oz = pack2(sz, sz);
    unsigned long long s2[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) s2[q] = mul2(ox, r.dx[q]);
#pragma unroll
    for (int q = 0; q < Q; ++q) s2[q] = fma2(oy, r.dy[q], s2[q]);
#pragma unroll
    for (int q = 0; q < Q; ++q) s2[q] = fma2(oz, r.dz[q], s2[q]);
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        float e0, e1;
        unpack2(s2[q], e0, e1);
        if (q < Q / 2 || Q == 1) m.x = fmaxf(m.x, fmaxf(fabsf(e0), fabsf(e1)));
        else m.y = fmaxf(m.y, fmaxf(fabsf(e0), fabsf(e1)));
    }
    return m;
}

// one group of the bank, g[0 .. group_floats).  Returns two values, for the lower and the upper half of the thread's pixels:
// <= 0 iff some (pixel, disk) pair of the group passed in that half.
template <int P, int FILTER>
__device__ __forceinline__ float2 group_margin(const float* g, const PixelRegs<P>& r) {
    if (FILTER == 0) {
        const float2 m = const_margin_min<P>(make_float4(g[0], g[1], g[2], g[3]), make_float4(g[4], g[5], g[6], g[7]), r, make_float2(INFINITY, INFINITY));
        return const_margin_min<P>(make_float4(g[8], g[9], g[10], g[11]), make_float4(g[12], g[13], g[14], g[15]), r, m);
    }
    float2 m = sphere_margin_max<P>(g[0], g[1], g[2], r, make_float2(0.f, 0.f));
    m = sphere_margin_max<P>(g[3], g[4], g[5], r, m);
    m = sphere_margin_max<P>(g[6], g[7], g[8], r, m);
    m = sphere_margin_max<P>(g[9], g[10], g[11], r, m);
    return make_float2(1.f - m.x, 1.f - m.y);
}

template <int P>
__device__ __forceinline__ void load_tile_rays(const float* __restrict__ rays, int n_pix, int tile, int tid, PixelRegs<P>& r) {
    float d[3][P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int pix = tile * (kThreads * P) + p * kThreads + tid;
        const bool ok = pix < n_pix;
        d[0][p] = ok ? rays[pix] : 0.f;
        d[1][p] = ok ? rays[(size_t)n_pix + pix] : 0.f;
        d[2][p] = ok ? rays[2 * (size_t)n_pix + pix] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < P / 2; ++q) {
        r.dx[q] = pack2(d[0][2 * q], d[0][2 * q + 1]);
        r.dy[q] = pack2(d[1][2 * q], d[1][2 * q + 1]);
        r.dz[q] = pack2(d[2][2 * q], d[2][2 * q + 1]);
    }
}

// append (thread slot, group) to the candidate queue.  (Plain per-thread atomics: a warp-aggregated append - activemask,
// shuffle - inside the filter loop also made the compiler drop the uniform registers.)
__device__ __forceinline__ void push_candidate(const ConstParams& prm, int tile, int slot, int group, float2 m) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        if (!((half ? m.y : m.x) <= 0.f)) continue;
        const int at = atomicAdd(prm.ctl, 1);
        if (at < prm.capacity) prm.queue[at] = make_uint2((unsigned)(2 * slot + half), (unsigned)group);
        else prm.flags[tile] = 1;           // k_const_fallback redoes this tile against this launch's records
    }
}

template <int P, int FILTER>
__global__ void __launch_bounds__(kThreads, FILTER == 1 ? SURF_CONST_SPHERE_CTAS : 2) k_filter_const(const __grid_constant__ ConstParams prm) {
    const int tid = threadIdx.x;
    for (int sgi = 0; sgi < 3; ++sgi) {
        const int4 sg = prm.seg[3 * blockIdx.x + sgi];
        for (int tile = sg.x; tile < sg.y; ++tile) {
            PixelRegs<P> r;
            load_tile_rays<P>(prm.rays, prm.n_pix, tile, tid, r);
            // Groups of two (plane filter) or four (sphere filter) disks, two groups per iteration with the records of the
            // next group fetched (LDCU) while the current one computes - the constant cache is cold at every launch.  The
            // branch taken for group k tests a filter result that finished long ago, so neither the FMNMX3 chain nor the
            // branch resolution sits on the critical path.  The host makes every segment an even number of groups and ends
            // it one group past its last (that group's result is never tested): code behind the loop costs the uniform
            // registers, too.  A segment padded to even length tests one group of the NEXT segment (or stale bank
            // contents) as well: a duplicate or spurious candidate, which k_narrow_queue filters again - never a miss.
            constexpr int GF = group_floats(FILTER);
            float2 m_prev = make_float2(INFINITY, INFINITY);
            float ga[GF], gc[GF];
#pragma unroll
            for (int j = 0; j < GF; ++j) ga[j] = c_recs[GF * sg.z + j];
#pragma unroll 1
            for (int k = sg.z; k < sg.w; k += 2) {
#pragma unroll
                for (int j = 0; j < GF; ++j) gc[j] = c_recs[GF * (k + 1) + j];
                const float2 m = group_margin<P, FILTER>(ga, r);
                if (fminf(m_prev.x, m_prev.y) <= 0.f) push_candidate(prm, tile, tile * kThreads + tid, prm.group0 + k - 1, m_prev);
#pragma unroll
                for (int j = 0; j < GF; ++j) ga[j] = c_recs[GF * (k + 2) + j];
                const float2 m2 = group_margin<P, FILTER>(gc, r);
                if (fminf(m.x, m.y) <= 0.f) push_candidate(prm, tile, tile * kThreads + tid, prm.group0 + k, m);
                m_prev = m2;
            }
        }
    }
}

struct NarrowParams {
    SetView sv;
    const CamState* cam;
    const float4* recs;              // the set's packed filter records (global memory)
    const float* rays;
    unsigned long long* zbuf;
    int n_pix;
    const uint2* queue;
    const int* ctl;
    int capacity;
    int group_size;                  // disks per queue entry: 2 (plane filter) or 4 (sphere filter)
    const float* spheres;            // sphere filter: the set's sphere records [count, 3] (the pairs that passed are found again first)
};

// one thread per (candidate, pixel pair of the candidate's half): the two pixels' rays, the sphere test of the group's
// disks again (sphere filter), the plane filter, and the exact reference-order hit test for what passes.  Consecutive threads share a candidate, so
// the queue entry and the records are broadcast loads.
template <int P>
__global__ void __launch_bounds__(256) k_narrow_queue(const __grid_constant__ NarrowParams prm) {
    constexpr int QH = P / 4;                    // pixel pairs per candidate (one half of a thread's pixels)
    const long long n = (long long)min(prm.ctl[0], prm.capacity) * QH;
    const Vec3 eye = v3(prm.cam->eye[0], prm.cam->eye[1], prm.cam->eye[2]);
    const float near_clip = prm.cam->near_clip, far_clip = prm.cam->far_clip;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const uint2 c = prm.queue[t / QH];
        const int q = (int)(c.x & 1u) * QH + (int)(t % QH);
        const int tile = (int)((c.x >> 1) / kThreads), tid = (int)((c.x >> 1) % kThreads);
        const int pix0 = tile * (kThreads * P) + 2 * q * kThreads + tid, pix1 = pix0 + kThreads;
        const bool ok0 = pix0 < prm.n_pix, ok1 = pix1 < prm.n_pix;
        PixelRegs<2> r;
        r.dx[0] = pack2(ok0 ? prm.rays[pix0] : 0.f, ok1 ? prm.rays[pix1] : 0.f);
        r.dy[0] = pack2(ok0 ? prm.rays[(size_t)prm.n_pix + pix0] : 0.f, ok1 ? prm.rays[(size_t)prm.n_pix + pix1] : 0.f);
        r.dz[0] = pack2(ok0 ? prm.rays[2 * (size_t)prm.n_pix + pix0] : 0.f, ok1 ? prm.rays[2 * (size_t)prm.n_pix + pix1] : 0.f);
        r.best_t[0] = r.best_t[1] = INFINITY;
        r.best_i[0] = r.best_i[1] = -1;
        const int first = (int)c.y * prm.group_size, last = min(first + prm.group_size, prm.sv.count);
        for (int i = first; i < last; ++i) {
            if (prm.spheres) {
                const float* sp = prm.spheres + 3 * (size_t)i;
                if (!(sphere_margin_max<2>(sp[0], sp[1], sp[2], r, make_float2(0.f, 0.f)).x >= 1.f)) continue;
            }
            const float4 A = prm.recs[2 * i], B = prm.recs[2 * i + 1];
            float e0, e1;
            unpack2(disk_margin2<2>(A, B, r, 0), e0, e1);
            if (e0 <= 0.f) narrow_one<2>(prm.sv, i, A, eye, near_clip, far_clip, r, 0);
            if (e1 <= 0.f) narrow_one<2>(prm.sv, i, A, eye, near_clip, far_clip, r, 1);
        }
        if (r.best_i[0] >= 0 && ok0)
            atomicMin(prm.zbuf + pix0, ((unsigned long long)float_order_key(r.best_t[0]) << 32) | (unsigned)r.best_i[0]);
        if (r.best_i[1] >= 0 && ok1)
            atomicMin(prm.zbuf + pix1, ((unsigned long long)float_order_key(r.best_t[1]) << 32) | (unsigned)r.best_i[1]);
    }
}

// tiles whose candidates overflowed the queue, per launch: filter + narrow phase inline (chunk_disks on the records in
// global memory).  One CTA per flagged (launch, tile); nothing flagged (the normal case): every CTA exits at once.
struct FallbackParams {
    NarrowParams np;
    const unsigned char* flags;      // [n_launches][n_tiles]
    int n_launches, n_tiles, recs_per_launch;
};
template <int P>
__global__ void __launch_bounds__(kThreads, 2) k_const_fallback(const __grid_constant__ FallbackParams fp) {
    const NarrowParams& prm = fp.np;
    if (prm.ctl[0] <= prm.capacity) return;
    const Vec3 eye = v3(prm.cam->eye[0], prm.cam->eye[1], prm.cam->eye[2]);
    const float near_clip = prm.cam->near_clip, far_clip = prm.cam->far_clip;
    const int tid = threadIdx.x;
    for (int item = blockIdx.x; item < fp.n_launches * fp.n_tiles; item += gridDim.x) {
        if (!fp.flags[item]) continue;
        const int launch = item / fp.n_tiles, tile = item - launch * fp.n_tiles;
        PixelRegs<P> r;
        load_tile_rays<P>(prm.rays, prm.n_pix, tile, tid, r);
#pragma unroll
        for (int p = 0; p < P; ++p) { r.best_t[p] = INFINITY; r.best_i[p] = -1; }
        const int local0 = launch * fp.recs_per_launch, count = min(fp.recs_per_launch, prm.sv.count - local0);
        IsectParams unused;
        chunk_disks<P, 2>(unused, prm.sv, prm.recs + 2 * (size_t)local0, local0, count, eye, near_clip, far_clip, r);
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int pix = tile * (kThreads * P) + p * kThreads + tid;
            if (r.best_i[p] >= 0 && pix < prm.n_pix)
                atomicMin(prm.zbuf + pix, ((unsigned long long)float_order_key(r.best_t[p]) << 32) | (unsigned)r.best_i[p]);
        }
    }
}

// sphere-filter records of a disk set: oc'_i = (o - c_i) / sqrt(|o - c_i|^2 - rs^2 - slack), three floats per disk.  rs is the
// plane filter's inflated radius (prep_disk: r + 2e-6 * scale bounds what the reference-order fp32 hit test can accept); the
// distance from c to the ray's line is at most the in-plane distance the hit test measures, so a hit implies
// |oc|^2 - (oc . d^)^2 <= rs^2.  slack = 16 u |oc|^2 (u = 2^-24) covers the fp32 evaluation of |oc' . d| >= 1: rounding of oc'
// (<= 2 u relative on the square), of the three-term dot product (<= 6 u), |d|^2 = 1 +- 4 u, with a margin.  A disk whose
// inflated sphere holds the eye (nothing to divide by) gets a record that never passes and goes on the `inside` list, which
// k_inside_disks tests against every pixel.
__global__ void __launch_bounds__(256) k_sphere_records(SetView sv, const CamState* __restrict__ cam, float* __restrict__ out,
                                                        int* __restrict__ n_inside, int* __restrict__ inside) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sv.count) return;
    const float* c = sv.pos + (size_t)i * sv.pos_stride;
    const double ox = cam->eye[0], oy = cam->eye[1], oz = cam->eye[2];
    const double x = ox - (double)c[0], y = oy - (double)c[1], z = oz - (double)c[2], r = fabs((double)sv.radius[i]);
    const double oc2 = x * x + y * y + z * z;
    const double scale = sqrt(ox * ox + oy * oy + oz * oz) + sqrt((double)c[0] * c[0] + (double)c[1] * c[1] + (double)c[2] * c[2]) + sqrt(oc2) + r;
    const double rs = r + 2e-6 * scale;
    const double den = oc2 - rs * rs * (1.0 + 1e-6) - 16.0 * 5.9604644775390625e-8 * oc2;
    float* o = out + 3 * (size_t)i;
    if (!(den > 1e-30 * oc2) || !(den > 0.0)) {          // also NaN / infinite inputs
        o[0] = o[1] = o[2] = 0.f;
        inside[atomicAdd(n_inside, 1)] = i;
        return;
    }
    const double inv = 1.0 / sqrt(den);
    o[0] = (float)(x * inv); o[1] = (float)(y * inv); o[2] = (float)(z * inv);
}

// disks on the `inside` list (the eye sits inside their bounding sphere) against every pixel: plane filter + exact test
struct InsideParams {
    SetView sv;
    const CamState* cam;
    const float4* recs;
    const float* rays;
    unsigned long long* zbuf;
    int n_pix, n_tiles;
    const int* n_inside;
    const int* inside;
};
template <int P>
__global__ void __launch_bounds__(kThreads, 2) k_inside_disks(const __grid_constant__ InsideParams prm) {
    const int n = *prm.n_inside;
    if (n == 0) return;
    const Vec3 eye = v3(prm.cam->eye[0], prm.cam->eye[1], prm.cam->eye[2]);
    const float near_clip = prm.cam->near_clip, far_clip = prm.cam->far_clip;
    const int tid = threadIdx.x;
    for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x) {
        PixelRegs<P> r;
        load_tile_rays<P>(prm.rays, prm.n_pix, tile, tid, r);
#pragma unroll
        for (int p = 0; p < P; ++p) { r.best_t[p] = INFINITY; r.best_i[p] = -1; }
#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            const int i = prm.inside[j];
            const float4 A = prm.recs[2 * i], B = prm.recs[2 * i + 1];
#pragma unroll
            for (int q = 0; q < P / 2; ++q) {
                float e0, e1;
                unpack2(disk_margin2<P>(A, B, r, q), e0, e1);
                if (e0 <= 0.f) narrow_one<P>(prm.sv, i, A, eye, near_clip, far_clip, r, 2 * q);
                if (e1 <= 0.f) narrow_one<P>(prm.sv, i, A, eye, near_clip, far_clip, r, 2 * q + 1);
            }
        }
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int pix = tile * (kThreads * P) + p * kThreads + tid;
            if (r.best_i[p] >= 0 && pix < prm.n_pix)
                atomicMin(prm.zbuf + pix, ((unsigned long long)float_order_key(r.best_t[p]) << 32) | (unsigned)r.best_i[p]);
        }
    }
}

// The constant bank is one per device: launches that use it are serialised across streams with an event (a stream
// under graph capture skips that: the captured work is ordered inside its graph).
struct ConstBankState {
    std::mutex mu;
    cudaEvent_t done[64] = {}, fork[64] = {}, join[64][3] = {};
    cudaStream_t side[64][3] = {};
    bool used[64] = {};
};
static ConstBankState g_bank;

bool const_path_fits(const Frame& f, const SetView& sv) {
    const int n_tiles = (f.n + kThreads * kConstP - 1) / (kThreads * kConstP);
    const long long n_launches = (sv.count + 2 * const_groups(0, 4) - 1) / (2 * const_groups(0, 4));
    return f.ws.cq != nullptr && f.ws.cq_inside != nullptr && n_launches * n_tiles <= (long long)f.ws.cq_flag_bytes;
}

int run_intersect_const(const Frame& f, const SetView& sv, int filter, cudaStream_t st) {
    if (sv.kind != KIND_DISK) return fail(SURF_ERR_BAD_ARG, "the constant-bank kernel takes disk sets");
    constexpr int P = kConstP;
    if (!const_path_fits(f, sv)) return fail(SURF_ERR_WORKSPACE, "workspace holds no candidate queue for this frame");
    ConstParams prm;
    prm.rays = f.ws.rays; prm.n_pix = f.n;
    prm.n_tiles = (f.n + kThreads * P - 1) / (kThreads * P);
    prm.queue = f.ws.cq; prm.ctl = f.ws.cq_ctl; prm.capacity = f.ws.cq_capacity;
    if (const char* cap_env = getenv("SURF_CONST_CAPACITY"))      // tests: a tiny queue sends tiles through k_const_fallback
        prm.capacity = std::max(0, std::min(prm.capacity, atoi(cap_env)));
    // small frames (row bands of a multi-GPU step): half-bank launches alternating between two streams hide the
    // kernel -> copy -> kernel bubbles (about 8 us per launch, 8 % of a 1/8-frame launch); large frames: whole-bank launches
    static const int halves_env = getenv("SURF_CONST_HALVES") ? atoi(getenv("SURF_CONST_HALVES")) : 0;      // tuning knob
    const int halves = halves_env == 1 || halves_env == 2 || halves_env == 4 ? halves_env : 2;
    // filter 0: 32-byte plane records, two per group; filter 1: 12-byte sphere records (computed here), four per group
    const int group_size = group_disks(filter);
    const size_t rec_bytes = filter == 1 ? 12 : 32, group_bytes = rec_bytes * group_size;
    const int per_launch = const_groups(filter, halves) * group_size;
    const int n_launches = (sv.count + per_launch - 1) / per_launch;
    int dev = 0;
    SURF_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(SURF_ERR_UNSUPPORTED, "device ordinal beyond 63");
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    SURF_CUDA(cudaStreamIsCapturing(st, &cap));
    const bool ordered = cap == cudaStreamCaptureStatusNone;
    std::lock_guard<std::mutex> lock(g_bank.mu);
    if (!g_bank.done[dev]) {
        SURF_CUDA(cudaEventCreateWithFlags(&g_bank.done[dev], cudaEventDisableTiming));
        SURF_CUDA(cudaEventCreateWithFlags(&g_bank.fork[dev], cudaEventDisableTiming));
        for (int k = 0; k < 3; ++k) {
            SURF_CUDA(cudaEventCreateWithFlags(&g_bank.join[dev][k], cudaEventDisableTiming));
            SURF_CUDA(cudaStreamCreateWithFlags(&g_bank.side[dev][k], cudaStreamNonBlocking));
        }
    }
    if (ordered && g_bank.used[dev]) SURF_CUDA(cudaStreamWaitEvent(st, g_bank.done[dev], 0));
    const float4* recs = f.ws.packed + sv.rec_off;
    const char* bank_src = (const char*)recs;
    if (filter == 1) {
        bank_src = (const char*)((float*)f.ws.circ + 3 * (size_t)sv.first);
    }
    void* pad = nullptr;
    if (filter == 1) SURF_CUDA(cudaGetSymbolAddress(&pad, g_pad_sphere));
    else SURF_CUDA(cudaGetSymbolAddress(&pad, g_pad_plane));
    static const int per_sm_env = getenv("SURF_CONST_CTAS") ? atoi(getenv("SURF_CONST_CTAS")) : 0;      // tuning knob
    const int per_sm = per_sm_env > 0 ? per_sm_env : (filter == 1 ? SURF_CONST_SPHERE_CTAS : 2);
    const int grid_max = std::min(sm_count() * per_sm, kConstGrid);
    timer_mark(0, 0, st);
    // control words + the overflow map in one memset (adjacent in the workspace)
    SURF_CUDA(cudaMemsetAsync(f.ws.cq_ctl, 0, 256 + (size_t)n_launches * prm.n_tiles, st));
    if (filter == 1) {
        k_sphere_records<<<(sv.count + 255) / 256, 256, 0, st>>>(sv, f.ws.cam, (float*)f.ws.circ + 3 * (size_t)sv.first, f.ws.cq_ctl + 2,
                                                                  f.ws.cq_inside);
        SURF_LAUNCHED("k_sphere_records");
    }
    const int n_streams = std::min(halves, n_launches);
    if (n_streams > 1) {
        SURF_CUDA(cudaEventRecord(g_bank.fork[dev], st));
        for (int k = 0; k + 1 < n_streams; ++k) SURF_CUDA(cudaStreamWaitEvent(g_bank.side[dev][k], g_bank.fork[dev], 0));
    }
    for (int j = 0; j < n_launches; ++j) {
        const int first = j * per_launch;
        const int count = std::min(per_launch, sv.count - first);
        const int n_groups = (count + group_size - 1) / group_size;
        const int part = j % halves;
        const int bank0 = part * (bank_groups(filter) / halves);                   // first bank group of this launch
        cudaStream_t s = part == 0 || n_streams == 1 ? st : g_bank.side[dev][part - 1];
        prm.group0 = first / group_size - bank0;
        prm.flags = f.ws.cq_flags + (size_t)j * prm.n_tiles;
        SURF_CUDA(cudaMemcpyToSymbolAsync(c_recs, bank_src + (size_t)first * rec_bytes, (size_t)count * rec_bytes, (size_t)bank0 * group_bytes,
                                          cudaMemcpyDeviceToDevice, s));
        if (count % group_size)      // fill the last group with records that never pass
            SURF_CUDA(cudaMemcpyToSymbolAsync(c_recs, pad, (size_t)(group_size - count % group_size) * rec_bytes,
                                              (size_t)bank0 * group_bytes + (size_t)count * rec_bytes, cudaMemcpyDeviceToDevice, s));
        const long long units = (long long)prm.n_tiles * n_groups;
        const int grid = (int)std::min<long long>(units, grid_max);
        for (int b = 0; b < grid; ++b) {
            const long long lo = units * b / grid, hi = units * (b + 1) / grid;       // hi > lo: grid <= units
            const int t0 = (int)(lo / n_groups), t1 = (int)((hi - 1) / n_groups);      // first and last tile touched
            const int g0 = (int)(lo - (long long)t0 * n_groups), g1 = (int)(hi - (long long)t1 * n_groups);
            int4* sg = prm.seg + 3 * b;
            // group ranges count bank groups, end one past the last group and hold an even number of groups: see the
            // loop of k_filter_const
            auto seg = [bank0](int ta, int tb, int ga, int gb) { return make_int4(ta, tb, bank0 + ga, bank0 + gb + 1 + ((gb + 1 - ga) & 1)); };
            if (t0 == t1) { sg[0] = seg(t0, t0 + 1, g0, g1); sg[1] = sg[2] = make_int4(0, 0, 0, 0); }
            else { sg[0] = seg(t0, t0 + 1, g0, n_groups); sg[1] = seg(t0 + 1, t1, 0, n_groups); sg[2] = seg(t1, t1 + 1, 0, g1); }
        }
        if (filter == 1) k_filter_const<P, 1><<<grid, kThreads, 0, s>>>(prm);
        else k_filter_const<P, 0><<<grid, kThreads, 0, s>>>(prm);
        SURF_LAUNCHED("k_filter_const");
    }
    for (int k = 0; k + 1 < n_streams; ++k) {
        SURF_CUDA(cudaEventRecord(g_bank.join[dev][k], g_bank.side[dev][k]));
        SURF_CUDA(cudaStreamWaitEvent(st, g_bank.join[dev][k], 0));
    }
    if (ordered) { SURF_CUDA(cudaEventRecord(g_bank.done[dev], st)); g_bank.used[dev] = true; }
    FallbackParams fp;
    NarrowParams& np = fp.np;
    np.sv = sv; np.cam = f.ws.cam; np.recs = recs; np.rays = f.ws.rays; np.zbuf = f.ws.zbuf; np.n_pix = f.n;
    np.queue = f.ws.cq; np.ctl = f.ws.cq_ctl; np.capacity = prm.capacity; np.group_size = group_size; np.spheres = filter == 1 ? (const float*)f.ws.circ + 3 * (size_t)sv.first : nullptr;
    fp.flags = f.ws.cq_flags; fp.n_launches = n_launches; fp.n_tiles = prm.n_tiles; fp.recs_per_launch = per_launch;
    k_narrow_queue<P><<<sm_count() * 16, 256, 0, st>>>(np);
    SURF_LAUNCHED("k_narrow_queue");
    k_const_fallback<P><<<sm_count() * 2, kThreads, 0, st>>>(fp);
    SURF_LAUNCHED("k_const_fallback");
    if (filter == 1) {
        InsideParams ip;
        ip.sv = sv; ip.cam = f.ws.cam; ip.recs = recs; ip.rays = f.ws.rays; ip.zbuf = f.ws.zbuf; ip.n_pix = f.n; ip.n_tiles = prm.n_tiles;
        ip.n_inside = f.ws.cq_ctl + 2; ip.inside = f.ws.cq_inside;
        k_inside_disks<P><<<std::min(prm.n_tiles, sm_count() * 2), kThreads, 0, st>>>(ip);
        SURF_LAUNCHED("k_inside_disks");
    }
    timer_mark(0, 1, st);
    static const bool debug = getenv("SURF_CONST_DEBUG") != nullptr;        // prints the queue fill; synchronises
    if (debug && ordered) {
        int ctl[2] = {0, 0};
        SURF_CUDA(cudaMemcpyAsync(ctl, f.ws.cq_ctl, sizeof(ctl), cudaMemcpyDeviceToHost, st));
        SURF_CUDA(cudaStreamSynchronize(st));
        fprintf(stderr, "[surf const] filter %d: %d pixels, %d disks, %d launches, %d candidates (capacity %d)\n", filter, f.n, sv.count,
                n_launches, ctl[0], f.ws.cq_capacity);
    }
    return SURF_OK;
}

}  // namespace surf
