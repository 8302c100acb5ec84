// surf_isect_const.cu - translation unit of libsurf_b200.so: k_intersect_const<P>, the camera-ray x disk intersection +
// z-buffer kernel whose filter records reach the FMA pipe through the UNIFORM datapath.
//
// Why: the packed filter of k_intersect (surf_intersect.cuh, chunk_disks) multiplies two pixel-pair registers by a
// warp-uniform scalar of the disk record.  With the record staged in shared memory that scalar arrives in a vector
// register (LDS.128), and an FFMA2 reading that register PLUS two register pairs needs three register-file cycles
// instead of two: the staged kernel tops out at 70 % of the FP32-FMA peak (tools/ubench/pipes.cu).  Blackwell's FFMA2
// also takes the scalar from a uniform register (SASS `FFMA2 R, R.F32x2.HI_LO, UR.F32, R.F32x2.HI_LO`), which costs
// no vector register-file port - but uniform registers can only be loaded from the constant bank (LDCU).  So the
// records are streamed through the 64 KB constant bank, 2048 disks at a time, by device-to-device copies between
// launches: same arithmetic, same filter + exact narrow phase, 114 -> 96 cycles per warp and disk
// (tools/ubench/uniform.cu: 70.0 % -> 83.7 % of peak).
//
// The compiler only keeps the records in uniform registers while the kernel is simple: a subroutine call (the slow path
// of an IEEE division) or a narrow phase of any size inside the filter loop turned all 200 FFMA2 of the kernel back to
// vector-register scalars (bisected on the SASS).  So the stage is split:
//   k_filter_const   the conservative filter only: all pixels of the frame x the <= 2048 records resident in the bank.
//                    The work grid [pixel tile][record group] is cut into equal contiguous ranges over the persistent CTAs
//                    (nothing is staged, so a range may start and stop at any group).  A (thread, group) whose filter
//                    minimum passes is appended to a candidate queue in the workspace (warp-aggregated atomic) - about one
//                    8-byte entry per pixel and frame on config E.
//   k_narrow_queue   one thread per candidate: per-pixel filter of the group's two disks, the exact reference-order hit
//                    test for the pairs that pass, 64-bit atomicMin into the z-buffer keys.
//   k_const_fallback tiles whose candidates did not fit the queue (dense close-ups) are flagged per launch and redone
//                    here with filter + narrow phase inline, records read from global memory; exits at once otherwise.
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

constexpr int kConstRecs = 2042;                      // 32-byte disk records per launch: the 64 KB constant bank less three groups
                                                      // that the filter loop may read (and ignore) behind a segment's last group
constexpr int kConstGrid = 3 * 160;                   // persistent grid: at most 3 CTAs per SM, 160 SMs
constexpr int kConstP = 8;                            // pixels per thread
__constant__ float4 c_recs[2 * (kConstRecs + 6)];
// pads an odd record count: n = (0, 0, 1), numer = 0, o - c = 0, -(r + slack)^2 = +inf - the margin is +inf (or NaN) for every ray
__device__ float4 g_pad_record[2] = {{0.f, 0.f, 1.f, 0.f}, {0.f, 0.f, 0.f, __builtin_huge_valf()}};

struct ConstParams {
    const float* rays;               // [3, n]
    int n_pix, n_tiles;
    int group0, n_groups;            // the records of this launch: set-local index / 2 of c_recs[0]; pairs of records in the
                                     // bank (an odd count is padded with a record that never passes)
    uint2* queue;                    // candidates: x = thread slot (tile * kThreads + tid), y = set-local group index
    int* ctl;                        // [0] entries appended (may exceed capacity), [1] flagged (launch, tile) pairs
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

// filter minimum of one disk over the P pixels of the thread; A, B are warp-uniform (uniform registers)
template <int P>
__device__ __forceinline__ float const_margin_min(const float4 A, const float4 B, const PixelRegs<P>& r, float m) {
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
        m = fminf(m, fminf(e0, e1));     // NaN-ignoring min: NaN margins are misses
    }
    return m;
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
__device__ __forceinline__ void push_candidate(const ConstParams& prm, int tile, int slot, int group) {
    const int at = atomicAdd(prm.ctl, 1);
    if (at < prm.capacity) prm.queue[at] = make_uint2((unsigned)slot, (unsigned)group);
    else prm.flags[tile] = 1;           // k_const_fallback redoes this tile against this launch's records
}

template <int P>
__global__ void __launch_bounds__(kThreads, 2) k_filter_const(const __grid_constant__ ConstParams prm) {
    const int tid = threadIdx.x;
    for (int sgi = 0; sgi < 3; ++sgi) {
        const int4 sg = prm.seg[3 * blockIdx.x + sgi];
        for (int tile = sg.x; tile < sg.y; ++tile) {
            PixelRegs<P> r;
            load_tile_rays<P>(prm.rays, prm.n_pix, tile, tid, r);
            // Groups of two disks, two groups per iteration with the records of the next group fetched (LDCU) while the
            // current one computes - the constant cache is cold at every launch.  The branch taken for group k tests a
            // filter minimum that finished long ago, so neither the FMNMX3 chain nor the branch resolution sits on the
            // critical path.  The host makes every segment an even number of groups and ends it one group past its last
            // (that group's minimum is never tested): code behind the loop costs the uniform registers, too.
            float m_prev = INFINITY;
            float4 a0 = c_recs[4 * sg.z], b0 = c_recs[4 * sg.z + 1], a1 = c_recs[4 * sg.z + 2], b1 = c_recs[4 * sg.z + 3];
#pragma unroll 1
            for (int k = sg.z; k < sg.w; k += 2) {
                const float4 c0 = c_recs[4 * k + 4], d0 = c_recs[4 * k + 5], c1 = c_recs[4 * k + 6], d1 = c_recs[4 * k + 7];
                float m = const_margin_min<P>(a0, b0, r, INFINITY);
                m = const_margin_min<P>(a1, b1, r, m);
                if (m_prev <= 0.f) push_candidate(prm, tile, tile * kThreads + tid, prm.group0 + k - 1);
                a0 = c_recs[4 * k + 8]; b0 = c_recs[4 * k + 9]; a1 = c_recs[4 * k + 10]; b1 = c_recs[4 * k + 11];
                float m2 = const_margin_min<P>(c0, d0, r, INFINITY);
                m2 = const_margin_min<P>(c1, d1, r, m2);
                if (m <= 0.f) push_candidate(prm, tile, tile * kThreads + tid, prm.group0 + k);
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
};

// one thread per candidate: per-pixel filter of the group's disks, exact narrow phase for the pairs that pass
template <int P>
__global__ void __launch_bounds__(256) k_narrow_queue(const __grid_constant__ NarrowParams prm) {
    const int n = min(prm.ctl[0], prm.capacity);
    const Vec3 eye = v3(prm.cam->eye[0], prm.cam->eye[1], prm.cam->eye[2]);
    const float near_clip = prm.cam->near_clip, far_clip = prm.cam->far_clip;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const uint2 c = prm.queue[e];
        const int tile = (int)(c.x / kThreads), tid = (int)(c.x % kThreads);
        PixelRegs<P> r;
        load_tile_rays<P>(prm.rays, prm.n_pix, tile, tid, r);
#pragma unroll
        for (int p = 0; p < P; ++p) { r.best_t[p] = INFINITY; r.best_i[p] = -1; }
        const int first = (int)c.y * 2, last = min(first + 2, prm.sv.count);
        for (int i = first; i < last; ++i) {
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

// tiles whose candidates overflowed the queue, per launch: filter + narrow phase inline (chunk_disks on the records in
// global memory).  One CTA per flagged (launch, tile); nothing flagged (the normal case): every CTA exits at once.
struct FallbackParams {
    NarrowParams np;
    const unsigned char* flags;      // [n_launches][n_tiles]
    int n_launches, n_tiles;
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
        const int local0 = launch * kConstRecs, count = min(kConstRecs, prm.sv.count - local0);
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

// The constant bank is one per device: launches that use it are serialised across streams with an event (a stream
// under graph capture skips that: the captured work is ordered inside its graph).
struct ConstBankState { std::mutex mu; cudaEvent_t done[64] = {}; bool used[64] = {}; };
static ConstBankState g_bank;

bool const_path_fits(const Frame& f, const SetView& sv) {
    const int n_tiles = (f.n + kThreads * kConstP - 1) / (kThreads * kConstP);
    const long long n_launches = (sv.count + kConstRecs - 1) / kConstRecs;
    return f.ws.cq != nullptr && n_launches * n_tiles <= (long long)f.ws.cq_flag_bytes;
}

int run_intersect_const(const Frame& f, const SetView& sv, cudaStream_t st) {
    if (sv.kind != KIND_DISK) return fail(SURF_ERR_BAD_ARG, "the constant-bank kernel takes disk sets");
    constexpr int P = kConstP;
    if (!const_path_fits(f, sv)) return fail(SURF_ERR_WORKSPACE, "workspace holds no candidate queue for this frame");
    ConstParams prm;
    prm.rays = f.ws.rays; prm.n_pix = f.n;
    prm.n_tiles = (f.n + kThreads * P - 1) / (kThreads * P);
    prm.queue = f.ws.cq; prm.ctl = f.ws.cq_ctl; prm.capacity = f.ws.cq_capacity;
    const int n_launches = (sv.count + kConstRecs - 1) / kConstRecs;
    int dev = 0;
    SURF_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(SURF_ERR_UNSUPPORTED, "device ordinal beyond 63");
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    SURF_CUDA(cudaStreamIsCapturing(st, &cap));
    const bool ordered = cap == cudaStreamCaptureStatusNone;
    std::lock_guard<std::mutex> lock(g_bank.mu);
    if (ordered) {
        if (!g_bank.done[dev]) SURF_CUDA(cudaEventCreateWithFlags(&g_bank.done[dev], cudaEventDisableTiming));
        if (g_bank.used[dev]) SURF_CUDA(cudaStreamWaitEvent(st, g_bank.done[dev], 0));
    }
    const float4* recs = f.ws.packed + sv.rec_off;
    void* pad = nullptr;
    SURF_CUDA(cudaGetSymbolAddress(&pad, g_pad_record));
    static const int per_sm = getenv("SURF_CONST_CTAS") ? atoi(getenv("SURF_CONST_CTAS")) : 2;      // tuning knob
    const int grid_max = std::min(sm_count() * per_sm, kConstGrid);
    timer_mark(0, 0, st);
    // control words + the overflow map in one memset (adjacent in the workspace)
    SURF_CUDA(cudaMemsetAsync(f.ws.cq_ctl, 0, 256 + (size_t)n_launches * prm.n_tiles, st));
    for (int j = 0; j < n_launches; ++j) {
        const int first = j * kConstRecs;
        prm.group0 = first / 2;
        const int count = std::min(kConstRecs, sv.count - first);
        prm.n_groups = (count + 1) / 2;
        prm.flags = f.ws.cq_flags + (size_t)j * prm.n_tiles;
        SURF_CUDA(cudaMemcpyToSymbolAsync(c_recs, recs + 2 * (size_t)first, (size_t)count * 32, 0, cudaMemcpyDeviceToDevice, st));
        if (count & 1) SURF_CUDA(cudaMemcpyToSymbolAsync(c_recs, pad, 32, (size_t)count * 32, cudaMemcpyDeviceToDevice, st));
        const int n_groups = prm.n_groups;
        const long long units = (long long)prm.n_tiles * n_groups;
        const int grid = (int)std::min<long long>(units, grid_max);
        for (int b = 0; b < grid; ++b) {
            const long long lo = units * b / grid, hi = units * (b + 1) / grid;       // hi > lo: grid <= units
            const int t0 = (int)(lo / n_groups), t1 = (int)((hi - 1) / n_groups);      // first and last tile touched
            const int g0 = (int)(lo - (long long)t0 * n_groups), g1 = (int)(hi - (long long)t1 * n_groups);
            int4* sg = prm.seg + 3 * b;
            // group ranges end one past the last group and hold an even number of groups: see the loop of k_filter_const
            auto seg = [](int ta, int tb, int ga, int gb) { return make_int4(ta, tb, ga, gb + 1 + ((gb + 1 - ga) & 1)); };
            if (t0 == t1) { sg[0] = seg(t0, t0 + 1, g0, g1); sg[1] = sg[2] = make_int4(0, 0, 0, 0); }
            else { sg[0] = seg(t0, t0 + 1, g0, n_groups); sg[1] = seg(t0 + 1, t1, 0, n_groups); sg[2] = seg(t1, t1 + 1, 0, g1); }
        }
        k_filter_const<P><<<grid, kThreads, 0, st>>>(prm);
        SURF_LAUNCHED("k_filter_const");
    }
    if (ordered) { SURF_CUDA(cudaEventRecord(g_bank.done[dev], st)); g_bank.used[dev] = true; }
    FallbackParams fp;
    NarrowParams& np = fp.np;
    np.sv = sv; np.cam = f.ws.cam; np.recs = recs; np.rays = f.ws.rays; np.zbuf = f.ws.zbuf; np.n_pix = f.n;
    np.queue = f.ws.cq; np.ctl = f.ws.cq_ctl; np.capacity = f.ws.cq_capacity;
    fp.flags = f.ws.cq_flags; fp.n_launches = n_launches; fp.n_tiles = prm.n_tiles;
    k_narrow_queue<P><<<sm_count() * 8, 256, 0, st>>>(np);
    SURF_LAUNCHED("k_narrow_queue");
    k_const_fallback<P><<<sm_count() * 2, kThreads, 0, st>>>(fp);
    SURF_LAUNCHED("k_const_fallback");
    timer_mark(0, 1, st);
    return SURF_OK;
}

}  // namespace surf
