// surf_kernels.cu - sm_100a kernels + the C ABI of libsurf_b200.so (include/surf_b200.h).
//
// Pipeline of one forward frame (all on the caller's stream, no host sync):
//   k_setup      1 thread    camera basis / image-plane scale -> CamState            (utils.py:402-456)
//   k_prep       O(M)        per-primitive exact plane constants + conservative filter records, packed
//                            contiguously per primitive set for TMA bulk staging       (utils.py:288-297)
//   k_raygen     O(N)        unit ray directions [3,N] + z-buffer key init             (utils.py:439-478)
//   k_intersect  O(M*N)      THE hot kernel: persistent CTAs, primitives streamed global->shared with
//                            cp.async.bulk + mbarrier (3-stage ring), P pixels per thread in registers,
//                            packed FFMA2 (fma.rn.f32x2) conservative disk filter, exact reference-order
//                            narrow phase on the rare candidates, per-pixel (depth,index) merged with a
//                            64-bit atomicMin so the [M,N] distance tensor never exists  (utils.py:481-512,
//                            renderer.py:170-189)
//   k_shade      O(N*L)      winner -> depth/nearest/pos/normal + Phong shading + tonemap (renderer.py:82-125,
//                            :266-340)
// Backward: k_backward O(N*L) recomputes the winning hit per pixel (surf_math.cuh backward_pixel), reduces
// light/material/colour gradients across the warp and CTA, primitive gradients with a warp-segmented
// reduction keyed by `nearest` before red.global.add; k_backward_finalize folds the double accumulators.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "surf_view.h"

namespace surf {

// ---------------------------------------------------------------------------------------------------
// error handling (thread-local string) / launch accounting and optional timers (process-wide counters)
// ---------------------------------------------------------------------------------------------------
static thread_local std::string g_error;
static int g_launches = 0;   // process-wide: autograd runs backward on its own thread

// optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg): a ring of
// event pairs per kernel kind so that a whole timed region can be averaged without synchronising inside it
constexpr int kTimerRing = 256;
struct KernelTimers {
    bool enabled = false;
    bool created = false;
    cudaEvent_t ev[3][kTimerRing][2] = {};
    long long count[3] = {0, 0, 0};     // launches recorded since timing was (re-)enabled
};
static KernelTimers g_timers;   // process-wide (see g_launches)
static void timer_mark(int which, int edge, cudaStream_t st) {
    if (!g_timers.enabled) return;
    if (!g_timers.created) {
        for (int k = 0; k < 3; ++k)
            for (int r = 0; r < kTimerRing; ++r)
                for (int e = 0; e < 2; ++e) cudaEventCreate(&g_timers.ev[k][r][e]);
        g_timers.created = true;
    }
    const int slot = (int)(g_timers.count[which] % kTimerRing);
    cudaEventRecord(g_timers.ev[which][slot][edge], st);
    if (edge == 1) ++g_timers.count[which];
}
static double timer_ms(int which, long long index) {
    const int slot = (int)(index % kTimerRing);
    if (cudaEventSynchronize(g_timers.ev[which][slot][1]) != cudaSuccess) return -1.0;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_timers.ev[which][slot][0], g_timers.ev[which][slot][1]) != cudaSuccess) return -1.0;
    return (double)ms;
}

static int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}
static int cuda_fail(cudaError_t e, const char* where) {
    g_error = std::string(where) + ": " + cudaGetErrorString(e);
    return SURF_ERR_CUDA;
}
#define SURF_CUDA(call)                                     \
    do {                                                    \
        cudaError_t e_ = (call);                            \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
    } while (0)
#define SURF_LAUNCHED(name)                                      \
    do {                                                         \
        ++g_launches;                                            \
        cudaError_t e_ = cudaPeekAtLastError();                  \
        if (e_ != cudaSuccess) return cuda_fail(e_, name);       \
    } while (0)

// ---------------------------------------------------------------------------------------------------
// workspace layout
// ---------------------------------------------------------------------------------------------------
struct Workspace {
    CamState* cam;
    float4* packed;              // plane-filter records, per set, 128-byte aligned (math_mode 1..3)
    float4* circ;                // level-1 screen-circle records, one float4 per primitive in global order
    float* rays;                 // [3, n] SoA unit directions (perspective)
    unsigned long long* zbuf;    // [n] packed (depth key << 32 | primitive index)
    double* acc;                 // backward scalar accumulators
    double* prim_acc;            // backward per-primitive accumulators [total_prims, 7]
    float* vis;                  // [L, n] shadow visibility
    float* gray;                 // [7, n] generic rays: origin xyz, direction xyz, t_max (orthographic / shadow rays)
    unsigned long long* zbuf2;   // [n] z-buffer keys of the shadow rays
    float* obound;               // [1] max |origin| over the generic rays of the launch (as float bits, atomicMax)
    size_t bytes;
};
constexpr int kMaxAccSlots = 512;

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t packed_bytes_bound(int total_prims) {
    // worst case: all triangles (4 float4 each) + 8 sets x 128-byte padding
    return (size_t)total_prims * 64 + kMaxSets * 128 + 256;
}

static void carve(void* base, int total_prims, int n_pix, int n_lights, bool shadow, Workspace* ws) {
    char* p = (char*)base;
    size_t off = 0;
    ws->cam = (CamState*)(p + off); off += align_up(sizeof(CamState), 256);
    ws->packed = (float4*)(p + off); off += align_up(packed_bytes_bound(total_prims), 256);
    ws->circ = (float4*)(p + off); off += align_up((size_t)total_prims * 16 + 256, 256);
    ws->rays = (float*)(p + off); off += align_up((size_t)3 * n_pix * sizeof(float), 256);
    ws->zbuf = (unsigned long long*)(p + off); off += align_up((size_t)n_pix * 8, 256);
    ws->acc = (double*)(p + off); off += align_up((size_t)kMaxAccSlots * 8, 256);
    ws->prim_acc = (double*)(p + off); off += align_up((size_t)total_prims * 7 * 8, 256);
    ws->vis = (float*)(p + off);
    if (shadow) off += align_up((size_t)n_lights * n_pix * sizeof(float), 256);
    // generic-ray buffers: always carved (orthographic frames need them too); 36 B per pixel
    ws->gray = (float*)(p + off); off += align_up((size_t)7 * n_pix * sizeof(float), 256);
    ws->zbuf2 = (unsigned long long*)(p + off); off += align_up((size_t)n_pix * 8, 256);
    ws->obound = (float*)(p + off); off += 256;
    ws->bytes = off;
}

// ---------------------------------------------------------------------------------------------------
// small PTX helpers: mbarrier, TMA bulk copy, packed f32x2 math
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// Blackwell packed fp32: one instruction, two lane-FMAs (SASS: FFMA2 / FMUL2)
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ---------------------------------------------------------------------------------------------------
// k_setup / k_prep / k_raygen
// ---------------------------------------------------------------------------------------------------
struct CamArgs {
    const float* eye; const float* at; const float* up;
    int proj, W, H;
    double fovy, focal;
    float near_clip, far_clip;
};

__global__ void k_setup(CamArgs a, CamState* cs) {
    if (threadIdx.x == 0 && blockIdx.x == 0)
        camera_setup(a.eye, a.at, a.up, a.proj, a.W, a.H, a.fovy, a.focal, a.near_clip, a.far_clip, cs);
}

__global__ void __launch_bounds__(256) k_prep(const __grid_constant__ SceneView sc, const CamState* __restrict__ cs,
                                              float4* __restrict__ packed) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= sc.total) return;
    const int s = find_set(sc, g);
    const SetView& sv = sc.sets[s];
    const int i = g - sv.first;
    const Vec3 o = v3(cs->eye[0], cs->eye[1], cs->eye[2]);
    F4 r[4];
    if (sv.kind == KIND_DISK) {
        prep_disk(ld3(sv.pos + (size_t)i * sv.pos_stride), ld3(sv.normal + (size_t)i * sv.normal_stride), sv.radius[i], o,
                  &r[0], &r[1]);
    } else if (sv.kind == KIND_PLANE) {
        prep_plane(ld3(sv.pos + (size_t)i * sv.pos_stride), ld3(sv.normal + (size_t)i * sv.normal_stride), o, &r[0]);
    } else if (sv.kind == KIND_SPHERE) {
        prep_sphere(ld3(sv.pos + (size_t)i * sv.pos_stride), sv.radius[i], o, &r[0]);
    } else {
        const float* f = sv.pos + (size_t)i * 3 * sv.pos_stride;
        prep_triangle(ld3(f), ld3(f + sv.pos_stride), ld3(f + 2 * sv.pos_stride),
                      ld3(sv.normal + (size_t)i * sv.normal_stride), o, &r[0], &r[1], &r[2], &r[3]);
    }
    const int nf4 = rec_f4(sv.kind);
    float4* dst = packed + sv.rec_off + (size_t)i * nf4;
    for (int k = 0; k < nf4; ++k) dst[k] = make_float4(r[k].x, r[k].y, r[k].z, r[k].w);
}

__global__ void __launch_bounds__(256) k_raygen(const CamState* __restrict__ cs, int pix0, int n,
                                                float* __restrict__ rays, float* __restrict__ ray_out,
                                                unsigned long long* __restrict__ zbuf) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    zbuf[k] = kMissKey;
    if (cs->proj == 0) {
        Vec3 d = pixel_ray_dir(*cs, pix0 + k);
        rays[k] = d.x; rays[(size_t)n + k] = d.y; rays[2 * (size_t)n + k] = d.z;
        if (ray_out) { ray_out[k] = d.x; ray_out[(size_t)n + k] = d.y; ray_out[2 * (size_t)n + k] = d.z; }
    } else if (k == 0 && ray_out) {
        ray_out[0] = cs->odir[0]; ray_out[1] = cs->odir[1]; ray_out[2] = cs->odir[2];
    }
}

__global__ void __launch_bounds__(256) k_prep_screen(const __grid_constant__ SceneView sc, const CamState* __restrict__ cs,
                                                     float4* __restrict__ circ) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= sc.total) return;
    const int s = find_set(sc, g);
    const F4 r = prep_screen(*cs, sc.sets[s], g - sc.sets[s].first);
    circ[g] = make_float4(r.x, r.y, r.z, r.w);
}

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
};

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
                float m = INFINITY;
#pragma unroll
                for (int g = 0; g < G; ++g) {
#pragma unroll
                    for (int q = 0; q < P / 2; ++q) {
                        float e0, e1;
                        unpack2(disk_margin2<P>(A[g], B[g], r, q), e0, e1);
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
            Vec3 d = ray_of<P>(r, p);
            float b = fmaf(A.z, d.z, fmaf(A.y, d.y, A.x * d.x));
            float t = A.w * rcp_approx(b);
            float c0 = fmaf(t, fmaf(W0.z, d.z, fmaf(W0.y, d.y, W0.x * d.x)), W0.w);
            float c1 = fmaf(t, fmaf(W1.z, d.z, fmaf(W1.y, d.y, W1.x * d.x)), W1.w);
            float c2 = fmaf(t, fmaf(W2.z, d.z, fmaf(W2.y, d.y, W2.x * d.x)), W2.w);
            any |= (c0 >= 0.f) & (c1 >= 0.f) & (c2 >= 0.f);
        }
        if (any) narrow<P>(prm, sv, local0 + i, A, eye, near_clip, far_clip, r);
    }
}

template <int P, int MODE>
__global__ void __launch_bounds__(kThreads, 2) k_intersect(const __grid_constant__ IsectParams prm) {
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
            if (best_i[p] >= 0) {
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
                        if (fmaf(dx, dx, sy) <= 0.f) {
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
                if (fmaf(dx, dx, sy) <= 0.f) {
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
    const long long n_items = (long long)prm.n_tiles * prm.n_chunks;
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

    constexpr int TILE = kThreads * P;
    RayRegs<P> r;
    int cur_tile = -1;
    auto flush = [&]() {
        if (cur_tile < 0) return;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int pix = cur_tile * TILE + p * kThreads + tid;
            if (r.best_i[p] >= 0 && pix < prm.n_pix)
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
                const bool ok = pix < prm.n_pix;
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

__global__ void __launch_bounds__(256) k_shade(const __grid_constant__ ShadeParams p) {
    __shared__ float sm[256][3];
    const int base = blockIdx.x * 256;
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

// shadow rays of light l (renderer.py:293-299): origin frag_pos + 0.1 L, direction L, t_max = |light - frag_pos|.
// Miss pixels get a null direction (no hits): their visibility never reaches an output (image is masked).
__global__ void __launch_bounds__(256) k_rays_shadow(const __grid_constant__ ShadowParams p, int l, float* __restrict__ gray,
                                                     unsigned long long* __restrict__ zbuf2, float* __restrict__ obound) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    float len = 0.f;
    if (k < p.n) {
        const size_t n = (size_t)p.n;
        zbuf2[k] = kMissKey;
        const unsigned long long key = p.zbuf[k];
        Vec3 so = v3(0.f, 0.f, 0.f), L = v3(0.f, 0.f, 0.f);
        float dist = 0.f;
        if (key != kMissKey) {
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
        gray[k] = so.x; gray[n + k] = so.y; gray[2 * n + k] = so.z;
        gray[3 * n + k] = L.x; gray[4 * n + k] = L.y; gray[5 * n + k] = L.z;
        gray[6 * n + k] = dist;
    }
    for (int off = 16; off > 0; off >>= 1) len = fmaxf(len, __shfl_xor_sync(0xffffffffu, len, off));
    if ((threadIdx.x & 31) == 0 && len > 0.f) atomicMax((int*)obound, __float_as_int(len));
}

// visible iff nothing was hit inside (0, |L|), or the nearest such hit is the fragment's own primitive (:306-309)
__global__ void __launch_bounds__(256) k_shadow_resolve(const unsigned long long* __restrict__ zbuf,
                                                        const unsigned long long* __restrict__ zbuf2, int n, float* __restrict__ vis_l) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const unsigned long long self = zbuf[k], hit = zbuf2[k];
    const bool visible = hit == kMissKey || self == kMissKey || (unsigned)(hit & 0xFFFFFFFFull) == (unsigned)(self & 0xFFFFFFFFull);
    vis_l[k] = visible ? 1.f : 0.f;
}

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

__global__ void __launch_bounds__(128) k_backward(const __grid_constant__ BackwardParams p) {
    __shared__ double cta_acc[kMaxAccSlots];
    for (int j = threadIdx.x; j < p.sm.total; j += blockDim.x) cta_acc[j] = 0.0;
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = k < p.n;
    const int kk = live ? k : p.n - 1;      // dead lanes shadow the last pixel with zero incoming gradients
    Vec3 o, d;
    pixel_ray(*p.cam, p.rays, p.n, p.pix0, kk, &o, &d);
    PixelGrads g;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        g.image[c] = (live && p.g_image) ? p.g_image[(size_t)kk * 3 + c] : 0.f;
        g.pos[c] = (live && p.g_pos) ? p.g_pos[(size_t)kk * 3 + c] : 0.f;
        g.normal[c] = (live && p.g_normal) ? p.g_normal[(size_t)kk * 3 + c] : 0.f;
    }
    g.depth = (live && p.g_depth) ? p.g_depth[kk] : 0.f;
    const float dep = p.depth[kk];
    const bool hit = dep <= p.cam->far_clip && dep >= p.cam->near_clip;
    float vis_l[16];
    const float* vis = nullptr;
    if (p.vis) {
        for (int l = 0; l < p.sc.n_lights && l < 16; ++l) vis_l[l] = p.vis[(size_t)l * p.n + kk];
        vis = vis_l;
    }
    DeviceSink sink(p, cta_acc);
    backward_pixel(p.sc, v3(p.cam->eye[0], p.cam->eye[1], p.cam->eye[2]), o, d, (int)p.nearest[kk], hit, p.fl, vis, g, sink);
    __syncthreads();
    for (int j = threadIdx.x; j < p.sm.total; j += blockDim.x)
        if (cta_acc[j] != 0.0) atomicAdd(p.acc + j, cta_acc[j]);
}

struct FinalizeParams {
    GradPtrs gp; SlotMap sm; const double* acc; int K, L, Cn, light_pos_stride;
    SceneView sc; const double* prim_acc;
};
__global__ void __launch_bounds__(128) k_backward_finalize(const __grid_constant__ FinalizeParams p) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < p.sc.total) {      // per-primitive accumulators -> fp32 leaves (caller's strides)
        const int s = find_set(p.sc, j);
        const SetView& sv = p.sc.sets[s];
        const int local = j - sv.first;
        const double* a = p.prim_acc + (size_t)j * 7;
        float* gpos = p.gp.prim_pos[s];
        if (gpos) {
            const size_t row = sv.kind == KIND_TRIANGLE ? (size_t)local * 3 * sv.pos_stride : (size_t)local * sv.pos_stride;
            for (int c = 0; c < 3; ++c) gpos[row + c] += (float)a[c];
        }
        float* gnr = p.gp.prim_normal[s];
        if (gnr && sv.kind != KIND_SPHERE)
            for (int c = 0; c < 3; ++c) gnr[(size_t)local * sv.normal_stride + c] += (float)a[3 + c];
        float* grd = p.gp.prim_radius[s];
        if (grd && sv.kind == KIND_SPHERE) grd[local] += (float)a[6];
        return;
    }
    j -= p.sc.total;
    if (j >= p.sm.total) return;
    const float v = (float)p.acc[j];
    if (j < p.sm.coeffs) { if (p.gp.albedo) p.gp.albedo[j - p.sm.albedo] += v; }
    else if (j < p.sm.light_pos) { if (p.gp.coeffs) p.gp.coeffs[j - p.sm.coeffs] += v; }
    else if (j < p.sm.atten) {
        const int q = j - p.sm.light_pos;
        if (p.gp.light_pos) p.gp.light_pos[(size_t)(q / 3) * p.light_pos_stride + q % 3] += v;
    }
    else if (j < p.sm.colors) { if (p.gp.atten) p.gp.atten[j - p.sm.atten] += v; }
    else if (j < p.sm.ambient) { if (p.gp.colors) p.gp.colors[j - p.sm.colors] += v; }
    else if (j < p.sm.gamma) { if (p.gp.ambient) p.gp.ambient[j - p.sm.ambient] += v; }
    else { if (p.gp.gamma) p.gp.gamma[0] += v; }
}

// ---------------------------------------------------------------------------------------------------
// render_splats_along_ray kernels (renderer.py:537-751)
// ---------------------------------------------------------------------------------------------------
struct SplatParams {
    SceneView sc;                 // lights (camera space, stride 3, in the workspace) / colours / materials
    const CamState* cam;
    const float* z; int z_stride;
    const float* normal; int normal_stride;
    const int* mat;
    const float* vis;             // [L, n] or null
    int n;
    ShadeFlags fl;
    float* image; float* depth; float* normal_out; float* pos;                       // forward outputs
    const float* g_image; const float* g_depth; const float* g_normal; const float* g_pos;   // backward inputs
    float* gz; float* gnormal;    // backward outputs (caller's strides)
    SlotMap sm; double* acc;
};

__global__ void k_splat_setup(CamArgs a, CamState* cs, const float* light_pos4, int n_lights, float* light_cc) {
    if (threadIdx.x == 0 && blockIdx.x == 0)
        camera_setup(a.eye, a.at, a.up, 0, a.W, a.H, a.fovy, a.focal, a.near_clip, a.far_clip, cs);
    __syncthreads();
    for (int l = threadIdx.x; l < n_lights; l += blockDim.x) {
        Vec3 v = light_to_camera(*cs, light_pos4 + 4 * (size_t)l);
        light_cc[3 * l] = v.x; light_cc[3 * l + 1] = v.y; light_cc[3 * l + 2] = v.z;
    }
}

__global__ void __launch_bounds__(256) k_splat_forward(const __grid_constant__ SplatParams p) {
    __shared__ float sm[256][3];
    const int base = blockIdx.x * 256;
    const int k = base + threadIdx.x;
    const bool live = k < p.n;
    SplatOut so = SplatOut();
    float nn[3] = {0.f, 0.f, 0.f};
    if (live) {
        float vis_l[16];
        const float* vis = nullptr;
        if (p.vis) {
            for (int l = 0; l < p.sc.n_lights && l < 16; ++l) vis_l[l] = p.vis[(size_t)l * p.n + k];
            vis = vis_l;
        }
        const float* np_ = p.normal + (size_t)k * p.normal_stride;
        nn[0] = np_[0]; nn[1] = np_[1]; nn[2] = np_[2];
        so = splat_pixel_forward(p.sc, *p.cam, k, p.z[(size_t)k * p.z_stride], v3(nn[0], nn[1], nn[2]),
                                 p.mat ? p.mat[k] : 0, p.fl, vis);
        if (p.depth) p.depth[k] = so.depth;
    }
    if (p.image) store3(p.image, sm, base, p.n, so.image);
    if (p.pos) store3(p.pos, sm, base, p.n, so.pos);
    if (p.normal_out) store3(p.normal_out, sm, base, p.n, nn);
}

__global__ void __launch_bounds__(128) k_splat_backward(const __grid_constant__ SplatParams p) {
    __shared__ double cta_acc[kMaxAccSlots];
    for (int j = threadIdx.x; j < p.sm.total; j += blockDim.x) cta_acc[j] = 0.0;
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = k < p.n;
    const int kk = live ? k : p.n - 1;
    PixelGrads g;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        g.image[c] = (live && p.g_image) ? p.g_image[(size_t)kk * 3 + c] : 0.f;
        g.pos[c] = (live && p.g_pos) ? p.g_pos[(size_t)kk * 3 + c] : 0.f;
        g.normal[c] = (live && p.g_normal) ? p.g_normal[(size_t)kk * 3 + c] : 0.f;
    }
    g.depth = (live && p.g_depth) ? p.g_depth[kk] : 0.f;
    float vis_l[16];
    const float* vis = nullptr;
    if (p.vis) {
        for (int l = 0; l < p.sc.n_lights && l < 16; ++l) vis_l[l] = p.vis[(size_t)l * p.n + kk];
        vis = vis_l;
    }
    BackwardParams bp_view;          // DeviceSink only reads the slot map from it
    bp_view.sm = p.sm;
    DeviceSink sink(bp_view, cta_acc);
    const float* np_ = p.normal + (size_t)kk * p.normal_stride;
    float gz, gn[3];
    splat_pixel_backward(p.sc, *p.cam, kk, p.z[(size_t)kk * p.z_stride], v3(np_[0], np_[1], np_[2]),
                         p.mat ? p.mat[kk] : 0, p.fl, vis, g, sink, &gz, gn);
    if (live) {
        if (p.gz) p.gz[(size_t)k * p.z_stride] += gz;
        if (p.gnormal)
            for (int c = 0; c < 3; ++c) p.gnormal[(size_t)k * p.normal_stride + c] += gn[c];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < p.sm.total; j += blockDim.x)
        if (cta_acc[j] != 0.0) atomicAdd(p.acc + j, cta_acc[j]);
}

struct SplatFinalizeParams { GradPtrs gp; SlotMap sm; const double* acc; const CamState* cam; int L; };
__global__ void __launch_bounds__(128) k_splat_finalize(const __grid_constant__ SplatFinalizeParams p) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p.sm.total) return;
    const float v = (float)p.acc[j];
    if (j < p.sm.coeffs) { if (p.gp.albedo) p.gp.albedo[j - p.sm.albedo] += v; }
    else if (j < p.sm.light_pos) { if (p.gp.coeffs) p.gp.coeffs[j - p.sm.coeffs] += v; }
    else if (j < p.sm.atten) {
        // camera -> world: l_cc = R^T l_xyz - l_w R^T eye  =>  d/dl_xyz = R g,  d/dl_w = -(R^T eye) . g
        const int q = j - p.sm.light_pos;
        const int l = q / 3, c = q % 3;
        if (p.gp.light_pos && c == 0) {
            const CamState& cs = *p.cam;
            const double g0 = p.acc[j], g1 = p.acc[j + 1], g2 = p.acc[j + 2];
            float* dst = p.gp.light_pos + 4 * (size_t)l;
            for (int r = 0; r < 3; ++r) dst[r] += (float)(cs.R[3 * r] * g0 + cs.R[3 * r + 1] * g1 + cs.R[3 * r + 2] * g2);
            double gw = 0.0;
            const double gi[3] = {g0, g1, g2};
            for (int i = 0; i < 3; ++i)
                gw -= ((double)cs.R[i] * cs.eye[0] + (double)cs.R[3 + i] * cs.eye[1] + (double)cs.R[6 + i] * cs.eye[2]) * gi[i];
            dst[3] += (float)gw;
        }
    }
    else if (j < p.sm.colors) { if (p.gp.atten) p.gp.atten[j - p.sm.atten] += v; }
    else if (j < p.sm.ambient) { if (p.gp.colors) p.gp.colors[j - p.sm.colors] += v; }
    else if (j < p.sm.gamma) { if (p.gp.ambient) p.gp.ambient[j - p.sm.ambient] += v; }
}

// d/d(image) of mean((image - target)^2) and the loss itself (inverse-rendering step, test_optimization.py:104)
__global__ void __launch_bounds__(256) k_mse_grad(const float* __restrict__ image, const float* __restrict__ target,
                                                  int count, float* __restrict__ g_image, double* __restrict__ loss_acc) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    float e = 0.f;
    if (j < count) {
        const float diff = image[j] - target[j];
        g_image[j] = 2.f * diff / (float)count;
        e = diff * diff;
    }
    e = warp_sum(e);
    __shared__ float part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += part[w];
        atomicAdd(loss_acc, (double)s / (double)count);
    }
}

// ---------------------------------------------------------------------------------------------------
// FP32 FMA-pipe microbenchmark (roofline denominator check)
// ---------------------------------------------------------------------------------------------------
template <bool PACKED>
__global__ void __launch_bounds__(256) k_fma_peak(int iters, float* out) {
    float a = 1.0f + threadIdx.x * 1e-7f, b = 0.999f;
    if (PACKED) {
        unsigned long long x[8];
        for (int j = 0; j < 8; ++j) x[j] = pack2(a + j, a - j);
        const unsigned long long bb = pack2(b, b), cc = pack2(1e-3f, 2e-3f);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = fma2(x[j], bb, cc);
        }
        float s = 0.f;
        for (int j = 0; j < 8; ++j) { float lo, hi; unpack2(x[j], lo, hi); s += lo + hi; }
        if (s == 12345.678f) out[0] = s;
    } else {
        float x[16];
        for (int j = 0; j < 16; ++j) x[j] = a + j;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = fmaf(x[j], b, 1e-3f);
        }
        float s = 0.f;
        for (int j = 0; j < 16; ++j) s += x[j];
        if (s == 12345.678f) out[0] = s;
    }
}

// ---------------------------------------------------------------------------------------------------
// host-side orchestration (device-pointer API)
// ---------------------------------------------------------------------------------------------------
static int g_sm_count = 0;
static int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

struct Frame {           // everything derived from (scene, camera, options) once per call
    SceneView sc;
    CamArgs cam;
    int pix0, n;
    ShadeFlags fl;
    bool shadow;
    Workspace ws;
};

static int make_frame(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* opt, void* workspace,
                      size_t workspace_bytes, Frame* f) {
    if (!scene || !camera || !opt) return fail(SURF_ERR_BAD_ARG, "null scene/camera/options");
    std::string err;
    if (!build_scene_view(*scene, &f->sc, &err)) return fail(SURF_ERR_BAD_ARG, err);
    if (!check_camera(*camera, &err)) return fail(SURF_ERR_BAD_ARG, err);
    const int N = camera->width * camera->height;
    f->pix0 = opt->pixel_begin;
    int p1 = opt->pixel_end;
    if (f->pix0 == 0 && p1 == 0) p1 = N;
    if (f->pix0 < 0 || p1 > N || p1 <= f->pix0) return fail(SURF_ERR_BAD_ARG, "bad pixel range");
    f->n = p1 - f->pix0;
    f->cam = CamArgs{camera->eye, camera->at, camera->up, camera->proj, camera->width, camera->height,
                     camera->fovy, camera->focal_length, camera->near_clip, camera->far_clip};
    f->fl = ShadeFlags{opt->double_sided, opt->use_quartic};
    f->shadow = opt->shadow != 0;
    if (f->sc.n_lights > 16 && f->shadow) return fail(SURF_ERR_UNSUPPORTED, "shadow supports at most 16 lights");
    if (!workspace) return fail(SURF_ERR_WORKSPACE, "null workspace");
    carve(workspace, f->sc.total, f->n, f->sc.n_lights, f->shadow, &f->ws);
    if (f->ws.bytes > workspace_bytes) return fail(SURF_ERR_WORKSPACE, "workspace too small; see surf_workspace_bytes");
    return SURF_OK;
}

template <int P, int MODE>
static int launch_intersect(const IsectParams& prm, int grid, size_t smem, cudaStream_t st) {
    auto kern = k_intersect<P, MODE>;
    SURF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    timer_mark(0, 0, st);
    kern<<<grid, kThreads, smem, st>>>(prm);
    timer_mark(0, 1, st);
    SURF_LAUNCHED("k_intersect");
    return SURF_OK;
}

template <int MODE>
static int run_intersect_rays(const struct Frame& f, unsigned long long* zbuf, cudaStream_t st);

template <int P>
static int launch_screen(const ScreenParams& prm, int grid, size_t smem, cudaStream_t st) {
    auto kern = k_intersect_screen<P>;
    SURF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    timer_mark(0, 0, st);
    kern<<<grid, kThreads, smem, st>>>(prm);
    timer_mark(0, 1, st);
    SURF_LAUNCHED("k_intersect_screen");
    return SURF_OK;
}

static int run_intersect_screen(const Frame& f, const SurfOptions* opt, cudaStream_t st) {
    ScreenParams prm;
    prm.sc = f.sc; prm.cam = f.ws.cam; prm.circ = f.ws.circ; prm.rays = f.ws.rays; prm.zbuf = f.ws.zbuf;
    prm.pix0 = f.pix0; prm.n_pix = f.n; prm.W = f.cam.W; prm.total = f.sc.total;
    int P = opt->pixels_per_thread ? opt->pixels_per_thread : 8;
    if (P != 4 && P != 8 && P != 16) return fail(SURF_ERR_BAD_ARG, "pixels_per_thread must be 4, 8 or 16 for math_mode 3");
    const int row0 = f.pix0 / f.cam.W, row1 = (f.pix0 + f.n - 1) / f.cam.W;
    prm.row0 = row0;
    prm.tiles_x = (f.cam.W + 8 * P - 1) / (8 * P);
    const int tiles_y = (row1 - row0 + 1 + 31) / 32;
    prm.n_tiles = prm.tiles_x * tiles_y;
    const int occ = P <= 8 ? 3 : 2;
    const int grid_max = sm_count() * occ;
    int chunk = opt->chunk_prims ? opt->chunk_prims : 1024;
    if (chunk < 32 || chunk > 2048 || chunk % 32) return fail(SURF_ERR_BAD_ARG, "chunk_prims must be a multiple of 32 in [32, 2048]");
    if (!opt->chunk_prims)
        while (chunk > 64 && (long long)((f.sc.total + chunk - 1) / chunk) * prm.n_tiles < 4LL * grid_max) chunk /= 2;
    prm.chunk = chunk;
    prm.n_chunks = (f.sc.total + chunk - 1) / chunk;
    const long long items = (long long)prm.n_tiles * prm.n_chunks;
    const int grid = (int)std::min<long long>(items, grid_max);
    const size_t smem = (size_t)kStages * chunk * sizeof(float4);
    if (P == 4) return launch_screen<4>(prm, grid, smem, st);
    if (P == 8) return launch_screen<8>(prm, grid, smem, st);
    return launch_screen<16>(prm, grid, smem, st);
}

static int run_intersect(const Frame& f, const SurfOptions* opt, cudaStream_t st) {
    if (f.cam.proj != 0) {
        if (opt->math_mode == 1) {       // exact-only fallback kept for cross-checking the filtered kernel
            k_intersect_generic<<<(f.n + 255) / 256, 256, 0, st>>>(f.sc, f.ws.cam, f.pix0, f.n, f.ws.zbuf);
            SURF_LAUNCHED("k_intersect_generic");
            return SURF_OK;
        }
        SURF_CUDA(cudaMemsetAsync(f.ws.obound, 0, 4, st));
        k_rays_ortho<<<(f.n + 255) / 256, 256, 0, st>>>(f.ws.cam, f.pix0, f.n, f.ws.gray, f.ws.obound);
        SURF_LAUNCHED("k_rays_ortho");
        return run_intersect_rays<0>(f, f.ws.zbuf, st);
    }
    if (opt->math_mode == 3) return run_intersect_screen(f, opt, st);
    IsectParams prm;
    prm.sc = f.sc; prm.cam = f.ws.cam; prm.packed = f.ws.packed; prm.rays = f.ws.rays; prm.zbuf = f.ws.zbuf;
    prm.n_pix = f.n;
    int P = opt->pixels_per_thread ? opt->pixels_per_thread : 8;
    if (P != 2 && P != 4 && P != 8) return fail(SURF_ERR_BAD_ARG, "pixels_per_thread must be 2, 4 or 8");
    const int tile = kThreads * P;
    prm.n_tiles = (f.n + tile - 1) / tile;
    // stage capacity: chunk_prims disk records (2 float4 each); keep >= 4x grid items for balance on small frames
    int chunk = opt->chunk_prims ? opt->chunk_prims : 1024;
    if (chunk < 32 || chunk > 2048 || chunk % 32) return fail(SURF_ERR_BAD_ARG, "chunk_prims must be a multiple of 32 in [32, 2048]");
    const int grid_max = sm_count() * 2;
    if (!opt->chunk_prims) {
        // pick the largest chunk whose item count splits over the persistent grid with <= 1.5% quantisation loss
        // (items are dealt as equal contiguous ranges: the slowest CTA runs ceil(items / grid) of them)
        auto items_for = [&](int ch) {
            long long items = 0;
            for (int s = 0; s < f.sc.n_sets; ++s) {
                const int ppc = (ch * 2) / rec_f4(f.sc.sets[s].kind);
                items += (f.sc.sets[s].count + ppc - 1) / ppc;
            }
            return items * prm.n_tiles;
        };
        int best = 64;
        double best_loss = 1e30;
        for (int ch = 1024; ch >= 64; ch /= 2) {
            const long long items = items_for(ch);
            const long long per = (items + grid_max - 1) / grid_max;
            const double loss = (double)per * grid_max / (double)items - 1.0;
            if (loss <= 0.015) { best = ch; best_loss = loss; break; }
            if (loss < best_loss) { best = ch; best_loss = loss; }
        }
        chunk = best;
    }
    prm.stage_f4 = chunk * 2;
    int nchunks = 0;
    for (int s = 0; s < kMaxSets; ++s) {
        prm.chunks_before[s] = nchunks;
        if (s < f.sc.n_sets) {
            const int ppc = prm.stage_f4 / rec_f4(f.sc.sets[s].kind);
            nchunks += (f.sc.sets[s].count + ppc - 1) / ppc;
        }
    }
    prm.chunks_before[kMaxSets] = nchunks;
    prm.n_chunks = nchunks;
    const long long items = (long long)prm.n_tiles * nchunks;
    const int grid = (int)std::min<long long>(items, grid_max);
    const size_t smem = (size_t)kStages * prm.stage_f4 * sizeof(float4);
    const int mode = opt->math_mode;
    if (mode < 0 || mode > 2) return fail(SURF_ERR_BAD_ARG, "math_mode must be 0..3");
#define SURF_DISPATCH(PP)                                                      \
    if (P == PP) {                                                             \
        if (mode == 0) return launch_intersect<PP, 0>(prm, grid, smem, st);    \
        if (mode == 1) return launch_intersect<PP, 1>(prm, grid, smem, st);    \
        return launch_intersect<PP, 2>(prm, grid, smem, st);                   \
    }
    SURF_DISPATCH(2)
    SURF_DISPATCH(4)
    SURF_DISPATCH(8)
#undef SURF_DISPATCH
    return fail(SURF_ERR_BAD_ARG, "unsupported pixels_per_thread");
}

template <int MODE>
static int run_intersect_rays(const Frame& f, unsigned long long* zbuf, cudaStream_t st) {
    k_prep_rays<<<(f.sc.total + 255) / 256, 256, 0, st>>>(f.sc, f.ws.obound, f.ws.packed);
    SURF_LAUNCHED("k_prep_rays");
    constexpr int P = 4;
    RayParams prm;
    prm.sc = f.sc; prm.cam = f.ws.cam; prm.packed = f.ws.packed; prm.gray = f.ws.gray; prm.zbuf = zbuf; prm.n_pix = f.n;
    const int tile = kThreads * P;
    prm.n_tiles = (f.n + tile - 1) / tile;
    const int grid_max = sm_count() * 2;
    int chunk = 1024;
    while (chunk > 64) {
        long long items = 0;
        for (int s = 0; s < f.sc.n_sets; ++s) {
            const int ppc = (chunk * 2) / rec_f4(f.sc.sets[s].kind);
            items += (f.sc.sets[s].count + ppc - 1) / ppc;
        }
        if (items * prm.n_tiles >= 4LL * grid_max) break;
        chunk /= 2;
    }
    prm.stage_f4 = chunk * 2;
    int nchunks = 0;
    for (int s = 0; s < kMaxSets; ++s) {
        prm.chunks_before[s] = nchunks;
        if (s < f.sc.n_sets) {
            const int ppc = prm.stage_f4 / rec_f4(f.sc.sets[s].kind);
            nchunks += (f.sc.sets[s].count + ppc - 1) / ppc;
        }
    }
    prm.chunks_before[kMaxSets] = nchunks;
    prm.n_chunks = nchunks;
    const long long items = (long long)prm.n_tiles * nchunks;
    const int grid = (int)std::min<long long>(items, grid_max);
    const size_t smem = (size_t)kStages * prm.stage_f4 * sizeof(float4);
    auto kern = k_intersect_rays<P, MODE>;
    SURF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreads, smem, st>>>(prm);
    SURF_LAUNCHED("k_intersect_rays");
    return SURF_OK;
}

static int run_common_prologue(const Frame& f, float* ray_out, cudaStream_t st, bool need_rays_and_zbuf) {
    k_setup<<<1, 32, 0, st>>>(f.cam, f.ws.cam);
    SURF_LAUNCHED("k_setup");
    if (need_rays_and_zbuf) {
        k_raygen<<<(f.n + 255) / 256, 256, 0, st>>>(f.ws.cam, f.pix0, f.n, f.ws.rays, ray_out, f.ws.zbuf);
        SURF_LAUNCHED("k_raygen");
    }
    return SURF_OK;
}

static int forward_impl(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* opt, void* workspace,
                        size_t workspace_bytes, const SurfOutputs* out, cudaStream_t st) {
    if (!out) return fail(SURF_ERR_BAD_ARG, "null outputs");
    Frame f;
    int rc = make_frame(scene, camera, opt, workspace, workspace_bytes, &f);
    if (rc) return rc;
    if ((rc = run_common_prologue(f, out->ray_dir, st, true))) return rc;
    if (f.cam.proj == 0 && opt->math_mode == 3) {
        k_prep_screen<<<(f.sc.total + 255) / 256, 256, 0, st>>>(f.sc, f.ws.cam, f.ws.circ);
        SURF_LAUNCHED("k_prep_screen");
    } else if (f.cam.proj == 0) {
        k_prep<<<(f.sc.total + 255) / 256, 256, 0, st>>>(f.sc, f.ws.cam, f.ws.packed);
        SURF_LAUNCHED("k_prep");
    }
    if ((rc = run_intersect(f, opt, st))) return rc;
    if (f.shadow) {
        ShadowParams sp{f.sc, f.ws.cam, f.ws.rays, f.ws.zbuf, f.ws.vis, f.pix0, f.n};
        if (opt->math_mode == 1) {       // exact-only brute force kept for cross-checking
            dim3 grid((f.n + 127) / 128, f.sc.n_lights);
            k_shadow<<<grid, 128, 0, st>>>(sp);
            SURF_LAUNCHED("k_shadow");
        } else {
            for (int l = 0; l < f.sc.n_lights; ++l) {
                SURF_CUDA(cudaMemsetAsync(f.ws.obound, 0, 4, st));
                k_rays_shadow<<<(f.n + 255) / 256, 256, 0, st>>>(sp, l, f.ws.gray, f.ws.zbuf2, f.ws.obound);
                SURF_LAUNCHED("k_rays_shadow");
                if ((rc = run_intersect_rays<1>(f, f.ws.zbuf2, st))) return rc;
                k_shadow_resolve<<<(f.n + 255) / 256, 256, 0, st>>>(f.ws.zbuf, f.ws.zbuf2, f.n, f.ws.vis + (size_t)l * f.n);
                SURF_LAUNCHED("k_shadow_resolve");
            }
        }
    }
    ShadeParams sh;
    sh.sc = f.sc; sh.cam = f.ws.cam; sh.rays = f.ws.rays; sh.zbuf = f.ws.zbuf; sh.vis = f.shadow ? f.ws.vis : nullptr;
    sh.pix0 = f.pix0; sh.n = f.n; sh.fl = f.fl;
    sh.image = out->image; sh.depth = out->depth; sh.normal = out->normal; sh.pos = out->pos;
    sh.nearest = (long long*)out->nearest;
    timer_mark(1, 0, st);
    k_shade<<<(f.n + 255) / 256, 256, 0, st>>>(sh);
    timer_mark(1, 1, st);
    SURF_LAUNCHED("k_shade");
    return SURF_OK;
}

static int backward_impl(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* opt, void* workspace,
                         size_t workspace_bytes, const int64_t* nearest, const float* depth, const SurfOutGrads* og,
                         const SurfSceneGrads* sg, cudaStream_t st, bool workspace_is_warm) {
    if (!nearest || !depth || !og || !sg) return fail(SURF_ERR_BAD_ARG, "null nearest/depth/out_grads/scene_grads");
    Frame f;
    int rc = make_frame(scene, camera, opt, workspace, workspace_bytes, &f);
    if (rc) return rc;
    const SlotMap sm = slot_map(f.sc.n_materials, f.sc.n_lights, f.sc.n_colors);
    if (sm.total > kMaxAccSlots) return fail(SURF_ERR_UNSUPPORTED, "too many materials/lights/colours for the backward accumulators");
    if (!workspace_is_warm) {
        // recompute camera state and rays (the forward call may have used a different workspace)
        k_setup<<<1, 32, 0, st>>>(f.cam, f.ws.cam);
        SURF_LAUNCHED("k_setup");
        k_raygen<<<(f.n + 255) / 256, 256, 0, st>>>(f.ws.cam, f.pix0, f.n, f.ws.rays, nullptr, f.ws.zbuf);
        SURF_LAUNCHED("k_raygen");
        if (f.shadow) return fail(SURF_ERR_UNSUPPORTED, "shadow backward needs the forward workspace (visibility)");
    }
    SURF_CUDA(cudaMemsetAsync(f.ws.acc, 0, sizeof(double) * kMaxAccSlots, st));
    SURF_CUDA(cudaMemsetAsync(f.ws.prim_acc, 0, sizeof(double) * 7 * (size_t)f.sc.total, st));
    BackwardParams bp;
    bp.sc = f.sc; bp.cam = f.ws.cam; bp.rays = f.ws.rays; bp.vis = f.shadow ? f.ws.vis : nullptr;
    bp.nearest = (const long long*)nearest; bp.depth = depth;
    bp.g_image = og->image; bp.g_depth = og->depth; bp.g_normal = og->normal; bp.g_pos = og->pos;
    bp.pix0 = f.pix0; bp.n = f.n; bp.fl = f.fl; bp.sm = sm; bp.acc = f.ws.acc; bp.prim_acc = f.ws.prim_acc;
    for (int s = 0; s < kMaxSets; ++s) {
        bp.gp.prim_pos[s] = sg->sets[s].pos; bp.gp.prim_normal[s] = sg->sets[s].normal; bp.gp.prim_radius[s] = sg->sets[s].radius;
    }
    bp.gp.light_pos = sg->light_pos; bp.gp.atten = sg->light_attenuation; bp.gp.ambient = sg->ambient;
    bp.gp.colors = sg->colors; bp.gp.albedo = sg->albedo; bp.gp.coeffs = sg->coeffs; bp.gp.gamma = sg->gamma;
    timer_mark(2, 0, st);
    k_backward<<<(f.n + 127) / 128, 128, 0, st>>>(bp);
    timer_mark(2, 1, st);
    SURF_LAUNCHED("k_backward");
    FinalizeParams fp{bp.gp, sm, f.ws.acc, f.sc.n_materials, f.sc.n_lights, f.sc.n_colors, f.sc.light_pos_stride,
                      f.sc, f.ws.prim_acc};
    k_backward_finalize<<<(f.sc.total + sm.total + 127) / 128, 128, 0, st>>>(fp);
    SURF_LAUNCHED("k_backward_finalize");
    return SURF_OK;
}

static int splat_frame(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* opt, const SurfSplats* sp,
                       void* workspace, size_t workspace_bytes, SplatParams* p, CamArgs* cam, Workspace* ws, float** light_cc) {
    if (!scene || !camera || !opt || !sp) return fail(SURF_ERR_BAD_ARG, "null scene/camera/options/splats");
    std::string err;
    SceneView sc;
    if (!build_scene_view(*scene, &sc, &err, true)) return fail(SURF_ERR_BAD_ARG, err);
    if (!check_camera(*camera, &err)) return fail(SURF_ERR_BAD_ARG, err);
    if (camera->proj != 0) return fail(SURF_ERR_UNSUPPORTED, "render_splats_along_ray is defined for the perspective frustum");
    if (scene->light_pos_stride != 4) return fail(SURF_ERR_BAD_ARG, "along-ray lights must be homogeneous [L,4] (torch.mm with the 4x4 view matrix)");
    if (sp->count != camera->width * camera->height) return fail(SURF_ERR_BAD_ARG, "one splat per pixel: count must equal width*height");
    if (!sp->z || !sp->normal) return fail(SURF_ERR_UNSUPPORTED, "splat depths and normals are required (normal estimation is not built)");
    if ((sp->z_stride != 1 && sp->z_stride != 3) || (sp->normal_stride != 3 && sp->normal_stride != 4))
        return fail(SURF_ERR_BAD_ARG, "z_stride must be 1 or 3, normal_stride 3 or 4");
    if (sc.n_lights > 16 && sp->light_vis) return fail(SURF_ERR_UNSUPPORTED, "light_vis supports at most 16 lights");
    if (!workspace) return fail(SURF_ERR_WORKSPACE, "null workspace");
    carve(workspace, 0, sp->count, sc.n_lights, false, ws);
    if (ws->bytes > workspace_bytes) return fail(SURF_ERR_WORKSPACE, "workspace too small; see surf_workspace_bytes");
    *light_cc = ws->rays;                       // the ray buffer is unused on this path: holds the L x 3 camera-space lights
    *cam = CamArgs{camera->eye, camera->at, camera->up, 0, camera->width, camera->height, camera->fovy,
                   camera->focal_length, camera->near_clip, camera->far_clip};
    p->sc = sc;
    p->sc.light_pos = *light_cc; p->sc.light_pos_stride = 3; p->sc.gamma = nullptr;
    p->cam = ws->cam;
    p->z = sp->z; p->z_stride = sp->z_stride; p->normal = sp->normal; p->normal_stride = sp->normal_stride;
    p->mat = sp->material_idx; p->vis = sp->light_vis; p->n = sp->count;
    p->fl = ShadeFlags{0, opt->use_quartic};
    p->image = p->depth = p->normal_out = p->pos = nullptr;
    p->g_image = p->g_depth = p->g_normal = p->g_pos = nullptr;
    p->gz = p->gnormal = nullptr;
    p->sm = slot_map(sc.n_materials, sc.n_lights, sc.n_colors);
    p->acc = ws->acc;
    if (p->sm.total > kMaxAccSlots) return fail(SURF_ERR_UNSUPPORTED, "too many materials/lights/colours for the backward accumulators");
    return SURF_OK;
}

static int splats_forward_impl(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* opt, const SurfSplats* sp,
                               void* workspace, size_t bytes, const SurfOutputs* out, cudaStream_t st) {
    if (!out) return fail(SURF_ERR_BAD_ARG, "null outputs");
    SplatParams p; CamArgs cam; Workspace ws; float* lcc;
    int rc = splat_frame(scene, camera, opt, sp, workspace, bytes, &p, &cam, &ws, &lcc);
    if (rc) return rc;
    k_splat_setup<<<1, 64, 0, st>>>(cam, ws.cam, scene->light_pos, scene->n_lights, lcc);
    SURF_LAUNCHED("k_splat_setup");
    p.image = out->image; p.depth = out->depth; p.normal_out = out->normal; p.pos = out->pos;
    timer_mark(1, 0, st);
    k_splat_forward<<<(p.n + 255) / 256, 256, 0, st>>>(p);
    timer_mark(1, 1, st);
    SURF_LAUNCHED("k_splat_forward");
    return SURF_OK;
}

static int splats_backward_impl(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* opt, const SurfSplats* sp,
                                void* workspace, size_t bytes, const SurfOutGrads* og, const SurfSceneGrads* sg,
                                const SurfSplatGrads* spg, cudaStream_t st) {
    if (!og || !sg || !spg) return fail(SURF_ERR_BAD_ARG, "null out_grads/scene_grads/splat_grads");
    SplatParams p; CamArgs cam; Workspace ws; float* lcc;
    int rc = splat_frame(scene, camera, opt, sp, workspace, bytes, &p, &cam, &ws, &lcc);
    if (rc) return rc;
    k_splat_setup<<<1, 64, 0, st>>>(cam, ws.cam, scene->light_pos, scene->n_lights, lcc);
    SURF_LAUNCHED("k_splat_setup");
    SURF_CUDA(cudaMemsetAsync(ws.acc, 0, sizeof(double) * kMaxAccSlots, st));
    p.g_image = og->image; p.g_depth = og->depth; p.g_normal = og->normal; p.g_pos = og->pos;
    p.gz = spg->z; p.gnormal = spg->normal;
    timer_mark(2, 0, st);
    k_splat_backward<<<(p.n + 127) / 128, 128, 0, st>>>(p);
    timer_mark(2, 1, st);
    SURF_LAUNCHED("k_splat_backward");
    SplatFinalizeParams fp;
    for (int s = 0; s < kMaxSets; ++s) fp.gp.prim_pos[s] = fp.gp.prim_normal[s] = fp.gp.prim_radius[s] = nullptr;
    fp.gp.light_pos = sg->light_pos; fp.gp.atten = sg->light_attenuation; fp.gp.ambient = sg->ambient;
    fp.gp.colors = sg->colors; fp.gp.albedo = sg->albedo; fp.gp.coeffs = sg->coeffs; fp.gp.gamma = nullptr;
    fp.sm = p.sm; fp.acc = ws.acc; fp.cam = ws.cam; fp.L = scene->n_lights;
    k_splat_finalize<<<(p.sm.total + 127) / 128, 128, 0, st>>>(fp);
    SURF_LAUNCHED("k_splat_finalize");
    return SURF_OK;
}

}  // namespace surf

// ===================================================================================================
// C ABI
// ===================================================================================================
using namespace surf;

struct SurfContext {
    int device;
    cudaStream_t stream;
    void* arena; size_t arena_bytes;
    uint64_t h2d, d2h;
};

extern "C" {

int surf_abi_version(void) { return SURF_ABI_VERSION; }
const char* surf_last_error(void) { return g_error.c_str(); }
int surf_last_launch_count(void) { return g_launches; }
void surf_set_kernel_timing(int32_t enabled) {
    g_timers.enabled = enabled != 0;
    for (int k = 0; k < 3; ++k) g_timers.count[k] = 0;
}
double surf_last_kernel_ms(int32_t which) {
    if (which < 0 || which > 2 || !g_timers.created || g_timers.count[which] == 0) return -1.0;
    return timer_ms(which, g_timers.count[which] - 1);
}
double surf_mean_kernel_ms(int32_t which, int32_t* launches) {
    if (launches) *launches = 0;
    if (which < 0 || which > 2 || !g_timers.created || g_timers.count[which] == 0) return -1.0;
    const long long n = g_timers.count[which];
    const long long first = n > kTimerRing ? n - kTimerRing : 0;
    double sum = 0.0;
    int used = 0;
    for (long long i = first; i < n; ++i) {
        const double ms = timer_ms(which, i);
        if (ms >= 0.0) { sum += ms; ++used; }
    }
    if (launches) *launches = used;
    return used ? sum / used : -1.0;
}

size_t surf_workspace_bytes(int32_t total_prims, int32_t n_pixels, int32_t n_lights, int32_t shadow) {
    Workspace ws;
    carve(nullptr, total_prims, n_pixels, n_lights, shadow != 0, &ws);
    return ws.bytes;
}

int surf_forward(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* options, void* workspace,
                 size_t workspace_bytes, const SurfOutputs* out, void* cuda_stream) {
    g_launches = 0;
    return forward_impl(scene, camera, options, workspace, workspace_bytes, out, (cudaStream_t)cuda_stream);
}

int surf_backward(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* options, void* workspace,
                  size_t workspace_bytes, const int64_t* nearest, const float* depth, const SurfOutGrads* out_grads,
                  const SurfSceneGrads* scene_grads, void* cuda_stream) {
    g_launches = 0;
    // options->forced_nearest == 2: the caller guarantees `workspace` still holds this frame's forward state
    const bool warm = options && options->forced_nearest == 2;
    return backward_impl(scene, camera, options, workspace, workspace_bytes, nearest, depth, out_grads, scene_grads,
                         (cudaStream_t)cuda_stream, warm);
}

int surf_splats_forward(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* options,
                        const SurfSplats* splats, void* workspace, size_t workspace_bytes, const SurfOutputs* out,
                        void* cuda_stream) {
    g_launches = 0;
    return splats_forward_impl(scene, camera, options, splats, workspace, workspace_bytes, out, (cudaStream_t)cuda_stream);
}

int surf_splats_backward(const SurfScene* scene, const SurfCamera* camera, const SurfOptions* options,
                         const SurfSplats* splats, void* workspace, size_t workspace_bytes, const SurfOutGrads* out_grads,
                         const SurfSceneGrads* scene_grads, const SurfSplatGrads* splat_grads, void* cuda_stream) {
    g_launches = 0;
    return splats_backward_impl(scene, camera, options, splats, workspace, workspace_bytes, out_grads, scene_grads,
                                splat_grads, (cudaStream_t)cuda_stream);
}

double surf_fma_peak(int32_t mode, int32_t iters, void* cuda_stream) {
    cudaStream_t st = (cudaStream_t)cuda_stream;
    float* out = nullptr;
    if (cudaMalloc(&out, 4) != cudaSuccess) return -1.0;
    const int grid = sm_count() * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {   // first pass warms up
        cudaEventRecord(e0, st);
        if (mode == 1) k_fma_peak<true><<<grid, 256, 0, st>>>(iters, out);
        else k_fma_peak<false><<<grid, 256, 0, st>>>(iters, out);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    if (cudaGetLastError() != cudaSuccess || ms <= 0.f) return -1.0;
    // lane-FMAs: both variants execute 16 lane-FMAs per thread per iteration
    const double lane_fma = (double)grid * 256.0 * (double)iters * 16.0;
    return lane_fma / (ms * 1e-3);
}

// ---------------------------------------------------------------------------------------------------
// host-pointer API
// ---------------------------------------------------------------------------------------------------
SurfContext* surf_context_create(int32_t device) {
    if (cudaSetDevice(device) != cudaSuccess) { g_error = "cudaSetDevice failed"; return nullptr; }
    SurfContext* c = new SurfContext();
    c->device = device; c->arena = nullptr; c->arena_bytes = 0; c->h2d = c->d2h = 0;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        g_error = "cudaStreamCreate failed";
        delete c;
        return nullptr;
    }
    return c;
}

void surf_context_destroy(SurfContext* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->arena) cudaFree(c->arena);
    cudaStreamDestroy(c->stream);
    delete c;
}

int surf_context_last_transfer(const SurfContext* c, uint64_t* h2d, uint64_t* d2h) {
    if (!c) return fail(SURF_ERR_BAD_ARG, "null context");
    if (h2d) *h2d = c->h2d;
    if (d2h) *d2h = c->d2h;
    return SURF_OK;
}

}  // extern "C"

namespace surf {

// bump allocator over the context arena
struct Bump {
    char* base; size_t off, cap;
    void* take(size_t bytes) {
        off = align_up(off, 256);
        void* p = base ? base + off : nullptr;
        off += bytes;
        return p;
    }
};

struct HostPlan {        // device mirrors of every host array of one call
    SurfScene dscene;
    SurfCamera dcam;
    SurfOutputs dout;
    SurfOutGrads dgout;
    SurfSceneGrads dgrads;
    void* workspace; size_t workspace_bytes;
    float* d_target; double* d_loss;
    size_t grads_begin, grads_end;     // arena range holding the gradient accumulators (zeroed per call)
};

static size_t set_pos_floats(const SurfPrimSet& s) {
    return (size_t)s.count * (s.kind == SURF_TRIANGLE ? 3 : 1) * s.pos_stride;
}

// Lays out the arena; with b.base == nullptr it only measures.
static void plan_host(const SurfScene& hs, const SurfCamera& hc, const SurfOptions& opt, int n, bool want_bwd,
                      bool has_target, Bump& b, HostPlan* pl) {
    pl->dscene = hs;
    pl->dcam = hc;
    int total = 0;
    for (int k = 0; k < hs.n_sets; ++k) {
        const SurfPrimSet& s = hs.sets[k];
        SurfPrimSet& d = pl->dscene.sets[k];
        total += s.count;
        d.pos = (const float*)b.take(set_pos_floats(s) * 4);
        d.normal = s.normal ? (const float*)b.take((size_t)s.count * s.normal_stride * 4) : nullptr;
        d.radius = s.radius ? (const float*)b.take((size_t)s.count * 4) : nullptr;
        d.material_idx = (const int32_t*)b.take((size_t)s.count * 4);
    }
    pl->dscene.light_pos = (const float*)b.take((size_t)hs.n_lights * hs.light_pos_stride * 4);
    pl->dscene.light_color_idx = (const int32_t*)b.take((size_t)hs.n_lights * 4);
    pl->dscene.light_attenuation = (const float*)b.take((size_t)hs.n_lights * 12);
    pl->dscene.ambient = (const float*)b.take(12);
    pl->dscene.colors = (const float*)b.take((size_t)hs.n_colors * 12);
    pl->dscene.albedo = (const float*)b.take((size_t)hs.n_materials * 12);
    pl->dscene.coeffs = (const float*)b.take((size_t)hs.n_materials * 12);
    pl->dscene.gamma = hs.gamma ? (const float*)b.take(4) : nullptr;
    pl->dcam.eye = (const float*)b.take(12);
    pl->dcam.at = (const float*)b.take(12);
    pl->dcam.up = (const float*)b.take(12);
    pl->dout.image = (float*)b.take((size_t)n * 12);
    pl->dout.depth = (float*)b.take((size_t)n * 4);
    pl->dout.normal = (float*)b.take((size_t)n * 12);
    pl->dout.pos = (float*)b.take((size_t)n * 12);
    pl->dout.nearest = (int64_t*)b.take((size_t)n * 8);
    pl->dout.ray_dir = (float*)b.take((size_t)(hc.proj == 0 ? n : 1) * 12);
    pl->workspace_bytes = surf_workspace_bytes(total, n, hs.n_lights, opt.shadow);
    pl->workspace = b.take(pl->workspace_bytes);
    pl->d_target = nullptr; pl->d_loss = nullptr;
    memset(&pl->dgout, 0, sizeof(pl->dgout));
    memset(&pl->dgrads, 0, sizeof(pl->dgrads));
    if (!want_bwd) return;
    pl->dgout.image = (const float*)b.take((size_t)n * 12);
    pl->dgout.depth = (const float*)b.take((size_t)n * 4);
    pl->dgout.normal = (const float*)b.take((size_t)n * 12);
    pl->dgout.pos = (const float*)b.take((size_t)n * 12);
    if (has_target) {
        pl->d_target = (float*)b.take((size_t)n * 12);
        pl->d_loss = (double*)b.take(8);
    }
    b.off = align_up(b.off, 256);
    pl->grads_begin = b.off;
    for (int k = 0; k < hs.n_sets; ++k) {
        const SurfPrimSet& s = hs.sets[k];
        pl->dgrads.sets[k].pos = (float*)b.take(set_pos_floats(s) * 4);
        pl->dgrads.sets[k].normal = s.normal ? (float*)b.take((size_t)s.count * s.normal_stride * 4) : nullptr;
        pl->dgrads.sets[k].radius = s.kind == SURF_SPHERE ? (float*)b.take((size_t)s.count * 4) : nullptr;
    }
    pl->dgrads.light_pos = (float*)b.take((size_t)hs.n_lights * hs.light_pos_stride * 4);
    pl->dgrads.light_attenuation = (float*)b.take((size_t)hs.n_lights * 12);
    pl->dgrads.ambient = (float*)b.take(12);
    pl->dgrads.colors = (float*)b.take((size_t)hs.n_colors * 12);
    pl->dgrads.albedo = (float*)b.take((size_t)hs.n_materials * 12);
    pl->dgrads.coeffs = (float*)b.take((size_t)hs.n_materials * 12);
    pl->dgrads.gamma = (float*)b.take(4);
    pl->grads_end = align_up(b.off, 256);
    b.off = pl->grads_end;
}

static int h2d(SurfContext* c, const void* dst, const void* src, size_t bytes) {
    if (!bytes) return SURF_OK;
    SURF_CUDA(cudaMemcpyAsync((void*)dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    c->h2d += bytes;
    return SURF_OK;
}
static int d2h(SurfContext* c, void* dst, const void* src, size_t bytes) {
    if (!bytes || !dst) return SURF_OK;
    SURF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    c->d2h += bytes;
    return SURF_OK;
}

static int host_call(SurfContext* c, const SurfScene* hs, const SurfCamera* hc, const SurfOptions* opt,
                     const SurfOutputs* hout, const SurfOutGrads* hgout, const float* target, float* loss,
                     const SurfSceneGrads* hgrads, bool want_bwd) {
    if (!c || !hs || !hc || !opt) return fail(SURF_ERR_BAD_ARG, "null context/scene/camera/options");
    SURF_CUDA(cudaSetDevice(c->device));
    g_launches = 0;
    c->h2d = c->d2h = 0;
    SceneView probe;
    std::string err;
    if (!build_scene_view(*hs, &probe, &err)) return fail(SURF_ERR_BAD_ARG, err);
    if (!check_camera(*hc, &err)) return fail(SURF_ERR_BAD_ARG, err);
    const int N = hc->width * hc->height;
    int p0 = opt->pixel_begin, p1 = opt->pixel_end;
    if (p0 == 0 && p1 == 0) p1 = N;
    if (p0 < 0 || p1 > N || p1 <= p0) return fail(SURF_ERR_BAD_ARG, "bad pixel range");
    const int n = p1 - p0;
    const bool has_target = want_bwd && target != nullptr;

    HostPlan pl;
    Bump measure{nullptr, 0, 0};
    plan_host(*hs, *hc, *opt, n, want_bwd, has_target, measure, &pl);
    const size_t need = align_up(measure.off, 256) + 256;
    if (need > c->arena_bytes) {
        if (c->arena) SURF_CUDA(cudaFree(c->arena));
        c->arena = nullptr; c->arena_bytes = 0;
        SURF_CUDA(cudaMalloc(&c->arena, need));
        c->arena_bytes = need;
    }
    Bump b{(char*)c->arena, 0, c->arena_bytes};
    plan_host(*hs, *hc, *opt, n, want_bwd, has_target, b, &pl);

    int rc;
    for (int k = 0; k < hs->n_sets; ++k) {
        const SurfPrimSet& s = hs->sets[k];
        const SurfPrimSet& d = pl.dscene.sets[k];
        if ((rc = h2d(c, d.pos, s.pos, set_pos_floats(s) * 4))) return rc;
        if (s.normal && (rc = h2d(c, d.normal, s.normal, (size_t)s.count * s.normal_stride * 4))) return rc;
        if (s.radius && (rc = h2d(c, d.radius, s.radius, (size_t)s.count * 4))) return rc;
        if ((rc = h2d(c, d.material_idx, s.material_idx, (size_t)s.count * 4))) return rc;
    }
    if ((rc = h2d(c, pl.dscene.light_pos, hs->light_pos, (size_t)hs->n_lights * hs->light_pos_stride * 4))) return rc;
    if ((rc = h2d(c, pl.dscene.light_color_idx, hs->light_color_idx, (size_t)hs->n_lights * 4))) return rc;
    if ((rc = h2d(c, pl.dscene.light_attenuation, hs->light_attenuation, (size_t)hs->n_lights * 12))) return rc;
    if ((rc = h2d(c, pl.dscene.ambient, hs->ambient, 12))) return rc;
    if ((rc = h2d(c, pl.dscene.colors, hs->colors, (size_t)hs->n_colors * 12))) return rc;
    if ((rc = h2d(c, pl.dscene.albedo, hs->albedo, (size_t)hs->n_materials * 12))) return rc;
    if ((rc = h2d(c, pl.dscene.coeffs, hs->coeffs, (size_t)hs->n_materials * 12))) return rc;
    if (hs->gamma && (rc = h2d(c, pl.dscene.gamma, hs->gamma, 4))) return rc;
    if ((rc = h2d(c, pl.dcam.eye, hc->eye, 12))) return rc;
    if ((rc = h2d(c, pl.dcam.at, hc->at, 12))) return rc;
    if ((rc = h2d(c, pl.dcam.up, hc->up, 12))) return rc;

    if ((rc = forward_impl(&pl.dscene, &pl.dcam, opt, pl.workspace, pl.workspace_bytes, &pl.dout, c->stream))) return rc;

    if (hout) {
        if ((rc = d2h(c, hout->image, pl.dout.image, (size_t)n * 12))) return rc;
        if ((rc = d2h(c, hout->depth, pl.dout.depth, (size_t)n * 4))) return rc;
        if ((rc = d2h(c, hout->normal, pl.dout.normal, (size_t)n * 12))) return rc;
        if ((rc = d2h(c, hout->pos, pl.dout.pos, (size_t)n * 12))) return rc;
        if ((rc = d2h(c, hout->nearest, pl.dout.nearest, (size_t)n * 8))) return rc;
        if ((rc = d2h(c, hout->ray_dir, pl.dout.ray_dir, (size_t)(hc->proj == 0 ? n : 1) * 12))) return rc;
    }
    if (want_bwd) {
        if (!hgrads) return fail(SURF_ERR_BAD_ARG, "null scene_grads");
        SurfOutGrads og = pl.dgout;
        if (has_target) {
            if ((rc = h2d(c, pl.d_target, target, (size_t)n * 12))) return rc;
            SURF_CUDA(cudaMemsetAsync(pl.d_loss, 0, 8, c->stream));
            k_mse_grad<<<(n * 3 + 255) / 256, 256, 0, c->stream>>>(pl.dout.image, pl.d_target, n * 3, (float*)og.image, pl.d_loss);
            SURF_LAUNCHED("k_mse_grad");
            og.depth = nullptr; og.normal = nullptr; og.pos = nullptr;
        } else {
            if (!hgout) return fail(SURF_ERR_BAD_ARG, "need out_grads or target_image");
            if (hgout->image) { if ((rc = h2d(c, og.image, hgout->image, (size_t)n * 12))) return rc; } else og.image = nullptr;
            if (hgout->depth) { if ((rc = h2d(c, og.depth, hgout->depth, (size_t)n * 4))) return rc; } else og.depth = nullptr;
            if (hgout->normal) { if ((rc = h2d(c, og.normal, hgout->normal, (size_t)n * 12))) return rc; } else og.normal = nullptr;
            if (hgout->pos) { if ((rc = h2d(c, og.pos, hgout->pos, (size_t)n * 12))) return rc; } else og.pos = nullptr;
        }
        SURF_CUDA(cudaMemsetAsync((char*)c->arena + pl.grads_begin, 0, pl.grads_end - pl.grads_begin, c->stream));
        if ((rc = backward_impl(&pl.dscene, &pl.dcam, opt, pl.workspace, pl.workspace_bytes, pl.dout.nearest,
                                pl.dout.depth, &og, &pl.dgrads, c->stream, true)))
            return rc;
        for (int k = 0; k < hs->n_sets; ++k) {
            const SurfPrimSet& s = hs->sets[k];
            if ((rc = d2h(c, hgrads->sets[k].pos, pl.dgrads.sets[k].pos, set_pos_floats(s) * 4))) return rc;
            if (s.normal && (rc = d2h(c, hgrads->sets[k].normal, pl.dgrads.sets[k].normal, (size_t)s.count * s.normal_stride * 4))) return rc;
            if (s.kind == SURF_SPHERE && (rc = d2h(c, hgrads->sets[k].radius, pl.dgrads.sets[k].radius, (size_t)s.count * 4))) return rc;
        }
        if ((rc = d2h(c, hgrads->light_pos, pl.dgrads.light_pos, (size_t)hs->n_lights * hs->light_pos_stride * 4))) return rc;
        if ((rc = d2h(c, hgrads->light_attenuation, pl.dgrads.light_attenuation, (size_t)hs->n_lights * 12))) return rc;
        if ((rc = d2h(c, hgrads->ambient, pl.dgrads.ambient, 12))) return rc;
        if ((rc = d2h(c, hgrads->colors, pl.dgrads.colors, (size_t)hs->n_colors * 12))) return rc;
        if ((rc = d2h(c, hgrads->albedo, pl.dgrads.albedo, (size_t)hs->n_materials * 12))) return rc;
        if ((rc = d2h(c, hgrads->coeffs, pl.dgrads.coeffs, (size_t)hs->n_materials * 12))) return rc;
        if (hs->gamma && (rc = d2h(c, hgrads->gamma, pl.dgrads.gamma, 4))) return rc;
    }
    double loss_d = 0.0;
    if (has_target && loss) {
        SURF_CUDA(cudaMemcpyAsync(&loss_d, pl.d_loss, 8, cudaMemcpyDeviceToHost, c->stream));
        c->d2h += 8;
    }
    SURF_CUDA(cudaStreamSynchronize(c->stream));
    if (has_target && loss) *loss = (float)loss_d;
    return SURF_OK;
}

}  // namespace surf

extern "C" {

int surf_render_host(SurfContext* ctx, const SurfScene* scene, const SurfCamera* camera, const SurfOptions* options,
                     const SurfOutputs* out) {
    return host_call(ctx, scene, camera, options, out, nullptr, nullptr, nullptr, nullptr, false);
}

int surf_render_backward_host(SurfContext* ctx, const SurfScene* scene, const SurfCamera* camera,
                              const SurfOptions* options, const SurfOutputs* out, const SurfOutGrads* out_grads,
                              const float* target_image, float* loss, const SurfSceneGrads* scene_grads) {
    return host_call(ctx, scene, camera, options, out, out_grads, target_image, loss, scene_grads, true);
}

}  // extern "C"
