// surf_runtime.cuh - part of libsurf_b200.so (included by every translation unit inside namespace surf; the globals are
// C++17 inline variables, so all translation units of the library share one instance).
// error handling, launch accounting, kernel timers, workspace layout
#pragma once

// ---------------------------------------------------------------------------------------------------
// error handling (thread-local string) / launch accounting and optional timers (process-wide counters)
// ---------------------------------------------------------------------------------------------------
inline thread_local std::string g_error;
inline std::atomic<int> g_launches{0};   // process-wide: autograd runs backward on its own thread

// optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg): a ring of
// event pairs per kernel kind so that a whole timed region can be averaged without synchronising inside it
constexpr int kTimerRing = 256;
struct KernelTimers {
    bool enabled = false;
    bool created = false;
    cudaEvent_t ev[3][kTimerRing][2] = {};
    long long count[3] = {0, 0, 0};     // launches recorded since timing was (re-)enabled
};
inline KernelTimers g_timers;   // process-wide (see g_launches)
inline void timer_mark(int which, int edge, cudaStream_t st) {
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
inline double timer_ms(int which, long long index) {
    const int slot = (int)(index % kTimerRing);
    if (cudaEventSynchronize(g_timers.ev[which][slot][1]) != cudaSuccess) return -1.0;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_timers.ev[which][slot][0], g_timers.ev[which][slot][1]) != cudaSuccess) return -1.0;
    return (double)ms;
}

inline int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}
inline int cuda_fail(cudaError_t e, const char* where) {
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
    float* gray;                 // [8, n] generic rays: origin xyz, direction xyz, t_max (orthographic / shadow rays);
                                 // row 7 = int slot_of[n]: where pixel k's compacted shadow ray is stored (-1: none)
    unsigned long long* zbuf2;   // [n] z-buffer keys of the shadow rays
    float* obound;               // [1] max |origin| over the generic rays of the launch (as float bits, atomicMax)
    double* loss_acc;            // [1] loss of a fused inverse-rendering step (surf_step_mse)
    float* gimg;                 // [n, 3] d(loss)/d(image) of a fused step (carved for step calls only)
    // candidate queue of the constant-bank intersection path (surf_isect_const.cu), frames above 256x256 pixels only:
    uint2* cq;                   // [cq_capacity] (thread slot, record group) pairs that passed the conservative filter
    int* cq_ctl;                 // [0] entries appended; 256 bytes, followed directly by
    unsigned char* cq_flags;     // [launch][tile] overflow map, cq_flag_bytes
    int* cq_inside;              // [total_prims] disks whose bounding sphere holds the eye (count in cq_ctl[2])
    int cq_capacity;
    size_t cq_flag_bytes;
    size_t bytes;
};
constexpr int kMaxAccSlots = 512;
constexpr int kMaxShadowLights = 254;      // per-light live-ray counters in the workspace / tile table of k_intersect_shadow

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline size_t packed_bytes_bound(int total_prims) {
    // worst case: all triangles (4 float4 each) + 8 sets x 128-byte padding
    return (size_t)total_prims * 64 + kMaxSets * 128 + 256;
}

// `generic_rays`: the frame traces rays with per-ray origins (orthographic camera, or shadow rays) and needs the
// generic-ray list + second key buffer (40 B per ray); `step`: a fused step also holds d(loss)/d(image).
// `queue`: append the candidate queue of the constant-bank intersection path (perspective frames above 256x256 pixels);
// it sits at the end, so the layout of everything else does not depend on it.
inline void carve(void* base, int total_prims, int n_pix, int n_lights, bool shadow, Workspace* ws, bool generic_rays = true,
                  bool step = false, bool queue = true) {
    char* p = (char*)base;
    size_t off = 0;
    ws->cam = (CamState*)(p + off); off += align_up(sizeof(CamState), 256);
    // with shadows: one set of disk records per light origin (k_prep_lights), 32 B per primitive and light
    ws->packed = (float4*)(p + off);
    off += align_up(packed_bytes_bound(total_prims) * (size_t)(shadow ? (n_lights > 1 ? n_lights : 1) : 1), 256);
    ws->circ = (float4*)(p + off); off += align_up((size_t)total_prims * 16 + 256, 256);
    ws->rays = (float*)(p + off); off += align_up((size_t)3 * n_pix * sizeof(float), 256);
    ws->zbuf = (unsigned long long*)(p + off); off += align_up((size_t)n_pix * 8, 256);
    ws->acc = (double*)(p + off); off += align_up((size_t)kMaxAccSlots * 8, 256);
    ws->prim_acc = (double*)(p + off); off += align_up((size_t)total_prims * 7 * 8, 256);
    ws->obound = (float*)(p + off); off += 1024;           // [0] origin bound, [1 ..] per-light live-ray counters
    ws->loss_acc = (double*)(p + off); off += 256;
    ws->vis = (float*)(p + off);
    if (shadow) off += align_up((size_t)n_lights * n_pix * sizeof(float), 256);
    // generic-ray buffers, 40 B per ray.  With shadows the rays of all lights are traced in one launch: n_pix * n_lights.
    const size_t n_rays = (generic_rays || shadow) ? (size_t)n_pix * (size_t)(shadow ? (n_lights > 1 ? n_lights : 1) : 1) : 0;
    ws->gray = (float*)(p + off); off += align_up((size_t)8 * n_rays * sizeof(float), 256);
    ws->zbuf2 = (unsigned long long*)(p + off); off += align_up(n_rays * 8, 256);
    ws->gimg = (float*)(p + off);
    if (step) off += align_up((size_t)3 * n_pix * sizeof(float), 256);
    ws->cq = nullptr; ws->cq_ctl = nullptr; ws->cq_flags = nullptr; ws->cq_inside = nullptr; ws->cq_capacity = 0; ws->cq_flag_bytes = 0;
    if (queue && n_pix > 256 * 256 && !generic_rays) {
        // 10 candidates per pixel (config E appends 1.6 per pixel with the plane filter, 7.7 with the sphere filter); one flag byte per (launch of at least 506 disks, tile of at least 2048 pixels)
        ws->cq_capacity = 10 * n_pix;
        ws->cq = (uint2*)(p + off); off += align_up((size_t)ws->cq_capacity * sizeof(uint2), 256);
        ws->cq_ctl = (int*)(p + off); off += 256;
        ws->cq_flag_bytes = align_up((size_t)(total_prims / 506 + 1) * (size_t)(n_pix / 2048 + 1), 256);
        ws->cq_flags = (unsigned char*)(p + off); off += ws->cq_flag_bytes;
        ws->cq_inside = (int*)(p + off); off += align_up((size_t)total_prims * sizeof(int) + 4, 256);
    }
    ws->bytes = off;
}

