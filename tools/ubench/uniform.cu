// Does the disk filter of k_intersect run faster when the warp-uniform record scalars come from UNIFORM registers
// (constant bank -> LDCU -> FFMA2 Rpair, UR.F32, Rpair) instead of vector registers (shared memory -> LDS.128 ->
// FFMA2 R.F32, Rpair, Rpair)?  Same arithmetic as chunk_disks MODE 0: P = 8 pixels per thread as 4 packed pairs,
// 10 FFMA2-class + 2 MUFU.RCP per pair and disk, one FMNMX3 per pair.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o uniform uniform.cu && ./uniform
#include <cstdio>
#include <cuda_runtime.h>
__constant__ float4 crecs[4096];          // 2048 records of 32 B = the whole 64 KB constant bank
__device__ __forceinline__ unsigned long long pack2(float a, float b){ unsigned long long r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void unpack2(unsigned long long v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c){ unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b){ unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ float rcpa(float x){ float r; asm("rcp.approx.ftz.f32 %0, %1;":"=f"(r):"f"(x)); return r;}

template <int Q>
__device__ __forceinline__ float margin_min(const float4 A, const float4 B, const unsigned long long* dx, const unsigned long long* dy, const unsigned long long* dz, float m) {
    const unsigned long long nx = pack2(A.x, A.x), ny = pack2(A.y, A.y), nz = pack2(A.z, A.z), nm = pack2(A.w, A.w);
    const unsigned long long ox = pack2(B.x, B.x), oy = pack2(B.y, B.y), oz = pack2(B.z, B.z), nr = pack2(B.w, B.w);
    unsigned long long b2[Q], t2[Q], rx[Q], ry[Q], rz[Q], e2[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) b2[q] = mul2(nx, dx[q]);
#pragma unroll
    for (int q = 0; q < Q; ++q) b2[q] = fma2(ny, dy[q], b2[q]);
#pragma unroll
    for (int q = 0; q < Q; ++q) b2[q] = fma2(nz, dz[q], b2[q]);
#pragma unroll
    for (int q = 0; q < Q; ++q) { float b0, b1; unpack2(b2[q], b0, b1); t2[q] = mul2(nm, pack2(rcpa(b0), rcpa(b1))); }
#pragma unroll
    for (int q = 0; q < Q; ++q) rx[q] = fma2(t2[q], dx[q], ox);
#pragma unroll
    for (int q = 0; q < Q; ++q) ry[q] = fma2(t2[q], dy[q], oy);
#pragma unroll
    for (int q = 0; q < Q; ++q) rz[q] = fma2(t2[q], dz[q], oz);
#pragma unroll
    for (int q = 0; q < Q; ++q) e2[q] = fma2(rx[q], rx[q], nr);
#pragma unroll
    for (int q = 0; q < Q; ++q) e2[q] = fma2(ry[q], ry[q], e2[q]);
#pragma unroll
    for (int q = 0; q < Q; ++q) e2[q] = fma2(rz[q], rz[q], e2[q]);
#pragma unroll
    for (int q = 0; q < Q; ++q) { float e0, e1; unpack2(e2[q], e0, e1); m = fminf(m, fminf(e0, e1)); }
    return m;
}

// SRC 0: records staged in shared memory (LDS.128, vector registers); SRC 1: records in the constant bank (LDCU, uniform registers)
template <int SRC, int Q>
__global__ void __launch_bounds__(256, 2) k(const float4* __restrict__ recs, const float* __restrict__ rays, int n, int reps, float* out, int* hits) {
    extern __shared__ float4 srecs[];
    if (SRC == 0) { for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) srecs[i] = recs[i]; __syncthreads(); }
    unsigned long long dx[Q], dy[Q], dz[Q];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int q = 0; q < Q; ++q) {
        const float* r = rays + ((size_t)t * Q + q) * 6;
        dx[q] = pack2(r[0], r[1]); dy[q] = pack2(r[2], r[3]); dz[q] = pack2(r[4], r[5]);
    }
    float acc = 0.f; int nh = 0;
    for (int rep = 0; rep < reps; ++rep) {
        float m_prev = INFINITY;
#pragma unroll 1
        for (int j = 0; j < n; j += 2) {
            float m = INFINITY;
            if (SRC == 0) { m = margin_min<Q>(srecs[2 * j], srecs[2 * j + 1], dx, dy, dz, m); m = margin_min<Q>(srecs[2 * j + 2], srecs[2 * j + 3], dx, dy, dz, m); }
            else { m = margin_min<Q>(crecs[2 * j], crecs[2 * j + 1], dx, dy, dz, m); m = margin_min<Q>(crecs[2 * j + 2], crecs[2 * j + 3], dx, dy, dz, m); }
            if (m_prev <= 0.f) { ++nh; acc += m_prev; }         // the rare branch of the real kernel, on the previous group
            m_prev = m;
        }
        if (m_prev <= 0.f) { ++nh; acc += m_prev; }
    }
    out[t] = acc; if (nh) atomicAdd(hits, nh);
}

// VARIANT kernels (constant bank only): 0 = loads at the top of the iteration (as k<1, Q>), 1 = two groups per iteration, the
// next group's records fetched while the current one computes, 2 = variant 0 plus a vector LDC "touch" eight groups ahead
template <int VARIANT, int Q>
__global__ void __launch_bounds__(256, 2) kc(const float* __restrict__ rays, int g0, int g1, int reps, float* out, int* hits) {
    unsigned long long dx[Q], dy[Q], dz[Q];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int q = 0; q < Q; ++q) {
        const float* r = rays + ((size_t)t * Q + q) * 6;
        dx[q] = pack2(r[0], r[1]); dy[q] = pack2(r[2], r[3]); dz[q] = pack2(r[4], r[5]);
    }
    if (VARIANT >= 5) {      // warm the constant cache: one load per 64-byte group, all in flight together
        int acc = 0;
#pragma unroll 16
        for (int k = g0; k < g1; ++k) acc ^= __float_as_int(crecs[4 * k].x);
        if (acc == 0x7fc12345) atomicAdd(hits, 1);
    }
    for (int rep = 0; rep < reps; ++rep) {
        float m_prev = INFINITY;
        if (VARIANT == 1 || VARIANT == 6) {
            float4 a0 = crecs[4 * g0], b0 = crecs[4 * g0 + 1], a1 = crecs[4 * g0 + 2], b1 = crecs[4 * g0 + 3];
#pragma unroll 1
            for (int k = g0; k < g1; k += 2) {
                const float4 c0 = crecs[4 * k + 4], d0 = crecs[4 * k + 5], c1 = crecs[4 * k + 6], d1 = crecs[4 * k + 7];
                float m = margin_min<Q>(a0, b0, dx, dy, dz, INFINITY);
                m = margin_min<Q>(a1, b1, dx, dy, dz, m);
                if (m_prev <= 0.f) atomicAdd(hits, 1);
                a0 = crecs[4 * k + 8]; b0 = crecs[4 * k + 9]; a1 = crecs[4 * k + 10]; b1 = crecs[4 * k + 11];
                float m2 = margin_min<Q>(c0, d0, dx, dy, dz, INFINITY);
                m2 = margin_min<Q>(c1, d1, dx, dy, dz, m2);
                if (m <= 0.f) atomicAdd(hits, 1);
                m_prev = m2;
            }
        } else if (VARIANT == 7) {       // four groups per iteration, two register sets
            float4 a0 = crecs[4 * g0], b0 = crecs[4 * g0 + 1], a1 = crecs[4 * g0 + 2], b1 = crecs[4 * g0 + 3];
#pragma unroll 1
            for (int k = g0; k < g1; k += 4) {
                float4 c0 = crecs[4 * k + 4], d0 = crecs[4 * k + 5], c1 = crecs[4 * k + 6], d1 = crecs[4 * k + 7];
                float m = margin_min<Q>(a0, b0, dx, dy, dz, INFINITY);
                m = margin_min<Q>(a1, b1, dx, dy, dz, m);
                if (m_prev <= 0.f) atomicAdd(hits, 1);
                a0 = crecs[4 * k + 8]; b0 = crecs[4 * k + 9]; a1 = crecs[4 * k + 10]; b1 = crecs[4 * k + 11];
                float m2 = margin_min<Q>(c0, d0, dx, dy, dz, INFINITY);
                m2 = margin_min<Q>(c1, d1, dx, dy, dz, m2);
                if (m <= 0.f) atomicAdd(hits, 1);
                c0 = crecs[4 * k + 12]; d0 = crecs[4 * k + 13]; c1 = crecs[4 * k + 14]; d1 = crecs[4 * k + 15];
                m = margin_min<Q>(a0, b0, dx, dy, dz, INFINITY);
                m = margin_min<Q>(a1, b1, dx, dy, dz, m);
                if (m2 <= 0.f) atomicAdd(hits, 1);
                a0 = crecs[4 * k + 16]; b0 = crecs[4 * k + 17]; a1 = crecs[4 * k + 18]; b1 = crecs[4 * k + 19];
                m2 = margin_min<Q>(c0, d0, dx, dy, dz, INFINITY);
                m2 = margin_min<Q>(c1, d1, dx, dy, dz, m2);
                if (m <= 0.f) atomicAdd(hits, 1);
                m_prev = m2;
            }
        } else if (VARIANT == 8) {       // two groups per iteration, both prefetched at the top, one set copied at the bottom
            float4 a0 = crecs[4 * g0], b0 = crecs[4 * g0 + 1], a1 = crecs[4 * g0 + 2], b1 = crecs[4 * g0 + 3];
#pragma unroll 1
            for (int k = g0; k < g1; k += 2) {
                const float4 c0 = crecs[4 * k + 4], d0 = crecs[4 * k + 5], c1 = crecs[4 * k + 6], d1 = crecs[4 * k + 7];
                const float4 e0 = crecs[4 * k + 8], f0 = crecs[4 * k + 9], e1 = crecs[4 * k + 10], f1 = crecs[4 * k + 11];
                float m = margin_min<Q>(a0, b0, dx, dy, dz, INFINITY);
                m = margin_min<Q>(a1, b1, dx, dy, dz, m);
                if (m_prev <= 0.f) atomicAdd(hits, 1);
                float m2 = margin_min<Q>(c0, d0, dx, dy, dz, INFINITY);
                m2 = margin_min<Q>(c1, d1, dx, dy, dz, m2);
                if (m <= 0.f) atomicAdd(hits, 1);
                m_prev = m2;
                a0 = e0; b0 = f0; a1 = e1; b1 = f1;
            }
        } else if (VARIANT == 3) {
            float4 a0 = crecs[4 * g0], b0 = crecs[4 * g0 + 1], a1 = crecs[4 * g0 + 2], b1 = crecs[4 * g0 + 3];
#pragma unroll 1
            for (int k = g0; k < g1; ++k) {
                const float4 c0 = crecs[4 * k + 4], d0 = crecs[4 * k + 5], c1 = crecs[4 * k + 6], d1 = crecs[4 * k + 7];
                float m = margin_min<Q>(a0, b0, dx, dy, dz, INFINITY);
                m = margin_min<Q>(a1, b1, dx, dy, dz, m);
                if (m_prev <= 0.f) atomicAdd(hits, 1);
                m_prev = m;
                a0 = c0; b0 = d0; a1 = c1; b1 = d1;
            }
        } else if (VARIANT == 4) {
            float4 a0 = crecs[4 * g0], b0 = crecs[4 * g0 + 1], a1 = crecs[4 * g0 + 2], b1 = crecs[4 * g0 + 3];
#pragma unroll 2
            for (int k = g0; k < g1; ++k) {
                const float4 c0 = crecs[4 * k + 4], d0 = crecs[4 * k + 5], c1 = crecs[4 * k + 6], d1 = crecs[4 * k + 7];
                float m = margin_min<Q>(a0, b0, dx, dy, dz, INFINITY);
                m = margin_min<Q>(a1, b1, dx, dy, dz, m);
                if (m_prev <= 0.f) atomicAdd(hits, 1);
                m_prev = m;
                a0 = c0; b0 = d0; a1 = c1; b1 = d1;
            }
        } else {
#pragma unroll 1
            for (int k = g0; k < g1; ++k) {
                if (VARIANT == 2) {
                    float x;
                    const float* ahead = (const float*)&crecs[4 * (k + 8)] + (threadIdx.x & 31) % 16 * 0;
                    asm volatile("{ .reg .u64 a; cvta.to.const.u64 a, %1; ld.const.f32 %0, [a]; }" : "=f"(x) : "l"(ahead));
                }
                float m = margin_min<Q>(crecs[4 * k], crecs[4 * k + 1], dx, dy, dz, INFINITY);
                m = margin_min<Q>(crecs[4 * k + 2], crecs[4 * k + 3], dx, dy, dz, m);
                if (m_prev <= 0.f) atomicAdd(hits, 1);
                m_prev = m;
            }
        }
    }
    out[t] = 0.f;
}

template <int VARIANT, int Q> void run_cold(const char* name, const float4* recs, const float* rays, float* out, int* hits) {
    const int grid = 148 * 2, launches = 24, n_groups = 1016;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaMemset(hits, 0, 4);
    kc<VARIANT, Q><<<grid, 256>>>(rays, 0, n_groups, 1, out, hits); cudaDeviceSynchronize();
    cudaMemset(hits, 0, 4);
    cudaEventRecord(e0);
    for (int l = 0; l < launches; ++l) {     // the copy invalidates the constant cache like the real sequence does
        cudaMemcpyToSymbolAsync(crecs, recs, 2 * 2048 * sizeof(float4), 0, cudaMemcpyDeviceToDevice, 0);
        kc<VARIANT, Q><<<grid, 256>>>(rays, 0, n_groups, 1, out, hits);
    }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int h; cudaMemcpy(&h, hits, 4, cudaMemcpyDeviceToHost);
    const double warp_disks_per_smsp = (double)grid * 8 * n_groups * 2 * launches / (148.0 * 4.0);
    const double cyc = ms * 1e-3 * 1.965e9 / warp_disks_per_smsp;
    printf("%-46s %8.3f ms  %.1f cycles per warp-disk per SMSP  [groups passing: %d, err %s]\n", name, ms, cyc, h, cudaGetErrorString(cudaGetLastError()));
}

template <int SRC, int Q> void run(const char* name, const float4* recs, const float* rays, int n, float* out, int* hits) {
    const int grid = 148 * 2, reps = 24;
    cudaFuncSetAttribute(k<SRC, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaMemset(hits, 0, 4);
    k<SRC, Q><<<grid, 256, SRC == 0 ? 65536 : 0>>>(recs, rays, n, 2, out, hits); cudaDeviceSynchronize();
    cudaMemset(hits, 0, 4);
    cudaEventRecord(e0); k<SRC, Q><<<grid, 256, SRC == 0 ? 65536 : 0>>>(recs, rays, n, reps, out, hits); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int h; cudaMemcpy(&h, hits, 4, cudaMemcpyDeviceToHost);
    const double warp_disks_per_smsp = (double)grid * 8 * n * reps / (148.0 * 4.0);
    const double cyc = ms * 1e-3 * 1.965e9 / warp_disks_per_smsp;
    const double tests = (double)grid * 256 * 2 * Q * n * reps;
    printf("%-34s %8.3f ms  %.1f cycles per warp-disk per SMSP (ideal %d)  %.3e tests/s = %.1f%% of FP32 peak  [groups passing: %d, err %s]\n", name, ms, cyc, 20 * Q,
           tests / (ms * 1e-3), 100.0 * tests * 10 / (ms * 1e-3) / (148.0 * 128 * 1.965e9), h, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int n = 2048, nthreads = 148 * 2 * 256;
    float4* h = new float4[2 * n];
    srand(1);
    auto rnd = []() { return (float)rand() / RAND_MAX; };
    for (int i = 0; i < n; ++i) {       // unit normal + numer | o - c, -(r + slack)^2 : small disks 4..6 units in front of the camera
        float nx = rnd() - 0.5f, ny = rnd() - 0.5f, nz = 0.5f + rnd(), l = sqrtf(nx * nx + ny * ny + nz * nz);
        nx /= l; ny /= l; nz /= l;
        float cx = 4 * (rnd() - 0.5f), cy = 4 * (rnd() - 0.5f), cz = -(4 + 2 * rnd()), r = 0.01f;
        h[2 * i] = make_float4(nx, ny, nz, cx * nx + cy * ny + cz * nz);
        h[2 * i + 1] = make_float4(-cx, -cy, -cz, -r * r);
    }
    float* hr = new float[(size_t)nthreads * 4 * 6];
    for (size_t i = 0; i < (size_t)nthreads * 4; ++i) {
        for (int e = 0; e < 2; ++e) { float x = rnd() - 0.5f, y = rnd() - 0.5f, z = -1.f, l = sqrtf(x * x + y * y + 1); hr[i * 6 + e] = x / l; hr[i * 6 + 2 + e] = y / l; hr[i * 6 + 4 + e] = z / l; }
    }
    float4* recs; float* rays; float* out; int* hits;
    cudaMalloc(&recs, 2 * n * sizeof(float4)); cudaMalloc(&rays, (size_t)nthreads * 4 * 6 * 4); cudaMalloc(&out, nthreads * 4); cudaMalloc(&hits, 4);
    cudaMemcpy(recs, h, 2 * n * sizeof(float4), cudaMemcpyHostToDevice);
    cudaMemcpy(rays, hr, (size_t)nthreads * 4 * 6 * 4, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(crecs, recs, 2 * n * sizeof(float4), 0, cudaMemcpyDeviceToDevice);
    run<0, 4>("shared memory -> vector regs, P=8", recs, rays, n, out, hits);
    run<1, 4>("constant bank -> uniform regs, P=8", recs, rays, n, out, hits);
    run<0, 2>("shared memory -> vector regs, P=4", recs, rays, n, out, hits);
    run<1, 2>("constant bank -> uniform regs, P=4", recs, rays, n, out, hits);
    run<0, 4>("shared memory -> vector regs, P=8", recs, rays, n, out, hits);
    run<1, 4>("constant bank -> uniform regs, P=8", recs, rays, n, out, hits);
    run_cold<0, 4>("cold launches, loads at the top, P=8", recs, rays, out, hits);
    run_cold<1, 4>("cold launches, next group prefetched, P=8", recs, rays, out, hits);
    run_cold<7, 4>("cold launches, 4 groups/iter rotating, P=8", recs, rays, out, hits);
    run_cold<8, 4>("cold launches, 2 groups/iter both prefetched + copy, P=8", recs, rays, out, hits);
    run_cold<0, 4>("cold launches, loads at the top, P=8", recs, rays, out, hits);
    return 0;
}
