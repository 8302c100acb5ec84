// surf_splats.cuh - part of libsurf_b200.so (single translation unit: included by surf_kernels.cu inside namespace surf).
// render_splats_along_ray kernels
#pragma once

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
    const float* pos_in;          // [n, 3] explicit positions or null
    float* gpos;                  // [n, 3] gradient of the explicit positions or null
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
        so = splat_pixel_forward(p.sc, *p.cam, k, p.pos_in ? 0.f : p.z[(size_t)k * p.z_stride],
                                 p.pos_in ? p.pos_in + 3 * (size_t)k : nullptr, v3(nn[0], nn[1], nn[2]),
                                 p.mat ? clamp_index(p.mat[k], p.sc.n_materials) : 0, p.fl, vis);
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
    BackwardParams bp_view;          // DeviceSink only reads the slot map from it
    bp_view.sm = p.sm;
    for (int base = blockIdx.x * blockDim.x; base < p.n; base += gridDim.x * blockDim.x) {   // persistent, see k_backward
        const int k = base + threadIdx.x;
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
        DeviceSink sink(bp_view, cta_acc);
        const float* np_ = p.normal + (size_t)kk * p.normal_stride;
        float gz, gpos[3], gn[3];
        splat_pixel_backward(p.sc, *p.cam, kk, p.pos_in ? 0.f : p.z[(size_t)kk * p.z_stride],
                             p.pos_in ? p.pos_in + 3 * (size_t)kk : nullptr, v3(np_[0], np_[1], np_[2]),
                             p.mat ? clamp_index(p.mat[kk], p.sc.n_materials) : 0, p.fl, vis, g, sink, &gz, gpos, gn);
        if (live) {
            if (p.gz && !p.pos_in) p.gz[(size_t)k * p.z_stride] += gz;
            if (p.gpos && p.pos_in)
                for (int c = 0; c < 3; ++c) p.gpos[(size_t)k * 3 + c] += gpos[c];
            if (p.gnormal)
                for (int c = 0; c < 3; ++c) p.gnormal[(size_t)k * p.normal_stride + c] += gn[c];
        }
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

