// surf_scatter.cuh - part of libsurf_b200.so (included by surf_kernels.cu inside namespace surf).
// The projection layer's scatter renderers: surfels of one view are projected into another camera and
// scattered onto its pixel grid.
//   k_project_surfels / _backward   world -> camera -> image plane -> pixel coordinates and destination index
//                                   (projection_layer.py:20-86: project_surfels, project_image_coordinates)
//   k_scatter_accum                 atomic scatter-add of weighted surfel data and of the weights
//   k_scatter_normalize             division by the accumulated weight: scatter_mean_dim0 (utils.py:146-175) and
//                                   scatter_weighted_blended_oit (utils.py:178-215, Weighted Blended OIT)
//   k_scatter_backward              gather form of the backward (no atomics): d/dx, d/dz, d/d(center_dist_2)
// The reference builds these from scatter_add_ on padded copies of every operand; here the destination index
// n_dst (= "dump") is simply skipped.
#pragma once

struct ProjectParams {
    int batch, n, pos_stride, W, H;
    float f, sx_px, sy_px, cx, cy;      // focal length; pixel scale -(W-1)/w, (H-1)/h; centre W/2, H/2
    const float* eye; const float* at; const float* up;
    long long eye_stride, at_stride, up_stride;
    const float* pos;                   // [B, n, pos_stride] world coordinates
    float* px_coord;                    // [B, n, 3]: pixel x, pixel y, depth -Z
    long long* px_idx;                  // [B, n]
    const float* g_px; float* g_pos;    // backward
};

// rows of the view matrix [R^T | -R^T eye] of scene b (lookat = inverse of [R | eye], utils.py:376-399)
__device__ __forceinline__ void view_rows(const ProjectParams& p, int b, float Rt[9], float t[3]) {
    CamState cs;
    camera_setup(p.eye + b * p.eye_stride, p.at + b * p.at_stride, p.up + b * p.up_stride, 0, p.W, p.H, 1.0, 1.0, 0.f, 1.f, &cs);
    for (int i = 0; i < 3; ++i) {
        Rt[3 * i] = cs.R[i]; Rt[3 * i + 1] = cs.R[3 + i]; Rt[3 * i + 2] = cs.R[6 + i];
        t[i] = -(cs.R[i] * cs.eye[0] + cs.R[3 + i] * cs.eye[1] + cs.R[6 + i] * cs.eye[2]);
    }
}

__global__ void __launch_bounds__(256) k_project_surfels(const __grid_constant__ ProjectParams p) {
    __shared__ float Rt[9], t[3];
    const int b = blockIdx.y;
    if (threadIdx.x == 0) view_rows(p, b, Rt, t);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const float* q = p.pos + ((size_t)b * p.n + i) * p.pos_stride;
    const float w = p.pos_stride == 4 ? q[3] : 1.f;
    const float X = Rt[0] * q[0] + Rt[1] * q[1] + Rt[2] * q[2] + t[0] * w;
    const float Y = Rt[3] * q[0] + Rt[4] * q[1] + Rt[5] * q[2] + t[1] * w;
    const float Z = Rt[6] * q[0] + Rt[7] * q[1] + Rt[8] * q[2] + t[2] * w;
    const float div = fabsf(Z) > 0.f ? Z : 1.f;                     // nonzero_divide
    const float x = p.f * (X / div), y = p.f * (Y / div);
    const float px = x * p.sx_px + p.cx, py = y * p.sy_px + p.cy;
    float* o = p.px_coord + ((size_t)b * p.n + i) * 3;
    o[0] = px; o[1] = py; o[2] = -Z;
    if (p.px_idx) {
        const long long ix = (long long)rintf(px - 0.5f), iy = (long long)rintf(py - 0.5f);     // torch.round: half to even
        const bool outside = iy < 0 || ix < 0 || iy >= p.H || ix >= p.W;
        p.px_idx[(size_t)b * p.n + i] = outside ? (long long)p.W * p.H : iy * p.W + ix;
    }
}

__global__ void __launch_bounds__(256) k_project_surfels_backward(const __grid_constant__ ProjectParams p) {
    __shared__ float Rt[9], t[3];
    const int b = blockIdx.y;
    if (threadIdx.x == 0) view_rows(p, b, Rt, t);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const float* q = p.pos + ((size_t)b * p.n + i) * p.pos_stride;
    const float w = p.pos_stride == 4 ? q[3] : 1.f;
    const float X = Rt[0] * q[0] + Rt[1] * q[1] + Rt[2] * q[2] + t[0] * w;
    const float Y = Rt[3] * q[0] + Rt[4] * q[1] + Rt[5] * q[2] + t[1] * w;
    const float Z = Rt[6] * q[0] + Rt[7] * q[1] + Rt[8] * q[2] + t[2] * w;
    const bool nz = fabsf(Z) > 0.f;
    const float div = nz ? Z : 1.f;
    const float* g = p.g_px + ((size_t)b * p.n + i) * 3;
    const float gx = g[0] * p.sx_px * p.f, gy = g[1] * p.sy_px * p.f;        // d/d(X/div), d/d(Y/div)
    const float gX = gx / div, gY = gy / div;
    const float gZ = -g[2] + (nz ? -(gx * X + gy * Y) / (div * div) : 0.f);
    float* o = p.g_pos + ((size_t)b * p.n + i) * p.pos_stride;
    o[0] += Rt[0] * gX + Rt[3] * gY + Rt[6] * gZ;
    o[1] += Rt[1] * gX + Rt[4] * gY + Rt[7] * gZ;
    o[2] += Rt[2] * gX + Rt[5] * gY + Rt[8] * gZ;
    if (p.pos_stride == 4) o[3] += t[0] * gX + t[1] * gY + t[2] * gZ;
}

struct ScatterParams {
    int batch, n, channels, n_dst, mode;      // mode 0: mean (weight 1), 1: weighted blended OIT
    int use_depth, use_center_dist;
    float alpha0, inv_2s2, z_scale, eps;      // 1 / (2 pi sigma^2), 1 / (2 sigma^2), extinction, 1e-8 (OIT) / 0 (mean)
    const float* x; const long long* idx; const float* z; const float* cd2;
    float* out; float* denom; unsigned char* mask;
    const float* g_out; float* g_x; float* g_z; float* g_cd2;
};

__device__ __forceinline__ float scatter_weight(const ScatterParams& p, size_t e) {
    if (p.mode == 0) return 1.f;
    float w = 1.f;
    if (p.use_center_dist) w *= p.alpha0 * expf(-p.cd2[e] * p.inv_2s2);
    if (p.use_depth) w *= expf(-p.z_scale * p.z[e]);
    return w;
}

__global__ void __launch_bounds__(256) k_scatter_accum(const __grid_constant__ ScatterParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= p.n) return;
    const size_t e = (size_t)b * p.n + i;
    const long long dst = p.idx[e];
    if (dst < 0 || dst >= p.n_dst) return;                 // the "dump" index of out-of-view surfels
    const float w = scatter_weight(p, e);
    const float* x = p.x + e * p.channels;
    float* o = p.out + ((size_t)b * p.n_dst + dst) * p.channels;
    for (int c = 0; c < p.channels; ++c) atomicAdd(o + c, x[c] * w);
    atomicAdd(p.denom + (size_t)b * p.n_dst + dst, w);
}

// out = nonzero_divide(out, denom, eps) (utils.py:87-95); mask = (denom == 0)
__global__ void __launch_bounds__(256) k_scatter_normalize(const __grid_constant__ ScatterParams p) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)p.batch * p.n_dst;
    if (j >= total) return;
    const float d = p.denom[j];
    const float div = (fabsf(d) > 0.f ? d : 1.f) + p.eps;
    for (int c = 0; c < p.channels; ++c) {
        p.out[j * p.channels + c] = p.out[j * p.channels + c] / div;
        if (p.mask) p.mask[j * p.channels + c] = d == 0.f ? 1 : 0;
    }
}

// gather form: surfel i reads the gradient and the normalised value of its destination pixel
__global__ void __launch_bounds__(256) k_scatter_backward(const __grid_constant__ ScatterParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= p.n) return;
    const size_t e = (size_t)b * p.n + i;
    const long long dst = p.idx[e];
    if (dst < 0 || dst >= p.n_dst) return;
    const size_t j = (size_t)b * p.n_dst + dst;
    const float d = p.denom[j];
    if (!(fabsf(d) > 0.f)) return;
    const float inv = 1.f / (d + p.eps);
    const float w = scatter_weight(p, e);
    const float* x = p.x + e * p.channels;
    const float* go = p.g_out + j * p.channels;
    const float* o = p.out + j * p.channels;
    float g_w = 0.f;
    for (int c = 0; c < p.channels; ++c) {
        if (p.g_x) p.g_x[e * p.channels + c] += go[c] * w * inv;
        g_w += go[c] * (x[c] - o[c]) * inv;
    }
    if (p.mode == 1) {
        if (p.g_z && p.use_depth) p.g_z[e] += g_w * w * (-p.z_scale);
        if (p.g_cd2 && p.use_center_dist) p.g_cd2[e] += g_w * w * (-p.inv_2s2);
    }
}

// ---------------------------------------------------------------------------------------------------
// projection_renderer_differentiable_fast (projection_layer.py:170-279): every surfel is splatted onto the four pixels
// around its projection with bilinear weights, each of the four corner scatters blended with Weighted Blended OIT
// (weights alpha(centre distance) * exp(-z_scale depth)), the four normalised results summed; then a separable
// Gaussian blur.  The reference runs scatter_weighted_blended_oit eight to twelve times (rgb, soft mask, depth x four
// corners) over padded copies; here one kernel accumulates all of it, one normalises, one gather-form kernel does the
// backward w.r.t. the surfel data AND the pixel coordinates (x, y, depth), and the blur is its own adjoint.
//   accumulators per corner k and pixel: A_k (sum of weights), C_k[ch] (data), M_k (soft mask), D_k (depth)
// ---------------------------------------------------------------------------------------------------
struct BilinearParams {
    int batch, n, channels, W, H;
    int use_depth, use_center_dist, want_depth;
    float alpha0, inv_2s2, z_scale, eps;
    const float* px;          // [B, n, 3] pixel x, pixel y, depth
    const float* x;           // [B, n, ch]
    float* acc;               // [B, 4, P, ch + 3]: C_k[ch], A_k, M_k, D_k
    float* out;               // [B, P, ch]
    float* mask;              // [B, P]
    float* depth;             // [B, P] or null
    const float* g_out; const float* g_mask; const float* g_depth;     // backward inputs
    float* g_x; float* g_px;                                            // backward outputs (added)
};

struct BilinearSurfel { int ix, iy; float fx, fy, depth, s; };
__device__ __forceinline__ BilinearSurfel bilinear_surfel(const BilinearParams& p, size_t e) {
    BilinearSurfel b;
    const float* q = p.px + e * 3;
    const float ax = q[0] - 0.5f, ay = q[1] - 0.5f;
    const float flx = floorf(ax), fly = floorf(ay);
    b.ix = (int)flx; b.iy = (int)fly;
    b.fx = ax - flx; b.fy = ay - fly;
    b.depth = q[2];
    float s = 1.f;
    if (p.use_center_dist) s *= p.alpha0 * expf(-(b.fx * b.fx + b.fy * b.fy) * p.inv_2s2);
    if (p.use_depth) s *= expf(-p.z_scale * b.depth);
    b.s = s;
    return b;
}
// corner k = (dx, dy) in the reference's order (0,0), (0,1), (1,0), (1,1); weight and destination pixel (-1: outside)
__device__ __forceinline__ void bilinear_corner(const BilinearParams& p, const BilinearSurfel& b, int k, float* w, int* dst) {
    const int dx = k >> 1, dy = k & 1;
    *w = (dx ? b.fx : 1.f - b.fx) * (dy ? b.fy : 1.f - b.fy);
    const int cx = b.ix + dx, cy = b.iy + dy;
    *dst = (cx < 0 || cy < 0 || cx >= p.W || cy >= p.H) ? -1 : cy * p.W + cx;
}

__global__ void __launch_bounds__(256) k_bilinear_accum(const __grid_constant__ BilinearParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, bi = blockIdx.y;
    if (i >= p.n) return;
    const size_t e = (size_t)bi * p.n + i;
    const BilinearSurfel b = bilinear_surfel(p, e);
    const int P = p.W * p.H, stride = p.channels + 3;
    const float* x = p.x + e * p.channels;
    for (int k = 0; k < 4; ++k) {
        float w; int dst;
        bilinear_corner(p, b, k, &w, &dst);
        if (dst < 0) continue;
        float* a = p.acc + (((size_t)bi * 4 + k) * P + dst) * stride;
        const float ws = w * b.s;
        for (int c = 0; c < p.channels; ++c) atomicAdd(a + c, x[c] * ws);
        atomicAdd(a + p.channels, b.s);
        atomicAdd(a + p.channels + 1, ws);
        if (p.want_depth) atomicAdd(a + p.channels + 2, b.depth * ws);
    }
}

__global__ void __launch_bounds__(256) k_bilinear_normalize(const __grid_constant__ BilinearParams p) {
    const int P = p.W * p.H, stride = p.channels + 3;
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (size_t)p.batch * P) return;
    const int bi = (int)(j / P), pix = (int)(j - (size_t)bi * P);
    float m = 0.f, d = 0.f;
    for (int c = 0; c < p.channels; ++c) p.out[j * p.channels + c] = 0.f;
    for (int k = 0; k < 4; ++k) {
        const float* a = p.acc + (((size_t)bi * 4 + k) * P + pix) * stride;
        const float A = a[p.channels];
        const float inv = 1.f / ((fabsf(A) > 0.f ? A : 1.f) + p.eps);
        for (int c = 0; c < p.channels; ++c) p.out[j * p.channels + c] += a[c] * inv;
        m += a[p.channels + 1] * inv;
        d += a[p.channels + 2] * inv;
    }
    p.mask[j] = m;
    if (p.depth) p.depth[j] = d;
}

__global__ void __launch_bounds__(256) k_bilinear_backward(const __grid_constant__ BilinearParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, bi = blockIdx.y;
    if (i >= p.n) return;
    const size_t e = (size_t)bi * p.n + i;
    const BilinearSurfel b = bilinear_surfel(p, e);
    const int P = p.W * p.H, stride = p.channels + 3;
    const float* x = p.x + e * p.channels;
    float g_s = 0.f, g_fx = 0.f, g_fy = 0.f, g_depth = 0.f;
    for (int k = 0; k < 4; ++k) {
        float w; int dst;
        bilinear_corner(p, b, k, &w, &dst);
        if (dst < 0) continue;
        const float* a = p.acc + (((size_t)bi * 4 + k) * P + dst) * stride;
        const float A = a[p.channels];
        if (!(fabsf(A) > 0.f)) continue;
        const float inv = 1.f / (A + p.eps);
        const size_t j = (size_t)bi * P + dst;
        float g_w = 0.f, g_sk = 0.f;        // d/d(bilinear weight), d/d(s) through this corner
        for (int c = 0; c < p.channels; ++c) {
            const float go = p.g_out ? p.g_out[j * p.channels + c] : 0.f;
            if (p.g_x) p.g_x[e * p.channels + c] += go * w * b.s * inv;
            g_w += go * x[c];
            g_sk += go * (x[c] * w - a[c] * inv);
        }
        const float gm = p.g_mask ? p.g_mask[j] : 0.f;
        g_w += gm;
        g_sk += gm * (w - a[p.channels + 1] * inv);
        if (p.want_depth && p.g_depth) {
            const float gd = p.g_depth[j];
            g_w += gd * b.depth;
            g_sk += gd * (b.depth * w - a[p.channels + 2] * inv);
            g_depth += gd * w * b.s * inv;
        }
        g_w *= b.s * inv;
        g_s += g_sk * inv;
        const int dx = k >> 1, dy = k & 1;
        g_fx += g_w * (dx ? 1.f : -1.f) * (dy ? b.fy : 1.f - b.fy);
        g_fy += g_w * (dy ? 1.f : -1.f) * (dx ? b.fx : 1.f - b.fx);
    }
    if (p.use_center_dist) {
        const float g_cd2 = g_s * b.s * (-p.inv_2s2);
        g_fx += 2.f * b.fx * g_cd2;
        g_fy += 2.f * b.fy * g_cd2;
    }
    if (p.use_depth) g_depth += g_s * b.s * (-p.z_scale);
    if (p.g_px) {
        float* g = p.g_px + e * 3;
        g[0] += g_fx; g[1] += g_fy; g[2] += g_depth;
    }
}

// one pass of the separable Gaussian blur of projection_layer.py:156-168 (zero padding): along W (axis 0) or H (axis 1)
// of a [B, H, W, C] image; taps [-half, half], weights exp(-d^2 / (2 sigma^2)) normalised to sum 1
__global__ void __launch_bounds__(256) k_blur_pass(const float* __restrict__ in, float* __restrict__ out, int batch, int H, int W,
                                                   int C, int axis, int half, float inv_2s2, float norm) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)batch * H * W * C;
    if (j >= total) return;
    const int c = (int)(j % C);
    const int xq = (int)((j / C) % W), yq = (int)((j / ((size_t)C * W)) % H);
    const size_t base = j - ((size_t)yq * W + xq) * C - c;          // start of this batch element's image
    float acc = 0.f;
    for (int d = -half; d <= half; ++d) {
        const int xx = axis == 0 ? xq + d : xq, yy = axis == 0 ? yq : yq + d;
        if (xx < 0 || yy < 0 || xx >= W || yy >= H) continue;
        acc = fmaf(expf(-(float)(d * d) * inv_2s2) * norm, in[base + ((size_t)yy * W + xx) * C + c], acc);
    }
    out[j] = acc;
}
