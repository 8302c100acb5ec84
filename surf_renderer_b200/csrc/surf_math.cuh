// surf_math.cuh - per-primitive / per-pixel math of the render path, written once as host+device inline
// functions so the CUDA kernels (surf_kernels.cu) and the CPU emulation used by the "not gpu" tests
// (tests/emul/) evaluate the very same expressions.
//
// The "x*" helpers are individually rounded IEEE fp32 operations (no FMA contraction): the reference is a
// chain of separate torch elementwise kernels, each of which rounds, so the parity-critical
// values (t, P, normals, shading) are evaluated in the reference's operation order:
//   normalize            diffrend/torch/utils.py:66-67,87-95,135-139
//   plane/disk/triangle  diffrend/torch/utils.py:281-366
//   sphere               diffrend/torch/utils.py:238-278
//   shading              diffrend/torch/renderer.py:82-125, 318-340
// torch.mm with K=3 (n.d, utils.py:292) is an FMA chain in index order on the CPU build measured here.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SURF_HD __host__ __device__ __forceinline__
#else
#define SURF_HD inline
#endif

namespace surf {

constexpr float kEps = 1e-10f;          // utils.py:135 normalize(eps=1e-10)
constexpr float kMissSentinel = 1001.f; // utils.py:271,323,363

#if defined(__CUDA_ARCH__)
SURF_HD float xmul(float a, float b) { return __fmul_rn(a, b); }
SURF_HD float xadd(float a, float b) { return __fadd_rn(a, b); }
SURF_HD float xsub(float a, float b) { return __fsub_rn(a, b); }
SURF_HD float xdiv(float a, float b) { return __fdiv_rn(a, b); }
SURF_HD float xsqrt(float a) { return __fsqrt_rn(a); }
SURF_HD float xfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#else
// host build: compiled with -ffp-contract=off so each operator rounds once
SURF_HD float xmul(float a, float b) { return a * b; }
SURF_HD float xadd(float a, float b) { return a + b; }
SURF_HD float xsub(float a, float b) { return a - b; }
SURF_HD float xdiv(float a, float b) { return a / b; }
SURF_HD float xsqrt(float a) { return sqrtf(a); }
SURF_HD float xfma(float a, float b, float c) { return fmaf(a, b, c); }
#endif

// approximate reciprocal of the conservative filters: MUFU.RCP on the device, 1/x in the host emulation
SURF_HD float approx_rcp(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
struct Vec3 {
    float x, y, z;
};
SURF_HD Vec3 v3(float x, float y, float z) { Vec3 r; r.x = x; r.y = y; r.z = z; return r; }
SURF_HD Vec3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
SURF_HD Vec3 vsub(Vec3 a, Vec3 b) { return v3(xsub(a.x, b.x), xsub(a.y, b.y), xsub(a.z, b.z)); }
SURF_HD Vec3 vadd(Vec3 a, Vec3 b) { return v3(xadd(a.x, b.x), xadd(a.y, b.y), xadd(a.z, b.z)); }
SURF_HD Vec3 vscale(float s, Vec3 a) { return v3(xmul(s, a.x), xmul(s, a.y), xmul(s, a.z)); }
SURF_HD Vec3 vneg(Vec3 a) { return v3(-a.x, -a.y, -a.z); }
// torch.sum(a*b, dim=-1) over three contiguous elements: ((p0+p1)+p2), each product rounded
SURF_HD float dot_seq(Vec3 a, Vec3 b) { return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z)); }
// torch.mm with K=3: fma(a2,b2, fma(a1,b1, a0*b0))
SURF_HD float dot_mm(Vec3 a, Vec3 b) { return xfma(a.z, b.z, xfma(a.y, b.y, xmul(a.x, b.x))); }
SURF_HD Vec3 cross3(Vec3 a, Vec3 b) {
    return v3(xsub(xmul(a.y, b.z), xmul(a.z, b.y)), xsub(xmul(a.z, b.x), xmul(a.x, b.z)),
              xsub(xmul(a.x, b.y), xmul(a.y, b.x)));
}
// fast (contractable) helpers for gradient math, where only 1e-4 relative parity is required
SURF_HD float fdot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SURF_HD Vec3 faxpy(float s, Vec3 a, Vec3 b) { return v3(s * a.x + b.x, s * a.y + b.y, s * a.z + b.z); }

// utils.py:135-139 normalize(): length = sqrt(sum(u^2 + eps)); zero length divides by one. Returns length.
SURF_HD Vec3 unit_eps(Vec3 u, float* len_out) {
    float s = xadd(xadd(xadd(xmul(u.x, u.x), kEps), xadd(xmul(u.y, u.y), kEps)), xadd(xmul(u.z, u.z), kEps));
    float len = xsqrt(s);
    float div = (fabsf(len) > 0.f) ? len : 1.f;
    if (len_out) *len_out = div;
    return v3(xdiv(u.x, div), xdiv(u.y, div), xdiv(u.z, div));
}

// ---------------------------------------------------------------------------------------------------
// scene views (plain pointers; device pointers in kernels, host pointers in the emulation)
// ---------------------------------------------------------------------------------------------------
enum { KIND_DISK = 0, KIND_PLANE = 1, KIND_SPHERE = 2, KIND_TRIANGLE = 3 };
constexpr int kMaxSets = 8;

struct SetView {
    int kind, count, first;   // first = global index of this set's primitive 0 (concatenation order)
    int rec_off;              // offset (float4 units) of this set's packed filter records in the workspace
    const float* pos; int pos_stride;
    const float* normal; int normal_stride;
    const float* radius;
    const int* mat;
};

struct SceneView {
    int n_sets, total;
    SetView sets[kMaxSets];
    int n_lights; const float* light_pos; int light_pos_stride; const int* light_color_idx;
    const float* light_atten; const float* ambient;
    int n_colors; const float* colors;
    int n_materials; const float* albedo; const float* coeffs;
    const float* gamma;
};

// float4 records per primitive in the packed (filter) buffer
SURF_HD int rec_f4(int kind) { return kind == KIND_TRIANGLE ? 4 : (kind == KIND_DISK ? 2 : 1); }

struct CamState {           // produced by camera_setup(); lives in the workspace
    float R[9];             // camera-to-world rotation, row-major; columns = camera x, y, z axes
    float eye[3];
    float odir[3];          // orthographic: the single ray direction unit(at-eye)
    float sx, sy;           // (float)(w/2), (float)(h/2)   (utils.py:455-456)
    float neg_focal;        // (float)(-focal_length)
    float near_clip, far_clip, far_plus1;
    double step_x, step_y;  // linspace steps 2/(W-1), -2/(H-1)
    int W, H, proj;
    int bad_index;          // set by k_prep when a material_idx / color_idx is out of range (the kernels clamp; the
                            // reference's index_select raises IndexError - surf_check_indices reports it)
};

struct F4 { float x, y, z, w; };
SURF_HD F4 f4(float x, float y, float z, float w) { F4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

// ---------------------------------------------------------------------------------------------------
// camera (utils.py:402-427 lookat_rot_inv; :439-478 generate_rays)
// ---------------------------------------------------------------------------------------------------
SURF_HD void camera_setup(const float* eye, const float* at, const float* up, int proj, int W, int H,
                          double fovy, double focal, float near_clip, float far_clip, CamState* cs) {
    Vec3 e = ld3(eye), a = ld3(at), u = ld3(up);
    Vec3 zc = unit_eps(vsub(e, a), nullptr);
    Vec3 un = unit_eps(u, nullptr);
    Vec3 xc = unit_eps(cross3(un, zc), nullptr);
    Vec3 yc = cross3(zc, xc);
    cs->R[0] = xc.x; cs->R[1] = yc.x; cs->R[2] = zc.x;
    cs->R[3] = xc.y; cs->R[4] = yc.y; cs->R[5] = zc.y;
    cs->R[6] = xc.z; cs->R[7] = yc.z; cs->R[8] = zc.z;
    cs->eye[0] = e.x; cs->eye[1] = e.y; cs->eye[2] = e.z;
    Vec3 od = unit_eps(vsub(a, e), nullptr);
    cs->odir[0] = od.x; cs->odir[1] = od.y; cs->odir[2] = od.z;
    double h = tan(fovy / 2) * 2 * focal;
    double w = h * ((double)W / (double)H);
    cs->sx = (float)(w / 2);
    cs->sy = (float)(h / 2);
    cs->neg_focal = (float)(-1.0 * focal);
    cs->near_clip = near_clip; cs->far_clip = far_clip;
    cs->far_plus1 = (float)((double)far_clip + 1.0);
    cs->step_x = W > 1 ? 2.0 / (double)(W - 1) : 0.0;
    cs->step_y = H > 1 ? -2.0 / (double)(H - 1) : 0.0;
    cs->W = W; cs->H = H; cs->proj = proj;
    cs->bad_index = 0;
}

// screen-plane coordinates of flat pixel `pix` (row-major), np.linspace in float64 then f32 scaling
SURF_HD void pixel_xy(const CamState& cs, int pix, float* x, float* y) {
    int row = pix / cs.W, col = pix - row * cs.W;
    double gx = (cs.W > 1 && col == cs.W - 1) ? 1.0 : -1.0 + (double)col * cs.step_x;
    double gy = (cs.H > 1 && row == cs.H - 1) ? -1.0 : 1.0 + (double)row * cs.step_y;
    *x = xmul((float)gx, cs.sx);
    *y = xmul((float)gy, cs.sy);
}

// perspective ray direction of a pixel (unit length)
SURF_HD Vec3 pixel_ray_dir(const CamState& cs, int pix) {
    float x, y;
    pixel_xy(cs, pix, &x, &y);
    Vec3 c = v3(x, y, cs.neg_focal);
    Vec3 d = v3(dot_mm(v3(cs.R[0], cs.R[1], cs.R[2]), c), dot_mm(v3(cs.R[3], cs.R[4], cs.R[5]), c),
                dot_mm(v3(cs.R[6], cs.R[7], cs.R[8]), c));
    float n = xsqrt(xadd(xadd(xmul(d.x, d.x), xmul(d.y, d.y)), xmul(d.z, d.z)));
    return v3(xdiv(d.x, n), xdiv(d.y, n), xdiv(d.z, n));
}

// orthographic ray origin of a pixel: pose * (x, y, 0, 1), divided by w (=1)
SURF_HD Vec3 pixel_ray_origin_ortho(const CamState& cs, int pix) {
    float x, y;
    pixel_xy(cs, pix, &x, &y);
    // rows of [R | eye] times (x, y, 0, 1): mm K=4 fma chain
    Vec3 o;
    o.x = xfma(cs.eye[0], 1.f, xfma(cs.R[2], 0.f, xfma(cs.R[1], y, xmul(cs.R[0], x))));
    o.y = xfma(cs.eye[1], 1.f, xfma(cs.R[5], 0.f, xfma(cs.R[4], y, xmul(cs.R[3], x))));
    o.z = xfma(cs.eye[2], 1.f, xfma(cs.R[8], 0.f, xfma(cs.R[7], y, xmul(cs.R[6], x))));
    return o;
}

// ---------------------------------------------------------------------------------------------------
// per-primitive preparation: exact plane constants + conservative filter records
// ---------------------------------------------------------------------------------------------------
struct PlaneConst { Vec3 n; float dist; float len; };   // unit normal, p.n, |raw normal| (eps-regularised)

SURF_HD PlaneConst plane_const(Vec3 p, Vec3 nraw) {
    PlaneConst pc;
    pc.n = unit_eps(nraw, &pc.len);
    pc.dist = dot_seq(p, pc.n);                    // utils.py:290
    return pc;
}
// numerator of t for a ray origin o:  dist - n.o   (utils.py:297)
SURF_HD float plane_numer(const PlaneConst& pc, Vec3 o) { return xsub(pc.dist, dot_seq(pc.n, o)); }

SURF_HD float f_round_up(double v) {     // smallest-ish float >= v
    float f = (float)v;
    if ((double)f < v) f = nextafterf(f, INFINITY);
    return f;
}
SURF_HD double dlen(double x, double y, double z) { return sqrt(x * x + y * y + z * z); }

// DISK: A = (n, numer)  B = (o-c, -(r+slack)^2).  Filter: |(o-c) + t d|^2 + B.w <= 0 with t ~ numer * rcp(n.d)
// (the negated radius rides as the addend of the first FMA of the squared distance)
SURF_HD void prep_disk(Vec3 p, Vec3 nraw, float r, Vec3 o, F4* A, F4* B) {
    PlaneConst pc = plane_const(p, nraw);
    *A = f4(pc.n.x, pc.n.y, pc.n.z, plane_numer(pc, o));
    double ocx = (double)o.x - p.x, ocy = (double)o.y - p.y, ocz = (double)o.z - p.z;
    double scale = dlen(o.x, o.y, o.z) + dlen(p.x, p.y, p.z) + dlen(ocx, ocy, ocz) + fabs((double)r);
    double rs = fabs((double)r) + 2e-6 * scale;
    *B = f4((float)ocx, (float)ocy, (float)ocz, -f_round_up(rs * rs * (1.0 + 1e-6)));
}
SURF_HD void prep_plane(Vec3 p, Vec3 nraw, Vec3 o, F4* A) {
    PlaneConst pc = plane_const(p, nraw);
    *A = f4(pc.n.x, pc.n.y, pc.n.z, plane_numer(pc, o));
}
// SPHERE: S = (o-c, |o-c|^2 - r^2 - slack).  Filter: (oc.d)^2 - S.w >= 0
SURF_HD void prep_sphere(Vec3 c, float r, Vec3 o, F4* S) {
    double ocx = (double)o.x - c.x, ocy = (double)o.y - c.y, ocz = (double)o.z - c.z;
    double oc2 = ocx * ocx + ocy * ocy + ocz * ocz;
    double cc = oc2 - (double)r * r;
    double slack = 8e-6 * (oc2 + (double)r * r) + 1e-30;
    float w = (float)(cc - slack);
    if ((double)w > cc - slack) w = nextafterf(w, -INFINITY);
    *S = f4((float)ocx, (float)ocy, (float)ocz, w);
}
// TRIANGLE, line form (rays on lines through o, any sign of t - the shadow rays of a light):
// A = (n, numer); W_i = (n x e_i, (o - v_i).(n x e_i) + slack).  Filter: t (d.W_i) + W_i.w >= 0, t ~ numer * rcp(n.d)
SURF_HD void prep_triangle_line(Vec3 v0, Vec3 v1, Vec3 v2, Vec3 nraw, Vec3 o, F4* A, F4* W0, F4* W1, F4* W2) {
    PlaneConst pc = plane_const(v0, nraw);
    *A = f4(pc.n.x, pc.n.y, pc.n.z, plane_numer(pc, o));
    const Vec3 vs[3] = {v0, v1, v2};
    F4* outs[3] = {W0, W1, W2};
    double emax = 0.0, el[3];
    double e[3][3];
    for (int i = 0; i < 3; ++i) {
        const Vec3& a = vs[i];
        const Vec3& b = vs[(i + 1) % 3];
        e[i][0] = (double)b.x - a.x; e[i][1] = (double)b.y - a.y; e[i][2] = (double)b.z - a.z;
        el[i] = dlen(e[i][0], e[i][1], e[i][2]);
        emax = el[i] > emax ? el[i] : emax;
    }
    double nx = pc.n.x, ny = pc.n.y, nz = pc.n.z;
    for (int i = 0; i < 3; ++i) {
        double wx = ny * e[i][2] - nz * e[i][1];
        double wy = nz * e[i][0] - nx * e[i][2];
        double wz = nx * e[i][1] - ny * e[i][0];
        const Vec3& v = vs[i];
        double ovx = (double)o.x - v.x, ovy = (double)o.y - v.y, ovz = (double)o.z - v.z;
        double k = ovx * wx + ovy * wy + ovz * wz;
        double scale = dlen(o.x, o.y, o.z) + dlen(v.x, v.y, v.z) + dlen(ovx, ovy, ovz) + emax;
        double slack = 4e-6 * el[i] * scale + 1e-30;
        *outs[i] = f4((float)wx, (float)wy, (float)wz, f_round_up(k + slack));
    }
}
// TRIANGLE, edge-function form (camera rays from the common origin o, hits at t >= 0).  With P = o + t d and
// t = numer / (n.d), the edge function (P - v_i).(n x e_i) = k_i + t (d.W_i) times (n.d) is LINEAR in d:
//   c_i (n.d) = (numer W_i + k_i n) . d,   and sign(n.d) = sign(numer) =: sg for every hit with t > 0,
// so the inside test of edge i is ONE dot product, no division:  G_i . d + slack >= 0,  G_i = sg (numer W_i + k_i n).
// Record: A = (n, numer) (exact narrow phase, shading), G_i = (G_i.xyz, slack_i).  3 x 3 FMAs per test instead of
// 16 + a reciprocal.  `t_nonneg` false (a negative near plane admits hits behind the origin): always-pass records.
SURF_HD void prep_triangle(Vec3 v0, Vec3 v1, Vec3 v2, Vec3 nraw, Vec3 o, F4* A, F4* W0, F4* W1, F4* W2, bool t_nonneg = true) {
    PlaneConst pc = plane_const(v0, nraw);
    const float numer_f = plane_numer(pc, o);
    *A = f4(pc.n.x, pc.n.y, pc.n.z, numer_f);
    F4* outs[3] = {W0, W1, W2};
    if (!t_nonneg) {
        for (int i = 0; i < 3; ++i) *outs[i] = f4(0.f, 0.f, 0.f, 1.f);
        return;
    }
    const Vec3 vs[3] = {v0, v1, v2};
    double emax = 0.0, el[3];
    double e[3][3];
    for (int i = 0; i < 3; ++i) {
        const Vec3& a = vs[i];
        const Vec3& b = vs[(i + 1) % 3];
        e[i][0] = (double)b.x - a.x; e[i][1] = (double)b.y - a.y; e[i][2] = (double)b.z - a.z;
        el[i] = dlen(e[i][0], e[i][1], e[i][2]);
        emax = el[i] > emax ? el[i] : emax;
    }
    const double nx = pc.n.x, ny = pc.n.y, nz = pc.n.z;
    // numer in double from the same rounded normal (the fp32 numer of the exact path differs by rounding: covered by the slack)
    const double numer = ((double)v0.x - o.x) * nx + ((double)v0.y - o.y) * ny + ((double)v0.z - o.z) * nz;
    const double sg = numer > 0.0 ? 1.0 : (numer < 0.0 ? -1.0 : 0.0);
    for (int i = 0; i < 3; ++i) {
        const double wx = ny * e[i][2] - nz * e[i][1];
        const double wy = nz * e[i][0] - nx * e[i][2];
        const double wz = nx * e[i][1] - ny * e[i][0];
        const Vec3& v = vs[i];
        const double ovx = (double)o.x - v.x, ovy = (double)o.y - v.y, ovz = (double)o.z - v.z;
        const double k = ovx * wx + ovy * wy + ovz * wz;
        const double gx = sg * (numer * wx + k * nx), gy = sg * (numer * wy + k * ny), gz = sg * (numer * wz + k * nz);
        const double scale = dlen(o.x, o.y, o.z) + dlen(v.x, v.y, v.z) + dlen(ovx, ovy, ovz) + emax;
        // slack of the line form (in units of the edge function, |n.d| <= 1) + rounding of G and of the fp32 dot
        // product + the fp32 rounding of the exact path's own plane constants relative to these double ones
        const double slack = 4e-6 * el[i] * scale + 2e-6 * (fabs(numer) * el[i] + fabs(k)) + 1e-6 * dlen(gx, gy, gz) + 1e-30;
        *outs[i] = f4((float)gx, (float)gy, (float)gz, f_round_up(slack));
    }
}
// ---------------------------------------------------------------------------------------------------
// level-1 screen-space record (perspective camera): a circle on the image plane z = -focal that contains the
// projection of the primitive's bounding sphere.  Any exact hit point P of the primitive lies inside that
// sphere, and the pixel's image-plane point is the projection of P, so a pixel outside the circle cannot hit.
//   record = (-u, -v, -rho^2, 0):   flagged  <=>  (x_p - u)^2 + (y_p - v)^2 - rho^2 <= 0
// rho = f R sqrt(1 + k^2) / (zc - R),  k = |c_xy| / zc  bounds |proj(c + delta) - proj(c)| over |delta| <= R.
// Primitives that reach the camera plane (zc - R <= 0) and planes are always flagged (-rho^2 = -inf).
// ---------------------------------------------------------------------------------------------------
SURF_HD F4 screen_circle(const CamState& cs, double cx, double cy, double cz, double R) {
    // camera coordinates: q = R^T (c - eye); columns of cs.R are the camera axes
    const double ex = cx - (double)cs.eye[0], ey = cy - (double)cs.eye[1], ez = cz - (double)cs.eye[2];
    const double qx = cs.R[0] * ex + cs.R[3] * ey + cs.R[6] * ez;
    const double qy = cs.R[1] * ex + cs.R[4] * ey + cs.R[7] * ez;
    const double qz = cs.R[2] * ex + cs.R[5] * ey + cs.R[8] * ez;
    const double f = -(double)cs.neg_focal;
    const double dist = dlen(ex, ey, ez);
    const double scale = dlen(cs.eye[0], cs.eye[1], cs.eye[2]) + dlen(cx, cy, cz) + dist + R;
    const double Rs = R + 4e-6 * scale;                     // fp32 noise of the exact path's hit point
    const double zc = -qz;
    if (!(zc - Rs > 1e-4 * (Rs + dist)) || !(f > 0.0)) return f4(0.f, 0.f, -INFINITY, 0.f);
    const double u = f * qx / zc, v = f * qy / zc;
    const double k2 = (qx * qx + qy * qy) / (zc * zc);
    double rho = f * Rs * sqrt(1.0 + k2) / (zc - Rs);
    rho = rho * (1.0 + 1e-4) + 2e-6 * ((double)fabsf(cs.sx) + (double)fabsf(cs.sy) + fabs(u) + fabs(v) + f);
    return f4((float)(-u), (float)(-v), -f_round_up(rho * rho * (1.0 + 1e-6)), 0.f);
}

SURF_HD F4 prep_screen(const CamState& cs, const SetView& sv, int i) {
    if (sv.kind == KIND_PLANE) return f4(0.f, 0.f, -INFINITY, 0.f);
    if (sv.kind == KIND_TRIANGLE) {
        const float* f = sv.pos + (size_t)i * 3 * sv.pos_stride;
        const double gx = ((double)f[0] + f[sv.pos_stride] + f[2 * sv.pos_stride]) / 3.0;
        const double gy = ((double)f[1] + f[sv.pos_stride + 1] + f[2 * sv.pos_stride + 1]) / 3.0;
        const double gz = ((double)f[2] + f[sv.pos_stride + 2] + f[2 * sv.pos_stride + 2]) / 3.0;
        double R = 0.0;
        for (int k = 0; k < 3; ++k) {
            const float* v = f + k * sv.pos_stride;
            const double d = dlen(v[0] - gx, v[1] - gy, v[2] - gz);
            R = d > R ? d : R;
        }
        return screen_circle(cs, gx, gy, gz, R);
    }
    const float* c = sv.pos + (size_t)i * sv.pos_stride;
    return screen_circle(cs, c[0], c[1], c[2], fabs((double)sv.radius[i]));
}
// scalar form of the level-1 test for pixel image-plane coordinates (x, y)
SURF_HD bool screen_filter(const F4& C, float x, float y) {
    const float dx = x + C.x, dy = y + C.y;
    return fmaf(dx, dx, fmaf(dy, dy, C.z)) <= 0.f;
}

// ---------------------------------------------------------------------------------------------------
// origin-independent filter records for rays with per-ray origins (orthographic camera, shadow rays).
// `obound` bounds |origin| over the rays of the launch; it only enters the conservative slack.
//   DISK: A = (n, p.n)  B = (-c, -(r+slack)^2)     PLANE: A           SPHERE: S = (c, (r+slack)^2)
//   TRIANGLE: A = (n, v0.n), W_i = (n x e_i, -v_i.(n x e_i) + slack)   (test: P.W_i.xyz + W_i.w >= 0)
// ---------------------------------------------------------------------------------------------------
SURF_HD void prep_disk_rays(Vec3 p, Vec3 nraw, float r, float obound, F4* A, F4* B) {
    PlaneConst pc = plane_const(p, nraw);
    *A = f4(pc.n.x, pc.n.y, pc.n.z, pc.dist);
    double scale = 2.0 * (double)obound + 2.0 * dlen(p.x, p.y, p.z) + fabs((double)r);
    double rs = fabs((double)r) + 4e-6 * scale;
    *B = f4(-p.x, -p.y, -p.z, -f_round_up(rs * rs * (1.0 + 1e-6)));
}
SURF_HD void prep_plane_rays(Vec3 p, Vec3 nraw, F4* A) {
    PlaneConst pc = plane_const(p, nraw);
    *A = f4(pc.n.x, pc.n.y, pc.n.z, pc.dist);
}
SURF_HD void prep_sphere_rays(Vec3 c, float r, float obound, F4* S) {
    double scale = 2.0 * (double)obound + 2.0 * dlen(c.x, c.y, c.z) + fabs((double)r);
    double rs = fabs((double)r) + 8e-6 * scale;
    *S = f4(c.x, c.y, c.z, f_round_up(rs * rs * (1.0 + 1e-5) + 1e-6 * scale * scale));   // hb^2 - |oc|^2 cancels: 2^-22 |oc|^2 noise
}
SURF_HD void prep_triangle_rays(Vec3 v0, Vec3 v1, Vec3 v2, Vec3 nraw, float obound, F4* A, F4* W0, F4* W1, F4* W2) {
    PlaneConst pc = plane_const(v0, nraw);
    *A = f4(pc.n.x, pc.n.y, pc.n.z, pc.dist);
    const Vec3 vs[3] = {v0, v1, v2};
    F4* outs[3] = {W0, W1, W2};
    double e[3][3], el[3], emax = 0.0;
    for (int i = 0; i < 3; ++i) {
        const Vec3& a = vs[i];
        const Vec3& b = vs[(i + 1) % 3];
        e[i][0] = (double)b.x - a.x; e[i][1] = (double)b.y - a.y; e[i][2] = (double)b.z - a.z;
        el[i] = dlen(e[i][0], e[i][1], e[i][2]);
        emax = el[i] > emax ? el[i] : emax;
    }
    const double nx = pc.n.x, ny = pc.n.y, nz = pc.n.z;
    for (int i = 0; i < 3; ++i) {
        const double wx = ny * e[i][2] - nz * e[i][1];
        const double wy = nz * e[i][0] - nx * e[i][2];
        const double wz = nx * e[i][1] - ny * e[i][0];
        const Vec3& v = vs[i];
        const double k = -(v.x * wx + v.y * wy + v.z * wz);
        const double scale = 2.0 * (double)obound + 3.0 * dlen(v.x, v.y, v.z) + emax;
        *outs[i] = f4((float)wx, (float)wy, (float)wz, f_round_up(k + 8e-6 * el[i] * scale + 1e-30));
    }
}
SURF_HD bool disk_filter_rays(const F4& A, const F4& B, Vec3 o, Vec3 d) {
    float no = fmaf(A.z, o.z, fmaf(A.y, o.y, A.x * o.x));
    float b = fmaf(A.z, d.z, fmaf(A.y, d.y, A.x * d.x));
    float t = (A.w - no) * approx_rcp(b);
    float rx = fmaf(t, d.x, o.x) + B.x, ry = fmaf(t, d.y, o.y) + B.y, rz = fmaf(t, d.z, o.z) + B.z;
    return fmaf(rz, rz, fmaf(ry, ry, fmaf(rx, rx, B.w))) <= 0.f;
}
SURF_HD bool sphere_filter_rays(const F4& S, Vec3 o, Vec3 d) {
    float ox = o.x - S.x, oy = o.y - S.y, oz = o.z - S.z;
    float hb = fmaf(oz, d.z, fmaf(oy, d.y, ox * d.x));
    float oc2 = fmaf(oz, oz, fmaf(oy, oy, ox * ox));
    return fmaf(hb, hb, S.w - oc2) >= 0.f;
}
SURF_HD bool triangle_filter_rays(const F4& A, const F4& W0, const F4& W1, const F4& W2, Vec3 o, Vec3 d) {
    float no = fmaf(A.z, o.z, fmaf(A.y, o.y, A.x * o.x));
    float b = fmaf(A.z, d.z, fmaf(A.y, d.y, A.x * d.x));
    float t = (A.w - no) * approx_rcp(b);
    float px = fmaf(t, d.x, o.x), py = fmaf(t, d.y, o.y), pz = fmaf(t, d.z, o.z);
    float c0 = fmaf(W0.z, pz, fmaf(W0.y, py, fmaf(W0.x, px, W0.w)));
    float c1 = fmaf(W1.z, pz, fmaf(W1.y, py, fmaf(W1.x, px, W1.w)));
    float c2 = fmaf(W2.z, pz, fmaf(W2.y, py, fmaf(W2.x, px, W2.w)));
    return (c0 >= 0.f) & (c1 >= 0.f) & (c2 >= 0.f);
}

// ---------------------------------------------------------------------------------------------------
// coarse (conservative) filters - scalar forms; the intersection kernel has packed f32x2 versions of the
// disk filter.  `rcp` is an approximate reciprocal on the device (MUFU.RCP), 1/x on the host.
// ---------------------------------------------------------------------------------------------------
SURF_HD bool disk_filter(const F4& A, const F4& B, Vec3 d) {
    float b = fmaf(A.z, d.z, fmaf(A.y, d.y, A.x * d.x));
    float t = A.w * approx_rcp(b);
    float rx = fmaf(t, d.x, B.x), ry = fmaf(t, d.y, B.y), rz = fmaf(t, d.z, B.z);
    float e = fmaf(rz, rz, fmaf(ry, ry, fmaf(rx, rx, B.w)));
    return e <= 0.f;
}
SURF_HD bool sphere_filter(const F4& S, Vec3 d) {
    float hb = fmaf(S.z, d.z, fmaf(S.y, d.y, S.x * d.x));
    return fmaf(hb, hb, -S.w) >= 0.f;
}
// edge-function form (prep_triangle): three dot products
SURF_HD bool triangle_filter(const F4&, const F4& W0, const F4& W1, const F4& W2, Vec3 d) {
    float c0 = fmaf(W0.z, d.z, fmaf(W0.y, d.y, fmaf(W0.x, d.x, W0.w)));
    float c1 = fmaf(W1.z, d.z, fmaf(W1.y, d.y, fmaf(W1.x, d.x, W1.w)));
    float c2 = fmaf(W2.z, d.z, fmaf(W2.y, d.y, fmaf(W2.x, d.x, W2.w)));
    return (c0 >= 0.f) & (c1 >= 0.f) & (c2 >= 0.f);
}
// line form (prep_triangle_line)
SURF_HD bool triangle_filter_line(const F4& A, const F4& W0, const F4& W1, const F4& W2, Vec3 d) {
    float b = fmaf(A.z, d.z, fmaf(A.y, d.y, A.x * d.x));
    float t = A.w * approx_rcp(b);
    float c0 = fmaf(t, fmaf(W0.z, d.z, fmaf(W0.y, d.y, W0.x * d.x)), W0.w);
    float c1 = fmaf(t, fmaf(W1.z, d.z, fmaf(W1.y, d.y, W1.x * d.x)), W1.w);
    float c2 = fmaf(t, fmaf(W2.z, d.z, fmaf(W2.y, d.y, W2.x * d.x)), W2.w);
    return (c0 >= 0.f) & (c1 >= 0.f) & (c2 >= 0.f);
}

// ---------------------------------------------------------------------------------------------------
// exact tests, reference operation order.  `numer` = dist - n.o for this ray's origin.
// ---------------------------------------------------------------------------------------------------
SURF_HD float plane_t(Vec3 n, float numer, Vec3 d) { return xdiv(numer, dot_mm(n, d)); }   // utils.py:292-297
SURF_HD Vec3 ray_point(Vec3 o, float t, Vec3 d) { return vadd(o, vscale(t, d)); }          // utils.py:235
SURF_HD float sq3_seq(Vec3 r) { return xadd(xadd(xmul(r.x, r.x), xmul(r.y, r.y)), xmul(r.z, r.z)); }

SURF_HD bool disk_exact(Vec3 n, float numer, Vec3 c, float r, Vec3 o, Vec3 d, float* t) {
    *t = plane_t(n, numer, d);
    Vec3 P = ray_point(o, *t, d);
    return sq3_seq(vsub(P, c)) <= xmul(r, r);                                              // utils.py:319-322
}
SURF_HD float edge_side(Vec3 e, Vec3 rel, Vec3 n) {                                        // utils.py:74-84,358
    Vec3 c = v3(xsub(xmul(e.y, rel.z), xmul(e.z, rel.y)), xadd(xmul(-e.x, rel.z), xmul(e.z, rel.x)),
                xsub(xmul(e.x, rel.y), xmul(e.y, rel.x)));
    return dot_seq(c, n);
}
SURF_HD bool triangle_exact(Vec3 n, float numer, Vec3 v0, Vec3 v1, Vec3 v2, Vec3 o, Vec3 d, float* t) {
    *t = plane_t(n, numer, d);
    Vec3 P = ray_point(o, *t, d);
    bool a = edge_side(vsub(v1, v0), vsub(P, v0), n) >= 0.f;
    bool b = edge_side(vsub(v2, v1), vsub(P, v1), n) >= 0.f;
    bool c = edge_side(vsub(v0, v2), vsub(P, v2), n) >= 0.f;
    return a & b & c;
}
// returns true when the ray hits the sphere in front of the origin; *t = nearest non-negative root.
// (Both roots negative: the reference reports a data-dependent phantom hit, utils.py:265-267 - treated as
// a miss here)  On a miss *t = 1001 like the reference's masked distance.
SURF_HD bool sphere_exact(Vec3 c, float r, Vec3 o, Vec3 d, float* t) {
    Vec3 oc = vsub(o, c);
    float qa = sq3_seq(d);
    float qb = xmul(2.f, dot_seq(oc, d));
    float qc = xsub(sq3_seq(oc), xmul(r, r));
    float disc = xsub(xmul(qb, qb), xmul(xmul(4.f, qa), qc));
    *t = kMissSentinel;
    if (!(disc >= 0.f)) return false;
    float root = xsqrt(disc);
    float inv = xdiv(1.f, xmul(2.f, qa));
    float t1 = xmul(xsub(-qb, root), inv);
    float t2 = xmul(xadd(-qb, root), inv);
    bool ok1 = t1 >= 0.f, ok2 = t2 >= 0.f;
    if (!ok1 && !ok2) return false;
    *t = (ok1 && ok2) ? fminf(t1, t2) : (ok1 ? t1 : t2);
    return true;
}

// ---------------------------------------------------------------------------------------------------
// primitive access
// ---------------------------------------------------------------------------------------------------
SURF_HD int find_set(const SceneView& sc, int idx) {
    int s = 0;
#pragma unroll
    for (int k = 1; k < kMaxSets; ++k)
        if (k < sc.n_sets && idx >= sc.sets[k].first) s = k;
    return s;
}

struct Fragment {
    float t;        // distance the reference reports for (pixel, primitive): masked depth on a hit
    float t_pos;    // distance used for the hit point (unmasked plane distance for planar primitives)
    Vec3 P, n;      // hit point, shading normal
    int set, local, kind, mat;
};

// geometry the reference's gather (renderer.py:185-189) yields for primitive `idx` on ray (o, d)
SURF_HD Fragment fragment_at(const SceneView& sc, int idx, Vec3 o, Vec3 d) {
    Fragment f;
    f.set = find_set(sc, idx);
    const SetView& sv = sc.sets[f.set];
    f.local = idx - sv.first;
    f.kind = sv.kind;
    f.mat = sv.mat[f.local];
    if (sv.kind == KIND_SPHERE) {
        Vec3 c = ld3(sv.pos + (size_t)f.local * sv.pos_stride);
        sphere_exact(c, sv.radius[f.local], o, d, &f.t);
        f.t_pos = f.t;
        f.P = ray_point(o, f.t, d);
        f.n = unit_eps(vsub(f.P, c), nullptr);                                            // utils.py:275
    } else {
        size_t prow = (sv.kind == KIND_TRIANGLE) ? (size_t)f.local * 3 * sv.pos_stride
                                                 : (size_t)f.local * sv.pos_stride;
        Vec3 p = ld3(sv.pos + prow);
        PlaneConst pc = plane_const(p, ld3(sv.normal + (size_t)f.local * sv.normal_stride));
        f.t_pos = plane_t(pc.n, plane_numer(pc, o), d);
        f.t = f.t_pos;
        f.P = ray_point(o, f.t_pos, d);
        f.n = pc.n;
    }
    return f;
}

// ---------------------------------------------------------------------------------------------------
// shading (renderer.py:82-125).  Returns sum over lights of the per-light colour (ambient included per
// light) before masking / relu / tonemap.
// ---------------------------------------------------------------------------------------------------
struct ShadeFlags { int double_sided, use_quartic; int raw_view; };   // raw_view: the view vector is used unnormalised
                                                                     // (render_splats_NDC passes cam_dir = -frag_pos, renderer.py:437)

SURF_HD float pow_like_torch(float base, float e) { return powf(base, e); }

// everything the fragment shader derives for one light at one fragment, in the reference's op order; shared by
// the forward shader and the backward so that the relu gates (D > 0, S > 0) are the same bits in both.
struct LightEval {
    Vec3 L, R, inc;          // unit light direction, reflected direction, incident (-L)
    float dl, ddiv, d2, pw, den, att, s, D, S;   // D, S: raw (before sign flip / relu)
    bool dl_nz, den_nz;
};
SURF_HD LightEval eval_light(const SceneView& sc, int l, Vec3 P, Vec3 n, Vec3 V, ShadeFlags fl) {
    LightEval e;
    Vec3 Lv = vsub(ld3(sc.light_pos + (size_t)l * sc.light_pos_stride), P);
    e.dl = xsqrt(sq3_seq(Lv));
    e.dl_nz = fabsf(e.dl) > 0.f;
    e.ddiv = e.dl_nz ? e.dl : 1.f;
    e.L = v3(xdiv(Lv.x, e.ddiv), xdiv(Lv.y, e.ddiv), xdiv(Lv.z, e.ddiv));
    e.d2 = xmul(e.dl, e.dl);
    e.pw = fl.use_quartic ? xmul(e.d2, e.d2) : e.d2;
    const float* at = sc.light_atten + 3 * l;
    e.den = xadd(xadd(at[0], xmul(e.dl, at[1])), xmul(e.pw, at[2]));
    e.den_nz = fabsf(e.den) > 0.f;
    e.att = xdiv(1.f, e.den_nz ? e.den : 1.f);
    e.D = dot_seq(n, vscale(e.att, e.L));
    e.inc = vneg(e.L);
    e.s = dot_seq(e.inc, n);
    e.R = vadd(vscale(xmul(-2.f, e.s), n), e.inc);
    e.S = dot_seq(V, e.R);
    return e;
}
SURF_HD float facing_sign(Vec3 V, Vec3 n) {
    float dp = dot_seq(V, n);
    return dp > 0.f ? 1.f : (dp < 0.f ? -1.f : (dp == 0.f ? 0.f : dp));
}

SURF_HD int clamp_index(int v, int count) { return v < 0 ? 0 : (v >= count ? count - 1 : v); }

SURF_HD void shade_pixel(const SceneView& sc, Vec3 eye, Vec3 P, Vec3 n, int mat, ShadeFlags fl,
                         const float* visibility /* [L] or null */, float rgb[3]) {
    mat = clamp_index(mat, sc.n_materials);
    const float* A = sc.albedo + 3 * mat;
    const float kd = sc.coeffs[3 * mat + 0], ks = sc.coeffs[3 * mat + 1], sh = sc.coeffs[3 * mat + 2];
    Vec3 V = unit_eps(vsub(eye, P), nullptr);
    const float sg = fl.double_sided ? facing_sign(V, n) : 1.f;
    float acc[3] = {0.f, 0.f, 0.f};
    for (int l = 0; l < sc.n_lights; ++l) {
        LightEval e = eval_light(sc, l, P, n, V, fl);
        float D = e.D, S = e.S;
        if (fl.double_sided) { D = xmul(sg, D); S = xmul(sg, S); }
        D = D > 0.f ? D : 0.f;
        S = S > 0.f ? S : 0.f;
        float scal = xadd(xmul(kd, D), xmul(ks, pow_like_torch(S, sh)));
        const float* col = sc.colors + 3 * clamp_index(sc.light_color_idx[l], sc.n_colors);
        float vis = visibility ? visibility[l] : 1.f;
        for (int c = 0; c < 3; ++c) {
            float tint = xmul(col[c], A[c]);
            if (visibility) tint = xmul(tint, vis);
            float v = xadd(xmul(scal, tint), xmul(sc.ambient[c], A[c]));
            acc[c] = l == 0 ? v : xadd(acc[c], v);
        }
    }
    rgb[0] = acc[0]; rgb[1] = acc[1]; rgb[2] = acc[2];
}

// compositing (renderer.py:330-340): mask by hit, relu, gamma
SURF_HD void composite(const float lit[3], bool hit, const float* gamma, float out[3]) {
    for (int c = 0; c < 3; ++c) {
        float v = hit ? lit[c] : 0.f;
        v = v > 0.f ? v : 0.f;
        out[c] = gamma ? powf(v, gamma[0]) : v;
    }
}

// ---------------------------------------------------------------------------------------------------
// backward (derived from the reference graph).  `Sink` receives the additive gradient
// contributions; the kernel's sink reduces + atomically adds, the emulation's sink adds into doubles.
//   sink.albedo(m, c, v) .coeff(m, c, v) .light_pos(l, c, v) .atten(l, c, v) .color(row, c, v)
//   sink.ambient(c, v) .gamma(v)
//   sink.end_light(l, colour_row)     called once per light by every pixel (warp-uniform point)
//   sink.end_pixel(set, local, idx, m, g7)   g7[0..2] = d/d(pos | v0 | centre), [3..5] = d/d(normal),
//                                            [6] = d/d(radius) (sphere); called once by every pixel
// ---------------------------------------------------------------------------------------------------
struct PixelGrads {           // incoming output gradients at one pixel
    float image[3]; float depth; float pos[3]; float normal[3];
};

// backward of composite + shading at one fragment (P, n, material m): accumulates material / light / colour /
// ambient / gamma gradients into the sink and ADDS d/dP and d/dn to *gP, *gn.
template <class Sink>
SURF_HD void backward_shading(const SceneView& sc, Vec3 eye, Vec3 P, Vec3 n, int m, bool hit, ShadeFlags fl,
                              const float* visibility, const float g_image[3], Sink& sink, Vec3* gP_io, Vec3* gn_io) {
    // Control flow is kept uniform across pixels (no early exits around sink calls): the device sink
    // reduces across the warp inside end_light()/end_pixel(), so every lane must reach them.  Inactive
    // pixels compute on (finite-or-not) garbage and contribute selected zeros.
    Vec3 gP = *gP_io, gn = *gn_io;
    m = clamp_index(m, sc.n_materials);
    const float* A = sc.albedo + 3 * m;
    const float kd = sc.coeffs[3 * m + 0], ks = sc.coeffs[3 * m + 1], sh = sc.coeffs[3 * m + 2];
    // forward recompute of the composite to get dLoss/dI (renderer.py:330-340 backwards)
    float lit[3] = {0.f, 0.f, 0.f};
    if (hit) shade_pixel(sc, eye, P, n, m, fl, visibility, lit);
    float gI[3];
    bool active = false;
    float g_gamma = 0.f;
    for (int c = 0; c < 3; ++c) {
        float Ic = lit[c];
        gI[c] = 0.f;
        if (hit && Ic > 0.f) {
            if (sc.gamma) {
                float gm = sc.gamma[0];
                gI[c] = g_image[c] * gm * powf(Ic, gm - 1.f);
                if (g_image[c] != 0.f) g_gamma += g_image[c] * powf(Ic, gm) * logf(Ic);
            } else {
                gI[c] = g_image[c];
            }
        }
        active |= (gI[c] != 0.f);
    }
    sink.gamma(g_gamma);

    float sv_len;
    Vec3 Vv = vsub(eye, P);
    Vec3 V = unit_eps(Vv, &sv_len);
    float sg = fl.double_sided ? facing_sign(V, n) : 1.f;
    if (!(sg == sg)) sg = 0.f;
    Vec3 gV = v3(0.f, 0.f, 0.f);
    for (int l = 0; l < sc.n_lights; ++l) {
        const LightEval ev = eval_light(sc, l, P, n, V, fl);
        const Vec3 L = ev.L, inc = ev.inc, R = ev.R;
        const float dl = ev.dl, d2 = ev.d2, pw = ev.pw, att = ev.att, s = ev.s;
        const bool dl_nz = ev.dl_nz, den_nz = ev.den_nz;
        const float* at = sc.light_atten + 3 * l;
        // gate values are the forward's bits (xmul like shade_pixel) so relu' agrees with the forward relu
        const float Dsg = fl.double_sided ? xmul(sg, ev.D) : ev.D;
        const float Ssg = fl.double_sided ? xmul(sg, ev.S) : ev.S;
        float Dp = Dsg > 0.f ? Dsg : 0.f;
        float Sp = Ssg > 0.f ? Ssg : 0.f;
        float spec = powf(Sp, sh);
        float scal = kd * Dp + ks * spec;
        const int crow = clamp_index(sc.light_color_idx[l], sc.n_colors);
        const float* col = sc.colors + 3 * crow;
        float vis = visibility ? visibility[l] : 1.f;
        float g_scal = 0.f;
        for (int c = 0; c < 3; ++c) {
            float tint = col[c] * A[c] * vis;
            g_scal += gI[c] * tint;
            float g_tint = gI[c] * scal * vis;
            sink.color(crow, c, active ? g_tint * A[c] : 0.f);
            sink.albedo(m, c, active ? g_tint * col[c] + gI[c] * sc.ambient[c] : 0.f);
            sink.ambient(c, active ? gI[c] * A[c] : 0.f);
        }
        sink.coeff(m, 0, active ? g_scal * Dp : 0.f);
        sink.coeff(m, 1, active ? g_scal * spec : 0.f);
        float g_spec = g_scal * ks;
        float g_Sp = 0.f, g_sh = 0.f;
        if (Sp > 0.f) {
            g_sh = g_spec * spec * logf(Sp);
            if (sh != 0.f) g_Sp = g_spec * sh * powf(Sp, sh - 1.f);
        }
        sink.coeff(m, 2, active ? g_sh : 0.f);
        float g_D = (active && Dsg > 0.f) ? g_scal * kd * sg : 0.f;
        float g_S = (active && Ssg > 0.f) ? g_Sp * sg : 0.f;
        // D = n . (att L)
        Vec3 gL = vscale(g_D * att, n);
        float g_att = g_D * fdot(n, L);
        // S = V . R ;  R = (-2 s) n + inc ; s = inc . n ; inc = -L
        Vec3 gR = vscale(g_S, V);
        float g_s = -2.f * fdot(gR, n);
        Vec3 g_inc = faxpy(g_s, n, gR);
        gL = v3(gL.x - g_inc.x, gL.y - g_inc.y, gL.z - g_inc.z);
        // att = 1/den
        float g_den = den_nz ? -g_att * att * att : 0.f;
        float dpw = fl.use_quartic ? 4.f * d2 * dl : 2.f * dl;
        float g_dl = g_den * (at[1] + at[2] * dpw);
        // L = Lv / dl ; dl = |Lv|
        Vec3 gLv = gL;
        if (dl_nz) {
            float gl_dot = fdot(gL, L);
            gLv = v3((gL.x - gl_dot * L.x) / dl + g_dl * L.x, (gL.y - gl_dot * L.y) / dl + g_dl * L.y,
                     (gL.z - gl_dot * L.z) / dl + g_dl * L.z);
        }
        if (active) {
            gn = faxpy(g_D * att, L, gn);
            gV = faxpy(g_S, R, gV);
            gn = faxpy(-2.f * s, gR, gn);
            gn = faxpy(g_s, inc, gn);
            gP = v3(gP.x - gLv.x, gP.y - gLv.y, gP.z - gLv.z);
        }
        sink.atten(l, 0, active ? g_den : 0.f);
        sink.atten(l, 1, active ? g_den * dl : 0.f);
        sink.atten(l, 2, active ? g_den * pw : 0.f);
        sink.light_pos(l, 0, active ? gLv.x : 0.f);
        sink.light_pos(l, 1, active ? gLv.y : 0.f);
        sink.light_pos(l, 2, active ? gLv.z : 0.f);
        sink.end_light(l, crow);
    }
    if (active) {   // V = Vv / sv_len
        float gv_dot = fdot(gV, V);
        gP = v3(gP.x - (gV.x - gv_dot * V.x) / sv_len, gP.y - (gV.y - gv_dot * V.y) / sv_len,
                gP.z - (gV.z - gv_dot * V.z) / sv_len);
    }

    *gP_io = gP;
    *gn_io = gn;
}

template <class Sink>
SURF_HD void backward_pixel(const SceneView& sc, Vec3 eye, Vec3 o, Vec3 d, int idx, bool hit,
                            ShadeFlags fl, const float* visibility, const PixelGrads& g, Sink& sink) {
    Fragment f = fragment_at(sc, idx, o, d);
    const SetView& sv = sc.sets[f.set];
    Vec3 P = f.P, n = f.n;
    Vec3 gP = v3(g.pos[0], g.pos[1], g.pos[2]);
    Vec3 gn = v3(g.normal[0], g.normal[1], g.normal[2]);
    const int m = f.mat;
    backward_shading(sc, eye, P, n, m, hit, fl, visibility, g.image, sink, &gP, &gn);

    const float g_depth = hit ? g.depth : 0.f;
    float out7[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (f.kind == KIND_SPHERE) {
        Vec3 c = ld3(sv.pos + (size_t)f.local * sv.pos_stride);
        float r = sv.radius[f.local];
        Vec3 G = vsub(P, c);
        float sG;
        Vec3 nn = unit_eps(G, &sG);
        float gdot = fdot(gn, nn);
        Vec3 gG = v3((gn.x - gdot * nn.x) / sG, (gn.y - gdot * nn.y) / sG, (gn.z - gdot * nn.z) / sG);
        gP = v3(gP.x + gG.x, gP.y + gG.y, gP.z + gG.z);
        Vec3 gc = vneg(gG);
        float gr = 0.f;
        if (hit) {   // t depends on (c, r) only on a real hit; a masked miss has the constant t = 1001
            float gt = g_depth + fdot(gP, d);
            float Gd = fdot(G, d);
            if (Gd != 0.f) {
                gc = faxpy(gt / Gd, G, gc);
                gr = gt * r / Gd;
            }
        }
        out7[0] = gc.x; out7[1] = gc.y; out7[2] = gc.z; out7[6] = gr;
    } else {
        size_t prow = (sv.kind == KIND_TRIANGLE) ? (size_t)f.local * 3 * sv.pos_stride
                                                 : (size_t)f.local * sv.pos_stride;
        Vec3 p = ld3(sv.pos + prow);
        PlaneConst pc = plane_const(p, ld3(sv.normal + (size_t)f.local * sv.normal_stride));
        float b = dot_mm(pc.n, d);
        float gt = g_depth + fdot(gP, d);
        float ga = gt / b;
        float gb = -gt * f.t_pos / b;
        Vec3 pmo = v3(p.x - o.x, p.y - o.y, p.z - o.z);
        gn = faxpy(ga, pmo, gn);
        gn = faxpy(gb, d, gn);
        Vec3 gp = vscale(ga, pc.n);
        float gdot = fdot(gn, pc.n);
        out7[0] = gp.x; out7[1] = gp.y; out7[2] = gp.z;
        out7[3] = (gn.x - gdot * pc.n.x) / pc.len;
        out7[4] = (gn.y - gdot * pc.n.y) / pc.len;
        out7[5] = (gn.z - gdot * pc.n.z) / pc.len;
    }
    sink.end_pixel(f.set, f.local, idx, m, out7);
}

// ---------------------------------------------------------------------------------------------------
// narrow phase shared by the intersection kernel and the emulation: exact hit test of primitive `local`
// of set `sv` for the ray (o, d).  (n, numer) are the exact plane constants for this origin (taken from
// the packed record for the common-origin perspective case, recomputed per ray otherwise).
// ---------------------------------------------------------------------------------------------------
SURF_HD bool exact_hit(const SetView& sv, int local, Vec3 n, float numer, Vec3 o, Vec3 d,
                       float near_clip, float far_clip, float* t) {
    bool hit;
    if (sv.kind == KIND_DISK) {
        hit = disk_exact(n, numer, ld3(sv.pos + (size_t)local * sv.pos_stride), sv.radius[local], o, d, t);
    } else if (sv.kind == KIND_PLANE) {
        *t = plane_t(n, numer, d);
        hit = true;
    } else if (sv.kind == KIND_TRIANGLE) {
        const float* f = sv.pos + (size_t)local * 3 * sv.pos_stride;
        hit = triangle_exact(n, numer, ld3(f), ld3(f + sv.pos_stride), ld3(f + 2 * sv.pos_stride), o, d, t);
    } else {
        hit = sphere_exact(ld3(sv.pos + (size_t)local * sv.pos_stride), sv.radius[local], o, d, t);
    }
    return hit && (near_clip <= *t) && (*t <= far_clip);                      // renderer.py:179
}

// exact plane constants of a planar primitive for an arbitrary ray origin (orthographic / shadow rays)
SURF_HD void plane_consts_for_origin(const SetView& sv, int local, Vec3 o, Vec3* n, float* numer) {
    if (sv.kind == KIND_SPHERE) { *n = v3(0.f, 0.f, 0.f); *numer = 0.f; return; }
    size_t prow = (sv.kind == KIND_TRIANGLE) ? (size_t)local * 3 * sv.pos_stride : (size_t)local * sv.pos_stride;
    PlaneConst pc = plane_const(ld3(sv.pos + prow), ld3(sv.normal + (size_t)local * sv.normal_stride));
    *n = pc.n;
    *numer = plane_numer(pc, o);
}

// ---------------------------------------------------------------------------------------------------
// per-pixel resolve (renderer.py:183-189, 284-340): winner -> depth / nearest / pos / normal / image
// ---------------------------------------------------------------------------------------------------
struct PixelOut { float image[3]; float depth; float normal[3]; float pos[3]; long long nearest; };

SURF_HD PixelOut resolve_pixel(const SceneView& sc, const CamState& cs, Vec3 o, Vec3 d,
                               unsigned long long key, ShadeFlags fl, const float* visibility) {
    PixelOut po;
    const bool hit = key != 0xFFFFFFFFFFFFFFFFull;
    const int idx = hit ? (int)(key & 0xFFFFFFFFull) : 0;      // argmin of an all-miss column is 0
    Fragment f = fragment_at(sc, idx, o, d);
    po.nearest = idx;
    po.depth = hit ? f.t : cs.far_plus1;
    po.normal[0] = f.n.x; po.normal[1] = f.n.y; po.normal[2] = f.n.z;
    po.pos[0] = f.P.x; po.pos[1] = f.P.y; po.pos[2] = f.P.z;
    float lit[3] = {0.f, 0.f, 0.f};
    if (hit) shade_pixel(sc, v3(cs.eye[0], cs.eye[1], cs.eye[2]), f.P, f.n, f.mat, fl, visibility, lit);
    composite(lit, hit, sc.gamma, po.image);
    return po;
}

// ---------------------------------------------------------------------------------------------------
// shadow rays (renderer.py:291-314): from frag_pos + 0.1 L toward light l against every primitive; visible
// iff nothing is hit strictly inside (0, |L|), or the nearest such hit is the fragment's own primitive.
// ---------------------------------------------------------------------------------------------------
SURF_HD float shadow_visibility(const SceneView& sc, Vec3 P, int self_idx, int l) {
    Vec3 Lv = vsub(ld3(sc.light_pos + (size_t)l * sc.light_pos_stride), P);
    float dist = xsqrt(sq3_seq(Lv));                                       // norm_p(., 2)
    Vec3 L = v3(xdiv(Lv.x, dist), xdiv(Lv.y, dist), xdiv(Lv.z, dist));
    Vec3 so = vadd(P, vscale(0.1f, L));
    float best_t = kMissSentinel;
    int best = 0;
    for (int s = 0; s < sc.n_sets; ++s) {
        const SetView& sv = sc.sets[s];
        for (int i = 0; i < sv.count; ++i) {
            Vec3 nn; float numer, t;
            plane_consts_for_origin(sv, i, so, &nn, &numer);
            bool hit = exact_hit(sv, i, nn, numer, so, L, -INFINITY, INFINITY, &t);
            if (hit && t > 0.f && t < dist && t < best_t) { best_t = t; best = sv.first + i; }
        }
    }
    return ((best_t == kMissSentinel) || (best == self_idx)) ? 1.f : 0.f;
}

// ---------------------------------------------------------------------------------------------------
// render_splats_along_ray (renderer.py:537-751): one splat per pixel at camera-space depth z on the pixel's ray,
// shaded in camera coordinates (eye at the origin, lights transformed by the view matrix).
// ---------------------------------------------------------------------------------------------------
// light l in camera coordinates: row l of torch.mm(light_pos [L,4], Mcam^T) with Mcam = [R^T | -R^T eye]
SURF_HD Vec3 light_to_camera(const CamState& cs, const float* l4) {
    Vec3 out;
    float* o = &out.x;
    for (int i = 0; i < 3; ++i) {
        const Vec3 col = v3(cs.R[i], cs.R[3 + i], cs.R[6 + i]);
        const float ti = -dot_seq(col, v3(cs.eye[0], cs.eye[1], cs.eye[2]));
        o[i] = xfma(l4[3], ti, xfma(l4[2], col.z, xfma(l4[1], col.y, xmul(l4[0], col.x))));
    }
    return out;
}
// splat position from its depth: Z = -relu(-z), X = -Z x / f, Y = -Z y / f (renderer.py:566-580); depth = |P|
SURF_HD void splat_fragment(const CamState& cs, int pix, float z, Vec3* P, float* depth, float* px, float* py) {
    const float Z = z < 0.f ? z : -0.f;
    float x, y;
    pixel_xy(cs, pix, &x, &y);
    const float f = -cs.neg_focal;
    *P = v3(xdiv(xmul(-Z, x), f), xdiv(xmul(-Z, y), f), Z);
    *depth = xsqrt(sq3_seq(*P));
    *px = x; *py = y;
}
struct SplatOut { float image[3]; float depth; float pos[3]; };
// `pos_opt` (explicit camera-space position, used by the supersampled path) overrides the z-derived position
SURF_HD SplatOut splat_pixel_forward(const SceneView& sc, const CamState& cs, int pix, float z, const float* pos_opt,
                                     Vec3 n, int mat, ShadeFlags fl, const float* visibility) {
    SplatOut so;
    Vec3 P;
    float x, y;
    if (pos_opt) { P = ld3(pos_opt); so.depth = xsqrt(sq3_seq(P)); }
    else splat_fragment(cs, pix, z, &P, &so.depth, &x, &y);
    float lit[3];
    shade_pixel(sc, v3(0.f, 0.f, 0.f), P, n, mat, fl, visibility, lit);
    for (int c = 0; c < 3; ++c) so.image[c] = lit[c] > 0.f ? lit[c] : 0.f;          // relu, no tonemap (:741)
    so.pos[0] = P.x; so.pos[1] = P.y; so.pos[2] = P.z;
    return so;
}
// backward of one splat: returns d/dz (or d/dpos when the position was explicit) and d/dnormal; light / material
// gradients go to the sink
template <class Sink>
SURF_HD void splat_pixel_backward(const SceneView& sc, const CamState& cs, int pix, float z, const float* pos_opt, Vec3 n,
                                  int mat, ShadeFlags fl, const float* visibility, const PixelGrads& g, Sink& sink,
                                  float* gz, float gpos_out[3], float gn_out[3]) {
    Vec3 P;
    float depth, x = 0.f, y = 0.f;
    if (pos_opt) { P = ld3(pos_opt); depth = xsqrt(sq3_seq(P)); }
    else splat_fragment(cs, pix, z, &P, &depth, &x, &y);
    Vec3 gP = v3(g.pos[0], g.pos[1], g.pos[2]);
    Vec3 gn = v3(g.normal[0], g.normal[1], g.normal[2]);
    backward_shading(sc, v3(0.f, 0.f, 0.f), P, n, mat, true, fl, visibility, g.image, sink, &gP, &gn);
    if (depth > 0.f) gP = faxpy(g.depth / depth, P, gP);
    const float f = -cs.neg_focal;
    const float gZ = gP.z - gP.x * x / f - gP.y * y / f;
    *gz = (!pos_opt && z < 0.f) ? gZ : 0.f;
    gpos_out[0] = gP.x; gpos_out[1] = gP.y; gpos_out[2] = gP.z;
    gn_out[0] = gn.x; gn_out[1] = gn.y; gn_out[2] = gn.z;
    sink.end_splat(mat);
}

// order-preserving float -> uint32 map for the packed z-buffer key (t_key << 32 | primitive index);
// atomicMin on the key = nearest depth, lowest index on exact ties (torch.min(0) semantics)
SURF_HD uint32_t float_order_key(float t) {
    uint32_t b;
#if defined(__CUDA_ARCH__)
    b = __float_as_uint(t);
#else
    union { float f; uint32_t u; } cv; cv.f = t; b = cv.u;
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
SURF_HD float float_from_order_key(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(b);
#else
    union { float f; uint32_t u; } cv; cv.u = b; return cv.f;
#endif
}
constexpr unsigned long long kMissKey = 0xFFFFFFFFFFFFFFFFull;

}  // namespace surf
