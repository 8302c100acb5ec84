// surf_emul.cpp - CPU emulation of the kernels' per-pixel math, TEST INFRASTRUCTURE ONLY.
//
// Compiles surf_renderer_b200/csrc/surf_math.cuh (the expressions the CUDA kernels evaluate) with g++ and
// drives it with plain loops, so the "not gpu" test-suite can check - without a GPU - that
//   * the conservative filters never reject a pair the exact test accepts,
//   * the exact tests / resolve / shading reproduce the oracle,
//   * the analytic backward matches the oracle's autograd.
// It is not part of the product: libsurf_b200.so never links or calls it, and the package has no CPU path.
// Build: g++ -O2 -ffp-contract=off -shared -fPIC (see tests/emul/build.py).
#include <cstring>
#include <string>
#include <vector>

#include "../../surf_renderer_b200/csrc/surf_view.h"

using namespace surf;

namespace {

struct Packed {
    std::vector<F4> rec;
};

// line_form: the records of a light origin (shadow rays: t of either sign), like k_prep_lights; else the camera's
void pack_all(const SceneView& sc, Vec3 o, Packed* pk, bool line_form = false, bool t_nonneg = true) {
    pk->rec.assign(packed_f4_total(sc), f4(0, 0, 0, 0));
    for (int s = 0; s < sc.n_sets; ++s) {
        const SetView& sv = sc.sets[s];
        for (int i = 0; i < sv.count; ++i) {
            F4* r = &pk->rec[sv.rec_off + (size_t)i * rec_f4(sv.kind)];
            if (sv.kind == KIND_DISK) {
                prep_disk(ld3(sv.pos + (size_t)i * sv.pos_stride), ld3(sv.normal + (size_t)i * sv.normal_stride),
                          sv.radius[i], o, r, r + 1);
            } else if (sv.kind == KIND_PLANE) {
                prep_plane(ld3(sv.pos + (size_t)i * sv.pos_stride), ld3(sv.normal + (size_t)i * sv.normal_stride), o, r);
            } else if (sv.kind == KIND_SPHERE) {
                prep_sphere(ld3(sv.pos + (size_t)i * sv.pos_stride), sv.radius[i], o, r);
            } else {
                const float* f = sv.pos + (size_t)i * 3 * sv.pos_stride;
                if (line_form)
                    prep_triangle_line(ld3(f), ld3(f + sv.pos_stride), ld3(f + 2 * sv.pos_stride),
                                       ld3(sv.normal + (size_t)i * sv.normal_stride), o, r, r + 1, r + 2, r + 3);
                else
                    prep_triangle(ld3(f), ld3(f + sv.pos_stride), ld3(f + 2 * sv.pos_stride),
                                  ld3(sv.normal + (size_t)i * sv.normal_stride), o, r, r + 1, r + 2, r + 3, t_nonneg);
            }
        }
    }
}

void pack_all_rays(const SceneView& sc, float obound, Packed* pk) {
    pk->rec.assign(packed_f4_total(sc), f4(0, 0, 0, 0));
    for (int s = 0; s < sc.n_sets; ++s) {
        const SetView& sv = sc.sets[s];
        for (int i = 0; i < sv.count; ++i) {
            F4* r = &pk->rec[sv.rec_off + (size_t)i * rec_f4(sv.kind)];
            if (sv.kind == KIND_DISK)
                prep_disk_rays(ld3(sv.pos + (size_t)i * sv.pos_stride), ld3(sv.normal + (size_t)i * sv.normal_stride), sv.radius[i], obound, r, r + 1);
            else if (sv.kind == KIND_PLANE)
                prep_plane_rays(ld3(sv.pos + (size_t)i * sv.pos_stride), ld3(sv.normal + (size_t)i * sv.normal_stride), r);
            else if (sv.kind == KIND_SPHERE)
                prep_sphere_rays(ld3(sv.pos + (size_t)i * sv.pos_stride), sv.radius[i], obound, r);
            else {
                const float* f = sv.pos + (size_t)i * 3 * sv.pos_stride;
                prep_triangle_rays(ld3(f), ld3(f + sv.pos_stride), ld3(f + 2 * sv.pos_stride),
                                   ld3(sv.normal + (size_t)i * sv.normal_stride), obound, r, r + 1, r + 2, r + 3);
            }
        }
    }
}

bool filter_pass_rays(const SetView& sv, const F4* r, Vec3 o, Vec3 d) {
    switch (sv.kind) {
        case KIND_DISK: return disk_filter_rays(r[0], r[1], o, d);
        case KIND_PLANE: return true;
        case KIND_SPHERE: return sphere_filter_rays(r[0], o, d);
        default: return triangle_filter_rays(r[0], r[1], r[2], r[3], o, d);
    }
}

bool filter_pass(const SetView& sv, const F4* r, Vec3 d, bool line_form = false) {
    switch (sv.kind) {
        case KIND_DISK: return disk_filter(r[0], r[1], d);
        case KIND_PLANE: return true;
        case KIND_SPHERE: return sphere_filter(r[0], d);
        default: return line_form ? triangle_filter_line(r[0], r[1], r[2], r[3], d) : triangle_filter(r[0], r[1], r[2], r[3], d);
    }
}

struct HostSink {
    std::vector<std::vector<double>> g_prim;   // per set: count*7
    std::vector<double> g_albedo, g_coeff, g_lpos, g_atten, g_color, g_amb;
    double g_gamma = 0;
    explicit HostSink(const SceneView& s) {
        g_prim.resize(s.n_sets);
        for (int k = 0; k < s.n_sets; ++k) g_prim[k].assign((size_t)s.sets[k].count * 7, 0.0);
        g_albedo.assign(s.n_materials * 3, 0); g_coeff.assign(s.n_materials * 3, 0);
        g_lpos.assign(s.n_lights * 3, 0); g_atten.assign(s.n_lights * 3, 0);
        g_color.assign(s.n_colors * 3, 0); g_amb.assign(3, 0);
    }
    // the interface backward_pixel() writes to
    void end_light(int, int) {}
    void end_splat(int) {}
    void end_pixel(int set, int local, int, int, const float* g7) {
        for (int k = 0; k < 7; ++k) g_prim[set][(size_t)local * 7 + k] += g7[k];
    }
    void albedo(int m, int c, float v) { g_albedo[m * 3 + c] += v; }
    void coeff(int m, int c, float v) { g_coeff[m * 3 + c] += v; }
    void light_pos(int l, int c, float v) { g_lpos[l * 3 + c] += v; }
    void atten(int l, int c, float v) { g_atten[l * 3 + c] += v; }
    void color(int r, int c, float v) { g_color[r * 3 + c] += v; }
    void ambient(int c, float v) { g_amb[c] += v; }
    void gamma(float v) { g_gamma += v; }
};

void pixel_ray(const CamState& cs, int pix, Vec3* o, Vec3* d) {
    if (cs.proj == 0) {
        *o = v3(cs.eye[0], cs.eye[1], cs.eye[2]);
        *d = pixel_ray_dir(cs, pix);
    } else {
        *o = pixel_ray_origin_ortho(cs, pix);
        *d = v3(cs.odir[0], cs.odir[1], cs.odir[2]);
    }
}

thread_local std::string g_err;

}  // namespace

extern "C" {

const char* emul_last_error() { return g_err.c_str(); }

// Forward.  Returns the number of (pixel, primitive) pairs the exact test accepted but the conservative
// filter rejected (must be 0), or <0 on error.  keys_out (optional) receives the packed z-buffer keys.
long long emul_forward(const SurfScene* scene, const SurfCamera* cam, const SurfOptions* opt,
                       const SurfOutputs* out, unsigned long long* keys_out) {
    SceneView sc;
    if (!build_scene_view(*scene, &sc, &g_err) || !check_camera(*cam, &g_err)) return -1;
    CamState cs;
    camera_setup(cam->eye, cam->at, cam->up, cam->proj, cam->width, cam->height, cam->fovy, cam->focal_length,
                 cam->near_clip, cam->far_clip, &cs);
    const int N = cam->width * cam->height;
    int p0 = opt->pixel_begin, p1 = opt->pixel_end;
    if (p0 == 0 && p1 == 0) p1 = N;
    ShadeFlags fl = {opt->double_sided, opt->use_quartic};
    Packed pk;
    const Vec3 eye = v3(cs.eye[0], cs.eye[1], cs.eye[2]);
    if (cs.proj == 0) pack_all(sc, eye, &pk, false, cs.near_clip >= 0.f);
    else {
        float ob = 0.f;
        for (int pix = p0; pix < p1; ++pix) {
            Vec3 oo = pixel_ray_origin_ortho(cs, pix);
            ob = fmaxf(ob, sqrtf(oo.x * oo.x + oo.y * oo.y + oo.z * oo.z));
        }
        pack_all_rays(sc, ob, &pk);
    }
    long long filter_misses = 0;
    std::vector<F4> circ;
    if (cs.proj == 0) {
        circ.resize(sc.total);
        for (int s = 0; s < sc.n_sets; ++s)
            for (int i = 0; i < sc.sets[s].count; ++i) circ[sc.sets[s].first + i] = prep_screen(cs, sc.sets[s], i);
    }
    for (int pix = p0; pix < p1; ++pix) {
        Vec3 o, d;
        pixel_ray(cs, pix, &o, &d);
        float px = 0.f, py = 0.f;
        pixel_xy(cs, pix, &px, &py);
        float best_t = INFINITY;
        int best = -1;
        for (int s = 0; s < sc.n_sets; ++s) {
            const SetView& sv = sc.sets[s];
            for (int i = 0; i < sv.count; ++i) {
                Vec3 n; float numer; bool pass = true;
                if (cs.proj == 0) {
                    const F4* r = &pk.rec[sv.rec_off + (size_t)i * rec_f4(sv.kind)];
                    pass = filter_pass(sv, r, d) && screen_filter(circ[sv.first + i], px, py);
                    n = v3(r[0].x, r[0].y, r[0].z); numer = r[0].w;
                } else {
                    plane_consts_for_origin(sv, i, o, &n, &numer);
                    pass = filter_pass_rays(sv, &pk.rec[sv.rec_off + (size_t)i * rec_f4(sv.kind)], o, d);
                }
                float t;
                bool hit = exact_hit(sv, i, n, numer, o, d, cs.near_clip, cs.far_clip, &t);
                if (hit && !pass) ++filter_misses;
                if (hit && t < best_t) { best_t = t; best = sv.first + i; }
            }
        }
        unsigned long long key = best < 0 ? kMissKey
                                          : (((unsigned long long)float_order_key(best_t)) << 32) | (unsigned)best;
        const int k = pix - p0;
        if (keys_out) keys_out[k] = key;
        float vis[16];
        const float* visp = nullptr;
        if (opt->shadow) {
            const int self = best < 0 ? 0 : best;
            Fragment fr = fragment_at(sc, self, o, d);
            for (int l = 0; l < sc.n_lights && l < 16; ++l) vis[l] = shadow_visibility(sc, fr.P, self, l);
            visp = vis;
        }
        PixelOut po = resolve_pixel(sc, cs, o, d, key, fl, visp);
        if (out->image) memcpy(out->image + 3 * (size_t)k, po.image, 12);
        if (out->depth) out->depth[k] = po.depth;
        if (out->normal) memcpy(out->normal + 3 * (size_t)k, po.normal, 12);
        if (out->pos) memcpy(out->pos + 3 * (size_t)k, po.pos, 12);
        if (out->nearest) out->nearest[k] = po.nearest;
        if (out->ray_dir) {
            const size_t n = (size_t)(p1 - p0);
            if (cs.proj == 0) { out->ray_dir[k] = d.x; out->ray_dir[n + k] = d.y; out->ray_dir[2 * n + k] = d.z; }
            else if (k == 0) { out->ray_dir[0] = d.x; out->ray_dir[1] = d.y; out->ray_dir[2] = d.z; }
        }
    }
    return filter_misses;
}

int emul_backward(const SurfScene* scene, const SurfCamera* cam, const SurfOptions* opt, const int64_t* nearest,
                  const float* depth, const SurfOutGrads* og, const SurfSceneGrads* sg) {
    SceneView sc;
    if (!build_scene_view(*scene, &sc, &g_err) || !check_camera(*cam, &g_err)) return -1;
    CamState cs;
    camera_setup(cam->eye, cam->at, cam->up, cam->proj, cam->width, cam->height, cam->fovy, cam->focal_length,
                 cam->near_clip, cam->far_clip, &cs);
    const int N = cam->width * cam->height;
    int p0 = opt->pixel_begin, p1 = opt->pixel_end;
    if (p0 == 0 && p1 == 0) p1 = N;
    ShadeFlags fl = {opt->double_sided, opt->use_quartic};
    HostSink hs(sc);
    HostSink& sink = hs;
    const Vec3 eye = v3(cs.eye[0], cs.eye[1], cs.eye[2]);
    for (int pix = p0; pix < p1; ++pix) {
        const int k = pix - p0;
        Vec3 o, d;
        pixel_ray(cs, pix, &o, &d);
        PixelGrads g;
        for (int c = 0; c < 3; ++c) {
            g.image[c] = og->image ? og->image[3 * (size_t)k + c] : 0.f;
            g.pos[c] = og->pos ? og->pos[3 * (size_t)k + c] : 0.f;
            g.normal[c] = og->normal ? og->normal[3 * (size_t)k + c] : 0.f;
        }
        g.depth = og->depth ? og->depth[k] : 0.f;
        const bool hit = depth[k] <= cs.far_clip && depth[k] >= cs.near_clip;
        float vis[16];
        const float* visp = nullptr;
        if (opt->shadow) {
            Fragment fr = fragment_at(sc, (int)nearest[k], o, d);
            for (int l = 0; l < sc.n_lights && l < 16; ++l) vis[l] = shadow_visibility(sc, fr.P, (int)nearest[k], l);
            visp = vis;
        }
        backward_pixel(sc, eye, o, d, (int)nearest[k], hit, fl, visp, g, sink);
    }
    for (int s = 0; s < sc.n_sets; ++s) {
        const SetView& sv = sc.sets[s];
        const SurfPrimSetGrads& pg = sg->sets[s];
        for (int i = 0; i < sv.count; ++i) {
            const double* a = &hs.g_prim[s][(size_t)i * 7];
            if (pg.pos) {
                size_t row = sv.kind == KIND_TRIANGLE ? (size_t)i * 3 * sv.pos_stride : (size_t)i * sv.pos_stride;
                for (int c = 0; c < 3; ++c) pg.pos[row + c] += (float)a[c];
            }
            if (pg.normal && sv.kind != KIND_SPHERE)
                for (int c = 0; c < 3; ++c) pg.normal[(size_t)i * sv.normal_stride + c] += (float)a[3 + c];
            if (pg.radius && sv.kind == KIND_SPHERE) pg.radius[i] += (float)a[6];
        }
    }
    for (int m = 0; m < sc.n_materials * 3; ++m) {
        if (sg->albedo) sg->albedo[m] += (float)hs.g_albedo[m];
        if (sg->coeffs) sg->coeffs[m] += (float)hs.g_coeff[m];
    }
    for (int l = 0; l < sc.n_lights; ++l)
        for (int c = 0; c < 3; ++c) {
            if (sg->light_pos) sg->light_pos[(size_t)l * sc.light_pos_stride + c] += (float)hs.g_lpos[l * 3 + c];
            if (sg->light_attenuation) sg->light_attenuation[l * 3 + c] += (float)hs.g_atten[l * 3 + c];
        }
    for (int r = 0; r < sc.n_colors * 3; ++r)
        if (sg->colors) sg->colors[r] += (float)hs.g_color[r];
    for (int c = 0; c < 3; ++c)
        if (sg->ambient) sg->ambient[c] += (float)hs.g_amb[c];
    if (sg->gamma) sg->gamma[0] += (float)hs.g_gamma;
    return 0;
}

// Counts shadow-ray (pixel, light, primitive) triples whose exact test reports an occluder but whose conservative
// per-ray-origin filter rejects the pair (must be 0).  `keys` are the z-buffer keys emul_forward produced.
long long emul_shadow_filter_misses(const SurfScene* scene, const SurfCamera* cam, const unsigned long long* keys) {
    SceneView sc;
    if (!build_scene_view(*scene, &sc, &g_err) || !check_camera(*cam, &g_err)) return -1;
    CamState cs;
    camera_setup(cam->eye, cam->at, cam->up, cam->proj, cam->width, cam->height, cam->fovy, cam->focal_length,
                 cam->near_clip, cam->far_clip, &cs);
    const int N = cam->width * cam->height;
    long long misses = 0;
    for (int l = 0; l < sc.n_lights; ++l) {
        std::vector<Vec3> so(N), dir(N);
        std::vector<float> tmax(N, 0.f);
        float ob = 0.f;
        for (int pix = 0; pix < N; ++pix) {
            if (keys[pix] == kMissKey) continue;
            Vec3 o, d;
            pixel_ray(cs, pix, &o, &d);
            Fragment f = fragment_at(sc, (int)(keys[pix] & 0xFFFFFFFFull), o, d);
            Vec3 Lv = vsub(ld3(sc.light_pos + (size_t)l * sc.light_pos_stride), f.P);
            float dist = xsqrt(sq3_seq(Lv));
            dir[pix] = v3(xdiv(Lv.x, dist), xdiv(Lv.y, dist), xdiv(Lv.z, dist));
            so[pix] = vadd(f.P, vscale(0.1f, dir[pix]));
            tmax[pix] = dist;
            ob = fmaxf(ob, sqrtf(so[pix].x * so[pix].x + so[pix].y * so[pix].y + so[pix].z * so[pix].z));
        }
        Packed pk, pkl;
        pack_all_rays(sc, ob, &pk);
        pack_all(sc, ld3(sc.light_pos + (size_t)l * sc.light_pos_stride), &pkl, true);
        for (int pix = 0; pix < N; ++pix) {
            if (keys[pix] == kMissKey) continue;
            for (int s = 0; s < sc.n_sets; ++s) {
                const SetView& sv = sc.sets[s];
                for (int i = 0; i < sv.count; ++i) {
                    Vec3 nn; float numer, t;
                    plane_consts_for_origin(sv, i, so[pix], &nn, &numer);
                    bool hit = exact_hit(sv, i, nn, numer, so[pix], dir[pix], -INFINITY, INFINITY, &t) && t > 0.f && t < tmax[pix];
                    if (hit && !filter_pass_rays(sv, &pk.rec[sv.rec_off + (size_t)i * rec_f4(sv.kind)], so[pix], dir[pix])) ++misses;
                    // k_intersect_shadow's filter: the camera-style filters with the LIGHT as the common origin and -L as
                    // the ray direction (records as k_prep_lights prepares them)
                    if (hit && !filter_pass(sv, &pkl.rec[sv.rec_off + (size_t)i * rec_f4(sv.kind)],
                                            v3(-dir[pix].x, -dir[pix].y, -dir[pix].z), true)) ++misses;
                }
            }
        }
    }
    return misses;
}

// ---- render_splats_along_ray emulation -----------------------------------------------------------
struct EmulSplats { int count; const float* z; int z_stride; const float* normal; int normal_stride; const int* mat;
                    const float* vis; const float* pos; };

static bool splat_setup(const SurfScene* scene, const SurfCamera* cam, SceneView* sc, CamState* cs, std::vector<float>* lcc) {
    SurfScene tmp = *scene;
    if (!build_scene_view(tmp, sc, &g_err, true) || !check_camera(*cam, &g_err)) return false;
    if (scene->light_pos_stride != 4) { g_err = "along-ray lights must be homogeneous [L,4]"; return false; }
    camera_setup(cam->eye, cam->at, cam->up, 0, cam->width, cam->height, cam->fovy, cam->focal_length,
                 cam->near_clip, cam->far_clip, cs);
    lcc->resize((size_t)sc->n_lights * 3);
    for (int l = 0; l < sc->n_lights; ++l) {
        Vec3 v = light_to_camera(*cs, scene->light_pos + 4 * (size_t)l);
        (*lcc)[3 * l] = v.x; (*lcc)[3 * l + 1] = v.y; (*lcc)[3 * l + 2] = v.z;
    }
    sc->light_pos = lcc->data(); sc->light_pos_stride = 3; sc->gamma = nullptr;
    return true;
}

int emul_splats_forward(const SurfScene* scene, const SurfCamera* cam, const SurfOptions* opt, const EmulSplats* sp,
                        const SurfOutputs* out) {
    SceneView sc; CamState cs; std::vector<float> lcc;
    if (!splat_setup(scene, cam, &sc, &cs, &lcc)) return -1;
    ShadeFlags fl = {0, opt->use_quartic};
    std::vector<float> vis(sc.n_lights);
    for (int k = 0; k < sp->count; ++k) {
        if (sp->vis) for (int l = 0; l < sc.n_lights; ++l) vis[l] = sp->vis[(size_t)l * sp->count + k];
        const float* nn = sp->normal + (size_t)k * sp->normal_stride;
        SplatOut so = splat_pixel_forward(sc, cs, k, sp->pos ? 0.f : sp->z[(size_t)k * sp->z_stride],
                                          sp->pos ? sp->pos + 3 * (size_t)k : nullptr, ld3(nn), sp->mat ? sp->mat[k] : 0, fl,
                                          sp->vis ? vis.data() : nullptr);
        memcpy(out->image + 3 * (size_t)k, so.image, 12);
        out->depth[k] = so.depth;
        memcpy(out->pos + 3 * (size_t)k, so.pos, 12);
        memcpy(out->normal + 3 * (size_t)k, nn, 12);
    }
    return 0;
}

int emul_splats_backward(const SurfScene* scene, const SurfCamera* cam, const SurfOptions* opt, const EmulSplats* sp,
                         const SurfOutGrads* og, const SurfSceneGrads* sg, float* gz, float* gnormal, float* gpos) {
    SceneView sc; CamState cs; std::vector<float> lcc;
    if (!splat_setup(scene, cam, &sc, &cs, &lcc)) return -1;
    ShadeFlags fl = {0, opt->use_quartic};
    HostSink hs(sc);
    std::vector<float> vis(sc.n_lights);
    for (int k = 0; k < sp->count; ++k) {
        if (sp->vis) for (int l = 0; l < sc.n_lights; ++l) vis[l] = sp->vis[(size_t)l * sp->count + k];
        PixelGrads g;
        for (int c = 0; c < 3; ++c) {
            g.image[c] = og->image ? og->image[3 * (size_t)k + c] : 0.f;
            g.pos[c] = og->pos ? og->pos[3 * (size_t)k + c] : 0.f;
            g.normal[c] = og->normal ? og->normal[3 * (size_t)k + c] : 0.f;
        }
        g.depth = og->depth ? og->depth[k] : 0.f;
        float gzk, gpk[3], gnk[3];
        splat_pixel_backward(sc, cs, k, sp->pos ? 0.f : sp->z[(size_t)k * sp->z_stride], sp->pos ? sp->pos + 3 * (size_t)k : nullptr,
                             ld3(sp->normal + (size_t)k * sp->normal_stride),
                             sp->mat ? sp->mat[k] : 0, fl, sp->vis ? vis.data() : nullptr, g, hs, &gzk, gpk, gnk);
        if (sp->pos) { for (int c = 0; c < 3; ++c) gpos[(size_t)k * 3 + c] += gpk[c]; }
        else gz[(size_t)k * sp->z_stride] += gzk;
        for (int c = 0; c < 3; ++c) gnormal[(size_t)k * sp->normal_stride + c] += gnk[c];
    }
    for (int m = 0; m < sc.n_materials * 3; ++m) {
        if (sg->albedo) sg->albedo[m] += (float)hs.g_albedo[m];
        if (sg->coeffs) sg->coeffs[m] += (float)hs.g_coeff[m];
    }
    for (int l = 0; l < sc.n_lights; ++l) {
        const double* gl = &hs.g_lpos[l * 3];
        if (sg->light_pos) {       // camera -> world: d/dl_xyz = R g, d/dl_w = -(R^T eye) . g
            for (int j = 0; j < 3; ++j)
                sg->light_pos[4 * (size_t)l + j] += (float)(cs.R[3 * j] * gl[0] + cs.R[3 * j + 1] * gl[1] + cs.R[3 * j + 2] * gl[2]);
            double gw = 0;
            for (int i = 0; i < 3; ++i)
                gw -= (cs.R[i] * cs.eye[0] + cs.R[3 + i] * cs.eye[1] + cs.R[6 + i] * cs.eye[2]) * gl[i];
            sg->light_pos[4 * (size_t)l + 3] += (float)gw;
        }
        for (int c = 0; c < 3; ++c)
            if (sg->light_attenuation) sg->light_attenuation[l * 3 + c] += (float)hs.g_atten[l * 3 + c];
    }
    for (int r = 0; r < sc.n_colors * 3; ++r)
        if (sg->colors) sg->colors[r] += (float)hs.g_color[r];
    for (int c = 0; c < 3; ++c)
        if (sg->ambient) sg->ambient[c] += (float)hs.g_amb[c];
    return 0;
}

}  // extern "C"
