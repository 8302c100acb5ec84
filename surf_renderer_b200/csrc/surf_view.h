// surf_view.h - validates the public C structs (include/surf_b200.h) and turns them into the SceneView the
// math/kernels consume.  Host-only helper shared by the CUDA library and the CPU emulation in tests/emul.
#pragma once
#include <string>

#include "../../include/surf_b200.h"
#include "surf_math.cuh"

namespace surf {

inline bool build_scene_view(const SurfScene& s, SceneView* v, std::string* err, bool shading_only = false) {
    const int n_sets = shading_only ? 0 : s.n_sets;       // along-ray splats: only lights / colours / materials
    if (!shading_only && (s.n_sets < 1 || s.n_sets > SURF_MAX_SETS)) { *err = "scene needs 1..8 primitive sets"; return false; }
    v->n_sets = n_sets;
    int first = 0, rec = 0;
    for (int k = 0; k < n_sets; ++k) {
        const SurfPrimSet& p = s.sets[k];
        SetView& o = v->sets[k];
        if (p.kind < 0 || p.kind > 3) { *err = "unknown primitive kind"; return false; }
        if (p.count < 1) { *err = "empty primitive set (the reference's torch.cat also rejects it)"; return false; }
        if (!p.pos || !p.material_idx) { *err = "primitive set without pos/material_idx"; return false; }
        if (p.pos_stride != 3 && p.pos_stride != 4) { *err = "pos_stride must be 3 or 4"; return false; }
        if (p.kind != SURF_SPHERE) {
            if (!p.normal) { *err = "planar primitive set without normals"; return false; }
            if (p.normal_stride != 3 && p.normal_stride != 4) { *err = "normal_stride must be 3 or 4"; return false; }
        }
        if ((p.kind == SURF_DISK || p.kind == SURF_SPHERE) && !p.radius) { *err = "missing radius"; return false; }
        o.kind = p.kind; o.count = p.count; o.first = first;
        o.rec_off = rec;
        o.pos = p.pos; o.pos_stride = p.pos_stride;
        o.normal = p.normal; o.normal_stride = p.normal_stride;
        o.radius = p.radius; o.mat = p.material_idx;
        first += p.count;
        // every set's packed records start on a 128-byte boundary (8 float4) for the TMA bulk copies
        rec += ((p.count * rec_f4(p.kind) + 7) / 8) * 8;
    }
    for (int k = n_sets; k < kMaxSets; ++k) { v->sets[k] = SetView(); v->sets[k].first = 0x7fffffff; }
    v->total = first;
    if (s.n_lights < 1 || !s.light_pos || !s.light_color_idx || !s.light_attenuation || !s.ambient) {
        *err = "lights: pos, color_idx, attenuation and ambient are required (renderer.py:266-274)";
        return false;
    }
    if (s.light_pos_stride != 3 && s.light_pos_stride != 4) { *err = "light_pos_stride must be 3 or 4"; return false; }
    if (!s.colors || s.n_colors < 1 || !s.albedo || !s.coeffs || s.n_materials < 1) {
        *err = "colors / materials.albedo / materials.coeffs are required (renderer.py:266,276-277)";
        return false;
    }
    v->n_lights = s.n_lights; v->light_pos = s.light_pos; v->light_pos_stride = s.light_pos_stride;
    v->light_color_idx = s.light_color_idx; v->light_atten = s.light_attenuation; v->ambient = s.ambient;
    v->n_colors = s.n_colors; v->colors = s.colors;
    v->n_materials = s.n_materials; v->albedo = s.albedo; v->coeffs = s.coeffs;
    v->gamma = s.gamma;
    return true;
}

inline int packed_f4_total(const SceneView& v) {
    const SetView& l = v.sets[v.n_sets - 1];
    return l.rec_off + ((l.count * rec_f4(l.kind) + 7) / 8) * 8;
}

inline bool check_camera(const SurfCamera& c, std::string* err) {
    if (c.proj != 0 && c.proj != 1) { *err = "Invalid projection type"; return false; }
    if (c.width < 1 || c.height < 1) { *err = "empty viewport"; return false; }
    if (!c.eye || !c.at || !c.up) { *err = "camera eye/at/up required"; return false; }
    return true;
}

}  // namespace surf
