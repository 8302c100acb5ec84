/* render_host.c - the C ABI of libsurf_b200.so (include/surf_b200.h) from plain C99: no Python, no torch.
 *
 * Builds a splat scene in host memory (splats on a sphere shell, two lights), renders it with the host-pointer
 * entry point surf_render_host (H2D, kernels, D2H inside the call), then runs one inverse-rendering step with
 * surf_render_backward_host against a constant target image, and prints checksums.  With an output path it also
 * writes the image as a binary PPM.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/render_host.c -Lsurf_renderer_b200 -lsurf_b200 \
 *       -Wl,-rpath,$PWD/surf_renderer_b200 -lm -o render_host && ./render_host 4000 128 128 out.ppm
 *
 * The scene replaces what the reference builds in test_optimization.py:634-655 (splat scene dict) and the call
 * replaces `res = render(scene); loss.backward()` (diffrend/torch/renderer.py:136, test_optimization.py:100-125).
 * tests/test_c_example.py compiles this file (CPU suite) and, on the GPU box, runs it and compares the checksums
 * with the Python path on the same scene.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "surf_b200.h"

/* the same generator as tests/test_c_example.py::lcg_scene: 32-bit LCG, 24-bit mantissa uniforms */
static uint32_t lcg_state = 12345u;
static float lcg_uniform(void) {
    lcg_state = lcg_state * 1664525u + 1013904223u;
    return (float)(lcg_state >> 8) * (1.0f / 16777216.0f);
}

#define CHECK(call)                                                           \
    do {                                                                      \
        int rc_ = (call);                                                     \
        if (rc_ != SURF_OK) {                                                 \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, surf_last_error()); \
            return 1;                                                         \
        }                                                                     \
    } while (0)

int main(int argc, char** argv) {
    const int M = argc > 1 ? atoi(argv[1]) : 4000;
    const int W = argc > 2 ? atoi(argv[2]) : 128;
    const int H = argc > 3 ? atoi(argv[3]) : 128;
    const char* ppm = argc > 4 ? argv[4] : NULL;
    const int N = W * H;
    if (surf_abi_version() != SURF_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 1; }

    /* ---- scene arrays (host) ---- */
    float* pos = malloc(sizeof(float) * 3 * M);
    float* nrm = malloc(sizeof(float) * 3 * M);
    float* rad = malloc(sizeof(float) * M);
    int32_t* mat = calloc(M, sizeof(int32_t));
    for (int i = 0; i < M; ++i) {
        float d[3], len = 0.f;
        do {                                           /* rejection-sample a direction */
            len = 0.f;
            for (int c = 0; c < 3; ++c) { d[c] = 2.f * lcg_uniform() - 1.f; len += d[c] * d[c]; }
        } while (len > 1.f || len < 1e-4f);
        len = sqrtf(len);
        for (int c = 0; c < 3; ++c) {
            pos[3 * i + c] = 0.5f * d[c] / len;
            nrm[3 * i + c] = d[c] / len + 0.05f * (2.f * lcg_uniform() - 1.f);
        }
        rad[i] = 0.03f;
    }
    float light_pos[8] = {20.f, 20.f, 20.f, 1.f, -15.f, 3.f, 15.f, 1.f};
    int32_t light_color[2] = {1, 2};
    float light_att[6] = {1.f, 0.f, 0.f, 1.f, 0.f, 0.f};
    float ambient[3] = {0.01f, 0.01f, 0.01f};
    float colors[9] = {0.f, 0.f, 0.f, 0.8f, 0.1f, 0.1f, 0.2f, 0.2f, 0.2f};
    float albedo[3] = {0.6f, 0.6f, 0.6f};
    float coeffs[3] = {0.5f, 0.4f, 8.0f};
    float gamma_v[1] = {0.8f};
    float eye[3] = {0.f, 0.f, 5.f}, at[3] = {0.f, 0.f, 0.f}, up[3] = {0.f, 1.f, 0.f};

    SurfScene sc;
    memset(&sc, 0, sizeof(sc));
    sc.n_sets = 1;
    sc.sets[0].kind = SURF_DISK; sc.sets[0].count = M;
    sc.sets[0].pos = pos; sc.sets[0].pos_stride = 3;
    sc.sets[0].normal = nrm; sc.sets[0].normal_stride = 3;
    sc.sets[0].radius = rad; sc.sets[0].material_idx = mat;
    sc.n_lights = 2; sc.light_pos = light_pos; sc.light_pos_stride = 4; sc.light_color_idx = light_color;
    sc.light_attenuation = light_att; sc.ambient = ambient;
    sc.n_colors = 3; sc.colors = colors;
    sc.n_materials = 1; sc.albedo = albedo; sc.coeffs = coeffs; sc.gamma = gamma_v;

    SurfCamera cam;
    memset(&cam, 0, sizeof(cam));
    cam.proj = 0; cam.width = W; cam.height = H;
    cam.fovy = 14.0 * 3.14159265358979323846 / 180.0; cam.focal_length = 1.0;
    cam.eye = eye; cam.at = at; cam.up = up; cam.near_clip = 0.1f; cam.far_clip = 1000.f;

    SurfOptions opt;
    memset(&opt, 0, sizeof(opt));

    /* ---- outputs (host) ---- */
    float* image = malloc(sizeof(float) * 3 * N);
    float* depth = malloc(sizeof(float) * N);
    int64_t* nearest = malloc(sizeof(int64_t) * N);
    SurfOutputs out;
    memset(&out, 0, sizeof(out));
    out.image = image; out.depth = depth; out.nearest = nearest;

    SurfContext* ctx = surf_context_create(0);
    if (!ctx) { fprintf(stderr, "surf_context_create: %s\n", surf_last_error()); return 1; }
    CHECK(surf_render_host(ctx, &sc, &cam, &opt, &out));

    double sum_image = 0.0, sum_depth = 0.0;
    long hits = 0, sum_nearest = 0;
    for (int k = 0; k < N; ++k) {
        sum_image += image[3 * k] + 2.0 * image[3 * k + 1] + 3.0 * image[3 * k + 2];
        if (depth[k] <= cam.far_clip) { ++hits; sum_depth += depth[k]; sum_nearest += (long)nearest[k]; }
    }
    printf("forward: hits %ld sum_image %.6f sum_depth %.6f sum_nearest %ld\n", hits, sum_image, sum_depth, sum_nearest);

    /* ---- one inverse-rendering step: d mean((image - target)^2) / d(pos, normal, albedo) ---- */
    float* target = malloc(sizeof(float) * 3 * N);
    for (int k = 0; k < 3 * N; ++k) target[k] = 0.25f;
    float* g_pos = calloc(3 * M, sizeof(float));
    float* g_nrm = calloc(3 * M, sizeof(float));
    float g_albedo[3] = {0.f, 0.f, 0.f};
    SurfSceneGrads sg;
    memset(&sg, 0, sizeof(sg));
    sg.sets[0].pos = g_pos; sg.sets[0].normal = g_nrm; sg.albedo = g_albedo;
    float loss = 0.f;
    CHECK(surf_render_backward_host(ctx, &sc, &cam, &opt, &out, NULL, target, &loss, &sg));
    double gp = 0.0, gn = 0.0;
    for (int k = 0; k < 3 * M; ++k) { gp += fabs(g_pos[k]); gn += fabs(g_nrm[k]); }
    uint64_t h2d = 0, d2h = 0;
    surf_context_last_transfer(ctx, &h2d, &d2h);
    printf("backward: loss %.8f sum|g_pos| %.6e sum|g_normal| %.6e g_albedo %.6e %.6e %.6e  (h2d %llu B, d2h %llu B)\n",
           loss, gp, gn, g_albedo[0], g_albedo[1], g_albedo[2], (unsigned long long)h2d, (unsigned long long)d2h);

    if (ppm) {
        FILE* f = fopen(ppm, "wb");
        if (f) {
            fprintf(f, "P6\n%d %d\n255\n", W, H);
            for (int k = 0; k < 3 * N; ++k) {
                float v = image[k] < 0.f ? 0.f : (image[k] > 1.f ? 1.f : image[k]);
                fputc((int)(255.f * v + 0.5f), f);
            }
            fclose(f);
        }
    }
    surf_context_destroy(ctx);
    free(pos); free(nrm); free(rad); free(mat); free(image); free(depth); free(nearest); free(target); free(g_pos); free(g_nrm);
    return 0;
}
