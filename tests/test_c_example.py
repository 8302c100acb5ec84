"""examples/render_host.c - the C ABI used from plain C99 (no Python, no torch in the process): compiled and linked
against the header and the library in the CPU suite; on the GPU box it runs and its checksums are compared with the
Python path on the same scene."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT

SRC = os.path.join(ROOT, 'examples', 'render_host.c')
PKG = os.path.join(ROOT, 'surf_renderer_b200')


def _compile(tmp_path):
    if shutil.which('gcc') is None:
        pytest.skip('gcc not available')
    from surf_renderer_b200 import build
    build.build()
    exe = str(tmp_path / 'render_host')
    subprocess.check_call(['gcc', '-std=c99', '-O2', '-Wall', '-Wextra', '-Werror', '-I' + os.path.join(ROOT, 'include'), SRC,
                           '-L' + PKG, '-lsurf_b200', '-Wl,-rpath,' + PKG, '-lm', '-o', exe])
    return exe


def lcg_scene(m, width, height):
    """The scene render_host.c builds, value for value (float32 arithmetic in the same order)."""
    f32 = np.float32
    state = np.uint64(12345)

    def uniform():
        nonlocal state
        state = (state * np.uint64(1664525) + np.uint64(1013904223)) & np.uint64(0xFFFFFFFF)
        return f32(int(state) >> 8) * f32(1.0 / 16777216.0)
    pos, nrm = np.zeros((m, 3), f32), np.zeros((m, 3), f32)
    for i in range(m):
        while True:
            d = [f32(2.0) * uniform() - f32(1.0) for _ in range(3)]
            ln = f32(0.0)
            for c in range(3):
                ln = f32(ln + f32(d[c] * d[c]))
            if not (ln > f32(1.0) or ln < f32(1e-4)):
                break
        ln = np.sqrt(ln, dtype=f32)
        for c in range(3):
            pos[i, c] = f32(f32(f32(0.5) * d[c]) / ln)
            nrm[i, c] = f32(f32(d[c] / ln) + f32(f32(0.05) * f32(f32(2.0) * uniform() - f32(1.0))))
    t = torch.tensor
    return {
        'camera': {'proj_type': 'perspective', 'viewport': [0, 0, width, height], 'fovy': 14.0 * 3.14159265358979323846 / 180.0,
                   'focal_length': 1.0, 'eye': t([0., 0., 5., 1.]), 'at': t([0., 0., 0., 1.]), 'up': t([0., 1., 0., 0.]),
                   'near': 0.1, 'far': 1000.0},
        'lights': {'pos': t([[20., 20., 20., 1.], [-15., 3., 15., 1.]]), 'color_idx': t([1, 2]),
                   'attenuation': t([[1., 0., 0.], [1., 0., 0.]]), 'ambient': t([0.01, 0.01, 0.01])},
        'colors': t([[0., 0., 0.], [0.8, 0.1, 0.1], [0.2, 0.2, 0.2]]),
        'materials': {'albedo': t([[0.6, 0.6, 0.6]]), 'coeffs': t([[0.5, 0.4, 8.0]])},
        'objects': {'disk': {'pos': t(pos), 'normal': t(nrm), 'radius': torch.full((m,), 0.03), 'material_idx': torch.zeros(m, dtype=torch.int64)}},
        'tonemap': {'type': 'gamma', 'gamma': t([0.8])},
    }


def test_c_example_compiles_and_links_against_the_header(tmp_path):
    exe = _compile(tmp_path)
    assert os.path.exists(exe)
    needed = subprocess.run(['ldd', exe], capture_output=True, text=True).stdout
    assert 'libsurf_b200.so' in needed and 'python' not in needed.lower() and 'torch' not in needed.lower()


@pytest.mark.gpu
def test_c_example_matches_the_python_path(tmp_path):
    import surf_renderer_b200
    import scene_io
    exe = _compile(tmp_path)
    m, w, h = 1500, 96, 80
    out = subprocess.run([exe, str(m), str(w), str(h), str(tmp_path / 'out.ppm')], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    fwd = re.search(r'forward: hits (\d+) sum_image (\S+) sum_depth (\S+) sum_nearest (\d+)', out.stdout)
    bwd = re.search(r'backward: loss (\S+) sum\|g_pos\| (\S+) sum\|g_normal\| (\S+) g_albedo (\S+) (\S+) (\S+)', out.stdout)
    assert fwd and bwd, out.stdout
    scene = lcg_scene(m, w, h)
    sc = scene_io.clone_scene(scene, device='cuda')
    for leaf in (sc['objects']['disk']['pos'], sc['objects']['disk']['normal'], sc['materials']['albedo']):
        leaf.requires_grad_(True)
    res = surf_renderer_b200.render(sc)
    hit = res['depth'] <= 1000.0
    img = res['image'].detach().double().cpu()
    assert int(hit.sum()) == int(fwd.group(1)) and int(hit.sum()) > 500
    assert int(res['nearest'][hit].sum()) == int(fwd.group(4))
    assert abs(float((img * torch.tensor([1., 2., 3.], dtype=torch.float64)).sum()) - float(fwd.group(2))) <= 1e-6 * abs(float(fwd.group(2))) + 1e-4
    assert abs(float(res['depth'].detach()[hit].double().sum()) - float(fwd.group(3))) <= 1e-6 * float(fwd.group(3))
    loss = ((res['image'] - 0.25) ** 2).mean()
    loss.backward()
    assert abs(float(loss.detach()) - float(bwd.group(1))) <= 1e-5 * float(bwd.group(1))
    gp = float(sc['objects']['disk']['pos'].grad.abs().double().sum())
    gn = float(sc['objects']['disk']['normal'].grad.abs().double().sum())
    assert abs(gp - float(bwd.group(2))) <= 1e-3 * gp and abs(gn - float(bwd.group(3))) <= 1e-3 * gn
    ga = sc['materials']['albedo'].grad.reshape(-1).cpu()
    for c in range(3):
        assert abs(float(ga[c]) - float(bwd.group(4 + c))) <= 1e-3 * abs(float(ga[c])) + 1e-7
    assert os.path.getsize(tmp_path / 'out.ppm') == len('P6\n%d %d\n255\n' % (w, h)) + 3 * w * h
