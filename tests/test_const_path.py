"""The constant-bank intersection path (csrc/surf_isect_const.cu: k_filter_const, k_narrow_queue, k_const_fallback,
k_inside_disks) - the default for disk sets of large single frames - against the staged kernel (math_mode 5) and the
oracle: every output of every pixel identical across the kernels; odd record counts, mixed scenes, the eye inside a
bounding sphere, a queue too small for the candidates, CUDA-graph capture.  CPU: the SASS of the filter kernel keeps its
record scalars in uniform registers."""
import os
import re
import shutil
import subprocess

import pytest
import torch

import parity
import scene_io
from surf_renderer_b200 import scenes as synth

KEYS = ('nearest', 'depth', 'image', 'pos', 'normal')


def _same(a, b):
    for k in KEYS:
        assert torch.equal(a[k], b[k]), '%s differs on %d pixels' % (k, int((a[k] != b[k]).reshape(a[k].shape[0] * a[k].shape[1], -1).any(1).sum()))


def _render_modes(scene, **params):
    import surf_renderer_b200
    sc = scene_io.clone_scene(scene, device='cuda')
    with torch.no_grad():
        out = {m: surf_renderer_b200.render(sc, _math_mode=m, **params) for m in (0, 6, 5)}
    torch.cuda.synchronize()
    return out


def test_filter_kernel_keeps_the_records_in_uniform_registers():
    """The point of k_filter_const is `FFMA2 R, R.F32x2, UR.F32, R.F32x2`: the record scalar comes from a uniform register.
    ptxas drops that for many innocent-looking changes of the kernel (see the comments there), so the built library is
    checked: at least 3/4 of the packed FMAs of both instantiations read a uniform register."""
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    from surf_renderer_b200 import build
    sass = subprocess.run([cuobjdump, '-sass', build.build()], capture_output=True, text=True, check=True).stdout
    counts, fn = {}, None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            fn = m.group(1)
        elif fn and 'k_filter_const' in fn and re.search(r'\b(FFMA2|FMUL2)\b', line):
            c = counts.setdefault(fn, [0, 0])
            c[0 if re.search(r'UR\d+\.F32', line) else 1] += 1
    assert len(counts) == 2, counts
    for fn, (ur, vec) in counts.items():
        assert ur >= 3 * vec and ur >= 90, (fn, ur, vec)


@pytest.mark.gpu
@pytest.mark.parametrize('m,size', [(3001, (320, 288)), (2718, (400, 300)), (8191, (512, 512)), (257, (300, 300))])
def test_gpu_const_path_identical_to_staged_kernel(m, size):
    """record counts that are not multiples of the group sizes (2 and 4), one and several bank loads, ragged last tile"""
    scene = synth.config_e(m=m, width=size[0], height=size[1], radius=0.02)
    out = _render_modes(scene)
    assert int((out[5]['depth'] < 1000).sum()) > 1000
    _same(out[0], out[5])
    _same(out[6], out[5])


@pytest.mark.gpu
def test_gpu_const_path_mixed_scene_and_oracle():
    """planes, spheres and few triangles-free sets go through the staged kernel, the disk set (>= 256) through the
    constant bank; both merge in the z-buffer.  Also against the oracle."""
    from oracle import torch_oracle
    scene = synth.random_mixed_scene(41, width=300, height=260, n_disk=700, n_plane=1, n_sphere=4, n_tri=0)
    out = _render_modes(scene)
    _same(out[0], out[5])
    _same(out[6], out[5])
    ref = torch_oracle.render(scene_io.clone_scene(scene))
    report = parity.compare_forward({k: v.cpu() for k, v in out[0].items() if v is not None}, ref, scene)
    print(report)


@pytest.mark.gpu
def test_gpu_const_path_eye_inside_bounding_spheres():
    """disks whose bounding sphere holds the eye have no sphere record: k_inside_disks tests them against every pixel"""
    scene = synth.config_e(m=1500, width=320, height=272, radius=0.02)
    disk = scene['objects']['disk']
    eye = scene['camera']['eye'][:3]
    # a large disk just in front of the eye, one behind it, one containing it in its plane
    disk['pos'][:3] = torch.stack((eye + torch.tensor([0.05, 0.0, -0.4]), eye + torch.tensor([0.0, 0.1, 0.3]), eye + torch.tensor([0.2, 0.0, 0.0])))
    disk['normal'][:3] = torch.tensor([[0.1, 0.0, 1.0], [0.0, 0.0, 1.0], [0.0, 1.0, 0.0]])
    radius = torch.full((1500,), 0.02)
    radius[:3] = torch.tensor([0.6, 0.5, 1.0])
    disk['radius'] = radius
    out = _render_modes(scene)
    assert int((out[5]['nearest'] == 0).sum()) > 5000          # the near disk covers a good part of the frame
    _same(out[0], out[5])


@pytest.mark.gpu
def test_gpu_const_path_queue_overflow_takes_the_fallback(monkeypatch):
    """a candidate queue that is far too small: the overflowing tiles are flagged and redone by k_const_fallback"""
    scene = synth.config_e(m=4000, width=352, height=300, radius=0.03)
    import surf_renderer_b200
    sc = scene_io.clone_scene(scene, device='cuda')
    with torch.no_grad():
        ref = surf_renderer_b200.render(sc, _math_mode=5)
        for cap in ('0', '777', '20000'):
            monkeypatch.setenv('SURF_CONST_CAPACITY', cap)
            _same(surf_renderer_b200.render(sc, _math_mode=0), ref)
            _same(surf_renderer_b200.render(sc, _math_mode=6), ref)
    torch.cuda.synchronize()


@pytest.mark.gpu
def test_gpu_const_path_replays_from_a_cuda_graph():
    """the two-stream sequence of constant-bank copies and launches is captured (fork / join inside the capture)"""
    import surf_renderer_b200
    scene = synth.config_e(m=6000, width=384, height=320, radius=0.02)
    sc = scene_io.clone_scene(scene, device='cuda')
    with torch.no_grad():
        ref = surf_renderer_b200.render(sc, _math_mode=5)
        for _ in range(2):
            surf_renderer_b200.render(sc)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = surf_renderer_b200.render(sc)
        for k in KEYS:
            out[k].zero_()
        g.replay()
        g.replay()
    torch.cuda.synchronize()
    _same(out, ref)


@pytest.mark.gpu
def test_gpu_const_path_gradients_match_the_staged_kernel():
    """the backward recomputes the winner from `nearest`: identical z-buffers give identical gradients"""
    import surf_renderer_b200
    scene = synth.config_e(m=5000, width=320, height=320, radius=0.02)
    grads = {}
    for mode in (0, 5):
        sc = scene_io.clone_scene(scene, device='cuda', requires_grad=True)
        res = surf_renderer_b200.render(sc, _math_mode=mode)
        (res['image'].sum() + res['depth'][res['depth'] < 1000].sum()).backward()
        grads[mode] = {k: sc['objects']['disk'][k].grad.clone() for k in ('pos', 'normal')}
    for k in ('pos', 'normal'):
        assert torch.allclose(grads[0][k], grads[5][k], rtol=1e-5, atol=1e-7 * float(grads[5][k].abs().max()))


@pytest.mark.gpu
def test_gpu_const_path_with_shadow_rays_and_double_sided():
    """the primary rays take the constant-bank path, the shadow pass its own kernels: same frame as with the staged kernel"""
    scene = synth.config_e(m=1800, width=336, height=288, radius=0.03)
    out = _render_modes(scene, shadow=True, double_sided=True)
    assert int((out[5]['depth'] < 1000).sum()) > 1000
    _same(out[0], out[5])
    _same(out[6], out[5])


@pytest.mark.gpu
def test_gpu_const_path_pixel_band_equals_the_rows_of_the_full_frame():
    """row bands (the multi-GPU sharding) through the constant-bank path: bit-identical to the full frame"""
    import surf_renderer_b200
    scene = synth.config_e(m=3000, width=512, height=384, radius=0.02)
    sc = scene_io.clone_scene(scene, device='cuda')
    with torch.no_grad():
        full = surf_renderer_b200.render(sc)
        n = 512 * 384
        lo, hi = n // 3, n // 3 + 70001            # a ragged band of more than 256 x 256 pixels
        (image, depth, _, _, nearest, _), _ = surf_renderer_b200.render_flat(sc, pixel_range=(lo, hi))
    torch.cuda.synchronize()
    for k, v in (('nearest', nearest), ('depth', depth), ('image', image)):
        assert torch.equal(v.reshape(hi - lo, -1), full[k].reshape(n, -1)[lo:hi]), k
