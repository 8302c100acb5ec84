"""ctypes driver for the CPU emulation of the kernel math (tests/emul/surf_emul.cpp) - test infrastructure."""
import ctypes as C
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from surf_renderer_b200 import _abi                      # noqa: E402
from surf_renderer_b200.marshal import Marshalled, make_options   # noqa: E402

_lib = None


def lib():
    global _lib
    if _lib is None:
        sys.path.insert(0, os.path.join(HERE, 'emul'))
        import build as emul_build
        _lib = C.CDLL(emul_build.build())
        _lib.emul_forward.restype = C.c_longlong
        _lib.emul_forward.argtypes = [C.POINTER(_abi.SurfScene), C.POINTER(_abi.SurfCamera), C.POINTER(_abi.SurfOptions),
                                      C.POINTER(_abi.SurfOutputs), C.c_void_p]
        _lib.emul_backward.restype = C.c_int
        _lib.emul_backward.argtypes = [C.POINTER(_abi.SurfScene), C.POINTER(_abi.SurfCamera), C.POINTER(_abi.SurfOptions),
                                       C.c_void_p, C.c_void_p, C.POINTER(_abi.SurfOutGrads), C.POINTER(_abi.SurfSceneGrads)]
        _lib.emul_last_error.restype = C.c_char_p
    return _lib


def forward(scene, **params):
    m = Marshalled(scene, 'cpu')
    n = m.n_pixels
    out = {'image': torch.empty(n, 3), 'depth': torch.empty(n), 'normal': torch.empty(n, 3), 'pos': torch.empty(n, 3),
           'nearest': torch.empty(n, dtype=torch.int64),
           'ray_dir': torch.empty(3, n if m.proj == 0 else 1)}
    co = _abi.SurfOutputs(*[out[k].data_ptr() for k in ('image', 'depth', 'normal', 'pos', 'nearest', 'ray_dir')])
    sc, cam, opt = m.c_scene(), m.c_camera(), make_options(params)
    keys = torch.empty(n, dtype=torch.int64)
    misses = lib().emul_forward(C.byref(sc), C.byref(cam), C.byref(opt), C.byref(co), keys.data_ptr())
    if misses < 0:
        raise RuntimeError(lib().emul_last_error().decode())
    if params.get('shadow', False) and m.proj == 0:
        lib().emul_shadow_filter_misses.restype = C.c_longlong
        lib().emul_shadow_filter_misses.argtypes = [C.POINTER(_abi.SurfScene), C.POINTER(_abi.SurfCamera), C.c_void_p]
        sm = lib().emul_shadow_filter_misses(C.byref(sc), C.byref(cam), keys.data_ptr())
        if sm < 0:
            raise RuntimeError(lib().emul_last_error().decode())
        misses += sm
    H, W = m.height, m.width
    res = {'image': out['image'].view(H, W, 3), 'depth': out['depth'].view(H, W), 'normal': out['normal'].view(H, W, 3),
           'pos': out['pos'].view(H, W, 3), 'nearest': out['nearest'].view(H, W), 'ray_dir': out['ray_dir']}
    return res, int(misses), m


def backward(m, params, nearest, depth, gouts):
    grads = [torch.zeros_like(t) for t in m.floats]
    sg = m.c_grads(grads)
    og = _abi.SurfOutGrads(*[(gouts[k].contiguous().data_ptr() if gouts.get(k) is not None else None)
                             for k in ('image', 'depth', 'normal', 'pos')])
    keep = [gouts[k].contiguous() for k in gouts if gouts[k] is not None]   # noqa: F841
    og = _abi.SurfOutGrads(*[(t.data_ptr() if t is not None else None) for t in
                             [gouts.get(k) for k in ('image', 'depth', 'normal', 'pos')]])
    sc, cam, opt = m.c_scene(), m.c_camera(), make_options(params)
    rc = lib().emul_backward(C.byref(sc), C.byref(cam), C.byref(opt), nearest.contiguous().data_ptr(),
                             depth.contiguous().data_ptr(), C.byref(og), C.byref(sg))
    if rc != 0:
        raise RuntimeError(lib().emul_last_error().decode())
    return dict(zip(m.names, grads))


class _EmulSplats(C.Structure):
    _fields_ = [('count', C.c_int), ('z', C.c_void_p), ('z_stride', C.c_int), ('normal', C.c_void_p),
                ('normal_stride', C.c_int), ('mat', C.c_void_p), ('vis', C.c_void_p), ('pos', C.c_void_p)]


def _splat_setup(scene, params, inp=None):
    from along_ray_program import build_inputs
    if inp is None:
        inp = build_inputs(scene, params, torch.device('cpu'))
    sc, cam, sp, opt = inp.structs(inp.floats, params)
    es = _EmulSplats(sp.count, sp.z, sp.z_stride, sp.normal, sp.normal_stride, sp.material_idx, sp.light_vis, sp.pos)
    L = lib()
    L.emul_splats_forward.restype = C.c_int
    L.emul_splats_forward.argtypes = [C.POINTER(_abi.SurfScene), C.POINTER(_abi.SurfCamera), C.POINTER(_abi.SurfOptions),
                                      C.POINTER(_EmulSplats), C.POINTER(_abi.SurfOutputs)]
    L.emul_splats_backward.restype = C.c_int
    L.emul_splats_backward.argtypes = [C.POINTER(_abi.SurfScene), C.POINTER(_abi.SurfCamera), C.POINTER(_abi.SurfOptions),
                                       C.POINTER(_EmulSplats), C.POINTER(_abi.SurfOutGrads), C.POINTER(_abi.SurfSceneGrads),
                                       C.c_void_p, C.c_void_p, C.c_void_p]
    return inp, sc, cam, es, opt


def splats_forward(scene, _inp=None, **params):
    inp, sc, cam, es, opt = _splat_setup(scene, params, _inp)
    n = inp.n
    out = {'image': torch.empty(n, 3), 'depth': torch.empty(n), 'normal': torch.empty(n, 3), 'pos': torch.empty(n, 3)}
    co = _abi.SurfOutputs(out['image'].data_ptr(), out['depth'].data_ptr(), out['normal'].data_ptr(), out['pos'].data_ptr(), None, None)
    if lib().emul_splats_forward(C.byref(sc), C.byref(cam), C.byref(opt), C.byref(es), C.byref(co)) != 0:
        raise RuntimeError(lib().emul_last_error().decode())
    H, W = inp.height, inp.width
    return {'image': out['image'].view(H, W, 3), 'depth': out['depth'].view(H, W), 'pos': out['pos'].view(H, W, 3),
            'normal': out['normal'].view(H, W, 3)}, inp


def splats_backward(scene, params, gouts, _inp=None):
    """gradients w.r.t. inp.floats (for explicit fragments: floats[0] = positions, floats[1] = normals)"""
    inp, sc, cam, es, opt = _splat_setup(scene, params, _inp)
    grads = [torch.zeros_like(t) for t in inp.floats]
    keep = {k: v.contiguous() for k, v in gouts.items()}
    og = _abi.SurfOutGrads(keep['image'].data_ptr(), keep['depth'].data_ptr(), keep['normal'].data_ptr(), keep['pos'].data_ptr())
    sg = _abi.SurfSceneGrads()
    sg.light_pos, sg.light_attenuation, sg.ambient = grads[2].data_ptr(), grads[3].data_ptr(), grads[4].data_ptr()
    sg.colors, sg.albedo, sg.coeffs = grads[5].data_ptr(), grads[6].data_ptr(), grads[7].data_ptr()
    gz = grads[0].data_ptr() + (8 if inp.z_stride == 3 else 0)
    if lib().emul_splats_backward(C.byref(sc), C.byref(cam), C.byref(opt), C.byref(es), C.byref(og), C.byref(sg), gz,
                                  grads[1].data_ptr(), grads[0].data_ptr()) != 0:
        raise RuntimeError(lib().emul_last_error().decode())
    return dict(zip(inp.names, grads))
