"""Drop-in for ``diffrend.torch.renderer.render_splats_along_ray`` (reference: renderer.py:537-751): one splat per
pixel at camera-space depth z along the pixel's ray - the differentiable path the GAN generator renders through
(GAN/gan.py:563-597).  Same scene dict; ``scene['objects']['disk']`` holds ``pos`` ([N] depths or [N,3] whose z
column is used), ``normal`` [N,3|4] in camera coordinates, ``material_idx`` [N] and optionally ``light_vis`` [L,N].
Returned keys: ``image [H,W,3]`` (no tonemap, like the reference), ``depth [H,W]``, ``pos [H,W,3]``, ``normal [H,W,3]``;
all autograd-connected to z, normals, lights and materials.  Built: samples == 1 with caller-provided normals.
Not built (NotImplementedError): normal estimation (``normal`` missing), supersampling (``samples > 1``),
``norm_depth_image_only``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi
from ._lib import check, lib
from .marshal import _as_float_tensor, _as_int_tensor, _scalar, make_options
from .renderer import _resolve_device, _stream_ptr, get_param_value


class _SplatInputs:
    def __init__(self, scene, device):
        cam = scene['camera']
        vp = cam['viewport']
        vp = vp.detach().cpu().numpy() if isinstance(vp, torch.Tensor) else np.asarray(vp)
        self.width, self.height = int(vp[2] - vp[0]), int(vp[3] - vp[1])
        self.n = self.width * self.height
        self.fovy, self.focal = _scalar(cam['fovy']), _scalar(cam['focal_length'])
        self.near, self.far = _scalar(cam.get('near', 0.1)), _scalar(cam.get('far', 1000.0))
        self.cam_vecs = {k: _as_float_tensor(cam[k], device).detach().reshape(-1)[:3].contiguous() for k in ('eye', 'at', 'up')}
        disk = scene['objects']['disk']
        if get_param_value('normal', disk, None) is None:
            raise NotImplementedError('normal estimation (utils.py:886-972) is not built: pass scene.objects.disk.normal')
        lights = scene['lights']
        self.names = ['objects/disk/pos', 'objects/disk/normal', 'lights/pos', 'lights/attenuation', 'lights/ambient',
                      'colors', 'materials/albedo', 'materials/coeffs']
        self.floats = [_as_float_tensor(v, device) for v in (disk['pos'], disk['normal'], lights['pos'], lights['attenuation'],
                                                             lights['ambient'], scene['colors'], scene['materials']['albedo'],
                                                             scene['materials']['coeffs'])]
        z = self.floats[0]
        if z.shape[0] != self.n:
            raise ValueError('render_splats_along_ray needs one splat per pixel: %d splats for %dx%d' % (z.shape[0], self.width, self.height))
        self.z_stride = 1 if z.dim() == 1 else int(z.shape[-1])
        if self.z_stride not in (1, 3):
            raise ValueError('disk.pos must be [N] or [N,3]')
        if self.floats[2].shape[-1] != 4:
            raise ValueError('lights.pos must be homogeneous [L,4] (it is multiplied by the 4x4 view matrix, renderer.py:709)')
        mi = disk.get('material_idx', None)
        self.mat = _as_int_tensor(mi, device) if mi is not None else None
        self.color_idx = _as_int_tensor(lights['color_idx'], device)
        lv = disk.get('light_vis', None)
        self.vis = _as_float_tensor(lv, device).detach() if lv is not None else None

    def structs(self, fl, params):
        z, nrm, lpos, att, amb, col, alb, cof = fl
        sc = _abi.SurfScene()
        sc.n_sets = 0
        sc.n_lights, sc.light_pos, sc.light_pos_stride = int(lpos.shape[0]), lpos.data_ptr(), 4
        sc.light_color_idx, sc.light_attenuation, sc.ambient = self.color_idx.data_ptr(), att.data_ptr(), amb.data_ptr()
        sc.n_colors, sc.colors = int(col.shape[0]), col.data_ptr()
        sc.n_materials = min(int(alb.shape[0]), int(cof.shape[0]))
        sc.albedo, sc.coeffs, sc.gamma = alb.data_ptr(), cof.data_ptr(), None
        cam = _abi.SurfCamera()
        cam.proj, cam.width, cam.height, cam.fovy, cam.focal_length = 0, self.width, self.height, self.fovy, self.focal
        cam.eye, cam.at, cam.up = (self.cam_vecs[k].data_ptr() for k in ('eye', 'at', 'up'))
        cam.near_clip, cam.far_clip = self.near, self.far
        sp = _abi.SurfSplats()
        sp.count = self.n
        sp.z = z.data_ptr() + (8 if self.z_stride == 3 else 0)          # column 2 of a [N,3] position array
        sp.z_stride, sp.normal, sp.normal_stride = self.z_stride, nrm.data_ptr(), int(nrm.shape[-1])
        sp.material_idx = self.mat.data_ptr() if self.mat is not None else None
        sp.light_vis = self.vis.data_ptr() if self.vis is not None else None
        return sc, cam, sp, make_options(params)


class _AlongRayFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, params, *floats):
        dev = floats[0].device
        n = inp.n
        sc, cam, sp, opt = inp.structs(floats, params)
        ws_bytes = lib().surf_workspace_bytes(0, n, sc.n_lights, 0)
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        image, depth = torch.empty(n, 3, device=dev), torch.empty(n, device=dev)
        pos, normal = torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev)
        out = _abi.SurfOutputs(image.data_ptr(), depth.data_ptr(), normal.data_ptr(), pos.data_ptr(), None, None)
        with torch.cuda.device(dev):
            check(lib().surf_splats_forward(C.byref(sc), C.byref(cam), C.byref(opt), C.byref(sp), workspace.data_ptr(),
                                            ws_bytes, C.byref(out), _stream_ptr()))
        ctx.inp, ctx.params, ctx.workspace = inp, params, workspace
        ctx.save_for_backward(*floats)
        return image, depth, pos, normal

    @staticmethod
    def backward(ctx, g_image, g_depth, g_pos, g_normal):
        floats = ctx.saved_tensors
        inp = ctx.inp
        sc, cam, sp, opt = inp.structs(floats, ctx.params)
        grads = [torch.zeros_like(t) if ctx.needs_input_grad[2 + i] else None for i, t in enumerate(floats)]

        def ptr(t):
            return None if t is None else t.data_ptr()
        gi, gd, gp, gn = (None if t is None else t.contiguous() for t in (g_image, g_depth, g_pos, g_normal))
        og = _abi.SurfOutGrads(ptr(gi), ptr(gd), ptr(gn), ptr(gp))
        sg = _abi.SurfSceneGrads()
        sg.light_pos, sg.light_attenuation, sg.ambient = ptr(grads[2]), ptr(grads[3]), ptr(grads[4])
        sg.colors, sg.albedo, sg.coeffs = ptr(grads[5]), ptr(grads[6]), ptr(grads[7])
        spg = _abi.SurfSplatGrads()
        if grads[0] is not None:
            spg.z = grads[0].data_ptr() + (8 if inp.z_stride == 3 else 0)
        spg.normal = ptr(grads[1])
        ws = ctx.workspace
        with torch.cuda.device(floats[0].device):
            check(lib().surf_splats_backward(C.byref(sc), C.byref(cam), C.byref(opt), C.byref(sp), ws.data_ptr(), ws.numel(),
                                             C.byref(og), C.byref(sg), C.byref(spg), _stream_ptr()))
        return (None, None) + tuple(grads)


def render_splats_along_ray(scene, **params):
    """Reference: diffrend/torch/renderer.py:537 ``render_splats_along_ray(scene, **params)``."""
    if get_param_value('samples', params, 1) > 1:
        raise NotImplementedError('supersampling (samples > 1, renderer.py:603-673) is not built')
    if get_param_value('norm_depth_image_only', params, False):
        raise NotImplementedError('norm_depth_image_only is not built for the along-ray renderer')
    dev = _resolve_device(scene)
    inp = _SplatInputs(scene, dev)
    image, depth, pos, normal = _AlongRayFn.apply(inp, dict(params), *inp.floats)
    H, W = inp.height, inp.width
    return {'image': image.view(H, W, 3), 'depth': depth.view(H, W), 'pos': pos.view(H, W, 3), 'normal': normal.view(H, W, 3)}
