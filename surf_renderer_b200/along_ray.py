"""Drop-in for ``diffrend.torch.renderer.render_splats_along_ray`` (reference: renderer.py:537-751): one splat per
pixel at camera-space depth z along the pixel's ray - the differentiable path the GAN generator renders through
(GAN/gan.py:563-597).  Same scene dict; ``scene['objects']['disk']`` holds ``pos`` ([N] depths or [N,3] whose z
column is used), optionally ``normal`` [N,3|4] in camera coordinates, ``material_idx`` [N] and optionally
``light_vis`` [L,N].  Returned keys: ``image [H,W,3]`` (no tonemap, like the reference), ``depth [H,W]``,
``pos [H,W,3]``, ``normal [H,W,3]``, all autograd-connected to z, normals, lights and materials.

Everything between the caller's tensors and the outputs runs in the CUDA kernels of csrc/surf_splats.cuh - nothing is
computed with torch ops here:

  * ``normal`` missing -> 3x3 stencil normal estimation in ``k_splat_normals`` (``normal_estimation_method`` 'plane',
    utils.py:886-923, or 'avg_normal', utils.py:854-883) with an analytic backward (``k_splat_normals_backward``);
  * ``samples`` K > 1 -> the K x K sub-pixel plane intersections of renderer.py:603-673 inside ``k_splat_forward``
    (outputs are [H K, W K, ...]);
  * ``norm_depth_image_only`` (renderer.py:677-686) -> depth min / max reduced in ``k_splat_forward``, normalised by
    ``k_splat_normdepth``;  ``orient_splats`` is a no-op in the reference and is accepted and ignored.

``render_splats_along_ray_batch`` renders a whole generator batch - depths [B,N], normals [B,N,3], camera eyes [B,3|4],
light positions [B,L,4] - in ONE call per direction (``surf_splats_forward_strided``), replacing the per-element Python
loop of gan.py:563-597.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi
from ._lib import check, lib
from .marshal import _as_float_tensor, _as_int_tensor, _scalar, make_options
from .renderer import _resolve_device, _stream_ptr, get_param_value

_METHODS = {'plane': 1, 'avg_normal': 2}


def _host_viewport(cam):
    vp = cam['viewport']
    vp = vp.detach().cpu().numpy() if isinstance(vp, torch.Tensor) else np.asarray(vp)
    return int(vp[2] - vp[0]), int(vp[3] - vp[1])


class _SplatCall:
    """Flat inputs of one call (single frame: batch = None, or a batch of B frames)."""

    ndc_stride = 0

    def __init__(self, scene, params, device, batch=None):
        cam = scene['camera']
        self.width, self.height = _host_viewport(cam)
        self.n_src = self.width * self.height
        self.fovy, self.focal = _scalar(cam['fovy']), _scalar(cam['focal_length'])
        self.near, self.far = _scalar(cam.get('near', 0.1)), _scalar(cam.get('far', 1000.0))
        self.batch = batch
        B = batch
        disk = scene['objects']['disk']
        lights = scene['lights']
        normal = get_param_value('normal', disk, None)
        self.estimate = 0
        if normal is None:
            method = get_param_value('normal_estimation_method', params, 'plane')
            if method not in _METHODS:
                raise ValueError("normal_estimation_method must be 'plane' or 'avg_normal'")   # 'quadric' maps to None upstream
            self.estimate = _METHODS[method]
        self.samples = int(get_param_value('samples', params, 1))
        self.norm_depth = bool(get_param_value('norm_depth_image_only', params, False))
        z = _as_float_tensor(disk['pos'], device)
        base = 1 if B is None else 2
        if z.dim() not in (base, base + 1) or (z.dim() == base + 1 and z.shape[-1] != 3):
            raise ValueError('disk.pos must be [N] or [N,3]' if B is None else 'disk.pos must be [B,N] or [B,N,3]')
        self.z_stride = 1 if z.dim() == base else 3
        if z.shape[base - 1] != self.n_src or (B is not None and z.shape[0] != B):
            raise ValueError('render_splats_along_ray needs one splat per pixel: %s splats for %dx%d'
                             % (tuple(z.shape), self.width, self.height))
        lp = _as_float_tensor(lights['pos'], device)
        if lp.shape[-1] != 4:
            raise ValueError('lights.pos must be homogeneous [L,4] (it is multiplied by the 4x4 view matrix, renderer.py:709)')
        self.n_lights = int(lp.shape[-2])
        self.names = ['objects/disk/pos', 'objects/disk/normal', 'lights/pos', 'lights/attenuation', 'lights/ambient',
                      'colors', 'materials/albedo', 'materials/coeffs']
        nrm = _as_float_tensor(normal, device) if normal is not None else None
        self.floats = [z, nrm, lp, _as_float_tensor(lights['attenuation'], device), _as_float_tensor(lights['ambient'], device),
                       _as_float_tensor(scene['colors'], device), _as_float_tensor(scene['materials']['albedo'], device),
                       _as_float_tensor(scene['materials']['coeffs'], device)]
        mi = disk.get('material_idx', None)
        self.mat = _as_int_tensor(mi, device) if mi is not None else None
        if self.samples > 1 and self.mat is None:
            raise AssertionError('supersampling needs material_idx (renderer.py:605)')
        self.color_idx = _as_int_tensor(lights['color_idx'], device)
        lv = disk.get('light_vis', None)
        self.vis = _as_float_tensor(lv, device).detach() if lv is not None else None
        eye = _as_float_tensor(cam['eye'], device).detach()
        self.eye_stride = 0
        if B is not None and eye.dim() == 2:
            if eye.shape[0] != B:
                raise ValueError('camera.eye batch dimension disagrees with disk.pos')
            eye = eye[:, :3].contiguous()
            self.eye_stride = 3
        else:
            eye = eye.reshape(-1)[:3].contiguous()
        self.cam_vecs = {'eye': eye}
        for k in ('at', 'up'):
            self.cam_vecs[k] = _as_float_tensor(cam[k], device).detach().reshape(-1)[:3].contiguous()
        K = max(1, self.samples)
        self.out_h, self.out_w = self.height * K, self.width * K
        self.n_out = self.n_src * K * K

    def _bstride(self, t, base_dim):
        """element stride between scenes of tensor t (0 when it carries no batch dimension)"""
        if self.batch is None or t is None or t.dim() == base_dim:
            return 0
        if t.shape[0] != self.batch:
            raise ValueError('batched along-ray scene: leading dimensions disagree')
        return int(t.stride(0))

    def structs(self, fl, params, norm_depth_ptr=None):
        z, nrm, lpos, att, amb, col, alb, cof = fl
        sc = _abi.SurfScene()
        sc.n_sets = 0
        sc.n_lights, sc.light_pos, sc.light_pos_stride = self.n_lights, lpos.data_ptr(), 4
        sc.light_color_idx, sc.light_attenuation, sc.ambient = self.color_idx.data_ptr(), att.data_ptr(), amb.data_ptr()
        sc.n_colors, sc.colors = int(col.shape[0]), col.data_ptr()
        sc.n_materials = min(int(alb.shape[0]), int(cof.shape[0]))
        sc.albedo, sc.coeffs, sc.gamma = alb.data_ptr(), cof.data_ptr(), None
        cam = _abi.SurfCamera()
        cam.proj, cam.width, cam.height, cam.fovy, cam.focal_length = 0, self.width, self.height, self.fovy, self.focal
        cam.eye, cam.at, cam.up = (self.cam_vecs[k].data_ptr() for k in ('eye', 'at', 'up'))
        cam.near_clip, cam.far_clip = self.near, self.far
        sp = _abi.SurfSplats()
        sp.count = self.n_src
        sp.z = z.data_ptr() + (8 if self.z_stride == 3 else 0)          # column 2 of a [N,3] position array
        sp.z_stride = self.z_stride
        if nrm is not None:
            sp.normal, sp.normal_stride = nrm.data_ptr(), int(nrm.shape[-1])
        sp.material_idx = self.mat.data_ptr() if self.mat is not None else None
        sp.light_vis = self.vis.data_ptr() if self.vis is not None else None
        sp.samples, sp.estimate_normals = self.samples, self.estimate
        sp.norm_depth = norm_depth_ptr
        lay = _abi.SurfSplatBatch()
        if self.batch is not None:
            lay.z = self._bstride(z, 1 if self.z_stride == 1 else 2)
            lay.normal = self._bstride(nrm, 2)
            lay.material_idx = self._bstride(self.mat, 1)
            lay.light_vis = self._bstride(self.vis, 2)
            lay.light_pos = self._bstride(lpos, 2)
            lay.eye = self.eye_stride
        return sc, cam, sp, lay, make_options(params)


class _AlongRayFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, call, params, *floats):
        dev = floats[0].device
        B = call.batch or 1
        n = call.n_out
        ws_bytes = lib().surf_splats_workspace_bytes(call.n_src, call.n_lights)
        workspace = torch.empty(B, ws_bytes, dtype=torch.uint8, device=dev)
        lead = (B,) if call.batch is not None else ()
        image, depth = torch.empty(lead + (n, 3), device=dev), torch.empty(lead + (n,), device=dev)
        pos, normal = torch.empty(lead + (n, 3), device=dev), torch.empty(lead + (n, 3), device=dev)
        norm_depth = torch.empty(lead + (n,), device=dev) if call.norm_depth else None
        sc, cam, sp, lay, opt = call.structs(floats, params, norm_depth.data_ptr() if norm_depth is not None else None)
        out = _abi.SurfOutputs(image.data_ptr(), depth.data_ptr(), normal.data_ptr(), pos.data_ptr(), None, None)
        with torch.cuda.device(dev):
            if call.batch is None:
                check(lib().surf_splats_forward(C.byref(sc), C.byref(cam), C.byref(opt), C.byref(sp), workspace.data_ptr(),
                                                ws_bytes, C.byref(out), _stream_ptr()))
            else:
                check(lib().surf_splats_forward_strided(B, C.byref(sc), C.byref(cam), C.byref(opt), C.byref(sp), C.byref(lay),
                                                        workspace.data_ptr(), ws_bytes, C.byref(out), _stream_ptr()))
        ctx.call, ctx.params, ctx.workspace = call, params, workspace
        ctx.save_for_backward(*[t for t in floats if t is not None])
        ctx.has = [t is not None for t in floats]
        if norm_depth is None:
            norm_depth = depth.new_empty(0)
        ctx.mark_non_differentiable(norm_depth)
        return image, depth, pos, normal, norm_depth

    @staticmethod
    def backward(ctx, g_image, g_depth, g_pos, g_normal, _g_nd):
        saved = list(ctx.saved_tensors)
        floats = [saved.pop(0) if h else None for h in ctx.has]
        call = ctx.call
        B = call.batch or 1
        sc, cam, sp, lay, opt = call.structs(floats, ctx.params)
        grads = [torch.zeros_like(t) if (t is not None and ctx.needs_input_grad[2 + i]) else None for i, t in enumerate(floats)]

        def ptr(t):
            return None if t is None else t.data_ptr()
        gi, gd, gp, gn = (None if t is None else t.contiguous() for t in (g_image, g_depth, g_pos, g_normal))
        og = _abi.SurfOutGrads(ptr(gi), ptr(gd), ptr(gn), ptr(gp))
        sg = _abi.SurfSceneGrads()
        sg.light_pos, sg.light_attenuation, sg.ambient = ptr(grads[2]), ptr(grads[3]), ptr(grads[4])
        sg.colors, sg.albedo, sg.coeffs = ptr(grads[5]), ptr(grads[6]), ptr(grads[7])
        spg = _abi.SurfSplatGrads()
        if grads[0] is not None and call.ndc_stride:
            spg.pos = grads[0].data_ptr()
        elif grads[0] is not None:
            spg.z = grads[0].data_ptr() + (8 if call.z_stride == 3 else 0)
        spg.normal = ptr(grads[1])
        ws = ctx.workspace
        with torch.cuda.device(floats[0].device):
            if call.batch is None:
                check(lib().surf_splats_backward(C.byref(sc), C.byref(cam), C.byref(opt), C.byref(sp), ws.data_ptr(), ws.shape[1],
                                                 C.byref(og), C.byref(sg), C.byref(spg), _stream_ptr()))
            else:
                check(lib().surf_splats_backward_strided(B, C.byref(sc), C.byref(cam), C.byref(opt), C.byref(sp), C.byref(lay),
                                                         ws.data_ptr(), ws.shape[1], C.byref(og), C.byref(sg), C.byref(spg),
                                                         _stream_ptr()))
        return (None, None) + tuple(grads)


def z_to_pcl_CC(z, camera):
    """Reference: diffrend/torch/renderer.py:484-507.  Camera-space point cloud [N,3] of one depth per pixel
    (z negative in front of the camera, clamped with -relu(-z)); differentiable in z, runs on z's device.  The GAN
    trainer imports it next to render / render_splats_along_ray (GAN/gan.py:28-29, :446)."""
    W, H = _host_viewport(camera)
    focal = _scalar(camera['focal_length'])
    h = np.tan(_scalar(camera['fovy']) / 2) * 2 * focal
    w = h * (W / H)
    Z = -torch.nn.functional.relu(-z)
    gx, gy = np.meshgrid(np.linspace(-1, 1, W), np.linspace(1, -1, H))
    gx *= w / 2
    gy *= h / 2
    x = torch.tensor(gx.ravel(), dtype=torch.float32, device=z.device)
    y = torch.tensor(gy.ravel(), dtype=torch.float32, device=z.device)
    return torch.stack((-Z * x / focal, -Z * y / focal, Z), dim=1)


def z_to_pcl_CC_batched(z, camera):
    """Reference: diffrend/torch/renderer.py:510-534, z is [B, H*W]; returns [B, H*W, 3]."""
    W, H = _host_viewport(camera)
    focal = _scalar(camera['focal_length'])
    h = np.tan(_scalar(camera['fovy']) / 2) * 2 * focal
    w = h * (W / H)
    Z = -torch.nn.functional.relu(-z)
    y, x = torch.meshgrid(torch.linspace(1, -1, H, device=z.device), torch.linspace(-1, 1, W, device=z.device), indexing='ij')
    x = (x * w / 2).flatten().repeat(z.shape[0], 1)
    y = (y * h / 2).flatten().repeat(z.shape[0], 1)
    return torch.stack((-Z * x / focal, -Z * y / focal, Z), dim=-1)


def _result(call, image, depth, pos, normal, norm_depth):
    H, W = call.out_h, call.out_w
    lead = (call.batch,) if call.batch is not None else ()
    if call.norm_depth:
        # renderer.py:677-686 returns pos / normal flat ([N,3]) in this branch
        return {'image': norm_depth.view(lead + (H, W)), 'depth': depth.view(lead + (H, W)), 'pos': pos, 'normal': normal}
    return {'image': image.view(lead + (H, W, 3)), 'depth': depth.view(lead + (H, W)), 'pos': pos.view(lead + (H, W, 3)),
            'normal': normal.view(lead + (H, W, 3))}


def render_splats_along_ray(scene, **params):
    """Reference: diffrend/torch/renderer.py:537 ``render_splats_along_ray(scene, **params)``."""
    dev = _resolve_device(scene)
    call = _SplatCall(scene, params, dev)
    return _result(call, *_AlongRayFn.apply(call, dict(params), *call.floats))


def render_splats_along_ray_batch(scene, **params):
    """A batch of along-ray frames in one call per direction (the loop body of GAN/gan.py:563-597 for every element):
    ``disk.pos`` [B,N] or [B,N,3]; optionally batched ``disk.normal`` [B,N,3|4], ``disk.material_idx`` [B,N],
    ``disk.light_vis`` [B,L,N], ``camera.eye`` [B,3|4], ``lights.pos`` [B,L,4]; everything else shared.  Returns the
    reference's keys with a leading batch dimension."""
    dev = _resolve_device(scene)
    pos = scene['objects']['disk']['pos']
    if not isinstance(pos, torch.Tensor) or pos.dim() < 2:
        raise ValueError('render_splats_along_ray_batch needs disk.pos of shape [B,N] or [B,N,3]')
    n_src = (lambda wh: wh[0] * wh[1])(_host_viewport(scene['camera']))
    if pos.dim() == 2 and pos.shape[0] == n_src and pos.shape[1] == 3:
        raise ValueError('disk.pos [N,3] is a single frame: use render_splats_along_ray')
    call = _SplatCall(scene, params, dev, batch=int(pos.shape[0]))
    return _result(call, *_AlongRayFn.apply(call, dict(params), *call.floats))


class _NDCCall(_SplatCall):
    """Flat inputs of render_splats_NDC: explicit fragment positions in normalised device coordinates."""

    def __init__(self, scene, params, device):
        cam = scene['camera']
        self.width, self.height = _host_viewport(cam)
        self.fovy, self.focal = _scalar(cam['fovy']), _scalar(cam.get('focal_length', 1.0))
        self.near, self.far = _scalar(cam['near']), _scalar(cam['far'])
        self.batch = None
        disk, lights = scene['objects']['disk'], scene['lights']
        pos = _as_float_tensor(disk['pos'], device)
        if pos.dim() != 2 or pos.shape[1] not in (3, 4):
            raise ValueError('render_splats_NDC: disk.pos must be [N,3] or [N,4] (normalised device coordinates)')
        self.n_src = self.n_out = int(pos.shape[0])
        if self.n_src != self.width * self.height:
            raise RuntimeError('render_splats_NDC needs one splat per pixel: %d splats for %dx%d (renderer.py:389)'
                               % (self.n_src, self.width, self.height))
        self.ndc_stride = int(pos.shape[1])
        self.z_stride, self.estimate, self.samples = 1, 0, 1
        self.norm_depth = bool(get_param_value('norm_depth_image_only', params, False))
        nrm = _as_float_tensor(disk['normal'], device)
        lp = _as_float_tensor(lights['pos'], device)
        if lp.shape[-1] != 4:
            raise ValueError('lights.pos must be homogeneous [L,4] (it is multiplied by the 4x4 view matrix, renderer.py:424)')
        self.n_lights = int(lp.shape[0])
        self.names = ['objects/disk/pos', 'objects/disk/normal', 'lights/pos', 'lights/attenuation', 'lights/ambient',
                      'colors', 'materials/albedo', 'materials/coeffs']
        self.floats = [pos, nrm, lp, _as_float_tensor(lights['attenuation'], device), _as_float_tensor(lights['ambient'], device),
                       _as_float_tensor(scene['colors'], device), _as_float_tensor(scene['materials']['albedo'], device),
                       _as_float_tensor(scene['materials']['coeffs'], device)]
        self.mat = _as_int_tensor(disk['material_idx'], device)
        self.color_idx = _as_int_tensor(lights['color_idx'], device)
        self.vis = None
        self.eye_stride = 0
        self.cam_vecs = {k: _as_float_tensor(cam[k], device).detach().reshape(-1)[:3].contiguous() for k in ('eye', 'at', 'up')}
        self.out_h, self.out_w = self.height, self.width

    def structs(self, fl, params, norm_depth_ptr=None):
        sc, cam, sp, lay, opt = super().structs(fl, params, norm_depth_ptr)
        sp.z, sp.z_stride = None, 1
        sp.pos, sp.ndc_stride = fl[0].data_ptr(), self.ndc_stride
        return sc, cam, sp, lay, opt


def render_splats_NDC(scene, **params):
    """Reference: diffrend/torch/renderer.py:358 ``render_splats_NDC(scene, **params)`` - one splat per pixel given in
    the camera's normalised device coordinates (``disk.pos`` [N,3|4]), unprojected with the inverse perspective of
    camera fovy / near / far (ops.py:50-68) and shaded in camera coordinates with the unnormalised view vector
    (renderer.py:437); ``double_sided``, ``use_quartic`` and ``norm_depth_image_only`` honoured.  Runs in the kernels of
    ``render_splats_along_ray`` (``SurfSplats.ndc_stride``); gradients reach pos, normals, lights and materials."""
    dev = _resolve_device(scene)
    call = _NDCCall(scene, params, dev)
    image, depth, pos, normal, norm_depth = _AlongRayFn.apply(call, dict(params), *call.floats)
    H, W = call.height, call.width
    if call.norm_depth:
        # renderer.py:392-401 returns the homogeneous [N,4] positions (w = 1 after the division) and the caller's normals
        return {'image': norm_depth.view(H, W), 'depth': depth.view(H, W),
                'pos': torch.cat((pos, torch.ones_like(pos[:, :1])), dim=1), 'normal': call.floats[1]}
    return {'image': image.view(H, W, 3), 'depth': depth.view(H, W), 'pos': pos.view(H, W, 3), 'normal': normal.view(H, W, 3)}
