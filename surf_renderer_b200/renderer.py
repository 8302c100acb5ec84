"""Drop-in for ``diffrend.torch.renderer.render`` (reference: diffrend/torch/renderer.py:136-355).

Same call: ``render(scene: dict, **params) -> dict`` with the reference's scene-dict schema and
kwargs (looked up like ``get_param_value``, diffrend/utils/utils.py:4-10).  Returned keys: ``image [H,W,3]``,
``depth [H,W]``, ``normal [H,W,3]``, ``pos [H,W,3]``, ``nearest [H,W] int64``, ``ray_dir`` and ``ray_dist``
(always ``None``: in the reference it is a last-tile artefact no caller reads).  ``image/depth/normal/pos`` are
autograd-connected to every float tensor of the scene except the camera (not differentiable in the reference
either, utils.py:476).  All compute runs in libsurf_b200.so on the current CUDA device and stream.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _abi
from ._lib import check, lib
from .marshal import Marshalled, MarshalledBatch, batched_scene_size, make_options


def get_param_value(key, dict_var, default_val, required=False):
    """diffrend/utils/utils.py:4-10."""
    if key in dict_var:
        return dict_var[key]
    elif required:
        raise ValueError('Missing required key {}'.format(key))
    return default_val


_KEEP_WORKSPACE_BYTES = 32 << 20
# SURF_B200_CHECK_INDICES=1 (or render(..., _check_indices=True)): after the forward, synchronise and raise IndexError
# when a material_idx / color_idx held in DEVICE memory was out of range (host-side index arrays are always checked).
_CHECK_INDICES = os.environ.get('SURF_B200_CHECK_INDICES', '0') not in ('', '0')


def _raise_on_bad_indices(workspace_ptr, params):
    if not params.get('_check_indices', _CHECK_INDICES):
        return
    if lib().surf_check_indices(workspace_ptr, _stream_ptr()) != 0:
        raise IndexError(lib().surf_last_error().decode('utf-8', 'replace'))


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _RenderFn(torch.autograd.Function):
    """forward: surf_forward; backward: surf_backward (recompute-the-winner analytic gradients)."""

    @staticmethod
    def forward(ctx, m, params, pixel_range, *floats):
        dev = floats[0].device
        n_total = m.n_pixels
        p0, p1 = pixel_range if pixel_range is not None else (0, n_total)
        n = p1 - p0
        opt = make_options(params, (p0, p1))
        shadow = int(opt.shadow)
        n_lights = int(floats[m.i_light_pos].shape[0])
        ws_bytes = lib().surf_workspace_bytes_ex(m.total_prims, n, n_lights, shadow, m.proj, 0)
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        image = torch.empty(n, 3, dtype=torch.float32, device=dev)
        depth = torch.empty(n, dtype=torch.float32, device=dev)
        normal = torch.empty(n, 3, dtype=torch.float32, device=dev)
        pos = torch.empty(n, 3, dtype=torch.float32, device=dev)
        nearest = torch.empty(n, dtype=torch.int64, device=dev)
        ray_dir = torch.empty(3, n if m.proj == 0 else 1, dtype=torch.float32, device=dev)
        out = _abi.SurfOutputs(image.data_ptr(), depth.data_ptr(), normal.data_ptr(), pos.data_ptr(),
                               nearest.data_ptr(), ray_dir.data_ptr())
        sc, cam = m.c_scene(floats), m.c_camera()
        with torch.cuda.device(dev):
            check(lib().surf_forward(C.byref(sc), C.byref(cam), C.byref(opt), workspace.data_ptr(), ws_bytes,
                                     C.byref(out), _stream_ptr()))
            _raise_on_bad_indices(workspace.data_ptr(), params)
        ctx.m, ctx.params, ctx.range = m, params, (p0, p1)
        # The backward needs the camera state, the rays and (with shadows) the visibility from the forward workspace.
        # Small workspaces ride along in the graph (saves two launches in backward); large ones are released now and the
        # backward recomputes camera + rays into a fresh one (k_setup + k_raygen, a few microseconds per megapixel).
        ctx.workspace = workspace if (shadow or ws_bytes <= _KEEP_WORKSPACE_BYTES) else None
        ctx.ws_bytes = ws_bytes
        ctx.save_for_backward(nearest, depth, *floats)
        ctx.mark_non_differentiable(nearest, ray_dir)
        return image, depth, normal, pos, nearest, ray_dir

    @staticmethod
    def backward(ctx, g_image, g_depth, g_normal, g_pos, _g_nearest, _g_ray):
        saved = ctx.saved_tensors
        nearest, depth, floats = saved[0], saved[1], saved[2:]
        m = ctx.m
        dev = depth.device
        opt = make_options(ctx.params, ctx.range)
        ws = ctx.workspace
        if ws is not None:
            opt.forced_nearest = 2      # ctx.workspace still holds this frame's camera state and rays
        else:
            ws = torch.empty(ctx.ws_bytes, dtype=torch.uint8, device=dev)
        grads = [torch.zeros_like(t) if ctx.needs_input_grad[3 + i] else None for i, t in enumerate(floats)]

        def c(t):
            return None if t is None else t.contiguous()

        g_image, g_depth, g_normal, g_pos = c(g_image), c(g_depth), c(g_normal), c(g_pos)
        og = _abi.SurfOutGrads(*[(t.data_ptr() if t is not None else None)
                                 for t in (g_image, g_depth, g_normal, g_pos)])
        sg = m.c_grads(grads)
        sc, cam = m.c_scene(floats), m.c_camera()
        with torch.cuda.device(dev):
            check(lib().surf_backward(C.byref(sc), C.byref(cam), C.byref(opt), ws.data_ptr(), ws.numel(),
                                      nearest.data_ptr(), depth.data_ptr(), C.byref(og), C.byref(sg),
                                      _stream_ptr()))
        return (None, None, None) + tuple(grads)


def _resolve_device(scene):
    """The scene's device, like the reference where every tensor lives on the device chosen at import
    (diffrend/torch/utils.py:7-15).  CPU scenes are moved to the current CUDA device."""
    def find(v):
        if isinstance(v, torch.Tensor) and v.is_cuda:
            return v.device
        if isinstance(v, dict):
            for x in v.values():
                d = find(x)
                if d is not None:
                    return d
        return None
    dev = find(scene)
    if dev is None:
        if not torch.cuda.is_available():
            raise RuntimeError('surf_renderer_b200.render needs a CUDA device (B200); there is no CPU path')
        dev = torch.device('cuda', torch.cuda.current_device())
    return dev


def render_flat(scene, pixel_range=None, **params):
    """render() on a contiguous flat pixel range [p0, p1) of the row-major grid (row bands for multi-GPU).
    Returns flat tensors (image [n,3], depth [n], normal [n,3], pos [n,3], nearest [n], ray_dir) and (H, W)."""
    if get_param_value('vis_stat', params, False):
        raise RuntimeError('Removed Support for vis_stat')                      # renderer.py:233-234
    dev = _resolve_device(scene)
    m = Marshalled(scene, dev)
    outs = _RenderFn.apply(m, dict(params), pixel_range, *m.floats)
    return outs, (m.height, m.width)


def render(scene, **params):
    """Render.  Reference: diffrend/torch/renderer.py:136 ``render(scene, **params)``.

    params honoured: double_sided, use_quartic, shadow, norm_depth_image_only.  Accepted and semantically no-ops
    here: tiled, tile_size (the pixel tiling only bounded the reference's [M,N] temporaries), backface_culling (the
    reference only labels the scene, renderer.py:152-159 - output unchanged).  vis_stat raises RuntimeError like
    the reference.
    """
    (image, depth, normal, pos, nearest, ray_dir), (H, W) = render_flat(scene, None, **params)
    if get_param_value('norm_depth_image_only', params, False):
        # renderer.py:245-260: depth normalised to [0, 1] with misses mapped to the minimum.  (In the reference
        # this branch only runs with tiled=False and for scenes whose intersectors ignore disable_normals; here
        # it is defined for every scene.)  Same arithmetic: blend, subtract, divide.
        im_depth = depth.view(H, W)
        far = float(scene['camera']['far'])
        min_depth = torch.min(im_depth)
        is_far = (im_depth >= far).float()
        norm = is_far * min_depth + (1 - is_far) * im_depth
        norm = (norm - min_depth) / (torch.max(im_depth) - min_depth)
        return {'image': norm, 'depth': im_depth, 'ray_dist': None, 'obj_dist': None, 'nearest': nearest.view(H, W),
                'ray_dir': ray_dir, 'valid_pixels': None, 'obj_pixel_count': None, 'pixel_obj_count': None,
                'valid_pixels_mask': None}
    return {
        'image': image.view(H, W, 3),
        'depth': depth.view(H, W),
        'normal': normal.view(H, W, 3),
        'pos': pos.view(H, W, 3),
        'ray_dist': None,
        'nearest': nearest.view(H, W),
        'ray_dir': ray_dir,
    }


# ---------------------------------------------------------------------------------------------------------------
# batches of independent scenes
# ---------------------------------------------------------------------------------------------------------------
class _RenderBatchFn(torch.autograd.Function):
    """surf_forward_batch / surf_backward_batch over a list of marshalled scenes: one host call per direction."""

    @staticmethod
    def forward(ctx, ms, params, *floats):
        dev = floats[0].device
        n = len(ms)
        opt = make_options(params)
        offs, o = [], 0
        for m in ms:
            offs.append(o)
            o += len(m.floats)
        scenes = (_abi.SurfScene * n)()
        cams = (_abi.SurfCamera * n)()
        outs = (_abi.SurfOutputs * n)()
        ws_ptrs = (C.c_void_p * n)()
        ws_sizes = (C.c_size_t * n)()
        results, workspaces = [], []
        for b, m in enumerate(ms):
            fl = floats[offs[b]:offs[b] + len(m.floats)]
            scenes[b], cams[b] = m.c_scene(fl), m.c_camera()
            npx = m.n_pixels
            ws_bytes = lib().surf_workspace_bytes_ex(m.total_prims, npx, int(fl[m.i_light_pos].shape[0]), int(opt.shadow), m.proj, 0)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            t = (torch.empty(npx, 3, device=dev), torch.empty(npx, device=dev), torch.empty(npx, 3, device=dev),
                 torch.empty(npx, 3, device=dev), torch.empty(npx, dtype=torch.int64, device=dev),
                 torch.empty(3, npx if m.proj == 0 else 1, device=dev))
            outs[b] = _abi.SurfOutputs(*[x.data_ptr() for x in t])
            ws_ptrs[b], ws_sizes[b] = ws.data_ptr(), ws_bytes
            results.append(t)
            workspaces.append(ws)
        with torch.cuda.device(dev):
            check(lib().surf_forward_batch(n, scenes, cams, C.byref(opt), ws_ptrs, ws_sizes, outs, _stream_ptr()))
        ctx.ms, ctx.params, ctx.offs, ctx.workspaces = ms, params, offs, workspaces
        ctx.n_float = len(floats)
        flat = [x for t in results for x in t]
        ctx.save_for_backward(*floats, *[t[4] for t in results], *[t[1] for t in results])
        ctx.mark_non_differentiable(*[t[4] for t in results], *[t[5] for t in results])
        return tuple(flat)

    @staticmethod
    def backward(ctx, *gouts):
        ms, n = ctx.ms, len(ctx.ms)
        saved = ctx.saved_tensors
        floats = saved[:ctx.n_float]
        nearest, depth = saved[ctx.n_float:ctx.n_float + n], saved[ctx.n_float + n:]
        dev = floats[0].device
        opt = make_options(ctx.params)
        opt.forced_nearest = 2
        grads = [torch.zeros_like(t) if ctx.needs_input_grad[2 + i] else None for i, t in enumerate(floats)]
        scenes = (_abi.SurfScene * n)()
        cams = (_abi.SurfCamera * n)()
        ogs = (_abi.SurfOutGrads * n)()
        sgs = (_abi.SurfSceneGrads * n)()
        ws_ptrs = (C.c_void_p * n)()
        ws_sizes = (C.c_size_t * n)()
        near_p = (C.c_void_p * n)()
        depth_p = (C.c_void_p * n)()
        keep = []
        for b, m in enumerate(ms):
            lo = ctx.offs[b]
            fl = floats[lo:lo + len(m.floats)]
            scenes[b], cams[b] = m.c_scene(fl), m.c_camera()
            g_image, g_depth, g_normal, g_pos = (None if g is None else g.contiguous() for g in gouts[6 * b:6 * b + 4])
            keep += [g_image, g_depth, g_normal, g_pos]
            ogs[b] = _abi.SurfOutGrads(*[(t.data_ptr() if t is not None else None) for t in (g_image, g_depth, g_normal, g_pos)])
            sgs[b] = m.c_grads(grads[lo:lo + len(m.floats)])
            ws_ptrs[b], ws_sizes[b] = ctx.workspaces[b].data_ptr(), ctx.workspaces[b].numel()
            near_p[b], depth_p[b] = nearest[b].data_ptr(), depth[b].data_ptr()
        with torch.cuda.device(dev):
            check(lib().surf_backward_batch(n, scenes, cams, C.byref(opt), ws_ptrs, ws_sizes, near_p, depth_p, ogs, sgs,
                                            _stream_ptr()))
        return (None, None) + tuple(grads)


class _RenderStridedFn(torch.autograd.Function):
    """surf_forward_strided / surf_backward_strided: a batch held as stacked tensors, one allocation per output and
    per gradient for the whole batch."""

    @staticmethod
    def forward(ctx, mb, params, *fulls):
        m, B = mb.m, mb.batch
        dev = fulls[0].device
        opt = make_options(params)
        npx = m.n_pixels
        views = mb.views0(fulls)
        ws_bytes = lib().surf_workspace_bytes_ex(m.total_prims, npx, int(views[m.i_light_pos].shape[0]), int(opt.shadow), m.proj, 0)
        ws_bytes = (ws_bytes + 255) // 256 * 256
        workspace = torch.empty(B, ws_bytes, dtype=torch.uint8, device=dev)
        image = torch.empty(B, npx, 3, device=dev)
        depth = torch.empty(B, npx, device=dev)
        normal = torch.empty(B, npx, 3, device=dev)
        pos = torch.empty(B, npx, 3, device=dev)
        nearest = torch.empty(B, npx, dtype=torch.int64, device=dev)
        ray_dir = torch.empty(B, 3, npx if m.proj == 0 else 1, device=dev)
        out = _abi.SurfOutputs(image.data_ptr(), depth.data_ptr(), normal.data_ptr(), pos.data_ptr(),
                               nearest.data_ptr(), ray_dir.data_ptr())
        sc, cam, lay = m.c_scene(views), m.c_camera(), mb.c_layout()
        with torch.cuda.device(dev):
            check(lib().surf_forward_strided(B, C.byref(sc), C.byref(cam), C.byref(lay), C.byref(opt),
                                             workspace.data_ptr(), ws_bytes, C.byref(out), _stream_ptr()))
        ctx.mb, ctx.params, ctx.workspace = mb, params, workspace
        ctx.save_for_backward(nearest, depth, *fulls)
        ctx.mark_non_differentiable(nearest, ray_dir)
        return image, depth, normal, pos, nearest, ray_dir

    @staticmethod
    def backward(ctx, g_image, g_depth, g_normal, g_pos, _g_nearest, _g_ray):
        saved = ctx.saved_tensors
        nearest, depth, fulls = saved[0], saved[1], saved[2:]
        mb = ctx.mb
        m, B = mb.m, mb.batch
        dev = depth.device
        opt = make_options(ctx.params)
        opt.forced_nearest = 2
        grads = [torch.zeros_like(t) if ctx.needs_input_grad[2 + i] else None for i, t in enumerate(fulls)]
        gs = [None if g is None else g.contiguous() for g in (g_image, g_depth, g_normal, g_pos)]
        og = _abi.SurfOutGrads(*[(t.data_ptr() if t is not None else None) for t in gs])
        sg = m.c_grads(grads)                       # base pointers; the layout strides apply to them as well
        sc, cam, lay = m.c_scene(mb.views0(fulls)), m.c_camera(), mb.c_layout()
        ws = ctx.workspace
        with torch.cuda.device(dev):
            check(lib().surf_backward_strided(B, C.byref(sc), C.byref(cam), C.byref(lay), C.byref(opt), ws.data_ptr(),
                                              ws.shape[1], nearest.data_ptr(), depth.data_ptr(), C.byref(og),
                                              C.byref(sg), _stream_ptr()))
        return (None, None) + tuple(grads)


def _render_strided(scene, params):
    dev = _resolve_device(scene)
    mb = MarshalledBatch(scene, dev)
    image, depth, normal, pos, nearest, ray_dir = _RenderStridedFn.apply(mb, dict(params), *mb.fulls)
    B, H, W = mb.batch, mb.m.height, mb.m.width
    if get_param_value('norm_depth_image_only', params, False):
        # renderer.py:245-260 per scene of the batch (GAN.get_real_samples passes the flag, gan.py:377-379): depth
        # normalised to [0, 1] with misses mapped to the scene's minimum; same arithmetic as render()
        im_depth = depth.view(B, H, W)
        lo = im_depth.amin(dim=(1, 2), keepdim=True)
        hi = im_depth.amax(dim=(1, 2), keepdim=True)
        is_far = (im_depth >= mb.m.far).float()
        norm = is_far * lo + (1 - is_far) * im_depth
        norm = (norm - lo) / (hi - lo)
        return {'image': norm, 'depth': im_depth, 'ray_dist': None, 'obj_dist': None, 'nearest': nearest.view(B, H, W),
                'ray_dir': ray_dir, 'valid_pixels': None, 'obj_pixel_count': None, 'pixel_obj_count': None,
                'valid_pixels_mask': None}
    return {'image': image.view(B, H, W, 3), 'depth': depth.view(B, H, W), 'normal': normal.view(B, H, W, 3),
            'pos': pos.view(B, H, W, 3), 'ray_dist': None, 'nearest': nearest.view(B, H, W), 'ray_dir': ray_dir}


def _stack_scenes(scenes):
    """A list of scene dicts with identical structure -> one batched scene dict (leaves that are the SAME tensor
    object in every scene stay shared, the others are stacked, differentiably), or None when the scenes differ in
    structure and must be rendered one by one."""
    first = scenes[0]

    def leaves(sc):
        out = []
        for kind, prim in sc['objects'].items():
            for f in sorted(prim):
                out.append(('objects', kind, f, prim[f]))
        for f in ('pos', 'attenuation', 'ambient', 'color_idx'):
            out.append(('lights', f, None, sc['lights'][f]))
        out.append(('colors', None, None, sc['colors']))
        for f in ('albedo', 'coeffs'):
            out.append(('materials', f, None, sc['materials'][f]))
        if 'tonemap' in sc:
            out.append(('tonemap', 'gamma', None, sc['tonemap']['gamma']))
        for f in ('eye', 'at', 'up'):
            out.append(('camera', f, None, sc['camera'][f]))
        return out

    def scalars(sc):
        cam = sc['camera']
        vp = cam['viewport']
        return (cam.get('proj_type'), tuple(int(x) for x in vp), float(cam['fovy']), float(cam['focal_length']),
                float(cam['near']), float(cam['far']), tuple(sc['objects'].keys()), 'tonemap' in sc,
                sc['tonemap']['type'] if 'tonemap' in sc else None)

    try:
        ref_leaves, ref_scalars = leaves(first), scalars(first)
        columns = [[v] for (_, _, _, v) in ref_leaves]
        for sc in scenes[1:]:
            if scalars(sc) != ref_scalars:
                return None
            lv = leaves(sc)
            if len(lv) != len(ref_leaves):
                return None
            for col, (a, b, c, v), (a0, b0, c0, v0) in zip(columns, lv, ref_leaves):
                if (a, b, c) != (a0, b0, c0):
                    return None
                if v is not v0:
                    if not isinstance(v, torch.Tensor):      # lists / numpy arrays / numbers (the demos mix them)
                        v = torch.as_tensor(np.asarray(v))
                    if not isinstance(col[0], torch.Tensor):
                        col[0] = torch.as_tensor(np.asarray(col[0]))
                    if v.shape != col[0].shape or v.dtype != col[0].dtype or v.device != col[0].device:
                        return None
                col.append(v)
    except (KeyError, TypeError, ValueError):
        return None                      # let the per-scene path raise the reference-style error
    batched = {'objects': {k: {} for k in first['objects']}, 'lights': {}, 'materials': {}, 'camera': dict(first['camera'])}
    for (a, b, c, v0), col in zip(ref_leaves, columns):
        shared = all(v is v0 for v in col)
        if not shared and col[0].dim() == 0:
            col = [v.reshape(1) for v in col]
        val = v0 if shared else torch.stack(col, 0)
        if a == 'objects':
            batched['objects'][b][c] = val
        elif a == 'colors':
            batched['colors'] = val
        elif a == 'tonemap':
            batched['tonemap'] = {'type': first['tonemap']['type'], 'gamma': val}
        else:
            batched[a][b] = val
    return batched


def render_batch(scenes, **params):
    """Render a batch of independent scenes with ONE library call per direction (the reference renders a GAN batch
    in a Python loop of render() calls, GAN/gan.py:326-377).

    `scenes` is either
      * a list of scene dicts -> returns a list of render()-style result dicts.  Scenes of identical structure
        (same primitive kinds and counts, same camera scalars - the GAN case) are stacked and rendered through the
        strided-batch entry points; otherwise each scene is marshalled on its own.  A tensor shared by several
        scenes (e.g. common lights / materials) receives the sum of its gradients, as autograd would; or
      * ONE scene dict whose tensors may carry a leading batch dimension B (splat positions [B,M,3], camera eyes
        [B,4], ...; everything else shared) -> returns one dict of stacked outputs (image [B,H,W,3], depth [B,H,W],
        normal, pos, nearest [B,H,W], ray_dir [B,3,n]).  This is the cheapest form: no per-scene host work at all."""
    if get_param_value('vis_stat', params, False):
        raise RuntimeError('Removed Support for vis_stat')
    if isinstance(scenes, dict):
        return _render_strided(scenes, params)
    if len(scenes) == 0:
        return []
    stacked = _stack_scenes(scenes) if len(scenes) > 1 else None
    if stacked is not None and batched_scene_size(stacked) != len(scenes):
        stacked = None                   # e.g. the same scene object repeated: nothing to stack along
    if get_param_value('norm_depth_image_only', params, False) and stacked is None:
        return [render(sc, **params) for sc in scenes]           # depth-only frames of differently shaped scenes
    if stacked is not None:
        res = _render_strided(stacked, params)
        keys = [k for k in ('image', 'depth', 'normal', 'pos', 'nearest', 'ray_dir') if res.get(k) is not None]
        cols = {k: torch.unbind(res[k], 0) for k in keys}
        return [dict({k: None for k in res}, **{k: cols[k][b] for k in keys}) for b in range(len(scenes))]
    dev = _resolve_device(scenes[0])
    ms = [Marshalled(sc, dev) for sc in scenes]
    flat = _RenderBatchFn.apply(ms, dict(params), *[t for m in ms for t in m.floats])
    out = []
    for b, m in enumerate(ms):
        image, depth, normal, pos, nearest, ray_dir = flat[6 * b:6 * b + 6]
        H, W = m.height, m.width
        out.append({'image': image.view(H, W, 3), 'depth': depth.view(H, W), 'normal': normal.view(H, W, 3),
                    'pos': pos.view(H, W, 3), 'ray_dist': None, 'nearest': nearest.view(H, W), 'ray_dir': ray_dir})
    return out
