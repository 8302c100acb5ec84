"""Scene dict (the reference's schema, SURVEY A.1 / renderer.py:136-355) -> flat tensors + C structs.

The render path reads: scene['camera'], ['lights'], ['colors'], ['materials'], ['objects'] (dict, in
insertion order - utils.py:486) and optionally ['tonemap'].  Values may be tensors, numpy arrays, lists
or python numbers (torch/render.py:88-107 make_torch_var produces tensors; the demos mix all of them).
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np
import torch

from . import _abi

PRIM_FIELDS = {'disk': ('pos', 'normal', 'radius'), 'plane': ('pos', 'normal'),
               'sphere': ('pos', 'radius'), 'triangle': ('face', 'normal')}


def _as_float_tensor(v, device):
    if isinstance(v, torch.Tensor):
        t = v
        if t.dtype != torch.float32:
            t = t.float()
        if t.device != device:
            t = t.to(device)
    else:
        t = torch.tensor(np.asarray(v, dtype=np.float64), dtype=torch.float32, device=device)
    return t if t.is_contiguous() else t.contiguous()


_INT_CACHE = {}          # (id(src), version, device) -> (weakref(src), int32 tensor); index arrays rarely change
_INT_CACHE_MAX = 256


def _as_int_tensor(v, device):
    """index arrays may arrive as long or float tensors, lists or numpy arrays (renderer.py:128,284).  The int32
    device copy of a tensor is cached until the tensor is modified in place or collected."""
    if isinstance(v, torch.Tensor):
        if v.dtype == torch.int32 and v.device == device and v.is_contiguous():
            return v                        # already in the library's format (e.g. a slice of a marshalled batch)
        key = (id(v), v._version, str(device))
        hit = _INT_CACHE.get(key)
        if hit is not None and hit[0]() is v:
            return hit[1]
        t = v.detach()
    else:
        key = None
        t = torch.tensor(np.asarray(v))
    t = t.long().to(torch.int32)          # .long() truncation first, like the reference's material_idx.long()
    if t.device != device:
        t = t.to(device)
    t = t.contiguous()
    if key is not None:
        if len(_INT_CACHE) >= _INT_CACHE_MAX:
            _INT_CACHE.clear()
        _INT_CACHE[key] = (weakref.ref(v), t)
    return t


def _host_index_range(v):
    """(min, max) of an index array held in HOST memory (list / numpy / CPU tensor), None for device tensors - those
    are range-checked by the kernels (clamped + flagged, see surf_check_indices) without a synchronisation."""
    if isinstance(v, torch.Tensor):
        if v.is_cuda or v.numel() == 0:
            return None
        t = v.detach().long()
        return int(t.min()), int(t.max())
    a = np.asarray(v)
    if a.size == 0:
        return None
    a = a.astype(np.int64)
    return int(a.min()), int(a.max())


def _scalar(v):
    if isinstance(v, torch.Tensor):
        return float(v.detach().cpu().reshape(-1)[0])
    return float(np.asarray(v).reshape(-1)[0])


class Marshalled:
    """Flat, C-ready view of one scene dict on one device."""

    def __init__(self, scene, device):
        device = torch.device(device)
        self.device = device
        cam = scene['camera']
        if 'proj_type' not in cam:
            raise KeyError('proj_type')
        proj = cam['proj_type']
        if proj in ('persp', 'perspective'):
            self.proj = 0
        elif proj in ('ortho', 'orthographic'):
            self.proj = 1
        else:
            raise ValueError('Invalid projection type')        # utils.py:470 falls through; Renderer:70 raises
        vp = cam['viewport']
        vp = vp.detach().cpu().numpy() if isinstance(vp, torch.Tensor) else np.asarray(vp)
        self.width, self.height = int(vp[2] - vp[0]), int(vp[3] - vp[1])
        self.fovy, self.focal = _scalar(cam['fovy']), _scalar(cam['focal_length'])
        self.near, self.far = _scalar(cam['near']), _scalar(cam['far'])
        # camera vectors are constants of the render (the reference cannot differentiate them: utils.py:476)
        self.cam_vecs = {k: _as_float_tensor(cam[k], device).detach().reshape(-1)[:3].contiguous()
                         for k in ('eye', 'at', 'up')}

        self.names, self.floats = [], []      # differentiable float inputs, canonical order
        self.ints = {}
        self._host_ranges = {}                # index array -> (min, max) when the caller's array is host data
        self.sets = []                        # (kind, count, pos_stride, normal_stride)

        def add(name, v):
            self.names.append(name)
            self.floats.append(_as_float_tensor(v, device))
            return len(self.floats) - 1

        objects = scene['objects']
        if len(objects) == 0 or len(objects) > _abi.SURF_MAX_SETS:
            raise ValueError('scene needs between 1 and %d primitive sets' % _abi.SURF_MAX_SETS)
        for kind, prim in objects.items():
            if kind not in _abi.KIND:
                raise KeyError(kind)                       # utils.py:488 intersection_fn[obj_type]
            slots = {}
            for f in PRIM_FIELDS[kind]:
                slots[f] = add('objects/%s/%s' % (kind, f), prim[f])
            geo = self.floats[slots['face' if kind == 'triangle' else 'pos']]
            count, pstride = int(geo.shape[0]), int(geo.shape[-1])
            nstride = int(self.floats[slots['normal']].shape[-1]) if 'normal' in slots else 0
            self.ints['objects/%s/material_idx' % kind] = _as_int_tensor(prim['material_idx'], device)
            self._host_ranges['objects/%s/material_idx' % kind] = _host_index_range(prim['material_idx'])
            self.sets.append((kind, count, pstride, nstride, slots))

        lights = scene['lights']
        self.i_light_pos = add('lights/pos', lights['pos'])
        self.i_atten = add('lights/attenuation', lights['attenuation'])
        self.i_ambient = add('lights/ambient', lights['ambient'])
        self.ints['lights/color_idx'] = _as_int_tensor(lights['color_idx'], device)
        self._host_ranges['lights/color_idx'] = _host_index_range(lights['color_idx'])
        self.i_colors = add('colors', scene['colors'])
        self.i_albedo = add('materials/albedo', scene['materials']['albedo'])
        self.i_coeffs = add('materials/coeffs', scene['materials']['coeffs'])
        self.i_gamma = None
        if 'tonemap' in scene:
            tm = scene['tonemap']
            if tm['type'] != 'gamma':
                raise ValueError('only the gamma tonemap exists (utils.py:430-432)')
            self.i_gamma = add('tonemap/gamma', tm['gamma'] if isinstance(tm['gamma'], torch.Tensor)
                               else [float(tm['gamma'])])
        self.total_prims = sum(s[1] for s in self.sets)
        self.n_pixels = self.width * self.height
        self._validate()

    def _validate(self):
        # the demos replace albedo with a 1-row table but keep SCENE_BASIC's 6-row coeffs
        # (full_diff_renderer_demo.py:63): rows are only ever gathered by material_idx, so use the smaller table
        n_mat = min(int(self.floats[self.i_albedo].shape[0]), int(self.floats[self.i_coeffs].shape[0]))
        n_col = int(self.floats[self.i_colors].shape[0])
        for (kind, count, _, _, _) in self.sets:
            mi = self.ints['objects/%s/material_idx' % kind]
            if mi.numel() != count:
                raise ValueError('%s: material_idx has %d entries for %d primitives' % (kind, mi.numel(), count))
        # Index ranges (the reference's index_select raises IndexError, renderer.py:284-286).  Host data (lists, numpy
        # arrays, CPU tensors) is checked here; device tensors cannot be read without a synchronisation - the kernels
        # clamp them and flag the workspace, which render(..., _check_indices=True) / SURF_B200_CHECK_INDICES=1 turns
        # into the same IndexError.
        for name, rng in self._host_ranges.items():
            if rng is None:
                continue
            hi = n_col if name == 'lights/color_idx' else n_mat
            if rng[0] < 0 or rng[1] >= hi:
                raise IndexError('%s out of range: [%d, %d] for %d rows' % (name, rng[0], rng[1], hi))

    # ------------------------------------------------------------------
    def c_scene(self, floats=None):
        """SurfScene over `floats` (defaults to the marshalled tensors; the autograd Function passes its args)."""
        fl = self.floats if floats is None else floats
        sc = _abi.SurfScene()
        sc.n_sets = len(self.sets)
        for k, (kind, count, pstride, nstride, slots) in enumerate(self.sets):
            ps = sc.sets[k]
            ps.kind, ps.count = _abi.KIND[kind], count
            ps.pos = fl[slots['face' if kind == 'triangle' else 'pos']].data_ptr()
            ps.pos_stride = pstride
            if 'normal' in slots:
                ps.normal = fl[slots['normal']].data_ptr()
                ps.normal_stride = nstride
            if 'radius' in slots:
                ps.radius = fl[slots['radius']].data_ptr()
            ps.material_idx = self.ints['objects/%s/material_idx' % kind].data_ptr()
        lp = fl[self.i_light_pos]
        sc.n_lights, sc.light_pos, sc.light_pos_stride = int(lp.shape[0]), lp.data_ptr(), int(lp.shape[-1])
        sc.light_color_idx = self.ints['lights/color_idx'].data_ptr()
        sc.light_attenuation = fl[self.i_atten].data_ptr()
        sc.ambient = fl[self.i_ambient].data_ptr()
        sc.n_colors, sc.colors = int(fl[self.i_colors].shape[0]), fl[self.i_colors].data_ptr()
        sc.n_materials = min(int(fl[self.i_albedo].shape[0]), int(fl[self.i_coeffs].shape[0]))
        sc.albedo, sc.coeffs = fl[self.i_albedo].data_ptr(), fl[self.i_coeffs].data_ptr()
        sc.gamma = fl[self.i_gamma].data_ptr() if self.i_gamma is not None else None
        return sc

    def c_camera(self):
        cam = _abi.SurfCamera()
        cam.proj, cam.width, cam.height = self.proj, self.width, self.height
        cam.fovy, cam.focal_length = self.fovy, self.focal
        cam.eye, cam.at, cam.up = (self.cam_vecs[k].data_ptr() for k in ('eye', 'at', 'up'))
        cam.near_clip, cam.far_clip = self.near, self.far
        return cam

    def c_grads(self, grads):
        """SurfSceneGrads over a list of gradient tensors aligned with self.floats (None = not wanted)."""
        sg = _abi.SurfSceneGrads()

        def ptr(i):
            return grads[i].data_ptr() if (i is not None and grads[i] is not None) else None

        for k, (kind, count, pstride, nstride, slots) in enumerate(self.sets):
            g = sg.sets[k]
            g.pos = ptr(slots.get('face' if kind == 'triangle' else 'pos'))
            g.normal = ptr(slots.get('normal'))
            g.radius = ptr(slots.get('radius')) if kind == 'sphere' else None
        sg.light_pos, sg.light_attenuation, sg.ambient = ptr(self.i_light_pos), ptr(self.i_atten), ptr(self.i_ambient)
        sg.colors, sg.albedo, sg.coeffs, sg.gamma = ptr(self.i_colors), ptr(self.i_albedo), ptr(self.i_coeffs), ptr(self.i_gamma)
        return sg


class MarshalledBatch:
    """Strided batch (include/surf_b200.h, SurfBatchLayout): ONE scene dict in which any tensor may carry a leading
    batch dimension B - splat positions [B,M,3], camera eyes [B,4], ... - and everything else is shared by all B
    scenes.  `m` is the Marshalled view of scene 0 (validation, slot order, base pointers); `fulls` are the whole
    (batched or shared) float tensors aligned with `m.floats`; `strides` their per-scene element strides."""

    _NDIM = {'pos': 2, 'normal': 2, 'radius': 1, 'face': 3}

    def __init__(self, scene, device):
        device = torch.device(device)
        self.batch = None
        owners = {}                       # id(scene-0 view) -> (full tensor, stride)

        def split(v, base_ndim, as_int=False):
            full = _as_int_tensor(v, device) if as_int else _as_float_tensor(v, device)
            if full.dim() == base_ndim + 1:
                if self.batch is None:
                    self.batch = int(full.shape[0])
                elif int(full.shape[0]) != self.batch:
                    raise ValueError('batched scene: leading dimensions %d and %d disagree' % (self.batch, full.shape[0]))
                view, stride = full[0], int(full.stride(0))
            elif full.dim() == base_ndim or (base_ndim == 1 and full.dim() == 0):
                view, stride = full, 0
            else:
                raise ValueError('batched scene: unexpected tensor rank %d (expected %d or %d)' % (full.dim(), base_ndim, base_ndim + 1))
            owners[id(view)] = (full, stride, view)
            return view

        s0 = {'objects': {}}
        self.int_strides = {}
        for kind, prim in scene['objects'].items():
            if kind not in PRIM_FIELDS:
                raise KeyError(kind)
            o = {f: split(prim[f], self._NDIM[f]) for f in PRIM_FIELDS[kind]}
            o['material_idx'] = split(prim['material_idx'], 1, as_int=True)
            self.int_strides['objects/%s/material_idx' % kind] = owners[id(o['material_idx'])][1]
            s0['objects'][kind] = o
        lights = scene['lights']
        s0['lights'] = {'pos': split(lights['pos'], 2), 'attenuation': split(lights['attenuation'], 2),
                        'ambient': split(lights['ambient'], 1), 'color_idx': split(lights['color_idx'], 1, as_int=True)}
        self.int_strides['lights/color_idx'] = owners[id(s0['lights']['color_idx'])][1]
        s0['colors'] = split(scene['colors'], 2)
        s0['materials'] = {'albedo': split(scene['materials']['albedo'], 2), 'coeffs': split(scene['materials']['coeffs'], 2)}
        if 'tonemap' in scene:
            tm = scene['tonemap']
            g = tm['gamma'] if isinstance(tm['gamma'], torch.Tensor) else [float(tm['gamma'])]
            s0['tonemap'] = {'type': tm['type'], 'gamma': split(g, 1)}
        cam = dict(scene['camera'])
        self.cam_strides = {}
        for k in ('eye', 'at', 'up'):
            cam[k] = split(cam[k], 1)
            self.cam_strides[k] = owners[id(cam[k])][1]
        s0['camera'] = cam
        if self.batch is None:
            raise ValueError('batched scene: no tensor carries a batch dimension')
        self.m = Marshalled(s0, device)
        for k in ('eye', 'at', 'up'):         # Marshalled re-slices the camera vectors; they must still alias the batch
            if self.m.cam_vecs[k].data_ptr() != cam[k].data_ptr():
                raise AssertionError('camera vector was copied while marshalling')
        self.fulls, self.strides = [], []
        for t in self.m.floats:
            full, stride, _ = owners[id(t)]
            self.fulls.append(full)
            self.strides.append(stride)

    def views0(self, fulls):
        """scene-0 views of the (autograd-unpacked) full tensors, for Marshalled.c_scene"""
        return [f[0] if st else f for f, st in zip(fulls, self.strides)]

    def c_layout(self):
        lay = _abi.SurfBatchLayout()
        m = self.m
        for k, (kind, _count, _ps, _ns, slots) in enumerate(m.sets):
            lay.set_pos[k] = self.strides[slots['face' if kind == 'triangle' else 'pos']]
            if 'normal' in slots:
                lay.set_normal[k] = self.strides[slots['normal']]
            if 'radius' in slots:
                lay.set_radius[k] = self.strides[slots['radius']]
            lay.set_material_idx[k] = self.int_strides['objects/%s/material_idx' % kind]
        lay.light_pos, lay.light_attenuation = self.strides[m.i_light_pos], self.strides[m.i_atten]
        lay.ambient, lay.light_color_idx = self.strides[m.i_ambient], self.int_strides['lights/color_idx']
        lay.colors, lay.albedo, lay.coeffs = self.strides[m.i_colors], self.strides[m.i_albedo], self.strides[m.i_coeffs]
        lay.gamma = self.strides[m.i_gamma] if m.i_gamma is not None else 0
        lay.eye, lay.at, lay.up = (self.cam_strides[k] for k in ('eye', 'at', 'up'))
        return lay


_BATCH_BASE_NDIM = {('objects', 'pos'): 2, ('objects', 'normal'): 2, ('objects', 'radius'): 1, ('objects', 'face'): 3,
                    ('objects', 'material_idx'): 1, ('lights', 'pos'): 2, ('lights', 'attenuation'): 2,
                    ('lights', 'ambient'): 1, ('lights', 'color_idx'): 1, ('colors', None): 2,
                    ('materials', 'albedo'): 2, ('materials', 'coeffs'): 2, ('tonemap', 'gamma'): 1,
                    ('camera', 'eye'): 1, ('camera', 'at'): 1, ('camera', 'up'): 1}


def batched_scene_size(scene):
    """Batch size B of a batched scene dict (the leading dimension of its batched tensors); None if nothing is batched."""
    def walk(group, d):
        for f, v in d.items():
            base = _BATCH_BASE_NDIM.get((group, f))
            if base is not None and isinstance(v, torch.Tensor) and v.dim() == base + 1:
                return int(v.shape[0])
        return None
    for kind, prim in scene['objects'].items():
        b = walk('objects', prim)
        if b is not None:
            return b
    for group in ('camera', 'lights', 'materials', 'tonemap'):
        if group in scene:
            b = walk(group, scene[group])
            if b is not None:
                return b
    v = scene.get('colors')
    return int(v.shape[0]) if isinstance(v, torch.Tensor) and v.dim() == 3 else None


def select_scenes(scene, index):
    """Sub-batch of a batched scene dict (see MarshalledBatch): every tensor that carries the batch dimension is
    indexed with `index` (a list / LongTensor / slice of scene numbers), shared tensors are passed through.  Used to
    shard a batch over ranks (dist.shard_scenes) - the selection is differentiable like any tensor indexing."""
    def pick(group, field, v):
        base = _BATCH_BASE_NDIM.get((group, field))
        if base is not None and isinstance(v, torch.Tensor) and v.dim() == base + 1:
            return v[index]
        return v
    out = {}
    for key, val in scene.items():
        if key == 'objects':
            out[key] = {kind: {f: pick('objects', f, v) for f, v in prim.items()} for kind, prim in val.items()}
        elif key == 'colors':
            out[key] = pick('colors', None, val)
        elif isinstance(val, dict):
            out[key] = {f: pick(key, f, v) for f, v in val.items()}
        else:
            out[key] = val
    return out


# process-wide default of the intersection-kernel variant (SurfOptions.math_mode); a call's `_math_mode` kwarg wins.
# 0 = ray-plane FFMA2 filter (default), 3 = screen-space level-1 test (fastest), 4 = dense small frames.  Also settable through the
# environment variable SURF_INTERSECT_MODE for drop-in use without touching call sites.
import os as _os

_DEFAULT_MATH_MODE = {'plane': 0, 'screen': 3, 'dense': 4}.get(_os.environ.get('SURF_INTERSECT_MODE', 'plane'), None)
if _DEFAULT_MATH_MODE is None:
    _DEFAULT_MATH_MODE = int(_os.environ['SURF_INTERSECT_MODE'])


def set_default_intersect_mode(mode):
    """'plane' (0, default), 'screen' (3) or 'dense' (4: small frames with wide splats); returns the previous value.
    All modes give bit-identical outputs."""
    global _DEFAULT_MATH_MODE
    prev = _DEFAULT_MATH_MODE
    _DEFAULT_MATH_MODE = {'plane': 0, 'screen': 3, 'dense': 4}.get(mode, mode)
    if _DEFAULT_MATH_MODE not in (0, 1, 2, 3, 4):
        _DEFAULT_MATH_MODE = prev
        raise ValueError("mode must be 'plane', 'screen', 'dense' or 0..4")
    return prev


def make_options(params, pixel_range=None, forced_nearest=False):
    opt = _abi.SurfOptions()
    opt.double_sided = int(bool(params.get('double_sided', False)))
    opt.use_quartic = int(bool(params.get('use_quartic', False)))
    opt.shadow = int(bool(params.get('shadow', False)))
    if pixel_range is not None:
        opt.pixel_begin, opt.pixel_end = int(pixel_range[0]), int(pixel_range[1])
    opt.forced_nearest = int(forced_nearest)
    opt.pixels_per_thread = int(params.get('_pixels_per_thread', 0))
    opt.chunk_prims = int(params.get('_chunk_prims', 0))
    opt.math_mode = int(params.get('_math_mode', _DEFAULT_MATH_MODE))
    return opt
