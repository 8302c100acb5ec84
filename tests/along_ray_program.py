"""TEST INFRASTRUCTURE (not part of the product path): a differentiable torch tensor program that prepares the
fragments of render_splats_along_ray - estimated normals (utils.py:854-923) and supersampled plane intersections
(renderer.py:603-673) - as explicit positions / normals.  The CPU emulation tests feed those to the emulated shading
kernels (tests/emul) and chain the gradients back through it with autograd; the product computes the same quantities
in CUDA (surf_renderer_b200/csrc/surf_splats.cuh: k_splat_normals, k_splat_forward, k_splat_normals_backward).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from surf_renderer_b200 import _abi
from surf_renderer_b200.marshal import _as_float_tensor, _as_int_tensor, _scalar, make_options
from surf_renderer_b200.renderer import get_param_value


class _SplatInputs:
    """Flat inputs of one call.  `explicit` = (pos [n,3], normal [n,3], material_idx [n] | None, light_vis [L,n] | None,
    H, W) when the fragments were prepared by the tensor program (estimated normals / supersampling)."""

    def __init__(self, scene, device, explicit=None):
        cam = scene['camera']
        vp = cam['viewport']
        vp = vp.detach().cpu().numpy() if isinstance(vp, torch.Tensor) else np.asarray(vp)
        self.width, self.height = int(vp[2] - vp[0]), int(vp[3] - vp[1])
        self.n = self.width * self.height
        self.fovy, self.focal = _scalar(cam['fovy']), _scalar(cam['focal_length'])
        self.near, self.far = _scalar(cam.get('near', 0.1)), _scalar(cam.get('far', 1000.0))
        self.cam_vecs = {k: _as_float_tensor(cam[k], device).detach().reshape(-1)[:3].contiguous() for k in ('eye', 'at', 'up')}
        disk = scene['objects']['disk']
        self.explicit = explicit is not None
        if self.explicit:
            e_pos, e_normal, e_mat, e_vis, self.height, self.width = explicit
            self.n = self.width * self.height
            disk = {'pos': e_pos, 'normal': e_normal, 'material_idx': e_mat, 'light_vis': e_vis}
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
        if self.z_stride not in (1, 3) or (self.explicit and self.z_stride != 3):
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
        if self.explicit:
            sp.pos = z.data_ptr()
        sp.z = z.data_ptr() + (8 if self.z_stride == 3 else 0)          # column 2 of a [N,3] position array
        sp.z_stride, sp.normal, sp.normal_stride = self.z_stride, nrm.data_ptr(), int(nrm.shape[-1])
        sp.material_idx = self.mat.data_ptr() if self.mat is not None else None
        sp.light_vis = self.vis.data_ptr() if self.vis is not None else None
        return sc, cam, sp, make_options(params)


# ---------------------------------------------------------------------------------------------------------------
# device-side tensor programs in front of the shading kernels (all differentiable torch ops on the scene's device)
# ---------------------------------------------------------------------------------------------------------------
def _unit(u, eps=1e-10):
    """utils.py:135-139 normalize: eps inside the sum, zero lengths divide by one."""
    length = torch.sqrt(torch.sum(u * u + eps, dim=-1, keepdim=True))
    return u / torch.where(length > 0, length, torch.ones_like(length))


def _image_plane(inp, device):
    """image-plane coordinates of the pixel grid: float64 linspace -> f32 -> scaled in f32 (renderer.py:570-577)"""
    h = np.tan(inp.fovy / 2) * 2 * inp.focal
    w = h * (inp.width / inp.height)
    gx, gy = np.meshgrid(np.linspace(-1, 1, inp.width), np.linspace(1, -1, inp.height))
    x = torch.tensor(gx.ravel(), dtype=torch.float32, device=device) * (w / 2)
    y = torch.tensor(gy.ravel(), dtype=torch.float32, device=device) * (h / 2)
    return x, y, w, h


def _neighbour_diffs(pos_hw):
    """utils.py:772-792 grad_spatial2d: differences to the 8 neighbours with reflect padding -> [8, H, W, 3]"""
    xp = torch.nn.functional.pad(pos_hw.permute(2, 0, 1)[None], (1, 1, 1, 1), mode='reflect')[0].permute(1, 2, 0)
    Hp, Wp = xp.shape[:2]
    centre = xp[1:-1, 1:-1, :]
    return torch.stack([xp[1 + dy:Hp + dy - 1, 1 + dx:Wp + dx - 1, :] - centre
                        for dy in (-1, 0, 1) for dx in (-1, 0, 1) if not (dx == 0 and dy == 0)], dim=0)


def _normals_plane_fit(pos_hw):
    """utils.py:886-923: least-squares plane through each splat and its 8 neighbours with n_z fixed to 1."""
    nd = _unit(_neighbour_diffs(pos_hw)).reshape(8, -1, 3)
    m = nd[:, :, :2].transpose(0, 1)                                    # [N, 8, 2]
    mtm = m.transpose(1, 2) @ m                                         # [N, 2, 2]
    a, b, c, d = mtm[:, 0, 0], mtm[:, 0, 1], mtm[:, 1, 0], mtm[:, 1, 1]
    det = a * d - b * c + 1e-12
    rhs = (m.transpose(1, 2) @ (-nd[:, :, 2].transpose(0, 1)[:, :, None]))[:, :, 0]        # [N, 2]
    nx = (d * rhs[:, 0] - b * rhs[:, 1]) / det
    ny = (-c * rhs[:, 0] + a * rhs[:, 1]) / det
    return _unit(torch.stack((nx, ny, torch.ones_like(nx)), dim=1))


def _normals_average(pos_hw):
    """utils.py:854-883 find_average_normal: mean of the 8 neighbour-difference cross products, clamped to [0, 1]."""
    nd = _unit(_neighbour_diffs(pos_hw))
    ring = ((4, 2), (2, 1), (1, 0), (0, 3), (3, 5), (5, 6), (6, 7), (7, 4))
    n = torch.stack([torch.linalg.cross(nd[i], nd[j], dim=-1) for i, j in ring], dim=0).mean(dim=0)
    return torch.clamp(_unit(n), 0.0, 1.0).reshape(-1, 3)


def _upsampled(x, H, W, C, K):
    """renderer.py:476-481 reshape_upsampled_data: [N, C, K*K] (sub-column major) -> [H*K * W*K, C] row-major"""
    return x.view(H, W, C, K, K).permute(0, 3, 1, 4, 2).contiguous().view(H * W * K * K, C)


def _prepare_fragments(scene, inp0, params, device):
    """positions / normals / materials / visibility of the (possibly supersampled) fragments, as tensors"""
    disk = scene['objects']['disk']
    z_in = _as_float_tensor(disk['pos'], device)
    x, y, w, h = _image_plane(inp0, device)
    H, W, focal = inp0.height, inp0.width, inp0.focal
    Z = -torch.relu(-(z_in if z_in.dim() == 1 else z_in[:, 2]))
    pos = torch.stack((-Z * x / focal, -Z * y / focal, Z), dim=1)
    normals = get_param_value('normal', disk, None)
    if normals is None:
        method = get_param_value('normal_estimation_method', params, 'plane')
        if method not in ('plane', 'avg_normal'):
            raise ValueError("normal_estimation_method must be 'plane' or 'avg_normal'")       # 'quadric' maps to None upstream
        normals = (_normals_plane_fit if method == 'plane' else _normals_average)(pos.view(H, W, 3))
    else:
        normals = _as_float_tensor(normals, device)[:, :3]
    mat = disk.get('material_idx', None)
    mat = _as_int_tensor(mat, device) if mat is not None else None
    vis = disk.get('light_vis', None)
    vis = _as_float_tensor(vis, device).detach() if vis is not None else None
    K = int(get_param_value('samples', params, 1))
    if K > 1:
        if mat is None:
            raise AssertionError('supersampling needs material_idx (renderer.py:605)')
        plane_d = torch.sum(pos * normals, dim=1)
        zz = torch.full_like(x, -focal)
        sub_w, sub_h = w / (K * W - 1), h / (K * H - 1)
        p_ss = []
        for deltax in np.linspace(-1, 1, K):
            xx = x + float(deltax * sub_w / 2)
            for deltay in np.linspace(1, -1, K):
                yy = y + float(deltay * sub_h / 2)
                ray = _unit(torch.stack((xx, yy, zz), dim=1))
                t = plane_d / torch.sum(ray * normals, dim=1)
                p_ss.append(t[:, None] * ray)
        pos = _upsampled(torch.stack(p_ss, dim=2), H, W, 3, K)
        normals = _upsampled(normals[:, :, None].expand(-1, -1, K * K).contiguous(), H, W, 3, K)
        mat = _upsampled(mat[:, None, None].expand(-1, 1, K * K).contiguous(), H, W, 1, K).view(-1).contiguous()
        if vis is not None:
            L = vis.shape[0]
            vis = _upsampled(vis.transpose(0, 1)[:, :, None].expand(-1, -1, K * K).contiguous(), H, W, L, K).transpose(0, 1).contiguous()
        H, W = H * K, W * K
    return pos.contiguous(), normals.contiguous(), mat, vis, H, W


def build_inputs(scene, params, device):
    """_SplatInputs of a call: the plain z-per-pixel form, or explicit fragments when normals are estimated / samples > 1"""
    disk = scene['objects']['disk']
    if get_param_value('normal', disk, None) is None or get_param_value('samples', params, 1) > 1:
        probe = _SplatInputs.__new__(_SplatInputs)          # camera-only view for the tensor program
        cam = scene['camera']
        vp = cam['viewport']
        vp = vp.detach().cpu().numpy() if isinstance(vp, torch.Tensor) else np.asarray(vp)
        probe.width, probe.height = int(vp[2] - vp[0]), int(vp[3] - vp[1])
        probe.fovy, probe.focal = _scalar(cam['fovy']), _scalar(cam['focal_length'])
        return _SplatInputs(scene, device, explicit=_prepare_fragments(scene, probe, params, device))
    return _SplatInputs(scene, device)


