"""Drop-ins for the projection layer's scatter renderers:

    scatter_mean_dim0              diffrend/torch/utils.py:146-175
    scatter_weighted_blended_oit   diffrend/torch/utils.py:178-215
    project_surfels                diffrend/torch/projection_layer.py:20-45
    project_image_coordinates      diffrend/torch/projection_layer.py:48-86
    projection_renderer            diffrend/torch/projection_layer.py:88-106

Same signatures, same outputs; the arithmetic runs in libsurf_b200.so (csrc/surf_scatter.cuh: one projection kernel,
an atomic scatter kernel + normalisation, gather-form backward kernels) instead of scatter_add_ over padded copies of
every operand.  Differentiable like the reference: w.r.t. the scattered data, the OIT depths / centre distances, and -
through project_image_coordinates' pixel coordinates - the surfel positions.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi
from ._lib import check, lib
from .marshal import _scalar
from .renderer import _stream_ptr


def _need_cuda(t):
    if not t.is_cuda:
        raise RuntimeError('surf_renderer_b200.projection needs CUDA tensors; there is no CPU path')


def _f32c(t):
    t = t if t.dtype == torch.float32 else t.float()
    return t if t.is_contiguous() else t.contiguous()


class _ScatterFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx, z, cd2, mode, sigma, z_scale, use_depth, use_center_dist):
        _need_cuda(x)
        x = _f32c(x)
        B, n, ch = x.shape
        idx = idx.long().contiguous()
        z = _f32c(z) if z is not None else None
        cd2 = _f32c(cd2) if cd2 is not None else None
        out = torch.empty(B, n, ch, dtype=torch.float32, device=x.device)
        denom = torch.empty(B, n, dtype=torch.float32, device=x.device)
        mask = torch.empty(B, n, ch, dtype=torch.uint8, device=x.device) if mode == 0 else None
        desc = _abi.SurfScatter(B, n, ch, n, mode, float(sigma), float(z_scale), int(bool(use_depth)), int(bool(use_center_dist)))
        ptr = lambda t: None if t is None else t.data_ptr()      # noqa: E731
        with torch.cuda.device(x.device):
            check(lib().surf_scatter_forward(C.byref(desc), x.data_ptr(), idx.data_ptr(), ptr(z), ptr(cd2), out.data_ptr(),
                                             denom.data_ptr(), ptr(mask), _stream_ptr()))
        ctx.desc = desc
        ctx.has = (z is not None, cd2 is not None)
        ctx.save_for_backward(x, idx, out, denom, *[t for t in (z, cd2) if t is not None])
        if mask is None:
            mask = torch.empty(0, dtype=torch.uint8, device=x.device)
        ctx.mark_non_differentiable(mask)
        return out, mask

    @staticmethod
    def backward(ctx, g_out, _g_mask):
        saved = list(ctx.saved_tensors)
        x, idx, out, denom = saved[:4]
        rest = saved[4:]
        z = rest.pop(0) if ctx.has[0] else None
        cd2 = rest.pop(0) if ctx.has[1] else None
        g_out = _f32c(g_out)
        g_x = torch.zeros_like(x) if ctx.needs_input_grad[0] else None
        g_z = torch.zeros_like(z) if (z is not None and ctx.needs_input_grad[2]) else None
        g_c = torch.zeros_like(cd2) if (cd2 is not None and ctx.needs_input_grad[3]) else None
        ptr = lambda t: None if t is None else t.data_ptr()      # noqa: E731
        with torch.cuda.device(x.device):
            check(lib().surf_scatter_backward(C.byref(ctx.desc), x.data_ptr(), idx.data_ptr(), ptr(z), ptr(cd2), out.data_ptr(),
                                              denom.data_ptr(), g_out.data_ptr(), ptr(g_x), ptr(g_z), ptr(g_c), _stream_ptr()))
        return g_x, None, g_z, g_c, None, None, None, None, None


def scatter_mean_dim0(x, idx):
    """utils.py:146-175.  x [B, n, C], idx [B, n] in [0, n] (n = dropped).  Returns (mean per destination [B, n, C],
    mask [B, n, C] bool: nothing landed there)."""
    out, mask = _ScatterFn.apply(x, idx, None, None, 0, 1.0, 0.0, False, False)
    return out, mask.bool()


def scatter_weighted_blended_oit(x, z, center_dist_2, idx, sigma=0.5, z_scale=2, use_depth=True, use_center_dist=True):
    """utils.py:178-215: Weighted Blended Order-Independent Transparency of the surfels landing on one pixel."""
    out, _ = _ScatterFn.apply(x, idx, z if use_depth else None, center_dist_2 if use_center_dist else None, 1, sigma, z_scale,
                              use_depth, use_center_dist)
    return out


def _projection_desc(pos, camera):
    vp = camera['viewport']
    vp = vp.detach().cpu().numpy() if isinstance(vp, torch.Tensor) else np.asarray(vp)
    W, H = int(vp[2] - vp[0]), int(vp[3] - vp[1])
    B, n, stride = pos.shape
    vecs, strides = [], []
    for k in ('eye', 'at', 'up'):
        v = torch.as_tensor(camera[k], dtype=torch.float32, device=pos.device).detach()
        if v.dim() == 1:
            v = v[:3].contiguous()
            strides.append(0)
        else:
            if v.shape[0] != B:
                raise ValueError('camera.%s batch dimension disagrees with the surfels' % k)
            v = v[:, :3].contiguous()
            strides.append(3)
        vecs.append(v)
    desc = _abi.SurfProjection(B, n, stride, W, H, _scalar(camera['fovy']), _scalar(camera['focal_length']),
                               vecs[0].data_ptr(), vecs[1].data_ptr(), vecs[2].data_ptr(), strides[0], strides[1], strides[2])
    return desc, vecs, W, H


class _ProjectFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, camera):
        _need_cuda(pos)
        pos = _f32c(pos)
        desc, vecs, W, H = _projection_desc(pos, camera)
        B, n, _ = pos.shape
        px = torch.empty(B, n, 3, dtype=torch.float32, device=pos.device)
        idx = torch.empty(B, n, dtype=torch.int64, device=pos.device)
        with torch.cuda.device(pos.device):
            check(lib().surf_project_surfels(C.byref(desc), pos.data_ptr(), px.data_ptr(), idx.data_ptr(), _stream_ptr()))
        ctx.desc, ctx.vecs = desc, vecs
        ctx.save_for_backward(pos)
        ctx.mark_non_differentiable(idx)
        return px, idx

    @staticmethod
    def backward(ctx, g_px, _g_idx):
        (pos,) = ctx.saved_tensors
        g_pos = torch.zeros_like(pos)
        g_px = _f32c(g_px)
        with torch.cuda.device(pos.device):
            check(lib().surf_project_surfels_backward(C.byref(ctx.desc), pos.data_ptr(), g_px.data_ptr(), g_pos.data_ptr(), _stream_ptr()))
        return g_pos, None


def project_image_coordinates(surfels, camera):
    """projection_layer.py:48-86.  surfels [B, N, 3|4] world coordinates -> (px_idx [B, N] int64 destination indices,
    W*H for surfels outside the frame; px_coord [B, N, 3] = pixel x, pixel y, depth), differentiable in px_coord."""
    px, idx = _ProjectFn.apply(surfels, camera)
    return idx, px


def project_surfels(surfel_pos_WC, camera):
    """projection_layer.py:20-45: image-plane coordinates (x, y) and depth z = -Z_cam of the surfels, [B, N, 3]."""
    _, px = project_image_coordinates(surfel_pos_WC, camera)
    vp = camera['viewport']
    vp = vp.detach().cpu().numpy() if isinstance(vp, torch.Tensor) else np.asarray(vp)
    W, H = float(vp[2] - vp[0]), float(vp[3] - vp[1])
    h = float(np.tan(_scalar(camera['fovy']) / 2) * 2 * _scalar(camera['focal_length']))
    w = h * (W / H)
    scale = torch.tensor([-w / (W - 1), h / (H - 1), 1.0], dtype=torch.float32, device=px.device)
    shift = torch.tensor([W / 2.0, H / 2.0, 0.0], dtype=torch.float32, device=px.device)
    return (px - shift) * scale


def projection_renderer(surfels, rgb, camera):
    """projection_layer.py:88-106: scatter the surfel data onto the pixels they project to (mean where several land).
    surfels [B, N, 3|4]; rgb [B, N, D] or [B, H, W, D].  Returns (image like rgb, mask like rgb)."""
    px_idx, _ = project_image_coordinates(surfels, camera)
    flat = rgb.reshape(rgb.size(0), -1, rgb.size(-1))
    out, mask = scatter_mean_dim0(flat, px_idx)
    return out.reshape(rgb.shape), mask.reshape(rgb.shape)


# ---------------------------------------------------------------------------------------------------------------
# projection_renderer_differentiable_fast (projection_layer.py:170-279)
# ---------------------------------------------------------------------------------------------------------------
class _BilinearOITFn(torch.autograd.Function):
    """the scatter stage: four bilinear corner scatters blended with Weighted Blended OIT, summed (one kernel)"""

    @staticmethod
    def forward(ctx, px_coord, x, W, H, use_depth, use_center_dist, compute_depth):
        _need_cuda(x)
        px_coord, x = _f32c(px_coord), _f32c(x)
        B, n, ch = x.shape
        desc = _abi.SurfBilinear(B, n, ch, W, H, 0.5, 2.0, int(bool(use_depth)), int(bool(use_center_dist)), int(bool(compute_depth)))
        acc = torch.empty(lib().surf_bilinear_acc_floats(C.byref(desc)), dtype=torch.float32, device=x.device)
        out = torch.empty(B, W * H, ch, dtype=torch.float32, device=x.device)
        mask = torch.empty(B, W * H, dtype=torch.float32, device=x.device)
        depth = torch.empty(B, W * H, dtype=torch.float32, device=x.device) if compute_depth else None
        with torch.cuda.device(x.device):
            check(lib().surf_bilinear_oit_forward(C.byref(desc), px_coord.data_ptr(), x.data_ptr(), acc.data_ptr(), out.data_ptr(),
                                                  mask.data_ptr(), depth.data_ptr() if depth is not None else None, _stream_ptr()))
        ctx.desc = desc
        ctx.save_for_backward(px_coord, x, acc)
        if depth is None:
            depth = out.new_empty(0)
        return out, mask, depth

    @staticmethod
    def backward(ctx, g_out, g_mask, g_depth):
        px_coord, x, acc = ctx.saved_tensors
        g_x = torch.zeros_like(x) if ctx.needs_input_grad[1] else None
        g_px = torch.zeros_like(px_coord) if ctx.needs_input_grad[0] else None
        ptr = lambda t: None if t is None else t.data_ptr()      # noqa: E731
        go = _f32c(g_out) if g_out is not None else None
        gm = _f32c(g_mask) if g_mask is not None else None
        gd = _f32c(g_depth) if (g_depth is not None and ctx.desc.compute_depth and g_depth.numel()) else None
        with torch.cuda.device(x.device):
            check(lib().surf_bilinear_oit_backward(C.byref(ctx.desc), px_coord.data_ptr(), x.data_ptr(), acc.data_ptr(), ptr(go), ptr(gm),
                                                   ptr(gd), ptr(g_x), ptr(g_px), _stream_ptr()))
        return g_px, g_x, None, None, None, None, None


class _BlurFn(torch.autograd.Function):
    """projection_layer.py:156-168 blur(): separable zero-padded Gaussian over a [B, H, W, C] image; self-adjoint"""

    @staticmethod
    def _run(img, sigma):
        img = _f32c(img)
        B, H, W, ch = img.shape
        scratch, out = torch.empty_like(img), torch.empty_like(img)
        with torch.cuda.device(img.device):
            check(lib().surf_gaussian_blur(img.data_ptr(), scratch.data_ptr(), out.data_ptr(), B, H, W, ch, float(sigma), _stream_ptr()))
        return out

    @staticmethod
    def forward(ctx, img, sigma):
        _need_cuda(img)
        ctx.sigma = sigma
        return _BlurFn._run(img, sigma)

    @staticmethod
    def backward(ctx, g):
        return _BlurFn._run(g, ctx.sigma), None


def blur(image_bhwc, blur_size):
    """Gaussian blur of a [B, H, W, C] image with sigma = blur_size * H / 6 (the reference's blur() takes the image as
    [B, C, H, W] and reads the size of its second-to-last axis, i.e. H)."""
    return _BlurFn.apply(image_bhwc, blur_size * image_bhwc.size(1) / 6)


def projection_renderer_differentiable_fast(surfels, rgb, camera, rotated_image=None, blur_size=0.15, use_depth=True,
                                            use_center_dist=True, compute_new_depth=False, blur_rotated_image=True,
                                            detach_mask=False, detach_mask2=False, detach_depth_merge=False):
    """projection_layer.py:170-279: project the surfels, splat them bilinearly with Weighted Blended OIT, blur, and merge
    with `rotated_image` through the soft mask.  Differentiable w.r.t. `rgb`, `rotated_image` and - through the pixel
    coordinates and the depth - the surfel positions.  Returns (out [B,H,W,C], {'mask', 'image1'[, 'depth']})."""
    _, px_coord = project_image_coordinates(surfels, camera)
    vp = camera['viewport']
    vp = vp.detach().cpu().numpy() if isinstance(vp, torch.Tensor) else np.asarray(vp)
    W, H = int(vp[2] - vp[0]), int(vp[3] - vp[1])
    rgb_in = rgb.reshape(rgb.size(0), -1, rgb.size(-1))
    if detach_depth_merge:
        px_coord = torch.cat((px_coord[..., :2], px_coord[..., 2:].detach()), dim=-1)
    rgb_out, soft_mask, depth_out = _BilinearOITFn.apply(px_coord, rgb_in, W, H, use_depth, use_center_dist, compute_new_depth)
    rgb_out = blur(rgb_out.view(*rgb.size()), blur_size)
    soft_mask = blur(soft_mask.view(*rgb.size()[:-1], 1), blur_size)
    # merge (projection_layer.py:247-279; elementwise, the reference's own formulation)
    soft_mask_nonzero = torch.where(soft_mask > 0, soft_mask, torch.ones_like(soft_mask)) + 1e-20
    rgb_out_normalized = torch.where(soft_mask > 0, rgb_out / soft_mask_nonzero, rgb_out)
    if rotated_image is not None:
        if blur_rotated_image:
            rotated_image = blur(rotated_image, blur_size)
        if detach_mask:
            out = torch.where(soft_mask > 1, rgb_out / soft_mask_nonzero.detach(), rgb_out + rotated_image * (1 - soft_mask.detach()))
        elif detach_mask2:
            soft_mask_detached = soft_mask.detach()
            out = soft_mask_detached * rgb_out_normalized + (1 - soft_mask_detached) * rotated_image
        else:
            out = torch.where(soft_mask > 1, rgb_out / soft_mask_nonzero, rgb_out + rotated_image * (1 - soft_mask))
    else:
        out = rgb_out_normalized
    proj_out = {'mask': soft_mask, 'image1': rgb_out_normalized}
    if compute_new_depth:
        depth_img = depth_out.view(*rgb.size()[:-1], 1)
        proj_out['depth'] = torch.where(soft_mask > 0, depth_img / soft_mask_nonzero, depth_img)
    return out, proj_out
