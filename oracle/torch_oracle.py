"""CPU oracle for the DiffRend ray-cast render path (TEST INFRASTRUCTURE ONLY).

This module is the parity checker for ``surf_renderer_b200``.  It restates, in
plain fp32 torch-CPU tensor ops, the arithmetic of the reference hot path

    diffrend/torch/renderer.py:136-355   render()
    diffrend/torch/renderer.py:82-125    fragment_shader()
    diffrend/torch/utils.py:439-478      generate_rays()
    diffrend/torch/utils.py:238-366      ray_{sphere,plane,disk,triangle}_intersection()
    diffrend/torch/utils.py:481-512      ray_object_intersections()

op for op (same operation order, same materialised ``[M, N]`` / ``[M, N, 3]``
intermediates, same arithmetic ``where`` blend), so that on the same inputs it
produces the same bits as the reference running on torch-CPU and autograd
yields the same gradients.  It is *not* a product path: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline / ``--impl
reference`` legs may import it.  The product (``surf_renderer_b200``) never
does and fails loudly when its CUDA library is missing.

Parity pinning: ``tests/golden/*.npz`` were produced by importing the real
reference from ``/root/reference`` (script: ``tests/golden/make_golden.py``);
``tests/test_oracle_golden.py`` checks this restatement bit-for-bit against
them, and (when ``/root/reference`` is present) against the live reference.
The shadow-ray branch (renderer.py:291-314) is pinned by ``sh_*.npz``
(``make_golden_shadow.py``: the stock branch run on CPU with the one attribute
``torch.cuda.FloatTensor`` it casts with pointed at ``torch.FloatTensor``),
``render_along_ray`` by ``ar_*.npz`` (``make_golden_along_ray.py``).

Extensions that the reference does not have (used by the tests only):
  * ``pixel_subset``: render only the listed flat pixel indices (pixels are
    independent, renderer.py:170-198) - lets the full-size configs be checked
    at sampled pixels without a multi-hour CPU run.
"""
from __future__ import annotations

import math

import numpy as np
import torch

MISS_SENTINEL = 1001  # hard-coded in the reference: utils.py:271,323,363


def _f32(x):
    """utils.py:19-22 ``tch_var_f`` without the device switch (oracle is CPU)."""
    if isinstance(x, torch.Tensor):
        return x
    # fp32 in normal use; tests may switch the default dtype to float64 for a high-precision variant
    return torch.tensor(np.asarray(x, dtype=np.float64), dtype=torch.get_default_dtype())


def blend(cond, a, b):
    """utils.py:58-59: arithmetic select; NaN/inf in the unselected arm leak."""
    c = cond.float()
    return c * a + (1 - c) * b


def safe_div(num, den, epsilon=0):
    """utils.py:87-95 ``nonzero_divide``: zero divisors are replaced by one."""
    nz = (torch.abs(den) > 0).float()
    return num / ((den * nz + (1 - nz)) + epsilon)


def lp_norm(u, p=2, eps=0):
    """utils.py:66-67 ``norm_p`` (eps is added per component inside the sum)."""
    return torch.pow(torch.sum(torch.pow(u, p) + eps, dim=-1), 1. / p)


def unit(u, eps=1e-10):
    """utils.py:135-139 ``normalize``."""
    length = lp_norm(u, 2, eps=eps)
    if u.dim() > 1:
        length = length[..., None]
    return safe_div(u, length)


def rowdot(a, b, axis=0):
    """utils.py:70-71 ``tensor_dot``."""
    return torch.sum(a * b, dim=axis)


def edge_cross(e, rel):
    """utils.py:74-84 ``tensor_cross_prod``: e [M,3] x rel [M,N,3]."""
    ex, ey, ez = e[:, 0][:, None], e[:, 1][:, None], e[:, 2][:, None]
    c0 = ey * rel[..., 2] - ez * rel[..., 1]
    c1 = -ex * rel[..., 2] + ez * rel[..., 0]
    c2 = ex * rel[..., 1] - ey * rel[..., 0]
    return torch.stack((c0, c1, c2), dim=2)


def mirror(incident, normal):
    """utils.py:218-224 ``reflect_ray``."""
    return -2 * torch.sum(incident * normal, dim=-1)[..., None] * normal + incident


def along_ray(origin, direction, dist):
    """utils.py:227-235 ``point_along_ray``: origin [1|N,3], direction [3,N], dist [M,N]."""
    return origin[None, ...] + dist[:, :, None] * direction.transpose(1, 0)[None, ...]


# --------------------------------------------------------------------------
# camera
# --------------------------------------------------------------------------
def camera_rotation(eye, at, up):
    """utils.py:402-427 ``lookat_rot_inv``: columns are the camera x, y, z axes."""
    if up.shape[-1] == 4:
        up = up[..., :3]
    if eye.shape[-1] == 4:
        eye = eye[..., :3] / eye[..., 3]
    if at.shape[-1] == 4:
        at = at[..., :3] / at[..., 3]
    zc = unit(eye - at)
    upn = unit(up)
    xc = unit(torch.linalg.cross(upn, zc))
    yc = torch.linalg.cross(zc, xc)
    return torch.stack((xc, yc, zc), dim=-1)


def camera_pose(eye, at, up):
    """utils.py:385-399 ``lookat_inv``: 4x4 camera-to-world matrix."""
    rot = camera_rotation(eye, at, up)
    top = torch.cat((rot, eye[..., :3][..., None]), dim=-1)
    bottom = _f32([0, 0, 0, 1.]).unsqueeze(0)
    return torch.cat((top, bottom), dim=-2)


def image_plane_grid(camera):
    """utils.py:440-456: pixel-centre-free screen grid, float64 linspace -> f32 -> scaled in f32."""
    vp = np.array(camera['viewport']) if isinstance(camera['viewport'], list) else camera['viewport']
    W, H = vp[2] - vp[0], vp[3] - vp[1]
    aspect = float(W) / float(H)
    gx, gy = np.meshgrid(np.linspace(-1, 1, int(W)), np.linspace(1, -1, int(H)))
    fovy = np.array(camera['fovy']) if isinstance(camera['fovy'], list) else camera['fovy']
    focal = np.array(camera['focal_length']) if isinstance(camera['focal_length'], list) else camera['focal_length']
    h = np.tan(fovy / 2) * 2 * focal
    w = h * aspect
    x = _f32(gx.ravel())
    y = _f32(gy.ravel())
    x *= w / 2
    y *= h / 2
    return x, y, focal, int(H), int(W)


def make_rays(camera):
    """utils.py:439-478 ``generate_rays`` -> (ray_orig, ray_dir [3,N], H, W)."""
    x, y, focal, H, W = image_plane_grid(camera)
    n = x.numel()
    eye, at, up = camera['eye'][:3], camera['at'][:3], camera['up'][:3]
    kind = camera['proj_type']
    if kind in ('ortho', 'orthographic'):
        direction = unit(at - eye)[:, None]
        origin = torch.stack((x, y, _f32(np.zeros(n)), _f32(np.ones(n))), dim=0)
        origin = torch.mm(camera_pose(eye=eye, at=at, up=up), origin)
        origin = (origin[:3] / origin[3][None, :]).permute(1, 0)
    elif kind in ('persp', 'perspective'):
        origin = eye[None, :]
        direction = torch.stack((x, y, _f32(-np.ones(n) * focal)), dim=0)
        direction = torch.mm(camera_rotation(eye=eye, at=at, up=up), direction)
        direction /= torch.sqrt(torch.sum(direction ** 2, dim=0))
    else:
        raise ValueError('Invalid projection type')
    return origin, direction, H, W


# --------------------------------------------------------------------------
# ray / primitive intersections.  Each returns (points [M,N,3], t [M,N], normals [M,N,3])
# --------------------------------------------------------------------------
def hit_plane(origin, direction, prim, want_normals=True):
    """utils.py:281-308."""
    p = prim['pos'][:, :3]
    n = unit(prim['normal'][:, :3])
    offset = torch.sum(p * n, dim=1)
    denom = torch.mm(n, direction)
    t = (offset.unsqueeze(-1) - torch.mm(n, origin.permute(1, 0))) / denom
    pts = along_ray(origin, direction, t)
    normals = n[:, None, :].repeat(1, pts.size()[1], 1) if want_normals else None
    return pts, t, normals


def hit_disk(origin, direction, prim, want_normals=True):
    """utils.py:311-326."""
    pts, t, normals = hit_plane(origin, direction, prim, want_normals)
    c = prim['pos'][:, :3]
    r = prim['radius']
    d2 = torch.sum((pts - c[:, None, :]) ** 2, dim=-1)
    inside = (d2 <= r[:, None] ** 2)
    return pts, blend(inside, t, MISS_SENTINEL), normals


def hit_triangle(origin, direction, prim, want_normals=True):
    """utils.py:329-366 (the plane call there never forwards disable_normals)."""
    face = prim['face']
    pts, t, normals_full = hit_plane(origin, direction, {'pos': face[:, 0, :], 'normal': prim['normal']})
    normals = normals_full[..., :3]
    r0 = pts - face[:, 0, :3][:, None, :]
    r1 = pts - face[:, 1, :3][:, None, :]
    r2 = pts - face[:, 2, :3][:, None, :]
    e01 = face[:, 1, :3] - face[:, 0, :3]
    e12 = face[:, 2, :3] - face[:, 1, :3]
    e20 = face[:, 0, :3] - face[:, 2, :3]
    in0 = torch.sum(edge_cross(e01, r0) * normals, dim=-1) >= 0
    in1 = torch.sum(edge_cross(e12, r1) * normals, dim=-1) >= 0
    in2 = torch.sum(edge_cross(e20, r2) * normals, dim=-1) >= 0
    inside = in0 * in1 * in2
    return pts, blend(inside, t, MISS_SENTINEL), normals_full


SAFE_SPHERE = False   # tests only: NaN-free variant for gradient checks (see hit_sphere)


def hit_sphere(origin, direction, prim, want_normals=True):
    """utils.py:238-278 (always returns normals).

    With SAFE_SPHERE (oracle-only switch, never set by the reference semantics tests) the arithmetic blends
    are replaced by selects and sqrt never sees a masked zero, so autograd yields finite sphere gradients
    (the reference's are NaN), and a sphere entirely behind the ray origin is a miss instead of
    the reference's data-dependent phantom hit."""
    if SAFE_SPHERE:
        return _hit_sphere_safe(origin, direction, prim)
    c = prim['pos'][:, :3]
    oc = origin[None, ...] - c[:, None, :]
    r = prim['radius']
    qa = torch.sum(direction ** 2, dim=0)
    qb = 2 * torch.sum(oc * direction.permute(1, 0)[None, ...], dim=-1)
    qc = (torch.sum(oc ** 2, dim=-1) - r[:, None] ** 2)
    disc = qb ** 2 - 4 * qa * qc
    real = disc >= 0
    disc = blend(real, disc, 0)
    root = torch.sqrt(disc)
    inv = 1. / (2 * qa)
    t1 = (-qb - root) * inv
    t2 = (-qb + root) * inv
    big = torch.max(torch.max(t1, t2)) + 1
    t1 = blend(real * (t1 >= 0), t1, big)
    t2 = blend(real * (t2 >= 0), t2, big)
    t, _ = torch.min(torch.stack((t1, t2), dim=2), dim=2)
    t = blend(real, t, MISS_SENTINEL)
    pts = along_ray(origin, direction, t)
    normals = unit(pts - c[:, None, :])
    return pts, t, normals


def _hit_sphere_safe(origin, direction, prim):
    c = prim['pos'][:, :3]
    oc = origin[None, ...] - c[:, None, :]
    r = prim['radius']
    qa = torch.sum(direction ** 2, dim=0)
    qb = 2 * torch.sum(oc * direction.permute(1, 0)[None, ...], dim=-1)
    qc = (torch.sum(oc ** 2, dim=-1) - r[:, None] ** 2)
    disc = qb ** 2 - 4 * qa * qc
    real = disc >= 0
    root = torch.sqrt(torch.where(real, disc, torch.ones_like(disc)))
    inv = 1. / (2 * qa)
    t1 = (-qb - root) * inv
    t2 = (-qb + root) * inv
    inf = torch.full_like(t1, float('inf'))
    t = torch.minimum(torch.where(real & (t1 >= 0), t1, inf), torch.where(real & (t2 >= 0), t2, inf))
    t = torch.where(torch.isfinite(t), t, torch.full_like(t, float(MISS_SENTINEL)))
    pts = along_ray(origin, direction, t)
    normals = unit(pts - c[:, None, :])
    return pts, t, normals


HIT_FN = {'disk': hit_disk, 'plane': hit_plane, 'sphere': hit_sphere, 'triangle': hit_triangle}


def hit_all(origin, direction, objects, want_normals=True):
    """utils.py:481-512: per-type results concatenated along M in dict order."""
    pts = t = normals = mat = None
    for kind in objects:
        p_k, t_k, n_k = HIT_FN[kind](origin, direction, objects[kind], want_normals)
        if pts is None:
            pts, t, normals, mat = p_k, t_k, n_k, objects[kind]['material_idx']
        else:
            pts = torch.cat((pts, p_k), dim=0)
            t = torch.cat((t, t_k), dim=0)
            if normals is not None:
                normals = torch.cat((normals, n_k), dim=0)
            mat = torch.cat((mat, objects[kind]['material_idx']), dim=0)
    return pts, t, normals, mat


# --------------------------------------------------------------------------
# shading
# --------------------------------------------------------------------------
def phong(frag_normals, to_light, to_eye, atten, coeffs, light_rgb, ambient, albedo,
          double_sided, use_quartic, visibility=None):
    """renderer.py:82-125 ``fragment_shader``: returns [L, N, 3] (ambient included per light)."""
    dist = torch.sqrt(torch.sum(to_light ** 2, dim=-1))[:, :, None]
    ldir = safe_div(to_light, dist)
    power = 2 if not use_quartic else 4
    falloff = safe_div(1, (atten[:, 0][:, None, None] +
                           dist * atten[:, 1][:, None, None] +
                           (dist ** power) * atten[:, 2][:, None, None]))
    lambert = rowdot(frag_normals, falloff * ldir, axis=-1)
    bounce = mirror(-ldir, frag_normals)
    gloss = rowdot(to_eye, bounce, axis=-1)
    if double_sided:
        facing = torch.sign(rowdot(to_eye, frag_normals, axis=-1))
        lambert = facing * lambert
        gloss = facing * gloss
    lambert = torch.nn.functional.relu(lambert)
    gloss = torch.nn.functional.relu(gloss)
    amb = ambient[None, None, :] * albedo[None, :, :]
    tint = light_rgb[:, None, :] * albedo[None, :, :]
    if visibility is not None:
        tint = tint * visibility[:, :, None]
    return (coeffs[:, 0][None, :, None] * lambert[:, :, None] +
            coeffs[:, 1][None, :, None] *
            (gloss[:, :, None] ** coeffs[:, 2][None, :, None])) * tint + amb


def _as_list(v):
    """renderer.py:128-133 ``get_as_list``."""
    if isinstance(v, np.ndarray):
        return v.tolist()
    if isinstance(v, torch.Tensor):
        return v.cpu().numpy().tolist()
    return v


# --------------------------------------------------------------------------
# render
# --------------------------------------------------------------------------
def _zbuffer_tile(origin, direction, objects, near, far, want_normals=True):
    """renderer.py:174-189 for one tile of rays."""
    pts, t, normals, mat = hit_all(origin, direction, objects, want_normals)
    ok = (near <= t) * (t <= far)
    masked = blend(ok, t, far + 1)
    depth, winner = masked.min(0)
    pick = winner[None, :, None].repeat(1, 1, 3)
    frag_n = torch.gather(normals, 0, pick)
    frag_p = torch.gather(pts, 0, pick)
    return depth, winner, frag_n, frag_p, mat, t


def render(scene, **params):
    """renderer.py:136-355.  Same kwargs; plus ``pixel_subset`` (oracle-only extension)."""
    camera = scene['camera']
    origin, direction, H, W = make_rays(camera)
    n_pix = H * W
    objects = scene['objects']
    subset = params.get('pixel_subset', None)
    if subset is not None:
        subset = torch.as_tensor(subset, dtype=torch.long)
        direction = direction[:, subset] if direction.shape[1] > 1 else direction
        origin = origin[subset] if origin.shape[0] > 1 else origin
        n_pix = subset.numel()
    if params.get('vis_stat', False):
        raise RuntimeError('Removed Support for vis_stat')
    persp = origin.shape[0] == 1

    near, far = camera['near'], camera['far']
    if params.get('tiled', True):
        tile = params.get('tile_size', 4096)
        parts = []
        for k in range(int(np.ceil(n_pix / tile))):
            lo, hi = k * tile, min((k + 1) * tile, n_pix)
            d_k = direction[:, lo:hi] if persp else direction
            o_k = origin if persp else origin[lo:hi]
            parts.append(_zbuffer_tile(o_k, d_k, objects, near, far))
        depth = torch.cat([p[0] for p in parts])
        winner = torch.cat([p[1] for p in parts])
        frag_n = torch.cat([p[2] for p in parts], dim=1)
        frag_p = torch.cat([p[3] for p in parts], dim=1)
        mat, last_t = parts[-1][4], parts[-1][5]
    else:
        depth, winner, frag_n, frag_p, mat, last_t = _zbuffer_tile(origin, direction, objects, near, far)

    if params.get('norm_depth_image_only', False):
        # renderer.py:245-260 (the reference reaches this only with tiled=False)
        im_depth = depth.view(H, W) if subset is None else depth.view(1, n_pix)
        min_depth = torch.min(im_depth)
        norm = blend(im_depth >= camera['far'], min_depth, im_depth)
        norm = (norm - min_depth) / (torch.max(im_depth) - min_depth)
        return {'image': norm, 'depth': im_depth, 'nearest': winner.view(*im_depth.shape), 'ray_dir': direction}

    lights = scene['lights']
    light_pos = lights['pos'][:, :3]
    light_rgb = scene['colors'][_as_list(lights['color_idx'])]
    atten = lights['attenuation']
    ambient = lights['ambient']
    frag_mat = torch.gather(mat.long(), 0, winner)
    albedo = torch.index_select(scene['materials']['albedo'], 0, frag_mat)
    coeffs = torch.index_select(scene['materials']['coeffs'], 0, frag_mat)

    visibility = None
    if params.get('shadow', False):
        visibility = _shadow_visibility(light_pos, frag_p, winner, objects, params, n_pix)

    eye3 = camera['eye'][None, None, :3]
    if not persp:
        # the reference uses camera['eye'] for the view vector in both projections (renderer.py:320)
        pass
    per_light = phong(frag_normals=frag_n,
                      to_light=light_pos[:, None, :] - frag_p,
                      to_eye=unit(eye3 - frag_p[:, :, :3]),
                      atten=atten, coeffs=coeffs, light_rgb=light_rgb, ambient=ambient, albedo=albedo,
                      double_sided=params.get('double_sided', False),
                      use_quartic=params.get('use_quartic', False),
                      visibility=visibility)
    shape2 = (H, W) if subset is None else (1, n_pix)
    im = torch.sum(per_light, dim=0).view(shape2[0], shape2[1], 3)
    depth = depth.view(*shape2)
    ok = (near <= depth) * (depth <= far)
    im = ok[:, :, None].float() * im
    im = torch.nn.functional.relu(im)
    if 'tonemap' in scene:
        tm = scene['tonemap']
        if tm['type'] == 'gamma':
            im = torch.pow(im, tm['gamma'])
        else:
            im = None  # utils.py:430-432 returns None for unknown types
    return {
        'image': im,
        'depth': depth,
        'normal': frag_n.view(shape2[0], shape2[1], 3),
        'pos': frag_p.view(shape2[0], shape2[1], 3),
        'ray_dist': last_t,
        'nearest': winner.view(*shape2),
        'ray_dir': direction,
    }


def _shadow_visibility(light_pos, frag_p, winner, objects, params, n_pix):
    """renderer.py:291-314 (the reference casts with torch.cuda.FloatTensor at :311; pinned by tests/golden/sh_*.npz)."""
    tile = params.get('tile_size', 4096)
    out = []
    for li in range(light_pos.shape[0]):
        to_l = (light_pos[li, None, :] - frag_p).squeeze(0).transpose(1, 0)      # [3, N]
        dist_l = lp_norm(to_l.transpose(1, 0), 2)                               # [N]
        to_l = to_l / dist_l
        start = frag_p.squeeze(0) + 0.1 * to_l.transpose(1, 0)                   # [N, 3]
        vis = []
        for k in range(int(np.ceil(n_pix / tile))):
            lo, hi = k * tile, min((k + 1) * tile, n_pix)
            _, t, _, _ = hit_all(start[lo:hi, :], to_l[:, lo:hi], objects, want_normals=False)
            ok = (t > 0) * (t < dist_l[lo:hi])
            t = blend(ok, t, MISS_SENTINEL)
            nearest_t, blocker = t.min(0)
            vis.append((((nearest_t == MISS_SENTINEL) + (blocker == winner[lo:hi])) > 0).float())
        out.append(torch.cat(vis))
    return torch.stack(out, dim=0)


# --------------------------------------------------------------------------
# render_splats_along_ray (renderer.py:537-751): one splat per pixel at depth z along the pixel's ray, shaded in
# camera coordinates (the GAN generator path, GAN/gan.py:563-597), with the 3x3-stencil normal estimation
# (utils.py:772-923) and the K x K supersampling (renderer.py:603-673).
# --------------------------------------------------------------------------
def view_matrix(eye, at, up):
    """utils.py:376-382 ``lookat``: world -> camera, the inverse of camera_pose."""
    return camera_pose(eye, at, up).inverse()


def _pad_reflect(x):
    """utils.py:748-769 ``pad2d(x, (1,1,1,1), 'reflect')`` for an [H, W, C] tensor."""
    xx = x[None, ...].transpose(3, 1)
    xx = torch.nn.ReflectionPad2d((1, 1, 1, 1))(xx).transpose(1, 3)
    return xx[0]


def neighbour_diffs(x):
    """utils.py:772-792 ``grad_spatial2d``: [8, H, W, C] differences to the 8 neighbours (reflect padding)."""
    xp = _pad_reflect(x)
    Hp, Wp = xp.shape[:2]
    centre = xp[1:-1, 1:-1, :]
    out = []
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx == 0 and dy == 0:
                continue
            out.append(xp[1 + dy:Hp + dy - 1, 1 + dx:Wp + dx - 1, :] - centre)
    return torch.stack(out, dim=0)


def normals_plane_fit(pos):
    """utils.py:886-923 ``estimate_surface_normals_plane_fit``: constrained least-squares plane through each splat."""
    nd = unit(neighbour_diffs(pos), 1e-10)
    nd = nd.view(nd.shape[0], -1, 3)
    M = nd[:, :, :2].transpose(1, 0)
    Mt = M.transpose(2, 1)
    MtM = Mt.matmul(M)
    det = MtM[:, 0, 0] * MtM[:, 1, 1] - MtM[:, 0, 1] * MtM[:, 1, 0]
    flip = torch.tensor([1, 0])
    adj = MtM.index_select(1, flip).transpose(2, 1).index_select(1, flip) * _f32([[1, -1], [-1, 1]])[None, ...]
    inv = adj / (det[:, None, None] + 1e-12)
    nxy = inv.matmul(Mt.matmul(-nd[..., 2].transpose(1, 0)[:, :, None])).squeeze()
    n = torch.cat([nxy, _f32(np.ones((nxy.shape[0], 1)))], dim=1).view(pos.shape)
    return unit(n)


def normals_average(pos):
    """utils.py:854-883 ``find_average_normal``."""
    nd = unit(neighbour_diffs(pos), 1e-10)
    order = ((4, 2), (2, 1), (1, 0), (0, 3), (3, 5), (5, 6), (6, 7), (7, 4))
    n = torch.stack([torch.cross(nd[a], nd[b], dim=-1) for a, b in order], dim=0)
    return torch.clamp(unit(torch.mean(n, dim=0), 1e-10), 0.0, 1.0)


def _upsampled(x, H, W, C, K):
    """renderer.py:476-481 ``reshape_upsampled_data``."""
    x = x.view(H, W, C, K, K)
    x = x.transpose(3, 1).transpose(3, 2).transpose(4, 3)
    return x.contiguous().view(H * W * K * K, C)


def render_along_ray(scene, **params):
    camera = scene['camera']
    vp = np.array(camera['viewport'])
    W, H = int(vp[2] - vp[0]), int(vp[3] - vp[1])
    aspect = W / H
    eye, at, up = camera['eye'][:3], camera['at'][:3], camera['up'][:3]
    mcam = view_matrix(eye=eye, at=at, up=up)
    splats = scene['objects']['disk']
    z_in = splats['pos']
    normals_cc = splats.get('normal', None)
    fovy, focal = camera['fovy'], camera['focal_length']
    h = np.tan(fovy / 2) * 2 * focal
    w = h * aspect
    if z_in.dim() == 1:
        Z = -torch.nn.functional.relu(-z_in)
    else:
        Z = -torch.nn.functional.relu(-z_in[:, 2])
    gx, gy = np.meshgrid(np.linspace(-1, 1, W), np.linspace(1, -1, H))
    gx *= w / 2
    gy *= h / 2
    x = _f32(gx.ravel())
    y = _f32(gy.ravel())
    X = -Z * x / focal
    Y = -Z * y / focal
    pos_cc = torch.stack((X, Y, Z), dim=1)
    if normals_cc is None:                                               # renderer.py:591-594
        method = params.get('normal_estimation_method', 'plane')
        est = {'plane': normals_plane_fit, 'avg_normal': normals_average}[method]
        normals_cc = est(pos_cc.view(H, W, 3))[..., :3].view(-1, 3)
    material_idx = splats['material_idx']
    visibility = splats.get('light_vis', None)
    samples = params.get('samples', 1)
    if samples > 1:                                                      # renderer.py:603-673
        plane_d = torch.sum(pos_cc * normals_cc[:, :3], dim=1)
        zz = _f32(np.ones(x.shape) * -focal)
        sub_w = w / (samples * W - 1)
        sub_h = h / (samples * H - 1)
        p_ss, n_ss, m_ss, v_ss = [], [], [], []
        if visibility is not None:
            visibility = visibility.transpose(1, 0)
        for deltax in np.linspace(-1, 1, samples):
            xx = x + deltax * sub_w / 2
            for deltay in np.linspace(1, -1, samples):
                yy = y + deltay * sub_h / 2
                ray = unit(torch.stack((xx, yy, zz), dim=1))
                t = plane_d / torch.sum(ray * normals_cc[:, :3], dim=1)
                p_ss.append(t[:, None] * ray)
                n_ss.append(normals_cc[:, :3])
                m_ss.append(material_idx[:, None])
                if visibility is not None:
                    v_ss.append(visibility)
        pos_cc = _upsampled(torch.stack(p_ss, dim=2), H, W, 3, samples)
        normals_cc = _upsampled(torch.stack(n_ss, dim=2), H, W, 3, samples)
        material_idx = _upsampled(torch.stack(m_ss, dim=2), H, W, 1, samples).view(-1)
        if visibility is not None:
            visibility = _upsampled(torch.stack(v_ss, dim=2), H, W, visibility.shape[1], samples).transpose(1, 0)
        H *= samples
        W *= samples
    depth = lp_norm(pos_cc[..., :3]).view(H, W)
    if params.get('norm_depth_image_only', False):                       # renderer.py:677-686
        lo = torch.min(depth)
        norm = blend(depth >= camera['far'], lo, depth)
        norm = (norm - lo) / (torch.max(depth) - lo)
        return {'image': norm, 'depth': depth, 'pos': pos_cc, 'normal': normals_cc}
    lights = scene['lights']
    light_rgb = scene['colors'][lights['color_idx']]
    light_cc = torch.mm(lights['pos'], mcam.transpose(1, 0))
    frag_n = normals_cc[:, :3]
    frag_p = pos_cc[:, :3]
    if material_idx is not None:
        albedo = torch.index_select(scene['materials']['albedo'], 0, material_idx)
        coeffs = torch.index_select(scene['materials']['coeffs'], 0, material_idx)
    else:
        albedo, coeffs = scene['materials']['albedo'], scene['materials']['coeffs']
    per_light = phong(frag_normals=frag_n, to_light=light_cc[:, None, :3] - frag_p[:, :3],
                      to_eye=-unit(frag_p[None, :, :3]), atten=lights['attenuation'], coeffs=coeffs,
                      light_rgb=light_rgb, ambient=lights['ambient'], albedo=albedo, double_sided=False,
                      use_quartic=params.get('use_quartic', False), visibility=visibility)
    im = torch.nn.functional.relu(torch.sum(per_light, dim=0).view(H, W, 3))
    return {'image': im, 'depth': depth, 'pos': pos_cc[..., :3].view(H, W, 3),
            'normal': normals_cc[..., :3].contiguous().view(H, W, 3)}


# --------------------------------------------------------------------------
# render_splats_NDC (renderer.py:358-474): one splat per pixel given in normalised device coordinates, unprojected
# with the inverse right-handed [-1, 1] perspective matrix (ops.py:7-68) and shaded in camera coordinates with the
# UNNORMALISED view vector -pos (renderer.py:437) - unlike render_along_ray, which normalises it.
# --------------------------------------------------------------------------
def unproject_matrix(fovy, aspect, near, far):
    """ops.py:50-58 ``inv_perspective_RH_NO``."""
    t = np.tan(fovy / 2.)
    m00, m11 = 1 / (aspect * t), 1 / t
    m22, m23 = (near + far) / (far - near), -2 * near * far / (far - near)
    return _f32([[1 / m00, 0, 0, 0], [0, 1 / m11, 0, 0], [0, 0, 0, -1], [0, 0, 1 / m23, -m22 / m23]])


def render_splats_ndc(scene, **params):
    camera = scene['camera']
    vp = np.array(camera['viewport'])
    W, H = int(vp[2] - vp[0]), int(vp[3] - vp[1])
    mcam = view_matrix(eye=camera['eye'][:3], at=camera['at'][:3], up=camera['up'][:3])
    minv = unproject_matrix(camera['fovy'], W / H, camera['near'], camera['far'])
    splats = scene['objects']['disk']
    pos_ndc, normals = splats['pos'], splats['normal']
    if pos_ndc.size()[-1] == 3:
        pos_ndc = torch.cat((pos_ndc, _f32(np.ones((pos_ndc.size()[0], 1)))), dim=1)
    pos_cc = torch.matmul(pos_ndc, minv.transpose(1, 0))
    pos_cc = pos_cc / pos_cc[..., 3][:, None]
    depth = lp_norm(pos_cc[..., :3]).view(H, W)
    if params.get('norm_depth_image_only', False):                       # renderer.py:392-401
        lo = torch.min(depth)
        norm = blend(depth >= camera['far'], lo, depth)
        norm = (norm - lo) / (torch.max(depth) - lo)
        return {'image': norm, 'depth': depth, 'pos': pos_cc, 'normal': normals}
    lights = scene['lights']
    light_rgb = scene['colors'][lights['color_idx']]
    light_cc = torch.mm(lights['pos'], mcam.transpose(1, 0))
    frag_n, frag_p = normals[:, :3], pos_cc[:, :3]
    albedo = torch.index_select(scene['materials']['albedo'], 0, splats['material_idx'])
    coeffs = torch.index_select(scene['materials']['coeffs'], 0, splats['material_idx'])
    per_light = phong(frag_normals=frag_n, to_light=light_cc[:, None, :3] - frag_p[:, :3], to_eye=-frag_p[:, :3],
                      atten=lights['attenuation'], coeffs=coeffs, light_rgb=light_rgb, ambient=lights['ambient'],
                      albedo=albedo, double_sided=params.get('double_sided', False),
                      use_quartic=params.get('use_quartic', False), visibility=None)
    im = torch.nn.functional.relu(torch.sum(per_light, dim=0).view(H, W, 3))
    return {'image': im, 'depth': depth, 'pos': pos_cc[:, :3].view(H, W, 3), 'normal': normals[:, :3].view(H, W, 3)}
