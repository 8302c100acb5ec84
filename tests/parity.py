"""Parity comparison helpers (tests only): candidate render output vs oracle/golden output.

Bar (BASELINE.json north_star): nearest index bit-exact except at documented epsilon-ties (oracle/margin64.py,
SURVEY A.7); image / depth / pos / normal within 1e-4 relative / 1e-5 absolute on non-tie pixels.
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import margin64

RTOL = 1e-4
ATOL = 1e-5
MAX_TIE_FRACTION = 5e-3     # of hit pixels
EPS_TANGENT = 5e-3          # grazing sphere hit: (r - closest approach) / r, see compare_forward


def _to_np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def ray_for_pixel(scene, ref_ray_dir, flat_index, oracle_origin=None):
    """float64 ray of a flat pixel from the oracle's fp32 rays."""
    rd = _to_np(ref_ray_dir).astype(np.float64)
    if rd.shape[1] > 1:
        d = rd[:, flat_index]
        o = _to_np(scene['camera']['eye']).astype(np.float64)[:3]
    else:
        d = rd[:, 0]
        o = _to_np(oracle_origin).astype(np.float64)[flat_index]
    return o, d


def compare_forward(cand, ref, scene, ortho_origins=None, rtol=RTOL, atol=ATOL, check_ray=True):
    """Returns a report dict; raises AssertionError on a parity violation."""
    near, far = float(scene['camera']['near']), float(scene['camera']['far'])
    c_near = _to_np(cand['nearest']).reshape(-1)
    r_near = _to_np(ref['nearest']).reshape(-1)
    c_depth = _to_np(cand['depth']).reshape(-1)
    r_depth = _to_np(ref['depth']).reshape(-1)
    c_hit, r_hit = c_depth <= far, r_depth <= far
    mism = np.nonzero((c_near != r_near) | (c_hit != r_hit))[0]
    ties = {}
    for k in mism:
        o, d = ray_for_pixel(scene, ref['ray_dir'], int(k), ortho_origins)
        ok, why = margin64.is_excused(scene, o, d, int(r_near[k]), bool(r_hit[k]), int(c_near[k]), bool(c_hit[k]), near, far)
        assert ok, ('pixel %d: nearest ref=%d(hit=%s, depth=%r) cand=%d(hit=%s, depth=%r) is not an epsilon-tie'
                    % (k, r_near[k], r_hit[k], r_depth[k], c_near[k], c_hit[k], c_depth[k]))
        ties[why] = ties.get(why, 0) + 1
    n_hit = max(1, int(r_hit.sum()))
    assert len(mism) <= max(2, MAX_TIE_FRACTION * n_hit), 'too many tie pixels: %d of %d hit pixels' % (len(mism), n_hit)
    good = np.ones(c_near.shape[0], dtype=bool)
    good[mism] = False
    worst = {}
    tangent = set()
    for key, width, hit_only in (('depth', 1, False), ('image', 3, False), ('pos', 3, True), ('normal', 3, True)):
        a = _to_np(cand[key]).reshape(-1, width).astype(np.float64)
        b = _to_np(ref[key]).reshape(-1, width).astype(np.float64)
        sel = good & r_hit if hit_only else good
        if key in ('pos', 'normal') and not hit_only:
            sel = good
        finite = np.isfinite(b).all(axis=1)
        err = np.abs(a - b) - (atol + rtol * np.abs(b))
        bad = np.nonzero(sel & finite & (err > 0).any(axis=1))[0]
        # Same winner, values apart: excused only for a GRAZING SPHERE hit.  The reference's fp32 quadratic
        # (utils.py:238-278) takes t from -b - sqrt(disc) with disc -> 0 at tangency, so the hit point - and with it
        # normal and shading - loses a factor 1/sqrt(2 (r - closest)/r) of accuracy in the reference itself (its own
        # vectorised sqrt is not correctly rounded).  Tangent tie: r - closest <= EPS_TANGENT * r in float64.
        for k in bad:
            o, d = ray_for_pixel(scene, ref['ray_dir'], int(k), ortho_origins)
            kind, i = margin64.locate(scene, int(r_near[k]))
            _, margin = margin64.eval_pair(scene, o, d, int(r_near[k]))
            radius = abs(float(_to_np(scene['objects'][kind]['radius'])[i])) if kind == 'sphere' else 0.0
            assert kind == 'sphere' and r_hit[k] and 0 <= margin <= EPS_TANGENT * radius, (
                '%s: pixel %d (winner %d, %s) differs by %.3g, beyond %g + %g*|ref|, and is not a grazing sphere hit'
                % (key, k, r_near[k], kind, np.abs(a[k] - b[k]).max(), atol, rtol))
            tangent.add(int(k))
        keep = sel & finite
        keep[list(tangent)] = False
        worst[key] = float(np.abs(a[keep] - b[keep]).max()) if keep.any() else 0.0
    if tangent:
        ties['tangent_sphere'] = len(tangent)
        good[list(tangent)] = False
        assert len(tangent) + len(mism) <= max(2, MAX_TIE_FRACTION * n_hit), 'too many tie pixels'
    mism = np.concatenate((mism, np.array(sorted(tangent), dtype=mism.dtype))) if tangent else mism
    # miss pixels report primitive 0's plane hit / normal: looser, they are far-away garbage by design
    if check_ray and 'ray_dir' in cand and cand['ray_dir'] is not None:
        a, b = _to_np(cand['ray_dir']).astype(np.float64), _to_np(ref['ray_dir']).astype(np.float64)
        assert a.shape == b.shape, (a.shape, b.shape)
        assert np.abs(a - b).max() <= 5e-7, 'ray_dir differs by %.3g' % np.abs(a - b).max()
    return {'mismatch_pixels': int(len(mism)), 'ties': ties, 'hit_pixels': int(r_hit.sum()), 'worst': worst,
            'good_mask': good}


def kink_mask(scene, ref, params=None, eps=2e-5):
    """[N] bool: pixels sitting on a relu / sign kink of the shader (renderer.py:104-115): |n.L|, |V.R| or |V.n|
    is rounding noise, so relu'(.) is decided by the last bit and the gradient is not defined there.  (SCENE_BASIC
    has light 6 exactly in the plane of disk 2, so every pixel of that disk is such a tie.)  Such pixels are
    dropped from the loss on both sides, like nearest-index ties."""
    params = params or {}
    P = _to_np(ref['pos']).reshape(-1, 3).astype(np.float64)
    n = _to_np(ref['normal']).reshape(-1, 3).astype(np.float64)
    eye = _to_np(scene['camera']['eye']).astype(np.float64)[:3]
    V = eye[None, :] - P
    V /= np.maximum(np.linalg.norm(V, axis=1, keepdims=True), 1e-300)
    lp = _to_np(scene['lights']['pos']).astype(np.float64)[:, :3]
    ks_any = bool((_to_np(scene['materials']['coeffs'])[:, 1] != 0).any())
    bad = np.zeros(P.shape[0], dtype=bool)
    with np.errstate(invalid='ignore', divide='ignore'):
        if params.get('double_sided', False):
            bad |= np.abs(np.sum(V * n, axis=1)) <= eps
        for l in range(lp.shape[0]):
            L = lp[l][None, :] - P
            L /= np.maximum(np.linalg.norm(L, axis=1, keepdims=True), 1e-300)
            nl = np.sum(n * L, axis=1)
            bad |= np.abs(nl) <= eps
            if ks_any:
                R = 2 * nl[:, None] * n - L
                bad |= np.abs(np.sum(V * R, axis=1)) <= eps
    return bad


def compare_grads(cand, ref, rtol=1e-4, atol_scale=1e-5, skip=()):
    """cand/ref: {leaf name: array}.  Tolerance: 1e-4 relative to the element, plus 1e-5 of the leaf's max |grad|
    (float atomics reorder the sum, so tiny elements carry the noise of the large ones)."""
    worst = {}
    for k, exp in ref.items():
        if k in skip:
            continue
        got = _to_np(cand[k]).astype(np.float64)
        exp = _to_np(exp).astype(np.float64)
        assert got.shape == exp.shape, (k, got.shape, exp.shape)
        scale = float(np.nanmax(np.abs(exp))) if exp.size else 0.0
        err = np.abs(got - exp)
        bound = atol_scale * max(scale, 1e-30) + rtol * np.abs(exp)
        bad = err > bound
        worst[k] = float(err.max() / max(scale, 1e-30)) if exp.size else 0.0
        assert not bad.any(), '%s: %d elements off, worst %.3g (max |ref| %.3g)' % (k, int(bad.sum()), err.max(), scale)
    return worst
