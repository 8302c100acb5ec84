"""float64 margin oracle (TEST INFRASTRUCTURE ONLY) - classifies nearest-index disagreements as epsilon-ties.

SURVEY A.7: two fp32 evaluations of the same scene may legitimately pick different winners at
  * depth ties  - two valid hits whose distances differ by fp32 noise,
  * rim / edge / tangent ties - the hit point sits on the primitive's boundary to within fp32 noise,
  * range ties  - t sits on near or far.
For a (pixel, primitive) pair this module evaluates, in float64 numpy, the ray distance and the signed
boundary margin (positive = inside) following the reference's geometry (diffrend/torch/utils.py:238-366).
"""
from __future__ import annotations

import numpy as np

EPS_DEPTH = 1e-5   # relative |t_a - t_b| <= EPS_DEPTH * t
EPS_EDGE = 4e-6    # relative boundary margin <= EPS_EDGE * t


def _np(x):
    return x.detach().cpu().numpy().astype(np.float64) if hasattr(x, 'detach') else np.asarray(x, dtype=np.float64)


def _unit(n):
    return n / np.sqrt(np.sum(n * n + 1e-10, axis=-1, keepdims=True))


def locate(scene, idx):
    """global primitive index -> (kind, local index) following dict insertion order (utils.py:486)."""
    first = 0
    for kind, prim in scene['objects'].items():
        cnt = int(prim['material_idx'].shape[0])
        if idx < first + cnt:
            return kind, idx - first
        first += cnt
    raise IndexError(idx)


def eval_pair(scene, origin, direction, idx):
    """(t, margin) in float64 for primitive `idx` and one ray.  margin = +inf for planes."""
    kind, i = locate(scene, int(idx))
    prim = scene['objects'][kind]
    o, d = origin, direction
    if kind == 'sphere':
        c = _np(prim['pos'])[i, :3]
        r = float(_np(prim['radius'])[i])
        oc = o - c
        a = d @ d
        b = 2 * (oc @ d)
        cc = oc @ oc - r * r
        disc = b * b - 4 * a * cc
        closest = np.linalg.norm(oc - (oc @ d) / a * d)
        margin = r - closest
        if disc < 0:
            return np.inf, margin
        t1, t2 = (-b - np.sqrt(disc)) / (2 * a), (-b + np.sqrt(disc)) / (2 * a)
        ts = [t for t in (t1, t2) if t >= 0]
        return (min(ts) if ts else np.inf), margin
    if kind == 'triangle':
        face = _np(prim['face'])[i, :, :3]
        p = face[0]
    else:
        p = _np(prim['pos'])[i, :3]
    n = _unit(_np(prim['normal'])[i, :3])
    denom = n @ d
    if denom == 0:
        return np.inf, -np.inf
    t = (p @ n - n @ o) / denom
    P = o + t * d
    if kind == 'plane':
        return t, np.inf
    if kind == 'disk':
        r = float(_np(prim['radius'])[i])
        return t, abs(r) - np.linalg.norm(P - p)
    margins = []
    for a_, b_ in ((0, 1), (1, 2), (2, 0)):
        e = face[b_] - face[a_]
        margins.append(np.cross(e, P - face[a_]) @ n / max(np.linalg.norm(e), 1e-300))
    return t, min(margins)


def is_excused(scene, origin, direction, idx_ref, hit_ref, idx_cand, hit_cand, near, far):
    """True when the disagreement (idx_ref, hit_ref) vs (idx_cand, hit_cand) at this ray is an epsilon-tie."""
    evals = []
    for idx, hit in ((idx_ref, hit_ref), (idx_cand, hit_cand)):
        if hit:
            evals.append(eval_pair(scene, origin, direction, idx))
    if not evals:
        return True, 'both-miss'
    for t, margin in evals:
        scale = max(abs(t), 1e-30) if np.isfinite(t) else 1.0
        if np.isfinite(margin) and abs(margin) <= EPS_EDGE * scale:
            return True, 'rim'
        if np.isfinite(t) and (abs(t - near) <= EPS_DEPTH * scale or abs(t - far) <= EPS_DEPTH * scale):
            return True, 'range'
    if len(evals) == 2:
        (ta, ma), (tb, mb) = evals
        if np.isfinite(ta) and np.isfinite(tb) and ma >= -EPS_EDGE * abs(ta) and mb >= -EPS_EDGE * abs(tb) \
                and abs(ta - tb) <= EPS_DEPTH * max(abs(ta), abs(tb)):
            return True, 'depth'
    return False, 'real'
