"""The bounding-sphere filter of the constant-bank intersection path (csrc/surf_isect_const.cu: k_sphere_records,
sphere_margin_max) restated in numpy float32, against the oracle's fp32 ray-disk test (reference utils.py:311-326): a
filter is only allowed to be CONSERVATIVE - every (ray, disk) pair the reference counts as a hit must pass it - on rays
aimed at the rim of the disks, where the slack decides.  Also the dense kernel's form (c0 from the staged plane record)."""
import numpy as np
import torch

from oracle import torch_oracle

U = 2.0 ** -24
F = np.float32


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def _sphere_records(eye, c, r):
    """k_sphere_records: oc / sqrt(|oc|^2 - rs^2 - slack) in double, rounded to float"""
    eye, c, r = eye.astype(np.float64), c.astype(np.float64), np.abs(r.astype(np.float64))
    oc = eye[None, :] - c
    oc2 = (oc ** 2).sum(1)
    scale = np.sqrt((eye ** 2).sum()) + np.sqrt((c ** 2).sum(1)) + np.sqrt(oc2) + r
    rs = r + 2e-6 * scale
    den = oc2 - rs * rs * (1.0 + 1e-6) - 16.0 * U * oc2
    assert (den > 0).all()
    return (oc / np.sqrt(den)[:, None]).astype(F), oc.astype(F), (-(rs * rs * (1.0 + 1e-6))).astype(F)


def _rim_rays(rng, eye, c, n, r, per_disk):
    """unit fp32 directions through in-plane points at radius r (1 + delta) of every disk"""
    m = c.shape[0]
    a = np.cross(n, rng.randn(m, 3))
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = np.cross(n, a)
    deltas = np.array([0.0, 1e-7, -1e-7, 1e-6, -1e-6, 1e-5, -1e-5, 1e-4, -1e-4, 1e-3, -1e-3, -0.5, 0.3, 1.0])
    phi = rng.rand(m, per_disk) * 2 * np.pi
    rho = r[:, None] * (1.0 + deltas[rng.randint(0, len(deltas), (m, per_disk))])
    pts = c[:, None, :] + rho[..., None] * (np.cos(phi)[..., None] * a[:, None, :] + np.sin(phi)[..., None] * b[:, None, :])
    v = (pts.reshape(-1, 3) - eye[None, :]).astype(F)
    return (v / np.sqrt((v * v).sum(1, dtype=F))[:, None]).astype(F)


def _scene(rng, m, dist, radius):
    eye = np.array([0.3, -0.2, dist], dtype=np.float64)
    c = rng.randn(m, 3) * 0.5
    n = rng.randn(m, 3)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    r = radius * (0.5 + rng.rand(m))
    return eye, c, n, r


def _oracle_hits(eye, c, n, r, d):
    prim = {'pos': torch.tensor(c, dtype=torch.float32), 'normal': torch.tensor(n, dtype=torch.float32), 'radius': torch.tensor(r, dtype=torch.float32)}
    _, t, _ = torch_oracle.hit_disk(torch.tensor(eye[None, :], dtype=torch.float32), torch.tensor(d.T.copy()), prim, want_normals=False)
    return (t != torch_oracle.MISS_SENTINEL).numpy()


def _check(seed, dist, radius):
    rng = np.random.RandomState(seed)
    n_hits = n_pass = 0
    for _ in range(12):
        eye, c, n, r = _scene(rng, 60, dist, radius)
        c32, n32, r32 = c.astype(F), n.astype(F), r.astype(F)
        d = _rim_rays(rng, eye.astype(F).astype(np.float64), c32.astype(np.float64), n32.astype(np.float64) / np.linalg.norm(n32, axis=1, keepdims=True),
                      r32.astype(np.float64), 48)
        hits = _oracle_hits(eye.astype(F), c32, n32, r32, d)                   # [M, N]
        rec, oc, neg_rs2 = _sphere_records(eye.astype(F), c32, r32)
        # k_filter_const<., 1>: s = fma(z, dz, fma(y, dy, x * dx)), pass iff |s| >= 1
        s = (rec[:, None, 0] * d[None, :, 0]).astype(F)
        s = _fma(np.broadcast_to(rec[:, None, 1], s.shape), np.broadcast_to(d[None, :, 1], s.shape), s)
        s = _fma(np.broadcast_to(rec[:, None, 2], s.shape), np.broadcast_to(d[None, :, 2], s.shape), s)
        passed = np.abs(s) >= F(1.0)
        assert not (hits & ~passed).any(), 'sphere filter (normalised form) rejects %d reference hits' % int((hits & ~passed).sum())
        # chunk_disks_dense<., true>: c0 = fma(|oc|^2, 1 - 2^-19, -rs^2) from the float record, pass iff s^2 - c0 >= 0
        oc2 = _fma(oc[:, 2], oc[:, 2], _fma(oc[:, 1], oc[:, 1], (oc[:, 0] * oc[:, 0]).astype(F)))
        nc0 = -_fma(oc2, np.full_like(oc2, F(1.0) - F(2.0 ** -19)), neg_rs2)
        s = (oc[:, None, 0] * d[None, :, 0]).astype(F)
        s = _fma(np.broadcast_to(oc[:, None, 1], s.shape), np.broadcast_to(d[None, :, 1], s.shape), s)
        s = _fma(np.broadcast_to(oc[:, None, 2], s.shape), np.broadcast_to(d[None, :, 2], s.shape), s)
        e = _fma(s, s, np.broadcast_to(nc0[:, None], s.shape))
        passed2 = e >= 0
        assert not (hits & ~passed2).any(), 'sphere filter (dense form) rejects %d reference hits' % int((hits & ~passed2).sum())
        n_hits += int(hits.sum())
        n_pass += int(passed.sum())
    assert n_hits > 5000                       # the rim rays do produce hits
    assert n_pass < 40 * n_hits                # and the filter still filters (random orientations: a few times the hits)
    return n_hits, n_pass


def test_sphere_filter_never_rejects_a_reference_hit_config_e_like():
    print(_check(1, 5.0, 0.005))


def test_sphere_filter_never_rejects_a_reference_hit_far_and_tiny():
    print(_check(2, 40.0, 0.002))


def test_sphere_filter_never_rejects_a_reference_hit_near_and_large():
    print(_check(3, 1.5, 0.2))
