"""Scene builder shared by make_golden_ndc.py and the tests (no reference dependency)."""
import numpy as np
import torch

from surf_renderer_b200 import scenes as synth


def ndc_scene(seed, W, H, homogeneous=False, mats=3):
    """one splat per pixel in normalised device coordinates: x, y on the pixel grid (jittered), z in the part of the
    [-1, 1] depth range that unprojects to a few units in front of the camera"""
    g = torch.Generator().manual_seed(seed)
    base = synth.random_mixed_scene(seed, width=W, height=H, homogeneous=True, n_mat=mats)
    n = W * H
    near, far = 0.5, 20.0
    base['camera'].update({'fovy': float(np.deg2rad(50.)), 'near': near, 'far': far, 'eye': torch.tensor([0.4, 0.9, 6.0, 1.0])})
    yy, xx = torch.meshgrid(torch.linspace(1, -1, H), torch.linspace(-1, 1, W), indexing='ij')
    depth = 2.5 + 3.0 * torch.rand(n, generator=g)                          # camera-space distance along -Z
    # z_ndc of the right-handed [-1, 1] perspective (ops.py:37-47): (m22 d + m23) / d
    m22, m23 = (near + far) / (far - near), -2 * near * far / (far - near)
    z = (m22 * depth + m23) / depth
    pos = torch.stack((xx.reshape(-1) + 0.02 * torch.randn(n, generator=g), yy.reshape(-1) + 0.02 * torch.randn(n, generator=g), z), 1)
    if homogeneous:         # [N,4] with a w that is not 1: the reference divides by the unprojected w
        w = 0.7 + 0.6 * torch.rand(n, 1, generator=g)
        pos = torch.cat((pos * w, w), 1)
    nrm = torch.randn(n, 3, generator=g)
    nrm[:, 2] = nrm[:, 2].abs() + 0.4
    nrm = nrm / nrm.norm(dim=1, keepdim=True) * (0.8 + 0.4 * torch.rand(n, 1, generator=g))    # not unit length
    if homogeneous:
        nrm[::5] *= -1      # some back-facing splats for double_sided
    # the view vector is NOT normalised on this path (renderer.py:437): |V|^shininess explodes for the usual exponents
    base['materials']['coeffs'] = torch.cat((0.3 + 0.6 * torch.rand(mats, 2, generator=g), 1.0 + 2.5 * torch.rand(mats, 1, generator=g)), 1)
    base['objects'] = {'disk': {'pos': pos, 'normal': nrm, 'material_idx': torch.randint(0, mats, (n,), generator=g)}}
    return base
