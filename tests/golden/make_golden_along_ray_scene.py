"""Scene builder shared by make_golden_along_ray.py and the tests (no reference dependency)."""
import numpy as np
import torch

from surf_renderer_b200 import scenes as synth


def along_ray_scene(seed, W, H, with_vis=False, pos3=False, mats=4, smooth_z=False):
    g = torch.Generator().manual_seed(seed)
    base = synth.random_mixed_scene(seed, width=W, height=H, homogeneous=True, n_mat=mats)
    n = W * H
    z = -(2.5 + 2.0 * torch.rand(n, generator=g))
    if smooth_z:      # a smooth depth field, so that estimated normals / subpixel planes are well conditioned
        yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing='ij')
        z = -(3.5 + 0.6 * torch.sin(2.2 * xx + 0.4) * torch.cos(1.7 * yy) + 0.25 * xx * yy).reshape(-1)
    else:
        z[::17] = 0.3                  # a few splats behind the camera plane: relu clamps them to Z = 0
    nrm = torch.randn(n, 3, generator=g)
    nrm[:, 2] = nrm[:, 2].abs() + 0.4
    nrm = nrm / nrm.norm(dim=1, keepdim=True) * (0.8 + 0.4 * torch.rand(n, 1, generator=g))    # not unit length
    disk = {'pos': torch.stack((torch.zeros(n), torch.zeros(n), z), 1) if pos3 else z, 'normal': nrm,
            'material_idx': torch.randint(0, mats, (n,), generator=g)}
    if with_vis:
        disk['light_vis'] = (torch.rand(base['lights']['pos'].shape[0], n, generator=g) > 0.3).float()
    base['objects'] = {'disk': disk}
    base['camera']['fovy'] = float(np.deg2rad(40.))
    base['camera']['eye'] = torch.tensor([0.4, 0.9, 6.0, 1.0])
    return base


