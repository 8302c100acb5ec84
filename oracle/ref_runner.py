"""Runs the UNMODIFIED reference renderer from oracle/_ref (see make_ref.py) on the host CPU.  Test / benchmark
infrastructure only.  The reference picks its device when `diffrend.torch.utils` is imported
(diffrend/torch/utils.py:7-15: CUDA whenever torch sees a GPU), so the process that imports it must hide the GPUs
first - `bench.py --impl reference` sets CUDA_VISIBLE_DEVICES="" before importing torch.
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, '_ref')


def available():
    return os.path.exists(os.path.join(REF, 'diffrend', 'torch', 'renderer.py'))


def load():
    """-> (render, module utils) of the reference; raises if oracle/_ref is absent or would run on a GPU"""
    if not available():
        raise RuntimeError('oracle/_ref is not populated: run `python oracle/make_ref.py` where /root/reference exists')
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):          # utils.py prints 'CUDA support ...' at import
        import diffrend.torch.utils as ref_utils
        from diffrend.torch.renderer import render
    if ref_utils.CUDA:
        raise RuntimeError('the reference selected CUDA tensors; import it in a process with CUDA_VISIBLE_DEVICES=""')
    return render, ref_utils


def scene_for_reference(scene):
    """Our synthetic scene dicts (plain python lists / CPU float tensors) -> the tensors the reference expects:
    float CPU tensors for the arrays, python numbers for the camera scalars (what make_torch_var produces)."""
    import numpy as np
    import torch

    def conv(v, integer=False):
        if isinstance(v, torch.Tensor):
            return v.detach().clone()
        a = np.asarray(v)
        return torch.tensor(a, dtype=torch.int64 if integer else torch.float32)

    out = {'camera': dict(scene['camera'])}
    for k in ('eye', 'at', 'up'):
        out['camera'][k] = conv(scene['camera'][k])
    out['lights'] = {k: conv(v, integer=(k == 'color_idx')) for k, v in scene['lights'].items()}
    out['colors'] = conv(scene['colors'])
    out['materials'] = {k: conv(v) for k, v in scene['materials'].items()}
    out['objects'] = {kind: {k: conv(v, integer=(k == 'material_idx')) for k, v in prim.items()}
                      for kind, prim in scene['objects'].items()}
    if 'tonemap' in scene:
        out['tonemap'] = {'type': scene['tonemap']['type'], 'gamma': conv(scene['tonemap']['gamma']).reshape(-1)}
    return out
