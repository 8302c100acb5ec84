import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def golden_cases():
    """fixtures of render() (make_golden.py)"""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR)
                  if f.endswith('.npz') and not f.startswith(('ar_', 'np_', 'b200_', 'pl_', 'ndc_')))


def along_ray_cases():
    """fixtures of render_splats_along_ray() (make_golden_along_ray.py)"""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith('.npz') and f.startswith('ar_'))


def numpy_twin_cases():
    """fixtures of the reference's numpy twin renderer (make_golden_numpy.py)"""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith('.npz') and f.startswith('np_'))


def ndc_cases():
    """fixtures of render_splats_NDC() (make_golden_ndc.py)"""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith('.npz') and f.startswith('ndc_'))
