"""Loader of libsurf_b200.so - the only compute path of this package.  There is no CPU / eager fallback:
if the library is missing or a call fails, the caller gets an exception."""
from __future__ import annotations

import ctypes as C
import os

from . import _abi

_PKG = os.path.dirname(os.path.abspath(__file__))
# SURF_B200_LIB points the loader at another build of the library (A/B runs of kernel variants); default = in-tree
LIB_PATH = os.environ.get('SURF_B200_LIB') or os.path.join(_PKG, 'libsurf_b200.so')
_lib = None


class SurfLibraryError(RuntimeError):
    pass


def lib():
    """ctypes handle with every symbol of include/surf_b200.h bound (built in-tree by build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SurfLibraryError(
                'libsurf_b200.so is not built (%s). Run `python -m surf_renderer_b200.build`; '
                'this package has no fallback path.' % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        _abi.bind(handle)
        if handle.surf_abi_version() != _abi.SURF_ABI_VERSION:
            raise SurfLibraryError('ABI version mismatch: library %d, python %d'
                                   % (handle.surf_abi_version(), _abi.SURF_ABI_VERSION))
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().surf_last_error().decode('utf-8', 'replace')
        if rc == -2:
            raise NotImplementedError(msg)
        if rc == -1:
            raise ValueError(msg)
        raise RuntimeError('libsurf_b200: %s (status %d)' % (msg, rc))
