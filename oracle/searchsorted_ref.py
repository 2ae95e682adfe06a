"""Loaders for the two CPU searchsorted checkers -- TEST INFRASTRUCTURE ONLY.

  c_oracle(a, v, side)            the plain-C restatement (oracle/_build/libsearchsorted_oracle.so)
  reference_searchsorted()        the reference's OWN C++ extension compiled by oracle/Makefile into
                                  oracle/_ref/ (build container only); returns a callable with the
                                  torchsearchsorted.searchsorted signature
"""
import ctypes as C
import importlib.util
import os
import subprocess

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
C_LIB = os.path.join(HERE, '_build', 'libsearchsorted_oracle.so')
REF_LIB = os.path.join(HERE, '_ref', 'searchsorted_ref_cpu.so')


def build_c_oracle():
    if not os.path.isfile(C_LIB) or os.path.getmtime(C_LIB) < os.path.getmtime(os.path.join(HERE, 'searchsorted_oracle.c')):
        subprocess.run(['make', '-C', HERE, '_build/libsearchsorted_oracle.so'], check=True, capture_output=True)
    return C_LIB


_c = None


def c_oracle(a: np.ndarray, v: np.ndarray, side: str = 'left') -> np.ndarray:
    global _c
    if _c is None:
        _c = C.CDLL(build_c_oracle())
        _c.ss_oracle.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int]
    a = np.ascontiguousarray(a, dtype=np.float32)
    v = np.ascontiguousarray(v, dtype=np.float32)
    res = np.empty((max(a.shape[0], v.shape[0]), v.shape[1]), dtype=np.int64)
    _c.ss_oracle(a.ctypes.data, a.shape[0], a.shape[1], v.ctypes.data, v.shape[0], v.shape[1], res.ctypes.data,
                 1 if side == 'left' else 0)
    return res


def reference_available() -> bool:
    return os.path.isfile(REF_LIB)


def reference_searchsorted():
    if not reference_available():
        raise RuntimeError('oracle/_ref/searchsorted_ref_cpu.so not built (needs /root/reference; run make -C oracle)')
    spec = importlib.util.spec_from_file_location('searchsorted_ref_cpu', REF_LIB)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    def searchsorted(a, v, out=None, side='left'):
        if out is None:
            out = torch.empty((max(a.shape[0], v.shape[0]), v.shape[1]), dtype=torch.long)
        mod.searchsorted_cpu_wrapper(a.contiguous(), v.contiguous(), out, 1 if side == 'left' else 0)
        return out

    return searchsorted
