"""ctypes binding of ``oracle/csr_oracle.c``.  TEST INFRASTRUCTURE ONLY."""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libcsr_oracle.so')
_lib = None


def build(force=False):
    """Compile the C oracle with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, 'csr_oracle.c')
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src))
    if force or stale:
        subprocess.check_call(['make', '-C', _HERE, '-B', 'libcsr_oracle.so'],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i64, p, dbl, i32 = (ctypes.c_int64, ctypes.c_void_p, ctypes.c_double,
                            ctypes.c_int)
        L.oracle_csr_matvecs.argtypes = [i64, i64, p, p, p, p, p, i32]
        L.oracle_csr_matvecs.restype = None
        L.oracle_remap_fused.argtypes = [i64, i64, p, p, p, p, p, p, i32, dbl,
                                         p, p, i32]
        L.oracle_remap_fused.restype = None
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _csr_parts(matrix):
    return (np.ascontiguousarray(matrix.indptr, dtype=np.int32),
            np.ascontiguousarray(matrix.indices, dtype=np.int32),
            np.ascontiguousarray(matrix.data, dtype=np.float64))


def csr_matvecs(matrix, X, threads=1):
    ap, aj, ax = _csr_parts(matrix)
    X = np.ascontiguousarray(X, dtype=np.float64)
    Y = np.empty((matrix.shape[0], X.shape[1]), dtype=np.float64)
    lib().oracle_csr_matvecs(matrix.shape[0], X.shape[1], _ptr(ap), _ptr(aj),
                             _ptr(ax), _ptr(X), _ptr(Y), int(threads))
    return Y


def remap_fused(matrix, frac_b, X, mode, threshold=0.0, valid=None,
                want_keep=False, threads=1):
    """mode: 0 raw, 1 frac_b branch, 2 masked branch (see csr_oracle.c)."""
    ap, aj, ax = _csr_parts(matrix)
    X = np.ascontiguousarray(X, dtype=np.float64)
    fb = None if frac_b is None else np.ascontiguousarray(frac_b, np.float64)
    vb = None if valid is None else np.ascontiguousarray(valid, np.uint8)
    Y = np.empty((matrix.shape[0], X.shape[1]), dtype=np.float64)
    keep = np.empty(Y.shape, dtype=np.uint8) if want_keep else None
    lib().oracle_remap_fused(matrix.shape[0], X.shape[1], _ptr(ap), _ptr(aj),
                             _ptr(ax), _ptr(fb), _ptr(X), _ptr(vb), int(mode),
                             float(threshold), _ptr(Y), _ptr(keep),
                             int(threads))
    return (Y, keep.astype(bool)) if want_keep else Y
