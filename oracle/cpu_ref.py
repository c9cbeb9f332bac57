"""ctypes wrapper of oracle/cpu_ref.c (checker / CPU baseline only - see oracle/__init__.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libcpu_ref.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "cpu_ref.c")):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = C.CDLL(LIB)
        _lib.lm_ref_krylov_block_step.restype = C.c_int64
        _lib.lm_ref_krylov_block_step.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                                  C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_int, C.c_int]
        _lib.lm_ref_localdensity.argtypes = [C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.lm_ref_max_threads.restype = C.c_int
    return _lib


def max_threads():
    return load().lm_ref_max_threads()


class CsrHam:
    def __init__(self, H):
        m = sp.csr_matrix(H)
        m.sort_indices()
        self.n = m.shape[0]
        self.rowptr = np.ascontiguousarray(m.indptr, np.int64)
        self.col = np.ascontiguousarray(m.indices, np.int32)
        self.val = np.ascontiguousarray(m.data, np.complex128)


def krylov_block_step(ham: CsrHam, psi, dt, krylovdim=30, tol=1e-12, maxiter=100, nthreads=0):
    """In-place: psi (n x M, Fortran order complex128) <- exp(-i H dt) psi, one Lanczos run per
    column (KrylovKitExp semantics).  Returns the number of matvecs."""
    assert psi.flags.f_contiguous and psi.dtype == np.complex128
    k = load().lm_ref_krylov_block_step(ham.n, ham.rowptr.ctypes.data, ham.col.ctypes.data, ham.val.ctypes.data,
                                        psi.shape[1], psi.ctypes.data, dt, krylovdim, tol, maxiter, nthreads)
    if k < 0:
        raise ValueError("`exponentiate` did not converge")
    return int(k)


def localdensity(psi, w=None, n_int=1):
    n, M = psi.shape
    rho = np.zeros(n // n_int)
    wp = None if w is None else np.ascontiguousarray(w, float)
    load().lm_ref_localdensity(n, M, n_int, psi.ctypes.data, None if wp is None else wp.ctypes.data, rho.ctypes.data)
    return rho
