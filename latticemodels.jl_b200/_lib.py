"""ctypes binding of include/lm_b200.h - the same C ABI the Julia glue ``ccall``s.

There is NO fallback: if ``lib/liblm_b200.so`` is missing or cannot be loaded every product
entry point raises (the CPU oracle under ``oracle/`` is test infrastructure and is never
imported from here).
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

LM_OK = 0
LM_C128, LM_C64 = 0, 1
METHOD_AUTO, METHOD_CHEBYSHEV, METHOD_TAYLOR, METHOD_LANCZOS, METHOD_TAYLOR_HORNER, METHOD_CHEBYSHEV_CLENSHAW = 0, 1, 2, 3, 4, 5
FIELD_LANDAU, FIELD_SYMMETRIC, FIELD_POINTFLUX_AXIAL, FIELD_POINTFLUX_SINGULAR = 1, 2, 3, 4


class ArgumentError(ValueError):
    """Mirrors Julia's ArgumentError (src/evolution.jl:152,239): raised for every non-zero
    status the library returns."""


class BackendUnavailable(RuntimeError):
    pass


_vp = C.c_void_p
_i32, _i64, _f64 = C.c_int32, C.c_int64, C.c_double
_pi32, _pi64, _pf64 = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)

# name -> argtypes ; every function returns int32 status unless listed in _SPECIAL
PROTOTYPES = {
    "lm_ctx_create": [_i32, _i32, _vp, C.POINTER(_vp)],
    "lm_ctx_destroy": [_vp],
    "lm_ctx_synchronize": [_vp],
    "lm_ctx_stream": [_vp, C.POINTER(_vp)],
    "lm_ctx_launch_count": [_vp, _pi64],
    "lm_timer_start": [_vp],
    "lm_timer_stop": [_vp, _pf64],
    "lm_comm_unique_id": [_vp],
    "lm_ctx_comm_init": [_vp, _vp, _i32, _i32],
    "lm_ctx_peer_handle": [_vp, _i64, _vp],
    "lm_ctx_peer_attach": [_vp, _vp],
    "lm_shard_range": [_i64, _i32, _i32, _pi64, _pi64],
    "lm_ham_create_csc": [_vp, _i64, _i32, _vp, _vp, _vp, _i32, C.POINTER(_vp)],
    "lm_ham_update_values": [_vp, _vp],
    "lm_ham_update_values_async": [_vp, _vp],
    "lm_ham_update_values_bcast": [_vp, _vp, C.c_int32],
    "lm_ham_create_bonds": [_vp, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, C.POINTER(_vp)],
    "lm_ham_set_fields": [_vp, _i32, _vp, _vp],
    "lm_ham_set_field_params": [_vp, _vp],
    "lm_ham_set_site_coords": [_vp, _vp],
    "lm_ham_set_row_block": [_vp, _i32],
    "lm_ham_set_lattice_dims": [_vp, _i32, _i32],
    "lm_ham_dims": [_vp, _pi64, _pi32, _pi64, _pi32],
    "lm_ham_get_csc": [_vp, _vp, _vp, _vp],
    "lm_ham_spectral_bounds": [_vp, _pf64, _pf64],
    "lm_ham_refine_bounds": [_vp, _i32, _f64],
    "lm_ham_destroy": [_vp],
    "lm_state_create_psi": [_vp, _i64, _i64, _vp, _vp, C.POINTER(_vp)],
    "lm_state_create_psi_synth": [_vp, _i64, _i64, _i64, C.c_uint64, C.POINTER(_vp)],
    "lm_state_column_norms2": [_vp, _vp],
    "lm_state_create_dense": [_vp, _i64, _vp, C.POINTER(_vp)],
    "lm_state_copy": [_vp, C.POINTER(_vp)],
    "lm_state_set_replicated": [_vp, _i32],
    "lm_state_dims": [_vp, _pi64, _pi64, _pi32],
    "lm_state_download_psi": [_vp, _vp],
    "lm_state_download_dense": [_vp, _vp],
    "lm_state_destroy": [_vp],
    "lm_eigs_lowest": [_vp, _i32, _f64, _i32, _i32, _pf64, _pf64, C.POINTER(_vp), _pi32],
    "lm_step": [_vp, _vp, _f64, _f64, _i32, _pi32],
    "lm_spmm_state": [_vp, _vp, _vp],
    "lm_spmm": [_vp, _vp, _vp, _i64, _i64],
    "lm_local_density": [_vp, _i32, _vp],
    "lm_currents_npairs": [_vp, _pi64],
    "lm_currents_pairs": [_vp, _vp, _vp],
    "lm_observables": [_vp, _vp, _vp, _vp],
    "lm_observables_async": [_vp, _vp, _i32, _i32],
    "lm_frame_wait": [_vp, _i32, _vp, _vp],
    "lm_currents_fromto": [_vp, _vp, _vp, _vp, _i32, _pf64],
    "lm_currents_from": [_vp, _vp, _vp, _i32, _vp],
    "lm_bond_currents": [_vp, _vp, _i64, _vp, _vp, _vp],
    "lm_local_expect": [_vp, _i32, _vp, _vp],
    "lm_operator_currents": [_vp, _vp, _vp, _vp],
}
_SPECIAL = {"lm_version": ([], _i32), "lm_last_error": ([], C.c_char_p)}

_lib = None


def library_path():
    return _build.LIB


def load():
    """Load liblm_b200.so (once).  Raises BackendUnavailable if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise BackendUnavailable(
            "liblm_b200.so not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`."
            " There is no CPU fallback for the evolution backend." % path)
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, args in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _i32
    for name, (args, res) in _SPECIAL.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    _lib = lib
    return lib


def check(status):
    if status != LM_OK:
        msg = load().lm_last_error()
        raise ArgumentError((msg.decode() if msg else "lm_b200 error") + " [status %d]" % status)


def ptr(a):
    """void* of a numpy array (None -> NULL)."""
    return None if a is None else a.ctypes.data_as(_vp)


def cdtype(precision):
    return np.complex128 if precision == LM_C128 else np.complex64
