"""Device context: one process drives one GPU (include/lm_b200.h, lm_ctx_*)."""
from __future__ import annotations

import ctypes as C
import os

from . import _lib

_PREC = {"c128": _lib.LM_C128, "complex128": _lib.LM_C128, "c64": _lib.LM_C64, "complex64": _lib.LM_C64,
         _lib.LM_C128: _lib.LM_C128, _lib.LM_C64: _lib.LM_C64}


class Context:
    def __init__(self, device=None, precision="c128", stream=None):
        lib = _lib.load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = int(device)
        self.precision = _PREC[precision]
        self.rank, self.nranks = 0, 1
        h = C.c_void_p()
        _lib.check(lib.lm_ctx_create(self.device, self.precision, C.c_void_p(stream) if stream else None, C.byref(h)))
        self.handle = h

    def synchronize(self):
        _lib.check(_lib.load().lm_ctx_synchronize(self.handle))

    def launch_count(self):
        n = C.c_int64()
        _lib.check(_lib.load().lm_ctx_launch_count(self.handle, C.byref(n)))
        return n.value

    def timer_start(self):
        _lib.check(_lib.load().lm_timer_start(self.handle))

    def timer_stop(self):
        ms = C.c_double()
        _lib.check(_lib.load().lm_timer_stop(self.handle, C.byref(ms)))
        return ms.value

    def stream(self):
        s = C.c_void_p()
        _lib.check(_lib.load().lm_ctx_stream(self.handle, C.byref(s)))
        return s.value

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _lib.check(_lib.load().lm_ctx_comm_init(self.handle, buf, rank, nranks))
        self.rank, self.nranks = rank, nranks

    def peer_handle(self, slot_doubles):
        buf = C.create_string_buffer(64)
        _lib.check(_lib.load().lm_ctx_peer_handle(self.handle, int(slot_doubles), buf))
        return buf.raw

    def peer_attach(self, all_handles: bytes):
        buf = C.create_string_buffer(bytes(all_handles), 64 * self.nranks)
        _lib.check(_lib.load().lm_ctx_peer_attach(self.handle, buf))

    def shard_range(self, M):
        return shard_range(M, self.rank, self.nranks)

    def close(self):
        if self.handle:
            _lib.load().lm_ctx_destroy(self.handle)
            self.handle = None


def shard_range(M, rank, nranks):
    """Contiguous column range of a rank; pure arithmetic, mirrors lm_shard_range so that the
    host logic is testable without the library."""
    return (M * rank) // nranks, (M * (rank + 1)) // nranks


def unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _lib.check(_lib.load().lm_comm_unique_id(buf))
    return buf.raw


_default = {}


def default_context(precision="c128"):
    key = _PREC[precision]
    if key not in _default:
        _default[key] = Context(precision=precision)
    return _default[key]
