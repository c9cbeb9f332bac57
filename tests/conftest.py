import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on a B200)")
    config.addinivalue_line("markers", "slow: larger CPU-only case")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _emulated_library():
    """LM_EMUL_LIB=<path>: run the gpu-marked parity tests against the CPU build of the whole library
    (tests/cpu_emul/build_emul_lib.py; every kernel executed by OS threads).  Test infrastructure:
    the swap happens HERE, in the test session - the product loader knows nothing about it."""
    path = os.environ.get("LM_EMUL_LIB")
    if not path:
        return False
    import ctypes as C
    from importlib import import_module
    import lm_b200  # noqa: F401  (registers the package)
    _lib = import_module("lm_b200._lib")
    if _lib._lib is None:
        lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
        for name, args in _lib.PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = args, C.c_int32
        for name, (args, res) in _lib._SPECIAL.items():
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = args, res
        _lib._lib = lib
    return True


def pytest_collection_modifyitems(config, items):
    if _have_gpu() or _emulated_library():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
