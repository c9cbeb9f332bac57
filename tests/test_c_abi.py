"""The C ABI seen from C: include/lm_b200.h compiled by gcc as plain C, a C program linked against
liblm_b200.so (tests/c_abi_smoke.c), and the hand-written ctypes prototype table checked against the
header's own declarations."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lm_b200.h")


def _compile(tmp_path, link=True):
    import lm_b200
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    lib = lm_b200.library_path()
    if link and not os.path.exists(lib):
        pytest.skip("liblm_b200.so not built")
    exe = str(tmp_path / "c_abi_smoke")
    cmd = [gcc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_abi_smoke.c")]
    cmd += (["-o", exe, "-L", os.path.dirname(lib), "-llm_b200", "-Wl,-rpath," + os.path.dirname(lib), "-lm"] if link else ["-c", "-o", exe + ".o"])
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    return exe


def test_header_compiles_as_c_and_links_against_the_library(tmp_path):
    """gcc -std=c11 -Wall -Wextra -Werror -pedantic on a C program that calls through the header and
    links every symbol it uses out of liblm_b200.so (no device needed to build)."""
    _compile(tmp_path, link=True)


def _header_prototypes():
    text = re.sub(r"/\*.*?\*/", " ", open(HEADER).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int32_t|const char\*)\s+(lm_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = [a.strip() for a in m.group(3).replace("\n", " ").split(",")]
        protos[m.group(2)] = (m.group(1), [] if args == ["void"] else args)
    return protos


def _ctype_class(arg):
    """'ptr' / 'i32' / 'i64' / 'u64' / 'f64' of a C parameter declaration."""
    if "*" in arg:
        return "ptr"
    for key, cls in (("uint64_t", "u64"), ("int64_t", "i64"), ("int32_t", "i32"), ("double", "f64")):
        if re.search(r"\b%s\b" % key, arg):
            return cls
    raise AssertionError("unclassified parameter %r" % arg)


def test_ctypes_prototype_table_matches_the_header():
    """Every function the header declares is in latticemodels.jl_b200/_lib.py PROTOTYPES with the same
    number of parameters and the same scalar / pointer classes (the table is what the Python host
    mirror - and by analogy the Julia ccall signatures - relies on)."""
    import ctypes as C
    from importlib import import_module
    import lm_b200  # noqa: F401
    _lib = import_module("lm_b200._lib")
    protos = _header_prototypes()
    assert len(protos) >= 50
    table = dict(_lib.PROTOTYPES)
    table.update({k: v[0] for k, v in _lib._SPECIAL.items()})
    assert set(protos) == set(table), (sorted(set(protos) - set(table)), sorted(set(table) - set(protos)))

    def cls(t):
        if t in (C.c_void_p, C.c_char_p) or isinstance(t, type(C.POINTER(C.c_int))):
            return "ptr"
        return {C.c_int32: "i32", C.c_int64: "i64", C.c_uint64: "u64", C.c_double: "f64"}[t]
    for name, (ret, args) in protos.items():
        got = [cls(t) for t in table[name]]
        want = [_ctype_class(a) for a in args]
        assert got == want, (name, got, want)


@pytest.mark.gpu
def test_c_program_runs_the_hot_path_through_the_c_abi(tmp_path):
    exe = _compile(tmp_path, link=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "c_abi_smoke OK" in res.stdout, res.stdout + res.stderr
