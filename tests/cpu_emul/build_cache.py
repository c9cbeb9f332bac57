"""Content-addressed cache of the g++ builds of the CPU execution harness (tests/cpu_emul/).

The fully unrolled stencil kernels take minutes to compile even at -O0; the emulated programs only change when the kernel
sources, the shim or the compiler flags change.  `cached(...)` keys a build product on the CONTENTS of its inputs (plus the
g++ version and the command line) and re-uses it from ``$LM_TEST_CACHE`` (default ``/tmp/lm_b200_test_cache``) - a stale
product can never be picked up, an unchanged one is not rebuilt.  TEST INFRASTRUCTURE ONLY.
"""
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "latticemodels.jl_b200", "csrc")


def cache_dir():
    d = os.environ.get("LM_TEST_CACHE") or "/tmp/lm_b200_test_cache"
    os.makedirs(d, exist_ok=True)
    return d


def _gxx_version():
    gxx = shutil.which("g++")
    return subprocess.run([gxx, "--version"], capture_output=True, text=True).stdout.splitlines()[0] if gxx else "none"


def source_files():
    """Everything an emulated build can include: the product's csrc/, the shim, the C header, the harness sources."""
    out = []
    for d in (CSRC, os.path.join(HERE, "shim"), HERE, os.path.join(ROOT, "include")):
        for name in sorted(os.listdir(d)):
            p = os.path.join(d, name)
            if os.path.isfile(p) and not name.endswith((".pyc", ".o", ".so")) and name != "gen_rtc_headers.inc":
                out.append(p)
    return out


def key_of(tag, files=None):
    h = hashlib.sha256()
    h.update(_gxx_version().encode())
    h.update(tag.encode())
    for p in (files if files is not None else source_files()):
        h.update(os.path.basename(p).encode())
        h.update(open(p, "rb").read())
    return h.hexdigest()[:20]


def cached(name, tag, build, files=None):
    """Path of the build product `name` for (tag, input contents); `build(path)` creates it when it is not cached yet."""
    base, ext = os.path.splitext(name)
    path = os.path.join(cache_dir(), "%s-%s%s" % (base, key_of(tag, files), ext))
    if os.path.exists(path):
        return path
    tmp = path + ".tmp%d" % os.getpid()
    build(tmp)
    os.replace(tmp, path)
    return path
