"""In-tree build of the CUDA shared library (sm_100a only)."""
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# translation units: the C ABI + the ELL kernels, and the instantiations of the register-tiled
# stencil kernel (large fully-unrolled kernels, kept apart so that api.cu rebuilds quickly)
UNITS = ["api.cu", "stencil.cu"] + ["stencil_k%d.cu" % i for i in range(9)]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "kernels.cuh"), os.path.join(CSRC, "stencil.cuh"), os.path.join(CSRC, "stencil_inst.cuh"), os.path.join(CSRC, "stencil_unit.inc"),
           os.path.join(CSRC, "taylor_roots.h"), os.path.join(HERE, "..", "include", "lm_b200.h")]
SRC = os.path.join(CSRC, "api.cu")
DEPS = [os.path.join(CSRC, u) for u in UNITS] + HEADERS
LIB = os.path.join(HERE, "lib", "liblm_b200.so")
OBJ_DIR = os.path.join(HERE, "lib", "obj")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build liblm_b200.so")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def _compile(unit, force, verbose, extra):
    src = os.path.join(CSRC, unit)
    obj = os.path.join(OBJ_DIR, unit.replace(".cu", ".o"))
    deps = [src] + HEADERS
    if not force and os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps):
        return obj
    cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-c", src, "-o", obj]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return obj


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> lib/liblm_b200.so with nvcc for sm_100a."""
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    extra = os.environ.get("LM_NVCC_EXTRA", "").split()
    with ThreadPoolExecutor(min(len(UNITS), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda u: _compile(u, force, verbose, extra), UNITS))
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-ldl"]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    return LIB
