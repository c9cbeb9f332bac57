"""CPU build of the WHOLE library for the test-suite: csrc/*.cu compiled by g++ through the CUDA
shim in ``shim/`` into a shared object with the same C ABI as liblm_b200.so, every kernel executed
by OS threads (see shim/cuda_runtime.h).  The only source change is mechanical: each
``kernel<<<grid, block, smem, stream>>>(args)`` becomes ``lm_emul::enqueue(grid, block, kernel, args)``.

TEST INFRASTRUCTURE ONLY: the product never loads this library (latticemodels.jl_b200/_lib.py loads
lib/liblm_b200.so or raises); tests/test_whole_library_cpu.py swaps it in explicitly to run the GPU
parity tests at small sizes without a GPU.
"""
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

import build_cache

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "latticemodels.jl_b200", "csrc")


def _match_back(s, i):
    """s[i] == '>': index of the matching '<'."""
    depth = 0
    while i >= 0:
        if s[i] == ">":
            depth += 1
        elif s[i] == "<":
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced template argument list")


def _match_fwd(s, i):
    """s[i] == '(': index of the matching ')'."""
    depth = 0
    while i < len(s):
        if s[i] == "(":
            depth += 1
        elif s[i] == ")":
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced argument list")


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite_launches(src):
    """kernel<T...><<<grid, block[, smem[, stream]]>>>(args) -> lm_emul::enqueue(grid, block, [](auto... a_) { kernel<T...>(a_...); }, args)"""
    out, pos, n = "", 0, 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            return out + src[pos:], n
        j = k - 1
        while src[j].isspace():
            j -= 1
        if src[j] == ">":
            j = _match_back(src, j) - 1
        while j >= 0 and (src[j].isalnum() or src[j] in "_:"):
            j -= 1
        kernel = src[j + 1:k].strip()
        e = src.index(">>>", k)
        cfg = _split_top(src[k + 3:e])
        a0 = e + 3
        while src[a0].isspace():
            a0 += 1
        assert src[a0] == "(", "launch without an argument list near: " + src[k - 40:k + 40]
        a1 = _match_fwd(src, a0)
        args = src[a0 + 1:a1].strip()
        call = "lm_emul::enqueue(%s, %s, [](auto... a_) { %s(a_...); }%s)" % (cfg[0], cfg[1], kernel, (", " + args) if args else "")
        out += src[pos:j + 1] + call
        pos = a1 + 1
        n += 1


def build(outdir, opt="-O1", verbose=False):
    """Returns the path of the emulated shared library: from the content-addressed cache (build_cache.py) when the kernel
    sources, the shim and the flags are unchanged, else built in ``outdir`` and added to the cache."""
    return build_cache.cached("liblm_b200_emul.so", "whole-library " + opt, lambda path: shutil.copy(_build(outdir, opt, verbose), path))


def _build(outdir, opt="-O1", verbose=False):
    gxx = shutil.which("g++")
    if gxx is None:
        raise RuntimeError("g++ not available")
    pkg = os.path.join(outdir, "pkg", "csrc")
    os.makedirs(pkg, exist_ok=True)
    os.makedirs(os.path.join(outdir, "include"), exist_ok=True)
    shutil.copy(os.path.join(ROOT, "include", "lm_b200.h"), os.path.join(outdir, "include", "lm_b200.h"))
    units, launches = [], 0
    for name in sorted(os.listdir(CSRC)):
        text = open(os.path.join(CSRC, name)).read()
        text, n = rewrite_launches(text)
        launches += n
        open(os.path.join(pkg, name), "w").write(text)
        if name.endswith(".cu"):
            units.append(name)
    assert launches >= 60, launches

    def cc(name):
        obj = os.path.join(outdir, name.replace(".cu", ".o"))
        cmd = [gxx, "-x", "c++", "-std=c++20", opt, "-w", "-pthread", "-fPIC", "-I", os.path.join(HERE, "shim"), "-c", os.path.join(pkg, name), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("g++ failed on %s:\n%s" % (name, (res.stdout + res.stderr)[-6000:]))
        return obj

    with ThreadPoolExecutor(min(len(units), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(cc, units))
    so = os.path.join(outdir, "liblm_b200_emul.so")
    res = subprocess.run([gxx, "-shared", "-pthread", "-o", so] + objs + ["-ldl"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return so


if __name__ == "__main__":
    import sys
    print(build(sys.argv[1] if len(sys.argv) > 1 else "/tmp/lm_emul_build", verbose=True))
