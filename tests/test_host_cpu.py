"""CPU-only tests: host mirror vs oracle, C-ABI library exports, sharding logic (no GPU compute)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import lm_b200 as lm
from oracle import fields as F
from oracle import lattice as L
from oracle import operators as OP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = {
    "tb44": (lambda: lm.tightbinding_hamiltonian(lm.SquareLattice(4, 4)),
             lambda: OP.tightbinding_hamiltonian(L.square_lattice(4, 4))),
    "qwz_landau": (lambda: lm.qwz(lm.SquareLattice(6, 5), field=lm.LandauGauge(0.1)),
                   lambda: OP.qwz(L.square_lattice(6, 5), field=F.LandauGauge(0.1))),
    "qwz_pbc": (lambda: lm.qwz(lm.SquareLattice(6, 6, boundaries=[("axis1", True)]), field=lm.LandauGauge(0.5)),
                lambda: OP.qwz(L.square_lattice(6, 6, periodic=(1,)), field=F.LandauGauge(0.5))),
    "haldane_sym": (lambda: lm.haldane(lm.HoneycombLattice(5, 4), 1.0, 0.2, 0.1, field=lm.SymmetricGauge(0.05)),
                    lambda: OP.haldane(L.honeycomb_lattice(5, 4), 1.0, 0.2, 0.1, field=F.SymmetricGauge(0.05))),
    "tb_axial": (lambda: lm.tightbinding_hamiltonian(lm.SquareLattice(10, 10), field=lm.PointFlux(0.13, (5.5, 5.5))),
                 lambda: OP.tightbinding_hamiltonian(L.square_lattice(10, 10), field=F.PointFlux(0.13, (5.5, 5.5)))),
    "tb_t123_twist_singular": (
        lambda: lm.tightbinding_hamiltonian(lm.SquareLattice(5, 6, boundaries=[("axis1", True), ("axis2", 0.7)]),
                                            t1=1, t2=0.3, t3=0.1, field=lm.PointFlux(0.2, (2.5, 2.5), gauge="singular")),
        lambda: OP.tightbinding_hamiltonian(L.square_lattice(5, 6, periodic=(1,), twists={2: 0.7}),
                                            t1=1, t2=0.3, t3=0.1, field=F.PointFlux(0.2, (2.5, 2.5), "singular"))),
    "sum_field": (lambda: lm.tightbinding_hamiltonian(lm.SquareLattice(7, 7), field=lm.LandauGauge(0.07) + lm.PointFlux(0.3, (3.5, 3.5))),
                  lambda: OP.tightbinding_hamiltonian(L.square_lattice(7, 7), field=F.LandauGauge(0.07) + F.PointFlux(0.3, (3.5, 3.5)))),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_host_assembly_matches_oracle(name):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        h = CASES[name][0]()
    o = CASES[name][1]()
    assert h.data.shape == o.shape
    assert abs(h.data - o).max() < 1e-15
    assert h.data.nnz == o.nnz


def test_site_order_and_nn_goldens():
    # src/core/latticevalue.jl:71-81 ; src/lattices/bravais/nearestneighbor.jl:149-153
    assert lm.SquareLattice(2, 2).x.tolist() == [1.0, 1.0, 2.0, 2.0]
    assert lm.SquareLattice(3, 3).y.tolist() == [1.0, 2.0, 3.0] * 3
    nn = lm.HoneycombLattice(5, 5).nearest_neighbor(1)
    assert [t.key() for t in nn] == [((1, 2), (0, -1)), ((1, 2), (-1, 0)), ((1, 2), (0, 0))]
    for n in (1, 2, 3):
        for mk, ok in ((lm.SquareLattice, L.square_lattice), (lm.HoneycombLattice, L.honeycomb_lattice)):
            a = [t.key() for t in mk(4, 4).nearest_neighbor(n)]
            b = [(t.site_indices, t.translate_uc) for t in L.nearest_neighbor(ok(4, 4), n)]
            assert a == b


def test_structure_cached_and_lazy():
    l = lm.SquareLattice(5, 5)
    h1 = lm.tightbinding_hamiltonian(l, field=lm.LandauGauge(0.1))
    h2 = lm.tightbinding_hamiltonian(l, field=lm.LandauGauge(0.2))
    assert h1.structure is h2.structure          # per-step closure cost = dictionary lookup
    assert abs(h1.data - h2.data).max() > 1e-3


def test_field_descriptor_layout():
    f = lm.LandauGauge(0.1) + lm.PointFlux(0.2, (1.5, 2.5)) + lm.PointFlux(0.3, (0, 1), gauge="singular")
    kinds, params = f.descriptor()
    assert kinds.tolist() == [1, 3, 4]
    assert params.tolist() == [[0.1, 0, 0], [0.2, 1.5, 2.5], [0.3, 0, 1]]
    with pytest.raises(lm.ArgumentError):
        lm.PointFlux(0.1, gauge="bogus")


def test_densitymatrix_handoff_matches_oracle_projector():
    from oracle import spectrum as SP
    H = lm.qwz(lm.SquareLattice(4, 4))
    pp = lm.densitymatrix(H, mu=0.0)
    P, Psi, w = SP.densitymatrix(OP.qwz(L.square_lattice(4, 4)), mu=0.0)
    assert pp.psi.shape == Psi.shape
    assert np.allclose(pp.dense(), P, atol=1e-12)


def test_shard_range_partition():
    for M in (1, 7, 64, 5000):
        for n in (1, 2, 3, 8):
            rs = [lm.shard_range(M, r, n) for r in range(n)]
            assert rs[0][0] == 0 and rs[-1][1] == M
            assert all(rs[k][1] == rs[k + 1][0] for k in range(n - 1))


def test_library_exports_every_header_symbol():
    """The C-ABI library loads and exports every symbol include/lm_b200.h declares (no compute)."""
    hdr = open(os.path.join(ROOT, "include", "lm_b200.h")).read()
    names = set(re.findall(r"\b(lm_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    if not os.path.exists(lm.library_path()):
        lm.build()
    lib = ctypes.CDLL(lm.library_path())
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    from importlib import import_module
    protos = import_module("lm_b200._lib").PROTOTYPES
    assert names - {"lm_version", "lm_last_error"} == set(protos)
    assert lib.lm_version() >= 100


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "latticemodels.jl_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn


def test_timesequence_semantics():
    # test/test_timedeps.jl:2-40 flavour: tolerant lookup, differentiate, integrate
    ts = lm.TimeSequence()
    for t in np.arange(0, 1.01, 0.1):
        ts[t] = np.array([t * t, 2 * t])
    assert np.allclose(ts[0.3 + 1e-12], [0.09, 0.6])
    d = ts.differentiate()
    assert np.allclose(d[0.15], [0.3, 2.0])
    i = ts.integrate()
    assert np.allclose(i[1.0][1], 1.0, atol=1e-12)
    with pytest.raises(KeyError):
        ts[0.35]


def test_timesequence_reference_testset():
    """test/test_timedeps.jl:2-40 re-expressed on the host mirror (src/timesequence.jl:6-137,193-265):
    length-mismatch error, site / mask indexing under a time interval, iteration, differentiate,
    tolerant keys, interval slicing, integrate, KeyError, empty / delete! by key and by interval;
    plus the two doctest series (:162-191, :218-247)."""
    TS = lm.TimeSequence
    l = lm.SquareLattice(3, 3)
    c = np.asarray(l.coords)
    x, y = c[:, 0], c[:, 1]
    xy, xly, site = x * y, x < y, 4                      # l[5] in 1-based Julia
    with pytest.raises(ValueError):
        TS([0.5], [xy, xy])
    rec = TS()
    rec[0] = xy
    rec[1] = xy
    rec[2] = xy
    assert rec.slice((0, np.inf), index=site) == TS(rec.timestamps(), [xy[site]] * 3)
    assert rec.slice((0, np.inf), index=xly) == TS(range(3), [xy[xly]] * 3)
    assert [(t, v.tolist()) for t, v in rec] == [(0, xy.tolist()), (1, xy.tolist()), (2, xy.tolist())]
    z = np.zeros(9)
    assert rec.differentiate() == TS([0.5, 1.5], [z, z])
    assert rec.slice((0, np.inf), index=site).differentiate() == TS([0.5, 1.5], [0, 0])
    rec2 = TS()
    rec2[0] = xy * 0
    rec2[1] = xy * 1
    rec2[2] = xy * 2
    assert (rec2[1e-9] == z).all() and (rec2[1 + 1e-9] == xy).all()
    assert rec2[(0.9, 2.1)] == TS([1, 2], [xy, xy * 2])
    assert rec.integrate() == rec2
    with pytest.raises(KeyError):
        rec2[0.5]
    g = np.arange(0, 101) / 10
    ts = TS(g, g ** 2)
    ts2 = ts.empty()
    assert isinstance(ts2, TS) and len(ts2) == 0
    for t in np.arange(70, 101) / 10:
        ts.delete(t)
    ts.delete((5, 7))
    for t in np.arange(0, 50) / 10:
        ts2[t] = t ** 2
    assert ts == ts2
    d = TS(g, g).differentiate()                          # f(t) = t -> f'(t) = 1 at the midpoints
    assert len(d) == 100 and abs(d.times[0] - 0.05) < 1e-15 and abs(d.times[-1] - 9.95) < 1e-12 and np.allclose(d.values(), 1.0)
    i = TS(g, g).integrate()                              # F(t) = t^2 / 2, first value zero
    assert len(i) == 101 and i.values()[0] == 0 and np.allclose(i.values(), g ** 2 / 2)
    u = TS()                                              # keys are kept in time order (searchsortedfirst)
    u[2] = 1.0
    u[0] = 2.0
    u[1] = 3.0
    assert u.times == [0.0, 1.0, 2.0] and u.values() == [2.0, 3.0, 1.0]
    with pytest.raises(RuntimeError):
        TS([0.0], [1.0]).differentiate()


GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch.distributed as dist
import lm_b200 as lm
from importlib import import_module
D = import_module("lm_b200.distributed")
from oracle import lattice as L, operators as OP, spectrum as SP, observables as OB
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"], rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()
H = OP.qwz(L.square_lattice(4, 4))
P, Psi, w = SP.densitymatrix(H, mu=0.0)
b, e = lm.shard_range(Psi.shape[1], rank, 2)
# each rank reduces ITS column shard (oracle arithmetic stands in for the device kernel here)
part_rho = OB.localdensity(OB.State(Psi[:, b:e], w[b:e], block=True), 2)
pairs = OB.site_adjacency(H, 2)
part_J = np.array([OB.density_current(H, OB.State(Psi[:, b:e], w[b:e], block=True), i, j, 2) for i, j in pairs])
tot = D.allreduce_host(np.concatenate([part_rho, part_J]))
full_rho = OB.localdensity(P, 2)
full_J = np.array([OB.density_current(H, P, i, j, 2) for i, j in pairs])
assert np.allclose(tot[:16], full_rho, atol=1e-13), "rho"
assert np.allclose(tot[16:], full_J, atol=1e-13), "J"
dist.destroy_process_group()
print("OK", rank)
'''


def test_column_sharding_allreduce_gloo_world2(tmp_path):
    """N > 1 path on CPU: column shards + one all-reduce of [rho | J] reproduce the unsharded
    observables."""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER % ROOT)
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), PORT=str(port), MASTER_ADDR="127.0.0.1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_compiled_stencil_masks_cover_reference_models():
    """The sparsity masks compiled into the register-tiled stencil kernels (LM_ST_MASK* in
    csrc/stencil.cuh) cover the stencils of the reference's own models, derived
    here from the oracle's Hamiltonians; periodic wrapping and field phases do not change them."""
    import re
    from oracle import fields as F, lattice as L, operators as OP, stencil as ST
    src = open(os.path.join(ROOT, "latticemodels.jl_b200", "csrc", "stencil.cuh")).read()
    masks = {int(k): int(v, 16) for k, v in re.findall(r"#define LM_ST_MASK(\d) (0x[0-9a-f]+)ull", src)}
    assert sorted(masks) == [0, 1, 2, 3, 4, 5]
    cases = [
        (OP.tightbinding_hamiltonian(L.square_lattice(6, 7)), 6, 7, 1, 0),
        (OP.tightbinding_hamiltonian(L.square_lattice(6, 7, periodic=(1, 2)), field=F.LandauGauge(0.1)), 6, 7, 1, 0),
        (OP.tightbinding_hamiltonian(L.square_lattice(6, 7), t1=1, t2=0.3), 6, 7, 1, 1),
        (OP.tightbinding_hamiltonian(L.honeycomb_lattice(5, 6)), 5, 6, 2, 2),
        (OP.qwz(L.square_lattice(5, 6), field=F.LandauGauge(0.2)), 5, 6, 2, 3),
        (OP.haldane(L.honeycomb_lattice(5, 6), 1.0, 0.2, 0.1), 5, 6, 2, 4),
        (OP.haldane(L.honeycomb_lattice(5, 6, periodic=(1, 2)), 1.0, 0.2, 0.1), 5, 6, 2, 4),
        # third-neighbour honeycomb hops and a QWZ model with diagonal hops stay within |d| <= 1: catch-all RC = 2 mask
        (OP.tightbinding_hamiltonian(L.honeycomb_lattice(5, 6), t1=1, t2=0.2, t3=0.1), 5, 6, 2, 5),
        (OP.qwz(L.square_lattice(5, 6)) + OP.tightbinding_hamiltonian(L.square_lattice(5, 6), n_int=2, t1=0, t2=0.3), 5, 6, 2, 5),
    ]
    for H, n1, n2, rc_want, mid in cases:
        rc, m = ST.stencil_mask(H, n1, n2)
        assert rc == rc_want and m is not None
        assert m & ~masks[mid] == 0, (mid, hex(m), hex(masks[mid]))
        # tight: the compiled mask adds nothing but the on-site (same-cell) block to the model's own pattern
        extra = masks[mid] & ~m
        assert mid in (1, 5) or extra & ~(((1 << (rc * rc)) - 1) << (4 * rc * rc)) == 0
    # patterns as the kernels see them: StPat<id> with multi-word masks (three / four rows per cell) and the
    # "purely imaginary" entries of the real / imaginary value class
    def word(tok):
        tok = tok.strip()
        return int(masks[int(tok[-1])] if tok.startswith("LM_ST_MASK") else (imags[int(tok[-1])] if tok.startswith("LM_ST_IMAG") else int(tok.rstrip("ul"), 16 if tok.startswith("0x") else 10)))
    imags = {int(k): int(v, 16) for k, v in re.findall(r"#define LM_ST_IMAG(\d) (0x[0-9a-f]+)ull", src)}
    pats = {}
    for pid, rc, m, im in re.findall(r"StPat<(\d+)> \{ static constexpr int rc = (\d); static constexpr st_mask_t mask = \{\{(.*?)\}\},\s*imag = \{\{(.*?)\}\}; \};", src, re.S):
        pats[int(pid)] = (int(rc), sum(word(t) << (64 * k) for k, t in enumerate(m.split(","))), sum(word(t) << (64 * k) for k, t in enumerate(im.split(","))))
    assert sorted(pats) == list(range(10)) and all(pats[k][1] == masks[k] for k in masks)
    # `qwz` itself has a diagonal on-site term: pattern 9 is its exact mask (9 entries per row), pattern 3 the full-block superset
    rcq, mq = ST.stencil_mask(OP.qwz(L.square_lattice(5, 6), field=F.LandauGauge(0.2)), 5, 6)
    assert rcq == 2 and mq == pats[9][1] and mq & ~pats[3][1] == 0 and pats[9][2] == pats[3][2]
    wide = [
        (OP.tightbinding_hamiltonian(L.kagome_lattice(5, 6), field=F.LandauGauge(0.1)), 5, 6, 3, 6),
        (OP.tightbinding_hamiltonian(L.kagome_lattice(5, 6, periodic=(1, 2))), 5, 6, 3, 6),
        (OP.tightbinding_hamiltonian(L.kagome_lattice(5, 6), t1=1, t2=0.3), 5, 6, 3, 7),
        (OP.kanemele(L.honeycomb_lattice(5, 6), 1.0, 0.2), 5, 6, 4, 8),
        (OP.kanemele(L.honeycomb_lattice(5, 6, periodic=(1, 2)), 1.0, 0.2, field=F.LandauGauge(0.05)), 5, 6, 4, 8),
    ]
    for H, n1, n2, rc_want, pid in wide:
        rc, m = ST.stencil_mask(H, n1, n2)
        assert rc == rc_want == pats[pid][0] and m is not None and m & ~pats[pid][1] == 0
        diag = sum(1 << (4 * rc * rc + a * rc + a) for a in range(rc))
        assert pats[pid][1] & ~m & ~diag == 0                     # tight up to the on-site diagonal
    # value classes: without a field every stored entry of the reference models is purely real, or purely
    # imaginary exactly where the pattern says so
    import scipy.sparse as sp
    for H, n1, n2, pid in [(OP.tightbinding_hamiltonian(L.square_lattice(6, 7)), 6, 7, 0), (OP.qwz(L.square_lattice(5, 6), 1.3), 5, 6, 3),
                           (OP.haldane(L.honeycomb_lattice(5, 6, periodic=(1, 2)), 1.0, 0.2, 0.1), 5, 6, 4),
                           (OP.tightbinding_hamiltonian(L.kagome_lattice(5, 6), t1=1, t2=0.3), 5, 6, 7),
                           (OP.kanemele(L.honeycomb_lattice(5, 6), 1.0, 0.2), 5, 6, 8)]:
        rc, _, imag = pats[pid]
        csr = sp.csr_matrix(H)
        for i in range(csr.shape[0]):
            ci, a = divmod(i, rc)
            for k in range(csr.indptr[i], csr.indptr[i + 1]):
                cj, b = divmod(int(csr.indices[k]), rc)
                d1, d2 = cj // n2 - ci // n2, cj % n2 - ci % n2
                d1 = d1 - n1 if d1 > n1 // 2 else (d1 + n1 if d1 < -(n1 // 2) else d1)
                d2 = d2 - n2 if d2 > n2 // 2 else (d2 + n2 if d2 < -(n2 // 2) else d2)
                bit = ((d1 + 1) * 3 + d2 + 1) * rc * rc + a * rc + b
                v = csr.data[k]
                assert (v.real == 0) if (imag >> bit) & 1 else (v.imag == 0), (pid, i, int(csr.indices[k]), v)
    # third-neighbour hops couple cells two apart: no |d| <= 1 mask
    assert ST.stencil_mask(OP.tightbinding_hamiltonian(L.square_lattice(8, 8), t1=1, t3=0.1), 8, 8)[1] is None
    # Haldane: 4 forward bonds from the A row, 5 from the B row (one correlator per bond)
    fw = ST.forward_entries(2, masks[4])
    assert [sum(1 for a, _, _ in fw if a == r) for r in (0, 1)] == [4, 5]
    # QWZ: 2 forward cells x 2 x 2 orbital pairs + the same-site orbital pair (never a current: mapped to -1)
    assert len(ST.forward_entries(1, masks[0])) == 2 and len(ST.forward_entries(2, masks[3])) == 9


def test_stencil_tile_logic_executes_on_cpu(tmp_path):
    """The register-tile body of the stencil kernels (st_tile + the compile-time slot / mask
    helpers of csrc/stencil.cuh) compiled with plain g++ through a CUDA shim and run on the CPU
    against an independently written reference loop: all five compiled patterns, tile shapes
    1x2, 2x2, 4x2, 4x4, with and without the product-form self term, complex128 and complex64 lanes."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    from concurrent.futures import ThreadPoolExecutor
    emul = os.path.join(ROOT, "tests", "cpu_emul")

    sys.path.insert(0, emul)
    import build_cache
    sys.path.pop(0)

    def group(g):
        flags = ["-std=c++20", "-O0", "-w", "-pthread", "-DLM_EMUL_GROUP=%d" % g]
        exe = build_cache.cached("stencil_emul_%d" % g, " ".join(flags), lambda path: subprocess.run(
            [gxx] + flags + ["-I", os.path.join(emul, "shim"), "-I", os.path.join(ROOT, "latticemodels.jl_b200", "csrc"),
                             os.path.join(emul, "stencil_emul.cpp"), "-o", path], check=True))
        return subprocess.run([exe], capture_output=True, text=True)
    with ThreadPoolExecutor(3) as ex:
        results = list(ex.map(group, range(3)))
    total = 0
    for res in results:
        assert res.returncode == 0 and res.stdout.startswith("OK "), res.stdout + res.stderr
        total += int(res.stdout.split()[1])
    assert total >= 450


def test_stencil_kernels_execute_on_cpu(tmp_path):
    """The WHOLE stencil kernels of csrc/stencil.cuh - k_apply_stencil_tma (default SpMM / propagator
    factor), k_apply_stencil and k_observe_stencil (fused localdensity + bond correlators) - executed
    on the CPU: every CUDA thread of a CTA is an OS thread (real __syncthreads, warp-shuffle
    mailboxes, atomics, mbarrier phase rule, alignment-checked cp.async.bulk).  All ten compiled
    patterns (one to four rows per unit cell) in the shapes the library launches, complex values and the
    real / imaginary class scalars, open / periodic / 3x3-torus / ragged lattices,
    complex128 and complex64, every MODE, and the column window + plain-store flag of the
    L2-resident strip schedule (columns outside the window must stay untouched)."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    from concurrent.futures import ThreadPoolExecutor
    emul = os.path.join(ROOT, "tests", "cpu_emul")

    sys.path.insert(0, emul)
    import build_cache
    sys.path.pop(0)

    def group(g):
        # the fully unrolled kernels are large: four pattern groups, built as separate programs in parallel
        # (and kept in the content-addressed build cache while the sources are unchanged)
        flags = ["-std=c++20", "-O0", "-w", "-pthread", "-DLM_EMUL_GROUP=%d" % g]
        exe = build_cache.cached("stencil_kernel_emul_%d" % g, " ".join(flags), lambda path: subprocess.run(
            [gxx] + flags + ["-I", os.path.join(emul, "shim"), "-I", os.path.join(ROOT, "latticemodels.jl_b200", "csrc"),
                             os.path.join(emul, "stencil_kernel_emul.cpp"), "-o", path], check=True))
        return subprocess.run([exe], capture_output=True, text=True, timeout=900)
    with ThreadPoolExecutor(4) as ex:
        results = list(ex.map(group, range(4)))
    total = 0
    for res in results:
        assert res.returncode == 0 and res.stdout.startswith("OK "), res.stdout + res.stderr
        total += int(res.stdout.split()[1])
    assert total >= 2000000


def test_timesequence_holds_arbitrary_values_and_nominal_keys():
    """the reference's TimeSequence{ET} stores arbitrary values (src/timesequence.jl:6-19) -
    Currents objects, LatticeValues, ragged tuples - and TimeSequence(f, evol_iter) keys its entries
    by the iterator's NOMINAL times (:41-43)."""
    import scipy.sparse as sp
    from lm_b200.observables import Currents, LatticeValue
    from lm_b200.timesequence import TimeSequence
    lat = lm.SquareLattice(2, 2)
    m = sp.csc_matrix(np.array([[0, 1., 0, 0], [-1., 0, 2., 0], [0, -2., 0, 0], [0, 0, 0, 0]]))
    ts = TimeSequence()
    ts[0.0] = Currents(m, lattice=lat)
    ts[0.5] = Currents(2 * m, lattice=lat)
    ts[1.0] = Currents(4 * m, lattice=lat)
    assert isinstance(ts[0.5], Currents) and ts[0.5] == Currents(2 * m, lattice=lat)
    d = ts.differentiate()
    assert isinstance(d[0.25], Currents) and d[0.25] == Currents(2 * m, lattice=lat) and d[0.75] == Currents(4 * m, lattice=lat)
    tv = TimeSequence([0.0, 1.0], [LatticeValue(lat, np.arange(4.0)), LatticeValue(lat, np.arange(4.0) + 2)])
    iv = tv.integrate()
    assert isinstance(iv[1.0], LatticeValue) and iv[1.0].lattice is lat and np.allclose(iv[1.0].values, np.arange(4.0) + 1)
    tt = TimeSequence([0.0, 1.0], [(np.zeros(3), np.zeros(5)), (np.ones(3), np.ones(5))])      # ragged (rho, J) tuples
    assert isinstance(tt[1.0], tuple) and len(tt[1.0][1]) == 5

    class It:                                   # an EvolutionIterator stand-in whose clock drifts off the nominal grid
        times = [0.0, 0.1, 0.2]

        def __iter__(self):
            class M:
                pass
            for k, t in enumerate(self.times):
                mo = M()
                mo.t = t + 1e-3 * k
                yield mo
    seq = TimeSequence(lambda mo: mo.t, It())
    assert seq.timestamps() == [0.0, 0.1, 0.2] and seq[0.2] == 0.2 + 2e-3


def test_currents_region_sums_accept_masks_on_the_host_path():
    """currentsfrom / currentsfromto on a materialised Currents take a boolean per-site mask
    or 1-based indices alike (`to_inds`, src/currents.jl:85-109)."""
    import scipy.sparse as sp
    from lm_b200.observables import Currents, currentsfrom, currentsfromto
    a = np.zeros((4, 4))
    a[0, 1], a[1, 2], a[2, 3] = 1.0, 2.0, 3.0
    c = Currents(sp.csc_matrix(a - a.T))
    mask = np.array([True, True, False, False])
    assert np.array_equal(currentsfrom(c, mask).values, currentsfrom(c, [1, 2]).values)
    assert np.array_equal(currentsfrom(c, [1, 2]).values, [0, 0, 2, 0])
    assert currentsfromto(c, mask) == currentsfromto(c, [1, 2]) == 2.0
    assert currentsfromto(c, mask, ~mask) == currentsfromto(c, [1, 2], [3, 4]) == 2.0


def test_run_time_specialisation_compiles_without_a_device():
    """csrc/stencil_rtc.cu: the library hands its own (embedded) kernel headers to NVRTC for a lattice pattern that is not among
    the compiled ones - here Kane-Mele with a spin-mixing nearest-neighbour term (four rows per cell), mask derived from the
    oracle Hamiltonian.  NVRTC needs no device, so the compilation itself (apply kernel of a product-form factor, fused
    observables kernel) is checked here; loading and launching are gpu tests (tests/test_zz_gpu_patterns.py)."""
    import ctypes as C
    from importlib import import_module
    from oracle import lattice as L, operators as OP, stencil as ST
    lib = import_module("lm_b200._lib").load()
    lib.lm_dbg_rtc_compile.restype = C.c_longlong
    lib.lm_dbg_rtc_compile.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.lm_dbg_rtc_error.restype = C.c_char_p
    lat = L.honeycomb_lattice(5, 6)
    sx = np.array([[0, 1], [1, 0]], complex)
    H = OP.kanemele(lat, 1.0, 0.2) + OP.construct_hamiltonian(lat, 2, [(0.3j * sx, L.nearest_neighbor(lat, 1))])
    rc, m = ST.stencil_mask(H, 5, 6)
    assert rc == 4 and m is not None
    m |= sum(1 << (4 * rc * rc + a * rc + a) for a in range(rc))
    words = (C.c_uint64 * 4)(*[(m >> (64 * k)) & (2 ** 64 - 1) for k in range(4)])
    zero = (C.c_uint64 * 4)(0, 0, 0, 0)
    n = lib.lm_dbg_rtc_compile(rc, words, zero, 0, 3)
    if n < 0 and b"libnvrtc" in lib.lm_dbg_rtc_error():
        pytest.skip("libnvrtc.so.12 not installed")
    assert n > 100000, lib.lm_dbg_rtc_error()[:2000]
    assert lib.lm_dbg_rtc_compile(rc, words, zero, 1, -1) > 10000, lib.lm_dbg_rtc_error()[:2000]      # complex64 observables kernel
