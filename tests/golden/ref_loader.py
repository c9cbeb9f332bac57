"""Loader of the golden series written by julia/make_golden.jl (the REFERENCE's own Evolution with
CachedExp(threshold=1e-14)).  The files live in tests/golden/ref/ and are absent until somebody with
Julia runs the script; tests that consume them skip otherwise."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "ref")

# (oracle Hamiltonian factory arguments, device Hamiltonian factory arguments) of the three configs
CONFIGS = ("config1", "config3s", "config4s")


def available(name):
    return os.path.exists(os.path.join(REF, name + "_meta.txt"))


def load(name):
    ns, npairs, nf, dt = open(os.path.join(REF, name + "_meta.txt")).read().split()
    ns, npairs, nf, dt = int(ns), int(npairs), int(nf), float(dt)
    f = lambda suffix, dtype: np.fromfile(os.path.join(REF, name + suffix), dtype=dtype)
    w0 = f("_w0.f64", "<f8")
    M = len(w0)
    psi0 = f("_psi0.c128", "<c16")
    N = psi0.size // M
    return dict(n_sites=ns, dt=dt, times=np.arange(nf) * dt,
                rho=f("_rho.f64", "<f8").reshape(nf, ns),
                pairs=f("_pairs.i64", "<i8").reshape(2, npairs).T,          # Julia wrote an n_pairs x 2 matrix column-major
                J=f("_J.f64", "<f8").reshape(nf, npairs),
                Psi0=psi0.reshape(M, N).T.copy(), w0=w0)


def hamiltonians(name):
    """(oracle t -> H(t), device-mirror t -> H(t), n_int) of a config (same definitions as julia/make_golden.jl)."""
    import lm_b200 as lm
    from oracle import fields as F, lattice as L, operators as OP
    if name == "config1":
        lo, ld = L.square_lattice(10, 10), lm.SquareLattice(10, 10)
        B = lambda t: 0.2 * min(t, 10.0) / 10.0
        return (lambda t: OP.tightbinding_hamiltonian(lo, field=F.PointFlux(B(t), (5.5, 5.5))),
                lambda t: lm.tightbinding_hamiltonian(ld, field=lm.PointFlux(B(t), (5.5, 5.5))), 1)
    if name == "config3s":
        lo, ld = L.square_lattice(12, 10), lm.SquareLattice(12, 10)
        B = lambda t: 0.1 * min(t, 1.0)
        return (lambda t: OP.qwz(lo, field=F.LandauGauge(B(t))), lambda t: lm.qwz(ld, field=lm.LandauGauge(B(t))), 2)
    if name == "config4s":
        lo, ld = L.honeycomb_lattice(9, 8, periodic=(1,)), lm.HoneycombLattice(9, 8, boundaries=[("axis1", True)])
        return (lambda t: OP.haldane(lo, 1.0, 0.2, 0.1, field=F.LandauGauge(0.03)),
                lambda t: lm.haldane(ld, 1.0, 0.2, 0.1, field=lm.LandauGauge(0.03)), 1)
    raise KeyError(name)
