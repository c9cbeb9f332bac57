#!/usr/bin/env python
"""Generates tests/golden/config3s_frames.npz and config4s_frames.npz - golden vectors for REDUCED
versions of BASELINE configs 3 and 4 (the full sizes are far beyond an exact CPU oracle):

  config3s  QWZ (m = 1) on a 12 x 10 square lattice, Landau field ramped B(t) = 0.1 min(t, 1),
            regenerated every step; Psi0 = the 60 lowest eigenvectors of H(0), T = 0 weights
  config4s  Haldane (t1 = 1, t2 = 0.2, m = 0.1) on a 9 x 8 honeycomb lattice with periodic axis 1 and a
            constant Landau field 0.03; Psi0 = a seeded random orthonormal block of 48 columns with
            fractional weights (the synthetic state bench.py uses)

20 steps of dt = 0.1 each, frames 0, 1, 5, 10, 20: localdensity and DensityCurrents over every bond.
ORACLE outputs (exact exp(-i H dt) products, reference stepping semantics), pinned for regression;
the inputs (Psi0, w0) are stored so that both sides start from the same state.

    python tests/golden/make_golden_small.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import evolution as EV, fields as F, lattice as L, observables as OB, operators as OP  # noqa: E402

KEEP = [0, 1, 5, 10, 20]
TS = np.arange(0, 21) * 0.1


def h3(t):
    return OP.qwz(L.square_lattice(12, 10), field=F.LandauGauge(0.1 * min(t, 1.0)))


def h4(t):
    return OP.haldane(L.honeycomb_lattice(9, 8, periodic=(1,)), 1.0, 0.2, 0.1, field=F.LandauGauge(0.03))


def run(name, h, n_int, Psi0, w0):
    pairs = OB.site_adjacency(h(0.0), n_int)
    rho, cur = [], []
    for k, (st, H, t) in enumerate(EV.Evolution(h, [Psi0], solver="exact", block=True)(TS)):
        if k in KEEP:
            ost = OB.State(st[0], w0, block=True)
            rho.append(OB.localdensity(ost, n_int))
            cur.append(np.array([OB.density_current(H, ost, i, j, n_int) for i, j in pairs]))
    np.savez_compressed(os.path.join(HERE, name), Psi0=Psi0, w0=w0, frames=np.array(KEEP), times=TS[KEEP],
                        rho=np.array(rho), J=np.array(cur), pairs=np.array(pairs))
    print("wrote %s:" % name, np.array(rho).shape, np.array(cur).shape)


if __name__ == "__main__":
    E, V = np.linalg.eigh(h3(0.0).toarray())
    run("config3s_frames.npz", h3, 2, np.ascontiguousarray(V[:, :60]), np.ones(60))
    rng = np.random.default_rng(1234)
    N4 = h4(0.0).shape[0]
    Q, _ = np.linalg.qr(rng.standard_normal((N4, 48)) + 1j * rng.standard_normal((N4, 48)))
    run("config4s_frames.npz", h4, 1, np.ascontiguousarray(Q), np.linspace(0.2, 1.0, 48))
