#!/usr/bin/env python
"""Generates tests/golden/config1_frames.npz - golden vectors for BASELINE config 1 (README
workflow: SquareLattice(10,10), PointFlux ramp through (5.5, 5.5), mu = 0 density matrix,
times 0:0.1:20) - from the CPU oracle (exact exp(-i H dt) products, reference stepping
semantics).  The reference itself (Julia) cannot run in the build container, so these are ORACLE
outputs pinned for regression; the oracle is in turn pinned to the reference's known answers by
tests/test_oracle_pins.py.  Inputs (Psi0, w0) are stored too: the Fermi level of this lattice is
10-fold degenerate, so the initial state must be shared, never recomputed.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import evolution as EV, fields as F, lattice as L, observables as OB, operators as OP, spectrum as SP  # noqa: E402

lat = L.square_lattice(10, 10)
TAU = 10.0


def h(t):
    return OP.tightbinding_hamiltonian(lat, field=F.PointFlux(0.2 * min(t, TAU) / TAU, (5.5, 5.5)))


P0, Psi0, w0 = SP.densitymatrix(h(0.0), mu=0.0)
ts = np.arange(0, 201) * 0.1
pairs = OB.site_adjacency(h(0.0), 1)
keep = [0, 1, 25, 50, 100, 101, 150, 200]
rho, cur = [], []
for k, (st, H, t) in enumerate(EV.Evolution(h, [P0], solver="exact")(ts)):
    if k in keep:
        rho.append(OB.localdensity(st[0], 1))
        cur.append(np.array([OB.density_current(H, st[0], i, j, 1) for i, j in pairs]))
np.savez_compressed(os.path.join(HERE, "config1_frames.npz"), Psi0=Psi0, w0=w0, frames=np.array(keep),
                    times=ts[keep], rho=np.array(rho), J=np.array(cur), pairs=np.array(pairs))
print("wrote config1_frames.npz:", np.array(rho).shape, np.array(cur).shape)
