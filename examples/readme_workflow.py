#!/usr/bin/env python
"""The reference's README workflow (README.md:25-55, BASELINE config 1) on the B200 backend:
SquareLattice(10,10), point flux ramped through the central plaquette, T = 0 half-filled density
matrix, unitary evolution over 0:0.1:20 with localdensity and DensityCurrents every frame.

    python examples/readme_workflow.py            # needs a CUDA device and the built library
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lm_b200 as lm  # noqa: E402

l = lm.SquareLattice(10, 10)
tau = 10.0


def h(B):
    return lm.tightbinding_hamiltonian(l, field=lm.PointFlux(B, (5.5, 5.5)))


P0 = lm.densitymatrix(h(0.0), mu=0.0)             # (Psi_occ, w) hand-off instead of a dense N x N matrix
ev = lm.Evolution(lm.B200Exp(tol=1e-12), lambda t: h(0.2 * min(t, tau) / tau), P0)
densities = lm.TimeSequence()
for P, H, t in ev(np.arange(0, 201) * 0.1):
    rho = lm.localdensity(P)
    cur = lm.Currents(lm.DensityCurrents(H, P))
    densities[t] = rho.values
    if abs(t - round(t)) < 1e-9 and int(round(t)) % 5 == 0:
        Is, Js, Vs = lm.findnz(cur)
        print("t = %5.1f  N = %.10f  max|J| = %.4e  (%d bonds carry current)" % (t, rho.values.sum(), np.abs(Vs).max() if len(Vs) else 0.0, len(Vs)))
charge_flow = densities.differentiate()
print("frames:", len(densities), " max |d rho/dt| at t=5.05:", np.abs(charge_flow[5.05]).max())
