"""Oracle: equilibrium initial states.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows (paths relative to /root/reference):
  diagonalize (dense eigen)     src/spectrum.jl:48-55
  projector(f, eig)             src/spectrum.jl:226-228   P = V * (f.(E) .* V')
  densfun                       src/spectrum.jl:230-232   T = 0 occupies E <= mu
  ensemble_densitymatrix        src/spectrum.jl:253-257
  fermisphere_densitymatrix     src/spectrum.jl:258-265
  groundstate                   src/spectrum.jl:205
"""
from __future__ import annotations

import numpy as np


def diagonalize(H):
    Hd = H.toarray() if hasattr(H, "toarray") else np.asarray(H)
    E, V = np.linalg.eigh(Hd)
    return E, V


def densfun(T, mu, statistics=1):
    """statistics: +1 FermiDirac, -1 BoseEinstein (Int(statistics) in the reference)."""
    def f(E):
        if E - mu == 0 and T == 0:
            return 1.0
        with np.errstate(over="ignore", divide="ignore"):
            x = (E - mu) / T if T != 0 else (np.inf if E > mu else -np.inf)
            return float(1.0 / (np.exp(x) + statistics))
    return f


def occupations(E, T=0.0, mu=0.0, statistics=1):
    f = densfun(T, mu, statistics)
    return np.array([f(e) for e in E])


def densitymatrix(H, T=0.0, mu=0.0, statistics=1):
    """Dense P0 = V diag(f(E)) V^dagger; also returns (Psi0, w) with the zero-weight columns
    dropped - the Psi-block hand-off of SURVEY.md section 8(a) row S1."""
    E, V = diagonalize(H)
    w = occupations(E, T, mu, statistics)
    P = (V * w[None, :]) @ V.conj().T
    keep = w != 0
    return P, V[:, keep].copy(), w[keep].copy()


def fermisphere(H, nparticles):
    E, V = diagonalize(H)
    Psi = V[:, :nparticles].copy()
    return Psi @ Psi.conj().T, Psi, np.ones(nparticles)


def groundstate(H):
    E, V = diagonalize(H)
    return V[:, 0].copy()
