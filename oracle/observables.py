"""Oracle: localdensity, DensityCurrents, Currents (scalar restatement).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows (paths relative to /root/reference):
  matrix_element              src/operators/bases.jl:37-39  (ket: psi[i] * conj(psi[j]))
  localdensity                src/operators/latticeutils.jl:41-45
  DensityCurrents.getindex    src/zoo/currents.jl:92-102  (_block :24-27, _avg :5-10)
  findnz(::AbstractCurrents)  src/currents.jl:159-172   (pairs i<j, |J| >= 1e-10)
  Currents(curr)              src/currents.jl:223-237   (all pairs)
  Currents(curr, bonds)       src/currents.jl:238-255   (listed bonds only)
  currentsfrom / -fromto      src/currents.jl:85-109
States: a ket (1-D), a dense density matrix (2-D square, ``block=False``) or a Psi block
(N x M with weights w: P = Psi diag(w) Psi^dagger).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

CURRENTS_EPS = 1e-10   # src/currents.jl:4


class State:
    def __init__(self, data, weights=None, block=False):
        self.data = np.asarray(data, dtype=complex)
        self.block = block or (self.data.ndim == 2 and self.data.shape[0] != self.data.shape[1])
        if self.block:
            m = self.data.shape[1]
            self.w = np.ones(m) if weights is None else np.asarray(weights, float)

    def elem(self, i, j):
        """matrix_element(state, i, j), 0-based."""
        d = self.data
        if d.ndim == 1:
            return d[i] * np.conj(d[j])
        if self.block:
            return np.sum(self.w * d[i, :] * np.conj(d[j, :]))
        return d[i, j]


def localdensity(state, n_int=1):
    st = state if isinstance(state, State) else State(state)
    dim = st.data.shape[0]
    ns = dim // n_int
    out = np.zeros(ns)
    for i in range(ns):
        out[i] = np.real(sum(st.elem(j, j) for j in range(i * n_int, (i + 1) * n_int)))
    return out


def density_current(H, state, i, j, n_int=1):
    """curr[i, j] with 1-based SITE indices (src/zoo/currents.jl:92-102)."""
    st = state if isinstance(state, State) else State(state)
    Hc = H.tocsr() if sp.issparse(H) else None
    out = 0.0
    for a in range(n_int):
        for b in range(n_int):
            ip = a + (i - 1) * n_int
            jp = b + (j - 1) * n_int
            t = Hc[ip, jp] if Hc is not None else H[ip, jp]
            out += 2 * np.imag(t * st.elem(jp, ip))
    return float(out)


def site_adjacency(H, n_int=1):
    """Site pairs (i < j, 1-based) with any stored H block entry - the only pairs that can
    carry current; equals AdjacencyMatrix(H) restricted to the upper triangle."""
    Hc = sp.coo_matrix(H)
    si = Hc.row // n_int
    sj = Hc.col // n_int
    mask = si < sj
    pairs = sorted(set(zip((sj[mask] + 1).tolist(), (si[mask] + 1).tolist())))
    # sort as findnz of a CSC matrix filtered by I < J: column (J) major, row (I) minor
    return [(i, j) for j, i in pairs]


def currents_findnz(H, state, n_int=1, pairs=None):
    """(Is, Js, Vs) like findnz(::AbstractCurrents): j outer, i < j inner, drop |J| < eps.
    ``pairs`` restricts the evaluation (all pairs if None - O(n^2), small lattices only)."""
    st = state if isinstance(state, State) else State(state)
    ns = st.data.shape[0] // n_int
    if pairs is None:
        pairs = [(i, j) for j in range(1, ns + 1) for i in range(1, j)]
    Is, Js, Vs = [], [], []
    for i, j in pairs:
        v = density_current(H, st, i, j, n_int)
        if abs(v) < CURRENTS_EPS:
            continue
        Is.append(i)
        Js.append(j)
        Vs.append(v)
    return np.array(Is, int), np.array(Js, int), np.array(Vs, float)


def currents_matrix(H, state, n_int=1, pairs=None):
    """Currents(curr) / Currents(curr, bonds): antisymmetric sparse matrix of site currents."""
    st = state if isinstance(state, State) else State(state)
    ns = st.data.shape[0] // n_int
    Is, Js, Vs = currents_findnz(H, st, n_int, pairs)
    m = sp.coo_matrix((np.concatenate([Vs, -Vs]),
                       (np.concatenate([Is, Js]) - 1, np.concatenate([Js, Is]) - 1)),
                      shape=(ns, ns)).tocsc()
    return m


def currents_from(H, state, src, n_int=1):
    """currentsfromto(curr, src): total current out of site ``src`` (src/currents.jl:85-109)."""
    st = state if isinstance(state, State) else State(state)
    ns = st.data.shape[0] // n_int
    return sum(density_current(H, st, src, j, n_int) for j in range(1, ns + 1) if j != src)


def localexpect(op, state, n_int):
    """localexpect(op, state) (src/operators/latticeutils.jl:13-20):
    [sum(op[j,k] * matrix_element(state, (i-1)N + k, (i-1)N + j) for j, k) for i in sites]."""
    st = state if isinstance(state, State) else State(state)
    op = np.asarray(op, complex)
    ns = st.data.shape[0] // n_int
    out = np.zeros(ns, complex)
    for i in range(ns):
        out[i] = sum(op[j, k] * st.elem(i * n_int + k, i * n_int + j) for j in range(n_int) for k in range(n_int))
    return out


def operator_current(H, state, op, i, j, n_int):
    """LocalOperatorCurrents.getindex (src/zoo/currents.jl:169-181), 1-based site indices."""
    st = state if isinstance(state, State) else State(state)
    Hc = H.tocsr() if sp.issparse(H) else np.asarray(H)
    O = np.asarray(op, complex)
    T = np.array([[Hc[(i - 1) * n_int + a, (j - 1) * n_int + b] for b in range(n_int)] for a in range(n_int)], complex)
    out = 0.0
    for a in range(n_int):
        for b in range(n_int):
            ot = sum(O[a, k] * T[k, b] for k in range(n_int))
            out += 2 * np.imag(ot * st.elem((j - 1) * n_int + b, (i - 1) * n_int + a))
    return float(out)
