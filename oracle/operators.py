"""Oracle: Hamiltonian assembly (scalar restatement of the OperatorBuilder path).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows (paths relative to /root/reference):
  expand_bond / Peierls factor   src/operators/builder.jl:282-285
  OperatorBuilder.setindex!      src/operators/builder.jl:296-309  (auto_hermitian)
  block scatter, zero skipping   src/operators/builder.jl:60-66
  add_term! family               src/operators/constructoperator.jl:4-47
  construct_operator             src/operators/constructoperator.jl:136-156
  tightbinding_hamiltonian       src/operators/constructoperator.jl:190-194
  qwz / haldane                  src/zoo/models.jl:130-137,139-170
  composite index                src/operators/system.jl:16 (internal index fastest)
"""
from __future__ import annotations

import cmath
import math

import numpy as np
import scipy.sparse as sp

from . import fields as F
from . import lattice as L


class Builder:
    """Accumulating builder: H[(i', j')] += v.  Mirrors `builder[is, js, factor=f] += B`."""

    def __init__(self, lat, n_int, field=None, auto_hermitian=True):
        self.lat = lat
        self.n = n_int
        self.field = (field or F.NoField()).adapt(lat)   # builder.jl:168-169 adapt_field
        self.auto_hermitian = auto_hermitian
        self.data = {}
        self.dim = len(lat) * n_int

    def _add_block(self, i, j, B, factor):
        n = self.n
        for a in range(n):
            for b in range(n):
                v = B[a, b]
                if v == 0:            # builder.jl:64  iszero(v) || ...
                    continue
                key = ((i - 1) * n + a, (j - 1) * n + b)   # 0-based composite
                self.data[key] = self.data.get(key, 0.0) + v * factor

    def add_bond(self, i, ri, j, rj, fac_j, B, fac_i=1.0 + 0.0j):
        """builder[s1, s2] += B  (builder.jl:282-309)."""
        field_fact = cmath.exp(-2j * math.pi * self.field.line_integral(ri, rj))
        total = fac_j * field_fact * np.conj(fac_i)
        B = np.asarray(B, dtype=complex)
        self._add_block(i, j, B, total)
        if self.auto_hermitian and i != j:
            self._add_block(j, i, B.conj().T, np.conj(total))

    def add_onsite(self, i, B, factor):
        self._add_block(i, i, np.asarray(B, dtype=complex), factor)

    def to_csc(self):
        if not self.data:
            return sp.csc_matrix((self.dim, self.dim), dtype=complex)
        keys = np.array(list(self.data.keys()), dtype=np.int64)
        vals = np.array(list(self.data.values()), dtype=complex)
        m = sp.coo_matrix((vals, (keys[:, 0], keys[:, 1])), shape=(self.dim, self.dim)).tocsc()
        m.sort_indices()
        return m


def _op_matrix(op, n_int):
    """op_to_matrix (builder.jl:260-276): number -> n * one(internal)."""
    if np.isscalar(op):
        return complex(op) * np.eye(n_int, dtype=complex)
    m = np.asarray(op, dtype=complex)
    if m.shape != (n_int, n_int):
        raise ValueError("matrix size does not match on-site dims")
    return m


def construct_hamiltonian(lat, n_int, terms, field=None):
    """construct_operator (constructoperator.jl:136-156) for the term kinds on the hot path:
    ``(op, number)``, ``(op, per-site array)`` [LatticeValue], ``(op, BravaisTranslation)``,
    ``(op, [BravaisTranslation...])`` [BravaisSiteMapping].  Returns scipy CSC complex128."""
    b = Builder(lat, n_int, field)
    for op, what in terms:
        if np.isscalar(op) and op == 0:           # add_pair_terms!: iszero(first(pair)) skip
            continue
        B = _op_matrix(op, n_int)
        if isinstance(what, L.BravaisTranslation):
            _add_translation(b, B, what)
        elif isinstance(what, (list, tuple)) and what and isinstance(what[0], L.BravaisTranslation):
            for tr in what:                        # constructoperator.jl:31-38
                _add_translation(b, B, tr)
        elif np.isscalar(what):
            for i in range(1, len(lat) + 1):       # constructoperator.jl:39-47
                b.add_onsite(i, B, what)
        else:
            vals = np.asarray(what)
            assert vals.shape == (len(lat),)
            for i in range(1, len(lat) + 1):       # constructoperator.jl:21-30
                b.add_onsite(i, B, vals[i - 1])
    return b.to_csc()


def _add_translation(b, B, tr):
    for i, ri, j, rj, fac in L.iterate_bonds(b.lat, tr):   # constructoperator.jl:14-20
        b.add_bond(i, ri, j, rj, fac, B)


def tightbinding_hamiltonian(lat, n_int=1, t1=1, t2=0, t3=0, field=None):
    """constructoperator.jl:190-194."""
    terms = []
    for t, n in ((t1, 1), (t2, 2), (t3, 3)):
        if t != 0:
            terms.append((t, L.nearest_neighbor(lat, n)))
    return construct_hamiltonian(lat, n_int, terms, field)


def qwz(lat, m=1, field=None):
    """src/zoo/models.jl:130-137."""
    return construct_hamiltonian(lat, 2, [
        (np.array([[1, 0], [0, -1]], dtype=complex), m),
        (np.array([[1, -1j], [-1j, -1]], dtype=complex) / 2, L.translation(axis=1, nu=lat.unitcell.nu)),
        (np.array([[1, -1], [1, -1]], dtype=complex) / 2, L.translation(axis=2, nu=lat.unitcell.nu)),
    ], field)


def haldane(lat, t1, t2, m=0, field=None):
    """src/zoo/models.jl:162-170."""
    ms = np.array([m if bas == 1 else -m for _, bas in lat.pointers], dtype=float)
    # construct_hamiltonian(sys, lattice .|> (...), t1 => NN(1), im*t2 => honeycomb_2nn):
    # a bare LatticeValue is the on-site term `1 => lv` (constructoperator.jl:76-89).
    return construct_hamiltonian(lat, 1, [
        (1, ms),
        (t1, L.nearest_neighbor(lat, 1)),
        (1j * t2, L.HONEYCOMB_2NN),
    ], field)


def kanemele(lat, t1, t2, field=None):
    """src/zoo/models.jl:188-194."""
    sz = np.array([[1, 0], [0, -1]], dtype=complex)
    return construct_hamiltonian(lat, 2, [
        (t1, L.nearest_neighbor(lat, 1)),
        (1j * t2 * sz, L.HONEYCOMB_2NN),
    ], field)
