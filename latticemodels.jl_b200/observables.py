"""Per-frame observables on device states.

Same names and semantics as the reference (paths relative to the reference repository):
  localdensity                 src/operators/latticeutils.jl:41-45
  DensityCurrents              src/zoo/currents.jl:81-102
  Currents(curr[, bonds])      src/currents.jl:223-255   (|J| < 1e-10 dropped, :4)
  findnz                       src/currents.jl:159-177
  currentsfrom / currentsfromto  src/currents.jl:85-109
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib
from .hamiltonian import DeviceHam, Hamiltonian
from .states import DeviceState, PsiProjector

CURRENTS_EPS = 1e-10


class LatticeValue:
    """Minimal ``LatticeValue``: values per site plus the lattice they live on."""

    def __init__(self, lattice, values):
        self.lattice, self.values = lattice, np.asarray(values)

    def __array__(self, dtype=None, copy=None):
        return self.values if dtype is None else self.values.astype(dtype)

    def __len__(self):
        return len(self.values)

    def __getitem__(self, i):
        return self.values[i]


def _device_state(state, ctx=None):
    if isinstance(state, DeviceState):
        return state
    return DeviceState.from_any(state, ctx)


def localdensity(state, n_int=None):
    """rho_i = sum_alpha Re P[i', i'] -> LatticeValue (all-reduced over ranks)."""
    ds = _device_state(state)
    n = n_int or ds.n_int or 1
    if ds.N % n:
        raise _lib.ArgumentError("State must be defined on a lattice")
    rho = np.zeros(ds.N // n)
    _lib.check(_lib.load().lm_local_density(ds.handle, n, _lib.ptr(rho)))
    return LatticeValue(ds.lattice, rho)


def localexpect(op, state, n_int=None):
    """``localexpect(op, state)``: expectation of the on-site operator ``op`` (n_int x n_int) on
    every site -> LatticeValue (complex in general, like the reference's un-real()ed sum)."""
    ds = _device_state(state)
    o = np.asarray(op, complex)
    n = n_int or ds.n_int or o.shape[0]
    if n < 2:
        raise _lib.ArgumentError("Cannot compute local expectation value for a state without internal degrees of freedom")
    if o.shape != (n, n):
        raise _lib.ArgumentError("Operator must be defined on the internal basis of the state")
    out = np.zeros(ds.N // n, complex)
    o_cm = np.asfortranarray(o)
    _lib.check(_lib.load().lm_local_expect(ds.handle, n, _lib.ptr(o_cm), _lib.ptr(out)))
    return LatticeValue(ds.lattice, out)


def _device_ham(ham, ctx, n_int):
    if isinstance(ham, Hamiltonian):
        return ham.device(ctx)
    if isinstance(ham, DeviceHam):
        return ham
    return DeviceHam.from_csc(ctx, ham, n_int or 1)


class DensityCurrents:
    """Lazy density currents ``DensityCurrents(hamiltonian, state)``."""

    def __init__(self, hamiltonian, state, n_int=None):
        self.state = _device_state(state)
        n = n_int or getattr(hamiltonian, "n_int", None) or self.state.n_int or 1
        self._ham_src = hamiltonian
        self.n_int = n
        dim = hamiltonian.structure.dim if isinstance(hamiltonian, Hamiltonian) else (
            hamiltonian.N if isinstance(hamiltonian, DeviceHam) else hamiltonian.shape[0])
        if dim != self.state.N:
            raise _lib.ArgumentError("Incompatible bases of the Hamiltonian and the state")
        self._dev = None
        self._values = None

    @property
    def lattice(self):
        return getattr(self._ham_src, "lattice", None) or self.state.lattice

    def _ham(self):
        if isinstance(self._ham_src, Hamiltonian):
            return self._ham_src.device(self.state.ctx)      # re-asserts this H's field
        if self._dev is None:
            self._dev = _device_ham(self._ham_src, self.state.ctx, self.n_int)
        return self._dev

    def pair_values(self):
        """(I, J, V) over H's site-level sparsity, I < J (1-based), unfiltered."""
        if self._values is None:
            dev = self._ham()
            I, J = dev.pairs()
            V = np.zeros(max(len(I), 1))
            _lib.check(_lib.load().lm_observables(dev.handle, self.state.handle, None, _lib.ptr(V)))
            self._values = (I, J, V[:len(I)])
        return self._values

    def __getitem__(self, ij):
        i, j = ij
        out = np.zeros(1)
        I = np.array([i - 1], np.int32)
        J = np.array([j - 1], np.int32)
        dev = self._ham()
        # lm_bond_currents takes indices in the base the Hamiltonian was created with (0 here)
        _lib.check(_lib.load().lm_bond_currents(dev.handle, self.state.handle, 1, _lib.ptr(I), _lib.ptr(J), _lib.ptr(out)))
        return float(out[0])

    def __len__(self):
        ns = self.state.N // self.n_int
        return ns * (ns - 1) // 2


class LocalOperatorCurrents(DensityCurrents):
    """``LocalOperatorCurrents(hamiltonian, state, op)`` (e.g. spin currents)."""

    def __init__(self, hamiltonian, state, op, n_int=None):
        super().__init__(hamiltonian, state, n_int)
        o = np.asarray(op, complex)
        if self.n_int < 2:
            raise _lib.ArgumentError("System expected to have internal degrees of freedom")
        if o.shape != (self.n_int, self.n_int):
            raise _lib.ArgumentError("Operator must be defined on the internal basis of the Hamiltonian.")
        self.op = np.asfortranarray(o)

    def pair_values(self):
        if self._values is None:
            dev = self._ham()
            I, J = dev.pairs()
            V = np.zeros(max(len(I), 1))
            _lib.check(_lib.load().lm_operator_currents(dev.handle, self.state.handle, _lib.ptr(self.op), _lib.ptr(V)))
            self._values = (I, J, V[:len(I)])
        return self._values

    def __getitem__(self, ij):
        i, j = ij
        I, J, V = self.pair_values()
        if i == j:
            return 0.0
        a, b, sgn = (i, j, 1.0) if i < j else (j, i, -1.0)
        hit = np.nonzero((I == a) & (J == b))[0]
        return sgn * float(V[hit[0]]) if len(hit) else 0.0


def findnz(curr):
    """(Is, Js, Vs) with Is < Js, |V| >= 1e-10, ordered like findnz of a CSC matrix."""
    if isinstance(curr, Currents):
        m = sp.triu(curr.currents, k=1).tocsc()
        m.sort_indices()
        coo = m.tocoo()
        order = np.lexsort((coo.row, coo.col))
        return coo.row[order] + 1, coo.col[order] + 1, coo.data[order]
    I, J, V = curr.pair_values()
    keep = np.abs(V) >= CURRENTS_EPS
    return I[keep], J[keep], V[keep]


class Currents:
    """Materialised antisymmetric site-current matrix ``Currents(lat, mat)``."""

    def __init__(self, curr, bonds=None, lattice=None):
        if isinstance(curr, DensityCurrents):       # incl. LocalOperatorCurrents
            I, J, V = curr.pair_values()
            if bonds is not None:
                # Currents(curr, bonds): only the listed (i, j) pairs (1-based), either order
                want = {(min(a, b), max(a, b)) for a, b in bonds}
                sel = np.array([(a, b) in want for a, b in zip(I.tolist(), J.tolist())], bool)
                I, J, V = I[sel], J[sel], V[sel]
            keep = np.abs(V) >= CURRENTS_EPS
            I, J, V = I[keep], J[keep], V[keep]
            ns = curr.state.N // curr.n_int
            self.lattice = curr.lattice
            self.currents = sp.coo_matrix((np.concatenate([V, -V]),
                                           (np.concatenate([I, J]) - 1, np.concatenate([J, I]) - 1)),
                                          shape=(ns, ns)).tocsc()
        else:
            self.lattice = lattice
            self.currents = sp.csc_matrix(curr)

    def __getitem__(self, ij):
        i, j = ij
        return float(self.currents[i - 1, j - 1])

    def __add__(self, o):
        return Currents(self.currents + o.currents, lattice=self.lattice)

    def __sub__(self, o):
        return Currents(self.currents - o.currents, lattice=self.lattice)

    def __mul__(self, k):
        return Currents(self.currents * k, lattice=self.lattice)

    __rmul__ = __mul__

    def __truediv__(self, k):
        return Currents(self.currents / k, lattice=self.lattice)

    def toarray(self):
        return self.currents.toarray()


def _site_mask(ns, region):
    """Region given as 1-based site indices or a boolean per-site mask -> uint8 mask."""
    r = np.asarray(region)
    if r.dtype == bool:
        if r.shape != (ns,):
            raise _lib.ArgumentError("region mask must have one entry per site")
        return np.ascontiguousarray(r, np.uint8)
    m = np.zeros(ns, np.uint8)
    m[np.atleast_1d(r.astype(int)) - 1] = 1
    return m


def currentsfrom(curr, src):
    """LatticeValue of the currents from region ``src`` (1-based site indices) to every other site."""
    if type(curr) is DensityCurrents:        # summed on the device: one LatticeValue crosses PCIe
        dev = curr._ham()
        ns = curr.state.N // curr.n_int
        out = np.zeros(ns)
        _lib.check(_lib.load().lm_currents_from(dev.handle, curr.state.handle, _lib.ptr(_site_mask(ns, src)), 0, _lib.ptr(out)))
        return LatticeValue(curr.lattice, out)
    c = curr if isinstance(curr, Currents) else Currents(curr)
    src = np.atleast_1d(np.asarray(src, int)) - 1
    m = c.currents.tocsr()
    out = np.asarray(m[src, :].sum(axis=0)).ravel()
    out[src] = 0
    return LatticeValue(c.lattice, out)


def currentsfromto(curr, src, dst=None):
    """Total current from region ``src`` to region ``dst`` (default: every other site),
    src/currents.jl:103-109.  Device currents are summed on the device (one double comes back)."""
    if type(curr) is DensityCurrents:
        dev = curr._ham()
        ns = curr.state.N // curr.n_int
        out = C.c_double()
        _lib.check(_lib.load().lm_currents_fromto(dev.handle, curr.state.handle, _lib.ptr(_site_mask(ns, src)),
                                                  None if dst is None else _lib.ptr(_site_mask(ns, dst)), 0, C.byref(out)))
        return float(out.value)
    c = curr if isinstance(curr, Currents) else Currents(curr)
    ns = c.currents.shape[0]
    src = np.atleast_1d(np.asarray(src, int)) - 1
    dst = np.setdiff1d(np.arange(ns), src) if dst is None else np.atleast_1d(np.asarray(dst, int)) - 1
    return float(c.currents.tocsr()[src, :][:, dst].sum())
