"""Per-frame observables on device states.

Same names and semantics as the reference (paths relative to the reference repository):
  localdensity                 src/operators/latticeutils.jl:41-45
  DensityCurrents              src/zoo/currents.jl:81-102
  Currents(curr[, bonds])      src/currents.jl:223-255   (|J| < 1e-10 dropped, :4)
  findnz                       src/currents.jl:159-177
  currentsfrom / currentsfromto  src/currents.jl:85-109
  SubCurrents, curr[region]    src/currents.jl:48-66,154-157
  Currents arithmetic / iteration / setindex!  src/currents.jl:30-47,122-152,179-190
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import norm as spla_norm

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

    # value semantics for TimeSequence (copy, differentiate / integrate keep the lattice)
    def copy(self):
        return LatticeValue(self.lattice, self.values.copy())

    def _other(self, o):
        return o.values if isinstance(o, LatticeValue) else o

    def __add__(self, o):
        return LatticeValue(self.lattice, self.values + self._other(o))

    __radd__ = __add__

    def __sub__(self, o):
        return LatticeValue(self.lattice, self.values - self._other(o))

    def __mul__(self, k):
        return LatticeValue(self.lattice, self.values * self._other(k))

    __rmul__ = __mul__

    def __truediv__(self, k):
        return LatticeValue(self.lattice, self.values / self._other(k))

    def __eq__(self, o):
        return isinstance(o, LatticeValue) and (o.lattice is self.lattice or o.lattice == self.lattice) and np.array_equal(o.values, self.values)

    __hash__ = None


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
        if not isinstance(ij, tuple):                # curr[region]: lazy SubCurrents (src/currents.jl:63-66)
            return SubCurrents(self, ij)
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
        if not isinstance(ij, tuple):
            return SubCurrents(self, ij)
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


def _to_inds(ns, region):
    """0-based site indices of a region: a boolean per-site mask or 1-based site numbers (``to_inds``)."""
    r = np.asarray(region)
    if r.dtype == bool:
        if r.shape != (ns,):
            raise _lib.ArgumentError("region mask must have one entry per site")
        return np.flatnonzero(r)
    return np.atleast_1d(r.astype(int)) - 1


class SubLattice:
    """``lat[indices]``: the sites of ``parent`` picked by 0-based ``indices`` (kept in that order)."""

    def __init__(self, parent, indices):
        self.parent, self.indices = parent, np.asarray(indices, int)

    def __len__(self):
        return len(self.indices)

    @property
    def coords(self):
        return np.asarray(self.parent.coords)[self.indices]

    def __eq__(self, o):
        return isinstance(o, SubLattice) and (o.parent is self.parent or o.parent == self.parent) and np.array_equal(o.indices, self.indices)

    __hash__ = None


def _sublattice(lat, inds):
    if isinstance(lat, SubLattice):
        return SubLattice(lat.parent, lat.indices[inds])
    return SubLattice(lat, inds)


class SubCurrents:
    """Lazy view of a currents object on a subset of its sites (``SubCurrents``,
    src/currents.jl:48-66): ``sub[i, j] = parent[indices[i], indices[j]]`` (1-based)."""

    def __init__(self, parent, region):
        ns = parent.state.N // parent.n_int if isinstance(parent, DensityCurrents) else len(parent.lattice)
        self.parent, self._nparent = parent, ns
        self.indices = _to_inds(ns, region)
        self.lattice = _sublattice(parent.lattice, self.indices)

    def __getitem__(self, ij):
        if not isinstance(ij, tuple):
            return SubCurrents(self, ij)
        i, j = ij
        return self.parent[int(self.indices[i - 1]) + 1, int(self.indices[j - 1]) + 1]

    def __len__(self):
        n = len(self.indices)
        return n * (n - 1) // 2

    def pair_values(self):
        """The parent's (I, J, V) restricted to pairs inside the subset, renumbered (I < J, 1-based)."""
        I, J, V = self.parent.pair_values()
        pos = np.zeros(self._nparent + 1, int)
        pos[self.indices + 1] = np.arange(1, len(self.indices) + 1)
        a, b = pos[I], pos[J]
        keep = (a > 0) & (b > 0)
        a, b, v = a[keep], b[keep], V[keep]
        flip = a > b
        return np.where(flip, b, a), np.where(flip, a, b), np.where(flip, -v, v)


class Currents:
    """Materialised antisymmetric site-current matrix ``Currents(lat, mat)`` (src/currents.jl:110-190).

    ``Currents(curr[, bonds])`` materialises a lazy currents object over H's own sparsity, dropping
    ``|J| < 1e-10`` (:223-255); ``Currents(lattice)`` is the empty matrix on a lattice (:122);
    ``Currents(matrix, lattice=...)`` wraps a matrix."""

    def __init__(self, curr, bonds=None, lattice=None):
        if isinstance(curr, (DensityCurrents, SubCurrents)):       # incl. LocalOperatorCurrents
            I, J, V = curr.pair_values()
            if bonds is not None:
                # Currents(curr, bonds): only the listed (i, j) pairs (1-based, numbered on the lattice
                # the bonds were made for - the parent's for a SubCurrents), either order
                want = {(min(a, b), max(a, b)) for a, b in bonds}
                if isinstance(curr, SubCurrents):
                    root, idx = curr, np.arange(len(curr.indices))
                    while isinstance(root, SubCurrents):
                        idx = root.indices[idx]
                        root = root.parent
                    gi, gj = idx[I - 1] + 1, idx[J - 1] + 1
                else:
                    gi, gj = I, J
                sel = np.array([(min(a, b), max(a, b)) in want for a, b in zip(gi.tolist(), gj.tolist())], bool)
                I, J, V = I[sel], J[sel], V[sel]
            keep = np.abs(V) >= CURRENTS_EPS
            I, J, V = I[keep], J[keep], V[keep]
            ns = len(curr.indices) if isinstance(curr, SubCurrents) else curr.state.N // curr.n_int
            self.lattice = curr.lattice
            self.currents = sp.coo_matrix((np.concatenate([V, -V]),
                                           (np.concatenate([I, J]) - 1, np.concatenate([J, I]) - 1)),
                                          shape=(ns, ns)).tocsc()
        elif hasattr(curr, "coords") and not hasattr(curr, "shape"):   # Currents(lattice)
            self.lattice = curr
            self.currents = sp.csc_matrix((len(curr), len(curr)))
        else:
            self.lattice = lattice
            self.currents = sp.csc_matrix(curr)
            if self.currents.shape[0] != self.currents.shape[1]:
                raise _lib.ArgumentError("currents matrix must be square")
            if lattice is not None and len(lattice) != self.currents.shape[0]:
                raise _lib.ArgumentError("currents matrix size does not match the lattice")

    def __getitem__(self, ij):
        if not isinstance(ij, tuple):                # Currents on the sub-lattice (:154-157)
            inds = _to_inds(self.currents.shape[0], ij)
            lat = _sublattice(self.lattice, inds) if self.lattice is not None else None
            return Currents(self.currents[inds, :][:, inds], lattice=lat)
        i, j = ij
        return float(self.currents[i - 1, j - 1])

    def __setitem__(self, ij, rhs):
        """``curr[site1, site2] = rhs`` also stores ``-rhs`` at ``[site2, site1]``; a nonzero current
        from a site to itself is refused with a warning (:138-152)."""
        i, j = ij
        if i == j and abs(rhs) > CURRENTS_EPS:
            import warnings
            warnings.warn("Attempt to assign nonzero current from site to self")
            return
        m = self.currents.tolil()
        m[i - 1, j - 1] = rhs
        m[j - 1, i - 1] = -rhs
        self.currents = m.tocsc()

    def __len__(self):
        n = self.currents.shape[0]
        return n * (n - 1) // 2

    def __iter__(self):
        """Every pair once, oriented along the flow: ``((a, b), c)`` with ``c >= 0`` (:36-47; sites as
        1-based indices)."""
        dense = self.currents.toarray()
        n = dense.shape[0]
        for i in range(n):
            for j in range(i + 1, n):
                c = dense[i, j]
                yield ((i + 1, j + 1), c) if c > 0 else ((j + 1, i + 1), -c)

    def _same_sites(self, o):
        if self.currents.shape != o.currents.shape:
            return False
        if self.lattice is None or o.lattice is None or self.lattice is o.lattice:
            return True
        eq = self.lattice == o.lattice
        return bool(eq) if isinstance(eq, (bool, np.bool_)) else self.lattice is o.lattice

    def __eq__(self, o):
        if not isinstance(o, Currents):
            return NotImplemented
        return self._same_sites(o) and (self.currents != o.currents).nnz == 0

    __hash__ = None

    def isapprox(self, o, rtol=None, atol=0.0):
        """``isapprox`` of the matrices in the Frobenius norm, Julia's default ``rtol = sqrt(eps)`` (:128-129)."""
        if not self._same_sites(o):
            return False
        rtol = math.sqrt(np.finfo(float).eps) if rtol is None else rtol
        na, nb = spla_norm(self.currents), spla_norm(o.currents)
        return spla_norm(self.currents - o.currents) <= max(atol, rtol * max(na, nb))

    def copy(self):
        return Currents(self.currents.copy(), lattice=self.lattice)

    def zero(self):
        return Currents(sp.csc_matrix(self.currents.shape), lattice=self.lattice)

    def _check(self, o):
        if not self._same_sites(o):
            raise _lib.ArgumentError("currents are defined on different sites")

    def __add__(self, o):
        self._check(o)
        return Currents(self.currents + o.currents, lattice=self.lattice)

    def __sub__(self, o):
        self._check(o)
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
    src = _to_inds(c.currents.shape[0], src)            # 1-based indices or a boolean mask (to_inds)
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
    src = _to_inds(ns, src)
    dst = np.setdiff1d(np.arange(ns), src) if dst is None else _to_inds(ns, dst)
    return float(c.currents.tocsr()[src, :][:, dst].sum())
