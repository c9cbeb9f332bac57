"""Host mirror of the reference's Hamiltonian constructors, producing DEVICE-resident
operators.

Same names / argument meaning as the reference (paths relative to the reference repository):
  construct_hamiltonian / construct_operator   src/operators/constructoperator.jl:136-168
  tightbinding_hamiltonian                     src/operators/constructoperator.jl:170-194
  qwz, haldane                                 src/zoo/models.jl:130-137,139-170
  Hamiltonian wrapper                          src/operators/system.jl:380-391
A ``Hamiltonian`` here is lazy: it holds the field-independent *structure* (directed bond
table, on-site blocks - built once per (lattice, terms) and cached on the lattice) plus a
gauge field.  ``t -> tightbinding_hamiltonian(l, field=LandauGauge(B(t)))`` therefore costs a
dictionary lookup per step on the host; the Peierls phases are regenerated on the device
(``lm_ham_set_field_params``).  ``.data`` assembles the scipy CSC matrix on demand (what the
reference's ``H.data`` is) for users who want the matrix itself.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib
from .fields import AbstractField, NoField
from .lattices import BravaisLattice, BravaisTranslation, NearestNeighbor


def _op_matrix(op, n):
    if np.isscalar(op):
        return complex(op) * np.eye(n, dtype=complex)
    m = np.asarray(op, dtype=complex)
    if m.shape != (n, n):
        raise _lib.ArgumentError("matrix size does not match on-site dims")
    return m


def _sig(x):
    if isinstance(x, BravaisTranslation):
        return ("tr",) + x.key()
    if isinstance(x, NearestNeighbor):
        return ("nn", x.n)
    if isinstance(x, (list, tuple)) and x and isinstance(x[0], BravaisTranslation):
        return ("trs",) + tuple(t.key() for t in x)
    a = np.asarray(x)
    return ("arr", a.shape, a.dtype.str, a.tobytes())


def _row_block(lat, n_int):
    """Rows of one Bravais unit cell (basis index innermost) when the lattice is unfiltered."""
    import os
    if os.environ.get("LM_ROW_BLOCK", "1") == "0":
        return None
    nb = getattr(lat, "nb", 1)
    if nb > 1 and len(lat) == lat.sizes[0] * lat.sizes[1] * nb:
        return nb * n_int
    return None


class Structure:
    """Field-independent part of a lattice Hamiltonian (OperatorBuilder restated as tables,
    src/operators/builder.jl:282-309, src/operators/constructoperator.jl:4-47)."""

    def __init__(self, lat: BravaisLattice, n_int: int, terms):
        n = n_int
        self.lat, self.n_int = lat, n
        self.n_sites = len(lat)
        src, dst, rs, rd, bf, amp = [], [], [], [], [], []
        onsite = None
        for op, what in terms:
            if np.isscalar(op) and op == 0:          # add_pair_terms!: zero terms are skipped
                continue
            B = _op_matrix(op, n)
            if isinstance(what, NearestNeighbor):
                what = lat.nearest_neighbor(what.n)
            if isinstance(what, BravaisTranslation):
                what = [what]
            if isinstance(what, (list, tuple)) and what and isinstance(what[0], BravaisTranslation):
                for tr in what:
                    t = lat.bonds(tr)
                    src.append(t.src), dst.append(t.dst), rs.append(t.r_src), rd.append(t.r_dst)
                    bf.append(t.bfac), amp.append(np.broadcast_to(B, (len(t), n, n)))
            else:
                vals = np.full(self.n_sites, what, dtype=complex) if np.isscalar(what) else np.asarray(what, dtype=complex)
                if vals.shape != (self.n_sites,):
                    raise _lib.ArgumentError("cannot interpret %r as on-lattice operator" % type(what))
                if onsite is None:
                    onsite = np.zeros((self.n_sites, n, n), complex)
                onsite += vals[:, None, None] * B[None]
        cat = lambda xs, shape, dt: (np.ascontiguousarray(np.concatenate(xs)) if xs else np.zeros(shape, dt))
        self.src = cat(src, (0,), np.int32).astype(np.int32)
        self.dst = cat(dst, (0,), np.int32).astype(np.int32)
        self.r_src = cat(rs, (0, 2), float)
        self.r_dst = cat(rd, (0, 2), float)
        self.bfac = cat(bf, (0,), complex)
        self.amp = cat(amp, (0, n, n), complex)           # amp[q, a, b]
        self.onsite = onsite
        self._dev = {}

    @property
    def dim(self):
        return self.n_sites * self.n_int

    def assemble(self, field: AbstractField):
        """scipy CSC of H for a given field (vectorised host assembly = what the reference
        does per call of t -> H(t), src/operators/builder.jl:296-309)."""
        n, N = self.n_int, self.dim
        rows, cols, vals = [], [], []
        if len(self.src):
            f = self.bfac * np.exp(-2j * np.pi * field.line_integral(self.r_src, self.r_dst))
            for a in range(n):
                for b in range(n):
                    v = self.amp[:, a, b]
                    nzm = v != 0
                    if not nzm.any():
                        continue
                    i, j = self.src[nzm].astype(np.int64) * n + a, self.dst[nzm].astype(np.int64) * n + b
                    rows.append(i), cols.append(j), vals.append(v[nzm] * f[nzm])
                    off = self.src[nzm] != self.dst[nzm]
                    rows.append(j[off]), cols.append(i[off]), vals.append(np.conj(v[nzm] * f[nzm])[off])
        if self.onsite is not None:
            for a in range(n):
                for b in range(n):
                    v = self.onsite[:, a, b]
                    nzm = v != 0
                    idx = np.nonzero(nzm)[0].astype(np.int64)
                    rows.append(idx * n + a), cols.append(idx * n + b), vals.append(v[nzm])
        if not rows:
            return sp.csc_matrix((N, N), dtype=complex)
        m = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N)).tocsc()
        m.sort_indices()
        return m

    def device(self, ctx):
        """The (single) device operator of this structure on context ``ctx``."""
        d = self._dev.get(id(ctx))
        if d is None:
            d = DeviceHam.from_structure(ctx, self)
            self._dev[id(ctx)] = d
        return d


class DeviceHam:
    """Owner of an ``lm_ham`` handle."""

    def __init__(self, ctx, handle, n_int, bond_mode):
        self.ctx, self.handle, self.n_int, self.bond_mode = ctx, handle, n_int, bond_mode
        self.field_key = None
        self._pairs = None
        N, ni, nnz, W = C.c_int64(), C.c_int32(), C.c_int64(), C.c_int32()
        _lib.check(_lib.load().lm_ham_dims(handle, C.byref(N), C.byref(ni), C.byref(nnz), C.byref(W)))
        self.N, self.nnz, self.W = N.value, nnz.value, W.value

    @classmethod
    def from_structure(cls, ctx, st: Structure):
        lib = _lib.load()
        h = C.c_void_p()
        # blocks are handed over column-major (Julia order): amp_cm[q, b, a] = amp[q, a, b]
        amp_cm = np.ascontiguousarray(np.transpose(st.amp, (0, 2, 1)))
        ons_cm = None if st.onsite is None else np.ascontiguousarray(np.transpose(st.onsite, (0, 2, 1)))
        _lib.check(lib.lm_ham_create_bonds(
            ctx.handle, st.n_sites, st.n_int, len(st.src), _lib.ptr(st.src), _lib.ptr(st.dst),
            _lib.ptr(st.r_src), _lib.ptr(st.r_dst), _lib.ptr(amp_cm), _lib.ptr(st.bfac),
            _lib.ptr(ons_cm), 0, C.byref(h)))
        d = cls(ctx, h, st.n_int, True)
        d.field_key = ((), b"")
        d.set_site_coords(st.lat.coords, row_block=_row_block(st.lat, st.n_int))
        if len(st.lat) == st.lat.sizes[0] * st.lat.sizes[1] * getattr(st.lat, "nb", 1):
            d.set_lattice_dims(*st.lat.sizes)       # unfiltered Bravais lattice: rows are cell-major
        return d

    def set_lattice_dims(self, n1, n2):
        """Declare the cell-major row order of an unfiltered n1 x n2 Bravais lattice (enables the
        register-tiled stencil kernel when the pattern is a compiled |d| <= 1 stencil)."""
        _lib.check(_lib.load().lm_ham_set_lattice_dims(self.handle, int(n1), int(n2)))

    def set_site_coords(self, coords, row_block=None):
        xy = np.ascontiguousarray(np.asarray(coords, float)[:, :2])
        if xy.shape != (self.N // self.n_int, 2):
            raise _lib.ArgumentError("site coordinates must be (n_sites, 2)")
        if row_block:
            _lib.check(_lib.load().lm_ham_set_row_block(self.handle, int(row_block)))
        _lib.check(_lib.load().lm_ham_set_site_coords(self.handle, _lib.ptr(xy)))

    @classmethod
    def from_csc(cls, ctx, mat, n_int=1, coords=None, lattice_dims=None):
        lib = _lib.load()
        m = sp.csc_matrix(mat)
        m.sort_indices()
        colptr = np.ascontiguousarray(m.indptr, np.int64)
        rowval = np.ascontiguousarray(m.indices, np.int64)
        nz = np.ascontiguousarray(m.data, _lib.cdtype(ctx.precision))
        h = C.c_void_p()
        _lib.check(lib.lm_ham_create_csc(ctx.handle, m.shape[0], n_int, _lib.ptr(colptr), _lib.ptr(rowval),
                                         _lib.ptr(nz), 0, C.byref(h)))
        d = cls(ctx, h, n_int, False)
        d.pattern = (colptr, rowval)
        if coords is not None:
            d.set_site_coords(coords)
        if lattice_dims is not None:
            d.set_lattice_dims(*lattice_dims)
        return d

    def update_values(self, nzval):
        nz = np.ascontiguousarray(nzval, _lib.cdtype(self.ctx.precision))
        if nz.shape != (self.nnz,):
            raise _lib.ArgumentError("update_values: expected %d stored entries" % self.nnz)
        _lib.check(_lib.load().lm_ham_update_values(self.handle, _lib.ptr(nz)))

    def ensure_field(self, field: AbstractField):
        kinds, params = field.descriptor()
        key = (tuple(kinds.tolist()), params.tobytes())
        if key == self.field_key:
            return
        lib = _lib.load()
        if self.field_key is not None and key[0] == self.field_key[0]:
            _lib.check(lib.lm_ham_set_field_params(self.handle, _lib.ptr(params)))
        else:
            _lib.check(lib.lm_ham_set_fields(self.handle, len(kinds), _lib.ptr(kinds), _lib.ptr(params)))
        self.field_key = key

    def to_csc(self):
        lib = _lib.load()
        colptr = np.zeros(self.N + 1, np.int64)
        rowval = np.zeros(max(self.nnz, 1), np.int64)
        nz = np.zeros(max(self.nnz, 1), _lib.cdtype(self.ctx.precision))
        _lib.check(lib.lm_ham_get_csc(self.handle, _lib.ptr(colptr), _lib.ptr(rowval), _lib.ptr(nz)))
        return sp.csc_matrix((nz[:self.nnz], rowval[:self.nnz], colptr), shape=(self.N, self.N))

    def refine_bounds(self, iters=60, margin=0.05):
        """Opt-in Lanczos tightening of the spectral enclosure (not rigorous; see lm_b200.h)."""
        _lib.check(_lib.load().lm_ham_refine_bounds(self.handle, int(iters), float(margin)))
        self._refined = True
        return self.spectral_bounds()

    def spectral_bounds(self):
        lo, hi = C.c_double(), C.c_double()
        _lib.check(_lib.load().lm_ham_spectral_bounds(self.handle, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def pairs(self):
        """Site pairs (I < J, 1-based) in the order the device returns their currents."""
        if self._pairs is None:
            n = C.c_int64()
            _lib.check(_lib.load().lm_currents_npairs(self.handle, C.byref(n)))
            I = np.zeros(max(n.value, 1), np.int32)
            J = np.zeros(max(n.value, 1), np.int32)
            _lib.check(_lib.load().lm_currents_pairs(self.handle, _lib.ptr(I), _lib.ptr(J)))
            self._pairs = (I[:n.value] + 1, J[:n.value] + 1)
        return self._pairs

    def __del__(self):
        try:
            if self.handle:
                _lib.load().lm_ham_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class Hamiltonian:
    """``Hamiltonian(sys, op)`` of the reference (src/operators/system.jl:380-391), lazy."""

    def __init__(self, structure: Structure, field: AbstractField):
        self.structure = structure
        self.field = field
        self._data = None

    @property
    def lattice(self):
        return self.structure.lat

    @property
    def n_int(self):
        return self.structure.n_int

    @property
    def data(self):
        if self._data is None:
            self._data = self.structure.assemble(self.field)
        return self._data

    def dense(self):
        return self.data.toarray()

    def device(self, ctx):
        d = self.structure.device(ctx)
        d.ensure_field(self.field)
        return d

    def __repr__(self):
        return "Hamiltonian(dim=%dx%d) on %r" % (self.structure.dim, self.structure.dim, self.lattice)


def construct_hamiltonian(lat: BravaisLattice, *args, field=None, **kw):
    """construct_hamiltonian(lat[, internal_dim], terms...; field).  Terms: ``(op, on_lattice)``
    pairs with ``op`` a number or n_int x n_int matrix and ``on_lattice`` a number, a per-site
    array (LatticeValue), a BravaisTranslation, a NearestNeighbor or a tuple of translations;
    a bare per-site array is the on-site term ``1 => values``; a bare matrix is ``mat => 1``."""
    args = list(args)
    n_int = 1
    if args and isinstance(args[0], (int, np.integer)) and not isinstance(args[0], bool):
        n_int = int(args.pop(0))
    terms = []
    for a in args:
        if isinstance(a, tuple) and len(a) == 2 and not isinstance(a[0], BravaisTranslation):
            terms.append(a)
        else:
            arr = np.asarray(a)
            if arr.ndim == 2 and arr.shape == (n_int, n_int):
                terms.append((arr, 1))
            else:
                terms.append((1, a))
    field = (field or NoField()).adapt(lat)
    key = ("structure", n_int, tuple((_sig(op), _sig(w)) for op, w in terms))
    st = lat._cache.get(key)
    if st is None:
        st = Structure(lat, n_int, terms)
        lat._cache[key] = st
    return Hamiltonian(st, field)


construct_operator = construct_hamiltonian


def tightbinding_hamiltonian(lat: BravaisLattice, n_int: int = 1, t1=1, t2=0, t3=0, field=None):
    """src/operators/constructoperator.jl:190-194."""
    terms = [(t, NearestNeighbor(k)) for t, k in ((t1, 1), (t2, 2), (t3, 3)) if t != 0]
    return construct_hamiltonian(lat, n_int, *terms, field=field)


def qwz(lat: BravaisLattice, m=1, field=None):
    """src/zoo/models.jl:130-137."""
    if lat.kind != "SquareLattice":
        raise _lib.ArgumentError("Invalid lattice type %s; expected SquareLattice" % lat.kind)
    return construct_hamiltonian(
        lat, 2,
        (np.array([[1, 0], [0, -1]], complex), m),
        (np.array([[1, -1j], [-1j, -1]], complex) / 2, BravaisTranslation(axis=1)),
        (np.array([[1, -1], [1, -1]], complex) / 2, BravaisTranslation(axis=2)),
        field=field)


def haldane(lat: BravaisLattice, t1, t2, m=0, field=None):
    """src/zoo/models.jl:162-170."""
    from .lattices import honeycomb_2nn
    if lat.kind != "HoneycombLattice":
        raise _lib.ArgumentError("Invalid lattice type %s; expected HoneycombLattice" % lat.kind)
    ms = np.where(lat.b == 0, float(m), -float(m))
    return construct_hamiltonian(lat, 1, (1, ms), (t1, NearestNeighbor(1)), (1j * t2, honeycomb_2nn), field=field)


def kanemele(lat: BravaisLattice, t1, t2, field=None):
    """src/zoo/models.jl:188-194: spin-1/2 honeycomb model, `t1 => NearestNeighbor(1)`, `im t2 sigma_z => honeycomb_2nn`."""
    from .lattices import honeycomb_2nn
    if lat.kind != "HoneycombLattice":
        raise _lib.ArgumentError("Invalid lattice type %s; expected HoneycombLattice" % lat.kind)
    sz = np.array([[1, 0], [0, -1]], complex)
    return construct_hamiltonian(lat, 2, (t1, NearestNeighbor(1)), (1j * t2 * sz, honeycomb_2nn), field=field)
