"""Host-side geometry for the device backend: Bravais lattices and directed bond tables.

This is the (vectorised numpy) producer of what the Julia host would hand to
``lm_ham_create_bonds``: site order, coordinates, (src, dst, r_src, r_dst_unwrapped, boundary
factor) per directed bond.  Conventions are the reference's (paths relative to the reference
repository):
  site order / coordinates   src/lattices/bravais/lattice.jl:101-111, unitcell.jl:119-122
  unit cells                 src/zoo/lattices.jl:131 (SquareLattice), :161 (HoneycombLattice)
  BravaisTranslation         src/lattices/bravais/bonds.jl:10-98
  bond iteration             src/core/bonds.jl:415-422
  boundaries / twist factor  src/core/boundaries.jl:156-165,251-289
  nearest-neighbour hops     src/lattices/bravais/nearestneighbor.jl:97-131,174-178
"""
from __future__ import annotations

import itertools
import math

import numpy as np


class BravaisTranslation:
    """``BravaisTranslation([a => b, ]uc)``; ``site_indices == (0, 0)`` keeps the sublattice."""

    def __init__(self, site_indices=(0, 0), uc=(0, 0), axis=0, dist=1):
        if axis:
            uc = [0] * max(axis, len(uc))
            uc[axis - 1] = dist
        self.site_indices = tuple(int(v) for v in site_indices)
        self.uc = tuple(int(v) for v in uc)
        if all(v == 0 for v in self.uc) and self.site_indices[0] == self.site_indices[1]:
            raise ValueError("bond connects site to itself")

    def key(self):
        return (self.site_indices, self.uc)

    def __eq__(self, other):
        return isinstance(other, BravaisTranslation) and self.key() == other.key()

    def __hash__(self):
        return hash(self.key())

    def __repr__(self):
        if self.site_indices == (0, 0):
            return "Bravais%s" % (list(self.uc),)
        return "%d => %d, %s" % (self.site_indices[0], self.site_indices[1], list(self.uc))


def Bravais(*uc):
    return BravaisTranslation((0, 0), uc)


class NearestNeighbor:
    def __init__(self, n=1):
        self.n = int(n)


class BondTable:
    """Directed bonds of one translation: 0-based site indices, coordinates of the source
    site and of the UNWRAPPED destination, and the boundary phase factor of the destination."""

    def __init__(self, src, dst, r_src, r_dst, bfac):
        self.src, self.dst, self.r_src, self.r_dst, self.bfac = src, dst, r_src, r_dst, bfac

    def __len__(self):
        return len(self.src)


class BravaisLattice:
    """Integer-sized 2-D Bravais lattice (optionally filtered by a predicate on coordinates)."""

    def __init__(self, translations, basis, sizes, boundaries=(), predicate=None, kind="Bravais"):
        self.a = np.asarray(translations, float)          # (2, 2) columns a1, a2
        self.basis = np.asarray(basis, float)             # (2, NB)
        self.sizes = tuple(int(s) for s in sizes)
        assert len(self.sizes) == 2, "the device backend covers 2-D lattices"
        self.nb = self.basis.shape[1]
        self.kind = kind
        n1, n2 = self.sizes
        j1, j2, b = np.meshgrid(np.arange(1, n1 + 1), np.arange(1, n2 + 1), np.arange(self.nb), indexing="ij")
        j1, j2, b = j1.ravel(), j2.ravel(), b.ravel()      # j2 fast, basis innermost
        coords = (self.basis[:, b] + self.a[:, [0]] * j1 + self.a[:, [1]] * j2).T
        keep = np.ones(len(j1), bool) if predicate is None else np.asarray(predicate(coords[:, 0], coords[:, 1]), bool)
        self.j1, self.j2, self.b = j1[keep], j2[keep], b[keep]
        self.coords = np.ascontiguousarray(coords[keep])
        self.imap = -np.ones((n1, n2, self.nb), np.int64)   # (j1-1, j2-1, b) -> site index
        self.imap[self.j1 - 1, self.j2 - 1, self.b] = np.arange(len(self.j1))
        # boundaries: list of (uc translation, theta); depth 1 (src/core/boundaries.jl:120-126)
        self.boundaries = [(tuple(int(v) for v in tr), float(th)) for tr, th in boundaries]
        self._cache = {}

    def __len__(self):
        return len(self.j1)

    @property
    def x(self):
        return self.coords[:, 0]

    @property
    def y(self):
        return self.coords[:, 1]

    def coordvalues(self):
        return self.coords[:, 0].copy(), self.coords[:, 1].copy()

    def site_index(self, j1, j2, b=1):
        """1-based lattice coordinates / basis index -> 1-based site index (or None)."""
        n1, n2 = self.sizes
        if not (1 <= j1 <= n1 and 1 <= j2 <= n2 and 1 <= b <= self.nb):
            return None
        i = int(self.imap[j1 - 1, j2 - 1, b - 1])
        return None if i < 0 else i + 1

    def _lookup(self, j1, j2, b):
        n1, n2 = self.sizes
        inside = (j1 >= 1) & (j1 <= n1) & (j2 >= 1) & (j2 <= n2)
        idx = -np.ones(len(j1), np.int64)
        idx[inside] = self.imap[j1[inside] - 1, j2[inside] - 1, b[inside]]
        return idx

    def bonds(self, tr: BravaisTranslation) -> BondTable:
        key = ("bonds", tr.key())
        if key in self._cache:
            return self._cache[key]
        a, bnew = tr.site_indices
        uc = list(tr.uc) + [0] * (2 - len(tr.uc))
        if len(tr.uc) > 2 and any(v != 0 for v in tr.uc[2:]):
            sel = np.zeros(len(self), bool)
        elif (a, bnew) == (0, 0):
            sel = np.ones(len(self), bool)
        else:
            sel = self.b == (a - 1)
        src = np.nonzero(sel)[0]
        d1, d2 = self.j1[src] + uc[0], self.j2[src] + uc[1]
        db = self.b[src] if (a, bnew) == (0, 0) else np.full(len(src), bnew - 1)
        r_dst = (self.basis[:, db] + self.a[:, [0]] * d1 + self.a[:, [1]] * d2).T
        dst = self._lookup(d1, d2, db)
        bfac = np.ones(len(src), complex)
        if self.boundaries:
            nbd = len(self.boundaries)
            rng = range(-1, 2)
            # CartesianIndices order: FIRST boundary index fastest (src/core/utils.jl:24-26)
            for tup_rev in itertools.product(*([rng] * nbd)):
                tup = tuple(reversed(tup_rev))
                if all(t == 0 for t in tup):
                    continue
                todo = dst < 0
                if not todo.any():
                    break
                e1, e2 = d1[todo].copy(), d2[todo].copy()
                for (trb, _), n in zip(self.boundaries, tup):   # nshifts: site - n * tr
                    e1 -= n * trb[0]
                    e2 -= n * trb[1]
                cand = self._lookup(e1, e2, db[todo])
                hit = cand >= 0
                where = np.nonzero(todo)[0][hit]
                dst[where] = cand[hit]
                bfac[where] = np.exp(1j * sum(th * n for (_, th), n in zip(self.boundaries, tup)))
        ok = dst >= 0
        table = BondTable(src[ok].astype(np.int32), dst[ok].astype(np.int32),
                          np.ascontiguousarray(self.coords[src[ok]]), np.ascontiguousarray(r_dst[ok]),
                          np.ascontiguousarray(bfac[ok]))
        self._cache[key] = table
        return table

    # ---- nearest neighbours -----------------------------------------------------------
    def nnhops(self, depth=2, limit=3):
        key = ("nn", depth, limit)
        if key in self._cache:
            return self._cache[key]
        cands = []          # (length, order, translation)
        order = 0
        rng = range(-depth, depth + 1)
        for c_rev in itertools.product(rng, rng):
            cvec = (c_rev[1], c_rev[0])                 # first component fastest
            shift = self.a @ np.array(cvec, float)
            for i in range(self.nb):
                for j in range(i, self.nb):
                    if i == j:
                        if i > 1:
                            continue
                        nz = [v for v in cvec if v != 0]
                        if not nz or nz[0] < 1:         # self hop / double counting
                            continue
                        t = BravaisTranslation((0, 0), cvec)
                    else:
                        t = BravaisTranslation((i + 1, j + 1), cvec)
                    r = float(np.linalg.norm(shift - self.basis[:, i] + self.basis[:, j]))
                    cands.append((r, order, t))
                    order += 1
        # group by length (rtol sqrt(eps)); within a shell keep discovery order, drop duplicates.
        # A shell only exists if it was among the `limit` shortest when FIRST met, which for
        # the lattices here equals the `limit` shortest shells overall.
        lens = sorted({round(r, 9) for r, _, _ in cands})[:limit]
        shells = [[] for _ in lens]
        for r, _, t in sorted(cands, key=lambda c: c[1]):
            for k, l in enumerate(lens):
                if abs(r - l) <= 1.5e-8 * max(r, l):
                    if t not in shells[k]:
                        shells[k].append(t)
        self._cache[key] = (lens, shells)
        return lens, shells

    def nearest_neighbor(self, n=1):
        lens, shells = self.nnhops()
        if n <= min(3, len(lens)):
            return shells[n - 1]
        lens, shells = self.nnhops(1 + math.ceil(math.sqrt(n)), n)
        return shells[n - 1]

    def __repr__(self):
        return "%d-site %s" % (len(self), self.kind)


def _axis_boundaries(sizes, boundaries):
    """``boundaries``: iterable of ``("axis1", True | theta)`` / ``(uc_vector, True | theta)``."""
    out = []
    for what, val in (boundaries or ()):
        if val is False or val is None:
            continue
        theta = 0.0 if val is True else float(val)
        if isinstance(what, str):
            ax = int(what.replace("axis", "").replace(":", ""))
            tr = [0] * len(sizes)
            tr[ax - 1] = sizes[ax - 1]
        else:
            tr = list(what)
        out.append((tuple(tr), theta))
    return out


def SquareLattice(n1, n2, boundaries=(), predicate=None):
    return BravaisLattice(np.eye(2), np.zeros((2, 1)), (n1, n2),
                          _axis_boundaries((n1, n2), boundaries), predicate, "SquareLattice")


def HoneycombLattice(n1, n2, boundaries=(), predicate=None):
    return BravaisLattice(np.array([[1.0, 0.5], [0.0, math.sqrt(3) / 2]]),
                          np.array([[0.0, 0.5], [0.0, math.sqrt(3) / 6]]), (n1, n2),
                          _axis_boundaries((n1, n2), boundaries), predicate, "HoneycombLattice")


def TriangularLattice(n1, n2, boundaries=(), predicate=None):
    """src/zoo/lattices.jl:145."""
    return BravaisLattice(np.array([[1.0, 0.5], [0.0, math.sqrt(3) / 2]]), np.zeros((2, 1)), (n1, n2),
                          _axis_boundaries((n1, n2), boundaries), predicate, "TriangularLattice")


def KagomeLattice(n1, n2, boundaries=(), predicate=None):
    """src/zoo/lattices.jl:209: three sites at (0, 0), (1/2, 0), (1/4, sqrt(3)/4) per unit cell."""
    return BravaisLattice(np.array([[1.0, 0.5], [0.0, math.sqrt(3) / 2]]),
                          np.array([[0.0, 0.5, 0.25], [0.0, 0.0, math.sqrt(3) / 4]]), (n1, n2),
                          _axis_boundaries((n1, n2), boundaries), predicate, "KagomeLattice")


# src/zoo/models.jl:139-145
honeycomb_2nn = (
    BravaisTranslation((1, 1), axis=1),
    BravaisTranslation((2, 2), axis=1, dist=-1),
    BravaisTranslation((1, 1), axis=2, dist=-1),
    BravaisTranslation((2, 2), axis=2),
    BravaisTranslation((1, 1), (-1, 1)),
    BravaisTranslation((2, 2), (1, -1)),
)
