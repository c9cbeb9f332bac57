"""Oracle: Bravais lattices, site order, translations, boundaries (scalar restatement).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows (paths relative to /root/reference):
  site order            src/lattices/bravais/lattice.jl:101-111 (add_bravaispointers!)
  pointer ordering      src/lattices/bravais/unitcell.jl:126-132
  site coordinates      src/lattices/bravais/unitcell.jl:113-122
  unit cells            src/zoo/lattices.jl:131 (Square), :161 (Honeycomb)
  translations          src/lattices/bravais/bonds.jl:10-98 (_destination_bp :87-96)
  bond iteration        src/core/bonds.jl:415-422
  boundaries            src/core/boundaries.jl:4-27,156-165,251-289
  NN detection          src/lattices/bravais/nearestneighbor.jl:97-131,174-178
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, field as dc_field

import numpy as np


@dataclass
class UnitCell:
    translations: np.ndarray  # (dim, NU) columns are unit vectors
    basissites: np.ndarray    # (dim, NB) columns are basis positions

    @property
    def nb(self):
        return self.basissites.shape[1]

    @property
    def nu(self):
        return self.translations.shape[1]

    def site_coords(self, latcoords, basindex):
        # src/lattices/bravais/unitcell.jl:119-122 (basindex is 1-based)
        return self.basissites[:, basindex - 1] + self.translations @ np.asarray(latcoords, dtype=float)


@dataclass(frozen=True)
class BravaisTranslation:
    """site_indices == (0, 0) means 'keep the sublattice' (bonds.jl:10-21)."""
    site_indices: tuple
    translate_uc: tuple

    def inv(self):
        a, b = self.site_indices
        return BravaisTranslation((b, a), tuple(-x for x in self.translate_uc))


def bravais(*uc):
    return BravaisTranslation((0, 0), tuple(int(x) for x in uc))


def translation(site_indices=None, uc=None, axis=0, dist=1, nu=2):
    if site_indices is None:
        site_indices = (0, 0)
    if uc is None:
        uc = [0] * nu
        uc[axis - 1] = dist
    return BravaisTranslation(tuple(site_indices), tuple(int(x) for x in uc))


@dataclass
class Boundary:
    """TwistedBoundary(translation, theta) (src/core/boundaries.jl:45-63); periodic = theta 0."""
    translate_uc: tuple
    theta: float = 0.0


@dataclass
class Lattice:
    unitcell: UnitCell
    pointers: list                      # [(latcoords tuple, basindex)] in site order
    boundaries: list = dc_field(default_factory=list)
    depth: int = 1
    sizes: tuple = ()
    kind: str = "bravais"

    def __post_init__(self):
        self._index = {p: i + 1 for i, p in enumerate(self.pointers)}
        self.coords = np.array([self.unitcell.site_coords(lc, b) for lc, b in self.pointers])

    def __len__(self):
        return len(self.pointers)

    def site_index(self, pointer):
        """1-based index or None (src/lattices/bravais/lattice.jl:69-78)."""
        return self._index.get(pointer)

    def with_boundaries(self, boundaries, depth=1):
        return Lattice(self.unitcell, list(self.pointers), list(boundaries), depth, self.sizes, self.kind)


def span_unitcells(unitcell, *sizes, boundaries=()):
    """src/lattices/bravais/lattice.jl:101-111,161-172.

    CartesianIndices(reverse(axes)) runs its FIRST index fastest, and the tuple is
    reversed again, so the LAST lattice axis runs fastest; basis index innermost.
    """
    ptrs = []
    rev_ranges = [range(1, s + 1) for s in reversed(sizes)]
    # itertools.product runs the LAST iterable fastest -> feed reversed order back.
    for latc_rev_slowfirst in itertools.product(*reversed(rev_ranges)):
        # latc_rev_slowfirst is (j1, j2, ...) with the last entry fastest.
        svec = tuple(latc_rev_slowfirst)
        for b in range(1, unitcell.nb + 1):
            ptrs.append((svec, b))
    lat = Lattice(unitcell, ptrs, [], 1, tuple(sizes))
    bl = []
    for b in boundaries:
        bl.append(b)
    if bl:
        lat = lat.with_boundaries(bl)
    return lat


def square_lattice(*sizes, periodic=(), twists=None):
    """SquareLattice(sz...) (src/zoo/lattices.jl:131).  ``periodic`` lists 1-based axes made
    periodic via the default translation `:axisN` = Bravais[... size_N ...]
    (src/lattices/bravais/lattice.jl:166-169); ``twists`` maps axis -> theta."""
    n = len(sizes)
    uc = UnitCell(np.eye(n), np.zeros((n, 1)))
    lat = span_unitcells(uc, *sizes)
    lat.kind = "square"
    return _apply_axis_boundaries(lat, sizes, periodic, twists)


def honeycomb_lattice(a, b, periodic=(), twists=None):
    """HoneycombLattice(a, b) (src/zoo/lattices.jl:161)."""
    uc = UnitCell(np.array([[1.0, 0.5], [0.0, math.sqrt(3) / 2]]),
                  np.array([[0.0, 0.5], [0.0, math.sqrt(3) / 6]]))
    lat = span_unitcells(uc, a, b)
    lat.kind = "honeycomb"
    return _apply_axis_boundaries(lat, (a, b), periodic, twists)


def triangular_lattice(a, b, periodic=(), twists=None):
    """TriangularLattice(a, b) (src/zoo/lattices.jl:145)."""
    uc = UnitCell(np.array([[1.0, 0.5], [0.0, math.sqrt(3) / 2]]), np.zeros((2, 1)))
    lat = span_unitcells(uc, a, b)
    lat.kind = "triangular"
    return _apply_axis_boundaries(lat, (a, b), periodic, twists)


def kagome_lattice(a, b, periodic=(), twists=None):
    """KagomeLattice(a, b) (src/zoo/lattices.jl:209)."""
    uc = UnitCell(np.array([[1.0, 0.5], [0.0, math.sqrt(3) / 2]]),
                  np.array([[0.0, 0.5, 0.25], [0.0, 0.0, math.sqrt(3) / 4]]))
    lat = span_unitcells(uc, a, b)
    lat.kind = "kagome"
    return _apply_axis_boundaries(lat, (a, b), periodic, twists)


def _apply_axis_boundaries(lat, sizes, periodic, twists):
    bl = []
    twists = twists or {}
    for ax in sorted(set(periodic) | set(twists)):
        tr = [0] * len(sizes)
        tr[ax - 1] = sizes[ax - 1]
        bl.append(Boundary(tuple(tr), float(twists.get(ax, 0.0))))
    return lat.with_boundaries(bl) if bl else lat


# --------------------------------------------------------------------------------------
# destinations / resolve_site
# --------------------------------------------------------------------------------------
def destination_pointer(tr: BravaisTranslation, pointer):
    """_destination_bp (src/lattices/bravais/bonds.jl:87-96). Returns None for NoSite."""
    latc, bas = pointer
    if tr.site_indices == (0, 0):
        new_bas = bas
    else:
        if tr.site_indices[0] != bas:
            return None
        new_bas = tr.site_indices[1]
    nu = len(latc)
    uc = list(tr.translate_uc) + [0] * max(0, nu - len(tr.translate_uc))
    if len(tr.translate_uc) > nu and any(x != 0 for x in tr.translate_uc[nu:]):
        return None
    return (tuple(latc[k] + uc[k] for k in range(nu)), new_bas)


def _shift(pointer, translate_uc, n):
    """nshifts(site, tr, n): site -= tr, n times (src/core/boundaries.jl:4-11)."""
    latc, bas = pointer
    return (tuple(latc[k] - n * translate_uc[k] for k in range(len(latc))), bas)


def resolve_site(lat: Lattice, pointer):
    """resolve_site(::LatticeWithMetadata, site) (src/core/boundaries.jl:251-289).

    Returns (index_1based, old_pointer, factor) or None.  ``old_pointer`` is the unwrapped
    site whose coordinates enter the Peierls integral (builder.jl:283).
    """
    if pointer is None:
        return None
    idx = lat.site_index(pointer)
    if idx is not None:
        return idx, pointer, 1.0 + 0.0j
    nb = len(lat.boundaries)
    if nb == 0:
        return None
    rng = range(-lat.depth, lat.depth + 1)
    # CartesianIndices: first index fastest.
    for tup_rev in itertools.product(*([rng] * nb)):
        tup = tuple(reversed(tup_rev))
        new_p = pointer
        for i in range(nb):
            new_p = _shift(new_p, lat.boundaries[i].translate_uc, tup[i])
        idx = lat.site_index(new_p)
        if idx is not None:
            if all(t == 0 for t in tup):
                return idx, pointer, 1.0 + 0.0j
            # findfactor for TwistedBoundary tuples (src/core/boundaries.jl:276-278)
            factor = np.exp(1j * sum(lat.boundaries[i].theta * tup[i] for i in range(nb)))
            return idx, pointer, complex(factor)
    return None


def iterate_bonds(lat: Lattice, tr: BravaisTranslation):
    """iterate(::AbstractTranslation) (src/core/bonds.jl:415-422).

    Yields (i, r_i, j, r_j_unwrapped, factor_j) with 1-based indices; the source site is
    always unwrapped with factor 1.
    """
    for i, p in enumerate(lat.pointers, start=1):
        dest = destination_pointer(tr, p)
        rs = resolve_site(lat, dest)
        if rs is None:
            continue
        j, old_p, fac = rs
        yield i, lat.coords[i - 1], j, lat.unitcell.site_coords(*old_p), fac


# --------------------------------------------------------------------------------------
# nearest-neighbour detection
# --------------------------------------------------------------------------------------
def _isapprox(a, b):
    # Julia isapprox default: rtol = sqrt(eps), atol = 0
    return abs(a - b) <= math.sqrt(np.finfo(float).eps) * max(abs(a), abs(b))


def detect_nnhops(uc: UnitCell, depth=2, limit=3):
    """src/lattices/bravais/nearestneighbor.jl:97-131 restated literally."""
    lens = []
    lists = []
    n = uc.nu
    rng = range(-depth, depth + 1)
    for c_rev in itertools.product(*([rng] * n)):
        C = tuple(reversed(c_rev))  # first index fastest
        U = uc.translations @ np.array(C, dtype=float)
        for i in range(1, uc.nb + 1):
            for j in range(i, uc.nb + 1):
                if i == j:
                    if i > 2:
                        continue
                    nzc = next((k for k, x in enumerate(C) if x != 0), None)
                    if nzc is None:
                        continue
                    if C[nzc] < 1:
                        continue
                R = U - uc.basissites[:, i - 1] + uc.basissites[:, j - 1]
                r = float(np.linalg.norm(R))
                k = next((l for l in range(len(lens)) if lens[l] > r or _isapprox(lens[l], r)), None)
                if k is None or not _isapprox(lens[k], r):
                    if k is None:
                        k = len(lens)
                    if k + 1 > limit:
                        continue
                    lens.insert(k, r)
                    lists.insert(k, [])
                tr = BravaisTranslation((0, 0), C) if i == j else BravaisTranslation((i, j), C)
                if tr not in lists[k]:
                    lists[k].append(tr)
    return lens, lists


def nearest_neighbor(lat: Lattice, n=1):
    """NearestNeighbor(N) adapted to a lattice: default nnbonds (getnnbonds,
    nearestneighbor.jl:174-178: depth 2, limit 3) if available, else detect deeper
    (nearestneighbor.jl:133-135)."""
    lens, lists = detect_nnhops(lat.unitcell)
    if n <= min(3, len(lens)):
        return lists[n - 1]
    lens, lists = detect_nnhops(lat.unitcell, 1 + math.ceil(math.sqrt(n)), n)
    return lists[n - 1]


# honeycomb_2nn (src/zoo/models.jl:139-145)
HONEYCOMB_2NN = [
    BravaisTranslation((1, 1), (1, 0)),
    BravaisTranslation((2, 2), (-1, 0)),
    BravaisTranslation((1, 1), (0, -1)),
    BravaisTranslation((2, 2), (0, 1)),
    BravaisTranslation((1, 1), (-1, 1)),
    BravaisTranslation((2, 2), (1, -1)),
]
