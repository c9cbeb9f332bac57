"""Oracle: cell-offset stencil mask of a lattice Hamiltonian (scalar restatement).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The device's register-tiled kernels (csrc/stencil.cuh) need the sparsity pattern of H as a bit
mask over the 9 unit-cell offsets |d1|, |d2| <= 1: with RC = rows per unit cell (basis sites x
orbitals) and rows ordered cell-major - the reference's site order, last lattice axis fastest,
basis / orbital index innermost (src/lattices/bravais/lattice.jl:101-111,
src/lattices/bravais/unitcell.jl:126-132, src/operators/system.jl:16) -

    bit (o * RC*RC + a * RC + b),  o = (d1 + 1) * 3 + (d2 + 1)
        <=>  some row a of a cell c has a stored entry in row b of cell c + (d1, d2)

(periodic images wrap to the short way round).  This module derives the mask from an oracle
Hamiltonian so that the compiled masks (LM_ST_MASK* in csrc/stencil.cuh) are pinned to the
reference's own stencils: square NN `Bravais[1,0]`, `Bravais[0,1]` (src/zoo/lattices.jl:52-55),
honeycomb NN (src/lattices/bravais/nearestneighbor.jl:149-153), `qwz` (src/zoo/models.jl:130-136),
`haldane` with `honeycomb_2nn` (src/zoo/models.jl:139-170).
"""
from __future__ import annotations

import scipy.sparse as sp


def stencil_mask(H, n1, n2):
    """(RC, mask) of H on an n1 x n2 grid of unit cells, or (RC, None) if an entry couples cells
    more than one apart."""
    m = sp.csr_matrix(H)
    N = m.shape[0]
    if N % (n1 * n2):
        raise ValueError("n1 * n2 does not divide the Hilbert dimension")
    rc = N // (n1 * n2)

    def wrap(d, n):
        if d > n // 2:
            d -= n
        if d < -(n // 2):
            d += n
        return d
    mask = 0
    for i in range(N):
        ci, a = divmod(i, rc)
        c1, c2 = divmod(ci, n2)
        for j in m.indices[m.indptr[i]:m.indptr[i + 1]]:
            cj, b = divmod(int(j), rc)
            d1, d2 = wrap(cj // n2 - c1, n1), wrap(cj % n2 - c2, n2)
            if abs(d1) > 1 or abs(d2) > 1:
                return rc, None
            o = (d1 + 1) * 3 + (d2 + 1)
            mask |= 1 << (o * rc * rc + a * rc + b)
    return rc, mask


def forward_entries(rc, mask):
    """[(a, o, b)] of the FORWARD direction of every bond: a later cell (offset (0,+1) or (+1,*)) or
    a later row of the same cell - what the device observables kernel walks."""
    out = []
    for a in range(rc):
        for o in range(4, 9):
            for b in range(rc):
                if (mask >> (o * rc * rc + a * rc + b)) & 1 and (o > 4 or b > a):
                    out.append((a, o, b))
    return out
