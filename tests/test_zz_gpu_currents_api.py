"""The remaining testsets of the reference's test/test_currents.jl on device currents, through the C
ABI and the host mirror of `Currents` / `SubCurrents` (src/currents.jl:30-66,110-190): iteration,
'materialised' currents, arithmetic, and the four equivalent ways to restrict currents to a region."""
import warnings

import numpy as np
import pytest

import lm_b200 as lm
from oracle import fields as F
from oracle import lattice as L
from oracle import observables as OB
from oracle import operators as OP
from oracle import spectrum as SP

pytestmark = pytest.mark.gpu


def _dense(m):
    return np.array(m.toarray() if hasattr(m, "toarray") else m, dtype=float)


@pytest.fixture(scope="module")
def setup():
    ctx = lm.default_context("c128")
    lo, l = L.square_lattice(4, 4), lm.SquareLattice(4, 4)
    H0o, H1o = OP.qwz(lo), OP.qwz(lo, field=F.LandauGauge(0.1))
    H1 = lm.qwz(l, field=lm.LandauGauge(0.1))
    P, Psi, w = SP.densitymatrix(H0o, mu=0.0)
    st = lm.DeviceState.from_psi(Psi, w, ctx=ctx, n_int=2)
    return ctx, l, H0o, H1o, H1, P, lm.DensityCurrents(H1, st)


def test_currents_iterator_and_materialised(setup):
    """test/test_currents.jl:34-52: iteration yields every pair once, oriented along the flow
    (value >= 0, equal to c[a, b]); Currents(lattice) with setindex!, the self-current warning,
    mc + mc == 2 mc + zero(mc)."""
    ctx, l, H0o, H1o, H1, P, dc = setup
    gs = SP.groundstate(H0o)
    c1 = lm.Currents(lm.DensityCurrents(H1, lm.DeviceState.from_psi(gs, ctx=ctx, n_int=2)))
    want = _dense(OB.currents_matrix(H1o, np.outer(gs, gs.conj()), 2))
    n = 0
    for (a, b), v in c1:
        assert v == c1[a, b] or (v == 0 and c1[a, b] == 0)
        assert v >= 0
        assert abs(v - want[a - 1, b - 1]) < 1e-13
        n += 1
    assert n == len(c1) == 16 * 15 // 2
    l2 = lm.SquareLattice(1, 2)
    m = lm.Currents(l2)
    m[1, 2] = 2
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        m[1, 1] = 1
    assert any("Attempt to assign nonzero current from site to self" in str(r.message) for r in rec)
    assert np.array_equal(m.toarray(), [[0, 2], [-2, 0]])
    mc = lm.Currents(dc)
    assert mc + mc == 2 * mc + mc.zero()
    assert mc.copy() == mc and (mc - mc) == mc.zero() and (mc / 2).isapprox(0.5 * mc)
    with pytest.raises(lm.ArgumentError):
        mc + m


def test_subcurrents_four_ways(setup):
    """test/test_currents.jl:54-63: Currents(dc)[x .< y] == Currents(dc[x .< y]) ==
    Currents(dc, adjacency)[x .< y] == Currents(dc[x .< y], adjacency), and all equal the oracle's
    current matrix restricted to the region."""
    ctx, l, H0o, H1o, H1, P, dc = setup
    c = np.asarray(l.coords)
    mask = c[:, 0] < c[:, 1]
    adj = [(i, j) for i, j in OB.site_adjacency(H1o, 2)]
    m1 = lm.Currents(dc)[mask]
    m2 = lm.Currents(dc[mask])
    m3 = lm.Currents(dc, adj)[mask]
    m4 = lm.Currents(dc[mask], adj)
    assert m1.isapprox(m2) and m1.isapprox(m3) and m1.isapprox(m4)
    inds = np.flatnonzero(mask)
    want = _dense(OB.currents_matrix(H1o, P, 2))[np.ix_(inds, inds)]
    want[np.abs(want) < 1e-10] = 0
    for m in (m1, m2, m3, m4):
        assert m.lattice == m1.lattice and len(m.lattice) == len(inds)
        assert np.abs(m.toarray() - want).max() < 1e-13
    sub = dc[mask]
    assert len(sub) == len(inds) * (len(inds) - 1) // 2
    a, b = 2, 5
    assert sub[a, b] == pytest.approx(dc[int(inds[a - 1]) + 1, int(inds[b - 1]) + 1], abs=0)
    assert sub[a, b] == -sub[b, a]
    # a region of a region, by 1-based site numbers
    nested = lm.Currents(sub[[1, 3, 4]])
    pick = inds[[0, 2, 3]]
    w2 = _dense(OB.currents_matrix(H1o, P, 2))[np.ix_(pick, pick)]
    w2[np.abs(w2) < 1e-10] = 0
    assert np.abs(nested.toarray() - w2).max() < 1e-13
    Is, Js, Vs = lm.findnz(sub)
    assert np.all(Is < Js) and np.all(np.abs(Vs) >= 1e-10)
    assert lm.currentsfromto(lm.Currents(dc), inds + 1) == pytest.approx(lm.currentsfromto(dc, mask), abs=1e-12)
