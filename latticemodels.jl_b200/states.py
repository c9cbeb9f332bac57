"""Evolved states: the Psi-block reformulation P = Psi diag(w) Psi' and device handles.

Mirrors (paths relative to the reference repository):
  projector / densfun / densitymatrix   src/spectrum.jl:216-232,253-265,290-333
  groundstate                           src/spectrum.jl:205
  EvolutionStateType, copy(state)       src/evolution.jl:36,193
The initial-state eigendecomposition runs once on the host (LAPACK via numpy) exactly like
the reference's ``diagonalize`` (src/spectrum.jl:48-55); what changes is the hand-off: the
occupied eigenvectors and their weights go to the device as a Psi block instead of being
multiplied out to a dense N x N matrix.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .context import default_context


class PsiProjector:
    """Host description of P = Psi diag(w) Psi' (Psi: N x M column-major friendly)."""

    def __init__(self, psi, weights=None, lattice=None, n_int=1):
        self.psi = np.asarray(psi, complex)
        if self.psi.ndim == 1:
            self.psi = self.psi[:, None]
        self.weights = None if weights is None else np.asarray(weights, float)
        self.lattice, self.n_int = lattice, n_int

    @property
    def shape(self):
        n = self.psi.shape[0]
        return (n, n)

    def dense(self):
        w = np.ones(self.psi.shape[1]) if self.weights is None else self.weights
        return (self.psi * w[None, :]) @ self.psi.conj().T


def diagonalize(ham):
    H = ham.data if hasattr(ham, "data") else ham
    Hd = H.toarray() if hasattr(H, "toarray") else np.asarray(H)
    return np.linalg.eigh(Hd)


def densitymatrix(ham, T=0.0, mu=0.0, N=None, statistics=1):
    """``densitymatrix(ham; T, mu, N, statistics)``: Fermi-Dirac (+1) / Bose-Einstein (-1)
    ensemble; ``N`` selects the Fermi sphere of the N lowest levels (T = 0)."""
    E, V = diagonalize(ham)
    lat = getattr(ham, "lattice", None)
    n_int = getattr(ham, "n_int", 1)
    if N is not None:
        if T != 0:
            raise _lib.ArgumentError("fixed-N ensembles at T > 0 are not covered by the device hand-off")
        if len(E) < N:
            raise _lib.ArgumentError("cannot build Fermi sphere with %d particles: only %d bands present" % (N, len(E)))
        return PsiProjector(V[:, :N].copy(), np.ones(N), lat, n_int)
    if T == 0:
        w = (E <= mu).astype(float)               # densfun(T = 0): occupies E <= mu
    else:
        with np.errstate(over="ignore"):
            w = 1.0 / (np.exp((E - mu) / T) + statistics)
    keep = w != 0
    return PsiProjector(V[:, keep].copy(), w[keep].copy(), lat, n_int)


def groundstate(ham):
    E, V = diagonalize(ham)
    return V[:, 0].copy()


def eigs_lowest(ham, n=10, tol=1e-10, ctx=None, max_iter=0, degree=0):
    """Lowest ``n`` eigenpairs ON THE DEVICE (``diagonalize(ham, :krylovkit; n)``, src/spectrum.jl:56-64, for sizes
    where the host eigen-decomposition is impossible): returns ``(E, DeviceState)`` - ascending eigenvalues and a
    device block of the n orthonormal eigenvectors (the Fermi sphere of the n lowest levels, ready for Evolution).
    ``eigs_lowest.info`` holds the residual norms and the iteration count of the last call."""
    ctx = ctx or default_context()
    dev = ham.device(ctx)
    E, res = np.zeros(n), np.zeros(n)
    h, it = C.c_void_p(), C.c_int32()
    _lib.check(_lib.load().lm_eigs_lowest(dev.handle, int(n), float(tol), int(max_iter), int(degree), E.ctypes.data_as(C.POINTER(C.c_double)),
                                          res.ctypes.data_as(C.POINTER(C.c_double)), C.byref(h), C.byref(it)))
    eigs_lowest.info = dict(residuals=res, iterations=it.value)
    return E, DeviceState(ctx, h, getattr(ham, "lattice", None), getattr(ham, "n_int", 1), n, (0, n))


def groundstate_device(ham, tol=1e-10, ctx=None):
    """``groundstate(ham)`` (src/spectrum.jl:205) on the device: ``(E0, DeviceState)`` with one column."""
    E, st = eigs_lowest(ham, 1, tol, ctx)
    return float(E[0]), st


class DeviceState:
    """Owner of an ``lm_state`` handle (a Psi block shard or a dense density matrix)."""

    def __init__(self, ctx, handle, lattice=None, n_int=1, full_M=None, col_range=None):
        self.ctx, self.handle, self.lattice, self.n_int = ctx, handle, lattice, n_int
        N, M, d = C.c_int64(), C.c_int64(), C.c_int32()
        _lib.check(_lib.load().lm_state_dims(handle, C.byref(N), C.byref(M), C.byref(d)))
        self.N, self.M, self.is_dense = N.value, M.value, bool(d.value)
        self.full_M = full_M if full_M is not None else self.M
        self.col_range = col_range if col_range is not None else (0, self.M)
        # every rank holds the whole block (a ket, shard=False, a one-rank job): its observables are
        # complete on each rank and must not be summed across ranks (lm_state_set_replicated)
        self.replicated = (not self.is_dense) and ctx.nranks > 1 and self.col_range == (0, self.full_M)
        if self.replicated:
            _lib.check(_lib.load().lm_state_set_replicated(handle, 1))

    # ---- constructors ---------------------------------------------------------------
    @classmethod
    def from_psi(cls, psi, weights=None, ctx=None, lattice=None, n_int=1, shard=True):
        ctx = ctx or default_context()
        psi = np.asarray(psi)
        if psi.ndim == 1:
            psi = psi[:, None]
        N, M = psi.shape
        b, e = ctx.shard_range(M) if (shard and ctx.nranks > 1) else (0, M)
        if e <= b:
            raise _lib.ArgumentError("rank %d owns no column of the %d-column block" % (ctx.rank, M))
        local = np.asfortranarray(psi[:, b:e], dtype=_lib.cdtype(ctx.precision))
        w = None if weights is None else np.ascontiguousarray(np.asarray(weights, float)[b:e])
        h = C.c_void_p()
        _lib.check(_lib.load().lm_state_create_psi(ctx.handle, N, e - b, _lib.ptr(local), _lib.ptr(w), C.byref(h)))
        return cls(ctx, h, lattice, n_int, M, (b, e))

    @classmethod
    def synthetic(cls, N, M, ctx=None, seed=1234, lattice=None, n_int=1, shard=True):
        """Device-generated block (lm_state_create_psi_synth): uniform complex in [-1, 1]^2, columns
        of norm ~ 1; a rank's shard holds the same values as those columns of the unsharded block."""
        ctx = ctx or default_context()
        b, e = ctx.shard_range(M) if (shard and ctx.nranks > 1) else (0, M)
        if e <= b:
            raise _lib.ArgumentError("rank %d owns no column of the %d-column block" % (ctx.rank, M))
        h = C.c_void_p()
        _lib.check(_lib.load().lm_state_create_psi_synth(ctx.handle, N, e - b, b, seed, C.byref(h)))
        return cls(ctx, h, lattice, n_int, M, (b, e))

    def column_norms2(self):
        """||psi_c||^2 of the local columns."""
        out = np.empty(self.M)
        _lib.check(_lib.load().lm_state_column_norms2(self.handle, _lib.ptr(out)))
        return out

    @classmethod
    def from_dense(cls, P, ctx=None, lattice=None, n_int=1):
        ctx = ctx or default_context()
        P = np.asfortranarray(P, dtype=_lib.cdtype(ctx.precision))
        if P.ndim != 2 or P.shape[0] != P.shape[1]:
            raise _lib.ArgumentError("density matrix must be square")
        h = C.c_void_p()
        _lib.check(_lib.load().lm_state_create_dense(ctx.handle, P.shape[0], _lib.ptr(P), C.byref(h)))
        return cls(ctx, h, lattice, n_int)

    @classmethod
    def from_any(cls, state, ctx=None, lattice=None, n_int=1):
        """EvolutionStateType dispatch (src/evolution.jl:36): vector -> Ket, square matrix ->
        density operator, PsiProjector -> Psi block, DeviceState -> copy."""
        if isinstance(state, DeviceState):
            return state.copy()
        if isinstance(state, PsiProjector):
            return cls.from_psi(state.psi, state.weights, ctx, state.lattice or lattice, state.n_int)
        arr = np.asarray(state)
        if arr.ndim == 1:
            return cls.from_psi(arr, None, ctx, lattice, n_int, shard=False)
        if arr.ndim == 2 and arr.shape[0] == arr.shape[1]:
            return cls.from_dense(arr, ctx, lattice, n_int)
        raise _lib.ArgumentError("invalid state type: %s" % type(state))

    # ---- access ------------------------------------------------------------------------
    def copy(self):
        h = C.c_void_p()
        _lib.check(_lib.load().lm_state_copy(self.handle, C.byref(h)))
        return DeviceState(self.ctx, h, self.lattice, self.n_int, self.full_M, self.col_range)

    def download(self):
        """Psi block (N x M_local) or dense P (N x N) as a numpy array."""
        dt = _lib.cdtype(self.ctx.precision)
        if self.is_dense:
            out = np.zeros((self.N, self.N), dt, order="F")
            _lib.check(_lib.load().lm_state_download_dense(self.handle, _lib.ptr(out)))
        else:
            out = np.zeros((self.N, self.M), dt, order="F")
            _lib.check(_lib.load().lm_state_download_psi(self.handle, _lib.ptr(out)))
        return out

    @property
    def data(self):
        d = self.download()
        return d[:, 0] if (not self.is_dense and self.M == 1) else d

    def dense(self):
        """Materialise the N x N density matrix (Psi diag(w) Psi' over the local columns)."""
        out = np.zeros((self.N, self.N), _lib.cdtype(self.ctx.precision), order="F")
        _lib.check(_lib.load().lm_state_download_dense(self.handle, _lib.ptr(out)))
        return out

    def __del__(self):
        try:
            if self.handle:
                _lib.load().lm_state_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
