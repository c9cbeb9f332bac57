"""Oracle: unitary evolution with the reference's stepping semantics.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows (paths relative to /root/reference):
  step!(evol, dt)          src/evolution.jl:238-250  (H evaluated at the OLD time;
                           |dt| < 1e-15 short-circuits; negative dt -> ArgumentError)
  EvolutionIterator        src/evolution.jl:252-275  (first dt = t[1] - evol.time)
  CachedExp step!          src/evolution.jl:69-78    (psi <- U psi ; P <- U P U')
  update_solver!           src/evolution.jl:83-92
  myexp!                   src/evolution.jl:93-128   (Taylor + scaling/squaring)
  KrylovKitExp             src/evolution.jl:140-154  -> KrylovKit.exponentiate (NOT in the
                           reference tree; published algorithm restated below)
The exact propagator (``exact_propagator``) is what the reference's own test pins both solvers
to (test/test_timedeps.jl:55-67: ``exp(-im * dt * Hd.data)`` at atol = 1e-10).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp


# --------------------------------------------------------------------------------------
# propagators
# --------------------------------------------------------------------------------------
def exact_propagator(H, dt):
    """U = exp(-i H dt) through a Hermitian eigendecomposition (accurate to ~1e-15)."""
    Hd = H.toarray() if sp.issparse(H) else np.asarray(H)
    E, V = np.linalg.eigh(Hd)
    return (V * np.exp(-1j * dt * E)[None, :]) @ V.conj().T


def _nextpow2(x):
    # Base.nextpow(2, x): smallest 2^n >= x with n a NON-NEGATIVE integer (x <= 1 -> 1)
    if x <= 1:
        return 1.0
    return float(2 ** math.ceil(math.log2(x)))


def myexp(A, factor, threshold=1e-10, nztol=1e-14):
    """myexp! restated (src/evolution.jl:93-128).  norm(A, Inf) on a matrix is the
    element-wise max-abs.  Returns (U, n_terms)."""
    sparse = sp.issparse(A)
    n = A.shape[0]
    P = sp.identity(n, dtype=complex, format="csc") if sparse else np.eye(n, dtype=complex)
    if sparse:
        mat_norm = abs(A).max() if A.nnz else 0.0
    else:
        mat_norm = np.abs(A).max() if A.size else 0.0
    if mat_norm == 0 or factor == 0:
        return P, 0
    scaling = _nextpow2(mat_norm * abs(factor))
    A = A / scaling
    delta = 1.0
    next_term = P.copy()
    k = 1
    while delta > threshold / scaling:
        next_term = A @ next_term
        next_term = next_term * (factor / k)
        if sparse:
            next_term = next_term.tocsc()
            next_term.data[np.abs(next_term.data) <= nztol] = 0   # droptol!(next_term, nztol)
            next_term.eliminate_zeros()
            delta = abs(next_term).max() if next_term.nnz else 0.0
        else:
            delta = np.abs(next_term).max()
        P = P + next_term
        k += 1
    for _ in range(int(round(math.log2(scaling)))):
        P = P @ P
    return P, k - 1


def _tridiag_exp_col(alpha, beta, tau):
    """exp(tau * T) e_1 for the Hermitian tridiagonal T (alpha diag, beta off-diag)."""
    m = len(alpha)
    T = np.diag(np.asarray(alpha, float))
    for i in range(m - 1):
        T[i, i + 1] = T[i + 1, i] = beta[i]
    E, Q = np.linalg.eigh(T)
    return Q @ (np.exp(tau * E) * Q[0, :])


def krylov_exponentiate(matvec, t, v, krylovdim=30, tol=1e-12, maxiter=100):
    """exp(t*A) v for Hermitian A by the Lanczos exponential integrator.

    Restates the published algorithm of ``KrylovKit.exponentiate`` (KrylovKit.jl, compat
    "0.4 - 0.9", /root/reference/Project.toml:10,26; source not vendored, version unpinned;
    call site src/evolution.jl:151; defaults krylovdim = 30, tol = 1e-12, maxiter = 100):
    expand the Lanczos factorisation A V_m = V_m T_m + beta_m v_{m+1} e_m^T one vector at a
    time, after each expansion estimate the error of ||v|| V_m exp(tau T_m) e_1 as
    beta_m |[exp(tau T_m)]_{m,1}| ||v||, stop as soon as it drops below tol; if the subspace
    is exhausted first, advance by the largest sub-step that meets the tolerance pro rata
    and restart from the new vector (at most ``maxiter`` restarts).
    Returns (w, converged, n_matvec).
    """
    v = np.asarray(v, dtype=complex)
    total = t
    done = 0.0 + 0.0j
    nmv = 0
    w = v.copy()
    for _ in range(maxiter):
        remaining = total - done
        if abs(remaining) == 0:
            return w, True, nmv
        beta0 = np.linalg.norm(w)
        if beta0 == 0:
            return w, True, nmv
        V = [w / beta0]
        alpha, beta = [], []
        converged = False
        for m in range(1, krylovdim + 1):
            r = matvec(V[-1])
            nmv += 1
            a = np.vdot(V[-1], r).real
            r = r - a * V[-1]
            if m > 1:
                r = r - beta[-1] * V[-2]
            # full re-orthogonalisation (KrylovKit default orth = ModifiedGramSchmidtIR)
            for q in V:
                r = r - np.vdot(q, r) * q
            b = np.linalg.norm(r)
            alpha.append(a)
            y = _tridiag_exp_col(alpha, beta, remaining)
            err = abs(b * y[-1]) * beta0
            if err <= tol * abs(remaining) / abs(total) or b < 1e-300:
                converged = True
                break
            if m < krylovdim:
                beta.append(b)
                V.append(r / b)
        if converged:
            w = beta0 * (np.array(V).T @ y)
            return w, True, nmv
        # subspace exhausted: shrink the step until the estimate meets the pro-rata tolerance
        tau = remaining
        for _ in range(60):
            tau = tau / 2
            y = _tridiag_exp_col(alpha, beta, tau)
            if abs(b * y[-1]) * beta0 <= tol * abs(tau) / abs(total):
                break
        w = beta0 * (np.array(V).T @ y)
        done += tau
    return w, False, nmv


# --------------------------------------------------------------------------------------
# the Evolution iterator
# --------------------------------------------------------------------------------------
class Evolution:
    """Stateful restatement of ``Evolution`` + ``EvolutionIterator`` (src/evolution.jl:183-275).

    ``hamiltonian``: matrix (dense/sparse) or callable t -> matrix.
    ``states``: list of vectors (kets), dense matrices (density matrices) or N x M blocks
      flagged by ``block=True`` (the Psi-block reformulation: every column is a ket).
    ``solver``: "exact" | "cachedexp" | "krylov".
    """

    def __init__(self, hamiltonian, states, solver="exact", block=False, **kw):
        self.ham = hamiltonian
        self.states = [np.array(s, dtype=complex, copy=True) for s in states]
        self.solver = solver
        self.block = block
        self.kw = kw
        self.time = 0.0
        self._U = None
        self._Ukey = None
        self.n_matvec = 0

    def _eval_ham(self, t):
        return self.ham(t) if callable(self.ham) else self.ham

    def _propagator(self, H, dt):
        key = (id(H), dt)
        if self._Ukey != key:
            if self.solver == "exact":
                self._U = exact_propagator(H, dt)
            else:
                U, _ = myexp(H, -1j * dt, **self.kw)
                self._U = U.toarray() if sp.issparse(U) else U
            self._Ukey = key
            self._Href = H            # keep H alive so id() stays unique
        return self._U

    def step(self, dt):
        if dt < -1e-15:
            raise ValueError("negative time step")
        H = self._eval_ham(self.time)
        if abs(dt) < 1e-15:
            return H
        for k, s in enumerate(self.states):
            if self.solver == "krylov":
                Hs = H
                mv = (lambda x: Hs @ x)
                cols = [s] if s.ndim == 1 else [s[:, c] for c in range(s.shape[1])]
                if s.ndim == 2 and not self.block:
                    raise ValueError("KrylovKitExp only evolves vectors (src/evolution.jl:150)")
                out = []
                for c in cols:
                    w, ok, nmv = krylov_exponentiate(mv, -1j * dt, c, **self.kw)
                    if not ok:
                        raise ValueError("`exponentiate` did not converge")
                    self.n_matvec += nmv
                    out.append(w)
                self.states[k] = out[0] if s.ndim == 1 else np.stack(out, axis=1)
            else:
                U = self._propagator(H, dt)
                if s.ndim == 1 or self.block:
                    self.states[k] = U @ s
                else:
                    self.states[k] = (U @ s) @ U.conj().T
        self.time += dt
        return H

    def __call__(self, times):
        times = list(times)
        for i, t in enumerate(times):
            dt = (t - times[i - 1]) if i > 0 else (t - self.time)
            H = self.step(dt)
            yield [s for s in self.states], H, float(self.time)
