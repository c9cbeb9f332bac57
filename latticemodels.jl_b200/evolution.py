"""Host mirror of the reference's time-domain driver with the device solver plugged in.

Same names, argument meaning and error behaviour as the reference (paths relative to the
reference repository):
  EvolutionSolver interface        src/evolution.jl:5-33
  Evolution / step!                src/evolution.jl:156-250
  EvolutionIterator                src/evolution.jl:252-275
  EvolutionTimestamp               src/evolution.jl:277-298
``B200Exp`` is the new ``EvolutionSolver`` subtype: it owns the device
Hamiltonian and calls ``lm_step``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib
from .context import default_context
from .hamiltonian import DeviceHam, Hamiltonian
from .states import DeviceState

_METHODS = {"auto": _lib.METHOD_AUTO, "chebyshev": _lib.METHOD_CHEBYSHEV, "taylor": _lib.METHOD_TAYLOR,
            "lanczos": _lib.METHOD_LANCZOS, "taylor_horner": _lib.METHOD_TAYLOR_HORNER,
            "chebyshev_clenshaw": _lib.METHOD_CHEBYSHEV_CLENSHAW}


class EvolutionSolver:
    """Abstract solver: update_solver(mat, dt, force), step(state, cache), evolution_cache(state)."""

    def update_solver(self, mat, dt, force=False):
        raise NotImplementedError

    def step(self, state, cache):
        raise NotImplementedError

    def evolution_cache(self, state):
        return None


class B200Exp(EvolutionSolver):
    """``B200Exp([ham]; tol=1e-12, method="auto", precision="c128")`` - device propagator.

    ``tol`` bounds the truncation error of exp(-i H dt) per step (the reference's own test
    pins its solvers to the exact exponential at atol 1e-10, test/test_timedeps.jl:55-67)."""

    def __init__(self, ham=None, tol=1e-12, method="auto", precision="c128", ctx=None, n_int=None, coords=None,
                 refine_bounds=False, lattice_dims=None):
        self.ctx = ctx or default_context(precision)
        self.tol, self.method = float(tol), _METHODS[method]
        self.dev = None
        self.dt = 0.0
        self.n_int = n_int
        self.coords = coords        # site coordinates for raw-matrix Hamiltonians (tile plan)
        self.lattice_dims = lattice_dims   # (n1, n2) of an unfiltered Bravais lattice (stencil kernel)
        self.refine_bounds = refine_bounds   # opt-in Lanczos tightening of the spectral enclosure
        self._mat = None            # last raw matrix (identity check, src/evolution.jl:86-88)
        self._csc_dev = None
        self.n_matvec = 0
        if ham is not None:
            self.update_solver(_eval_ham(ham, 0.0), 0.0)

    def update_solver(self, mat, dt, force=False):
        self.dt = float(dt)
        if isinstance(mat, Hamiltonian):
            self.dev = mat.device(self.ctx)     # regenerates phases on device if the field changed
            if self.refine_bounds and not getattr(self.dev, "_refined", False):
                self.dev.refine_bounds()
            return
        if isinstance(mat, DeviceHam):
            self.dev = mat
            return
        if not force and mat is self._mat and self._csc_dev is not None:
            self.dev = self._csc_dev
            return
        m = sp.csc_matrix(mat)
        m.sort_indices()
        d = self._csc_dev
        if d is not None and d.N == m.shape[0] and d.nnz == m.nnz and \
                np.array_equal(d.pattern[0], m.indptr) and np.array_equal(d.pattern[1], m.indices):
            d.update_values(m.data)             # same sparsity pattern: upload nzval only
        else:
            d = DeviceHam.from_csc(self.ctx, m, self.n_int or 1, coords=self.coords, lattice_dims=self.lattice_dims)
        self._csc_dev, self._mat, self.dev = d, mat, d

    def step(self, state, cache=None):
        if not isinstance(state, DeviceState):
            raise _lib.ArgumentError("B200Exp evolves device states; got %s" % type(state))
        nmv = C.c_int32()
        _lib.check(_lib.load().lm_step(self.dev.handle, state.handle, self.dt, self.tol, self.method, C.byref(nmv)))
        self.n_matvec = nmv.value
        return state


def _eval_ham(ham, t):
    """_eval_ham (src/evolution.jl:42-47)."""
    if callable(ham) and not isinstance(ham, (Hamiltonian, DeviceHam)):
        return ham(t)
    return ham


class EvolutionTimestamp:
    """Destructures as ``(states..., H, t)`` (src/evolution.jl:277-298)."""

    def __init__(self, H, states, names, t):
        self.H, self.states, self._names, self.t = H, tuple(states), names, float(t)

    def __iter__(self):
        yield from self.states
        yield self.H
        yield self.t

    def __getitem__(self, k):
        if isinstance(k, str):
            if self._names is None or k not in self._names:
                raise KeyError(k)
            return self.states[self._names.index(k)]
        return self.states[k]

    @property
    def state(self):
        if len(self.states) != 1:
            raise _lib.ArgumentError("`.state` needs exactly one evolved state")
        return self.states[0]

    def __getattr__(self, name):
        names = self.__dict__.get("_names")
        if names and name in names:
            return self.states[names.index(name)]
        raise AttributeError(name)


class Evolution:
    """``Evolution([solver, ]hamiltonian, states...; timedomain, namedstates...)``.

    Stateful iterator: keeps the current time and mutates its (copied) states in place; call it
    with a time grid to iterate (src/evolution.jl:183-218)."""

    def __init__(self, *args, timedomain=None, **namedstates):
        args = list(args)
        if args and (isinstance(args[0], EvolutionSolver) or (isinstance(args[0], type) and issubclass(args[0], EvolutionSolver))):
            solver = args.pop(0)
            if isinstance(solver, type):
                solver = solver()
        else:
            solver = B200Exp()
        if not args:
            raise _lib.ArgumentError("No Hamiltonian provided")
        self.hamiltonian = args.pop(0)
        self.solver = solver
        if args and namedstates:
            raise _lib.ArgumentError("Do not use named and unnamed states together")
        if not args and not namedstates:
            raise _lib.ArgumentError("No states provided")
        self._names = list(namedstates) if namedstates else None
        raw = list(namedstates.values()) if namedstates else args
        h0 = _eval_ham(self.hamiltonian, 0.0)
        lat = getattr(h0, "lattice", None)
        n_int = getattr(h0, "n_int", None) or getattr(solver, "n_int", None) or 1
        if getattr(solver, "n_int", None) is None:
            solver.n_int = n_int
        self.states = []
        for s in raw:                       # copy(state) + evolution_cache (src/evolution.jl:190-194)
            ds = DeviceState.from_any(s, solver.ctx, lat, n_int)
            self.states.append((ds, solver.evolution_cache(ds)))
        self.time = 0.0
        self._timedomain = timedomain

    def step(self, dt):
        """step!(evol, dt) (src/evolution.jl:238-250): H is evaluated at the OLD time and
        returned; |dt| < 1e-15 does not step."""
        if dt < -1e-15:
            raise _lib.ArgumentError("negative time step")
        H = _eval_ham(self.hamiltonian, self.time)
        if abs(dt) < 1e-15:
            return H
        self.solver.update_solver(H, dt, force=False)
        for state, cache in self.states:
            self.solver.step(state, cache)
        self.time += dt
        return H

    def __call__(self, ts):
        return EvolutionIterator(self, ts)

    def __iter__(self):
        if self._timedomain is None:
            raise _lib.ArgumentError("no time domain: call the Evolution object with one")
        return iter(EvolutionIterator(self, self._timedomain))


class EvolutionIterator:
    def __init__(self, evol, times):
        self.evol, self.times = evol, [float(t) for t in times]

    def __len__(self):
        return len(self.times)

    def __iter__(self):
        ev = self.evol
        for i, t in enumerate(self.times):
            dt = (t - self.times[i - 1]) if i > 0 else (t - ev.time)     # src/evolution.jl:269
            H = ev.step(dt)
            yield EvolutionTimestamp(H, [s for s, _ in ev.states], ev._names, ev.time)
