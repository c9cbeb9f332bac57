"""TimeSequence frame collector (host side; src/timesequence.jl:6-61,109-125,193-265) and the
asynchronous device frame sink that feeds it without a host stall per frame."""
from __future__ import annotations

import math

import numpy as np


_ATOL = math.sqrt(np.finfo(float).eps)        # the reference's key tolerance: isapprox(t, atol = sqrt(eps()))


def _in(t, rng):
    lo, hi = rng
    return lo <= t <= hi


def _same(a, b):
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        return np.shape(a) == np.shape(b) and bool(np.all(np.asarray(a) == np.asarray(b)))
    return bool(a == b)


def _num(v):
    """Operand of the calculus: lists become arrays, everything else keeps its own ``*`` / ``+``."""
    return np.asarray(v) if isinstance(v, (list, tuple)) else v


class TimeSequence:
    """Time-ordered ``t => value`` dictionary (src/timesequence.jl:6-61).

    ``TimeSequence()`` is empty, ``TimeSequence(times, values)`` wraps two equally long sequences
    (``ValueError`` otherwise, the reference's ``ArgumentError``, :9-12), and
    ``TimeSequence(f, evol[, times])`` is ``[f(moment) for moment in evol(times)]`` (:41-43).
    Keys match within ``sqrt(eps)`` (:85-88); new keys are inserted in time order (:89-99).
    A closed interval ``(lo, hi)`` stands for the reference's ``t = lo .. hi``."""

    def __init__(self, f=None, evol=None, times=None):
        self.times, self.snapshots = [], []
        if f is None:
            return
        if callable(f):
            # TimeSequence(f, evol_iter) = TimeSequence(evol_iter.times, [f(moment) for moment in evol_iter]):
            # the keys are the NOMINAL times of the iterator (src/timesequence.jl:41-43), not the accumulated clock
            it = evol(times) if times is not None else evol
            vals = [self._own(f(moment)) for moment in it]
            keys = [float(t) for t in getattr(it, "times", [])]
            if len(keys) != len(vals):
                raise ValueError("Keys/values length mismatch:\n%d timestamps, %d snapshots" % (len(keys), len(vals)))
            self.times, self.snapshots = keys, vals
            return
        ts, vs = list(f), list(evol if evol is not None else [])
        if len(ts) != len(vs):
            raise ValueError("Keys/values length mismatch:\n%d timestamps, %d snapshots" % (len(ts), len(vs)))
        self.times = [float(t) for t in ts]
        self.snapshots = [self._own(v) for v in vs]

    @staticmethod
    def _own(v):
        """The reference's TimeSequence{ET} holds ARBITRARY values (Currents, LatticeValue, tuples ...):
        arrays are copied, objects with a ``copy`` method are copied through it, anything else is kept."""
        if isinstance(v, np.ndarray):
            return v.copy()
        if np.isscalar(v) or isinstance(v, (tuple, str)) or v is None:
            return v
        if isinstance(v, list):
            return list(v)
        cp = getattr(v, "copy", None)
        return cp() if callable(cp) else v

    # ---- dictionary interface (:54-58, 85-137) ----
    def timestamps(self):
        return self.times

    def timerange(self):
        return (self.times[0], self.times[-1])

    def values(self):
        return self.snapshots

    def _find(self, t):
        for k, tk in enumerate(self.times):
            if abs(tk - t) <= _ATOL:
                return k
        return None

    def get(self, t, default=None):
        k = self._find(t)
        return default if k is None else self.snapshots[k]

    def __setitem__(self, t, value):
        v = self._own(value)
        k = self._find(t)
        if k is None:
            k = int(np.searchsorted(np.asarray(self.times, dtype=float), float(t), side="left"))   # searchsortedfirst
            self.times.insert(k, float(t))
            self.snapshots.insert(k, v)
        else:
            self.snapshots[k] = v

    def __getitem__(self, t):
        if isinstance(t, tuple) and len(t) == 2:            # tseq[t = lo .. hi]
            return self.slice(t)
        k = self._find(t)
        if k is None:
            raise KeyError(t)
        return self.snapshots[k]

    def slice(self, t=(-math.inf, math.inf), index=None):
        """``tseq[args...; t = lo .. hi]`` (:106-120): the entries inside the closed interval, each value
        optionally indexed by ``index`` (a site number, a mask, ...)."""
        pick = (lambda v: v) if index is None else (lambda v: v[index] if hasattr(v, "__getitem__") and not isinstance(v, (list, tuple)) else np.asarray(v)[index])
        keep = [k for k, tk in enumerate(self.times) if _in(tk, t)]
        return TimeSequence([self.times[k] for k in keep], [pick(self.snapshots[k]) for k in keep])

    def delete(self, t):
        """``delete!(tseq, t)`` for a number (:100-106, a missing key is ignored) or
        ``delete!(tseq; t = lo .. hi)`` for an interval ``(lo, hi)`` (:121-126)."""
        if isinstance(t, tuple):
            drop = [k for k, tk in enumerate(self.times) if _in(tk, t)]
        else:
            k = self._find(t)
            drop = [] if k is None else [k]
        for k in reversed(drop):
            del self.times[k]
            del self.snapshots[k]
        return self

    def __contains__(self, t):
        return self._find(t) is not None

    def __len__(self):
        return len(self.times)

    def __iter__(self):
        return iter(zip(self.times, self.snapshots))

    def __eq__(self, other):
        if not isinstance(other, TimeSequence):
            return NotImplemented
        return self.times == other.times and len(self.snapshots) == len(other.snapshots) and \
            all(_same(a, b) for a, b in zip(self.snapshots, other.snapshots))

    __hash__ = None

    def copy(self):
        return TimeSequence(list(self.times), self.snapshots)

    def empty(self):
        return TimeSequence()

    def map(self, f):
        return TimeSequence(list(self.times), [f(v) for v in self.snapshots])

    def __repr__(self):
        head = "TimeSequence with %d entr%s" % (len(self), "y" if len(self) == 1 else "ies")
        if len(self) >= 2:
            head += ", timestamps in range %g .. %g" % self.timerange()
        return head

    # ---- calculus (:193-265) ----
    def differentiate_(self):
        """``differentiate!``: symmetric differences, keys moved to the interval midpoints (:193-207)."""
        if len(self) < 2:
            raise RuntimeError("Cannot differentiate TimeSequence of length %d" % len(self))
        td, vs = self.times, self.snapshots
        for i in range(1, len(td)):
            dt = td[i] - td[i - 1]
            a, b = _num(vs[i]), _num(vs[i - 1])             # the values' own arithmetic (arrays, Currents, LatticeValue-likes)
            vs[i - 1] = (1 / dt) * a + (-1 / dt) * b
            td[i - 1] += dt / 2
        td.pop()
        vs.pop()
        return self

    def differentiate(self):
        return self.copy().differentiate_()

    def integrate_(self):
        """``integrate!``: trapezoidal rule, first value zero, keys kept (:251-264)."""
        if len(self) < 2:
            raise RuntimeError("Cannot integrate TimeSequence of length %d" % len(self))
        td, vs = self.times, self.snapshots
        for i in range(1, len(td)):
            dt = td[i] - td[i - 1]
            vs[i - 1] = (dt / 2) * _num(vs[i]) + (dt / 2) * _num(vs[i - 1])
        last = vs.pop()
        vs.insert(0, 0 * last)
        for i in range(1, len(td)):
            vs[i] = vs[i - 1] + vs[i]
        return self

    def integrate(self):
        return self.copy().integrate_()


class AsyncFrameSink:
    """Double-buffered collector of ``(localdensity, DensityCurrents)`` frames: ``push`` enqueues
    the fused reductions of the current device state and the device->host copy of the frame on
    a second stream and returns at once, so the next propagation step overlaps the copy; the
    frame is read back one ``push`` later (or by ``finish``).  Restates
    ``TimeSequence(f, evol, times)`` (src/timesequence.jl:41-43) for the common
    ``f = moment -> (localdensity(P), DensityCurrents(H, P))`` without a stall per frame.

        sink = AsyncFrameSink()
        for t in ts:
            sol.update_solver(H(t), dt); sol.step(state)
            sink.push(sol.dev, state, t)
        rho_seq, cur_seq = sink.finish()        # two TimeSequence objects
    """

    def __init__(self, want_currents=True):
        from . import _lib
        self._lib = _lib
        self.want_currents = bool(want_currents)
        self.rho, self.currents = TimeSequence(), TimeSequence()
        self._pending = []          # (slot, t, ctx, n_sites, npairs)
        self._next = 0

    def _collect(self):
        slot, t, ctx, ns, npairs = self._pending.pop(0)
        rho = np.zeros(ns)
        J = np.zeros(max(npairs, 1))
        self._lib.check(self._lib.load().lm_frame_wait(ctx.handle, slot, self._lib.ptr(rho),
                                                       self._lib.ptr(J) if self.want_currents else None))
        self.rho[t] = rho
        if self.want_currents:
            self.currents[t] = J[:npairs]

    def push(self, dev, state, t):
        if len(self._pending) == 2:
            self._collect()                          # frees the slot this push is about to reuse
        slot = self._next
        self._next ^= 1
        self._lib.check(self._lib.load().lm_observables_async(dev.handle, state.handle, slot, 1 if self.want_currents else 0))
        npairs = len(dev.pairs()[0]) if self.want_currents else 0
        self._pending.append((slot, float(t), state.ctx, state.N // dev.n_int, npairs))

    def finish(self):
        while self._pending:
            self._collect()
        return self.rho, self.currents
