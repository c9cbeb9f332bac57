"""TimeSequence frame collector (host side; src/timesequence.jl:6-61,109-125,193-265) and the
asynchronous device frame sink that feeds it without a host stall per frame."""
from __future__ import annotations

import math

import numpy as np


class TimeSequence:
    def __init__(self, f=None, evol=None, times=None):
        self.times, self.snapshots = [], []
        if f is not None:
            it = evol(times) if times is not None else evol
            for moment in it:                       # [f(moment) for moment in evol]
                self[moment.t] = f(moment)

    def __setitem__(self, t, value):
        v = np.array(value, copy=True) if not np.isscalar(value) else value
        for k, tk in enumerate(self.times):
            if abs(tk - t) < math.sqrt(np.finfo(float).eps):
                self.snapshots[k] = v
                return
        self.times.append(float(t))
        self.snapshots.append(v)

    def __getitem__(self, t):
        for tk, v in zip(self.times, self.snapshots):
            if abs(tk - t) < math.sqrt(np.finfo(float).eps):   # tolerant lookup
                return v
        raise KeyError(t)

    def __len__(self):
        return len(self.times)

    def __iter__(self):
        return iter(zip(self.times, self.snapshots))

    def differentiate(self):
        out = TimeSequence()
        for k in range(len(self.times) - 1):
            dt = self.times[k + 1] - self.times[k]
            out[(self.times[k + 1] + self.times[k]) / 2] = (np.asarray(self.snapshots[k + 1]) - np.asarray(self.snapshots[k])) / dt
        return out

    def integrate(self):
        out = TimeSequence()
        acc = np.zeros_like(np.asarray(self.snapshots[0], dtype=float))
        out[self.times[0]] = acc.copy()
        for k in range(1, len(self.times)):
            dt = self.times[k] - self.times[k - 1]
            acc = acc + (np.asarray(self.snapshots[k]) + np.asarray(self.snapshots[k - 1])) / 2 * dt
            out[self.times[k]] = acc.copy()
        return out


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
