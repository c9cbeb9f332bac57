"""TimeSequence frame collector (host side; src/timesequence.jl:6-61,109-125,193-265)."""
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
