"""Gauge fields: descriptors handed to the device phase kernel, plus a vectorised host
evaluation used only when the user asks for the assembled matrix (``Hamiltonian.data``).

Mirrors (paths relative to the reference repository):
  NoField / FieldSum      src/operators/magneticfield.jl:43-44,100-119
  LandauGauge             src/zoo/magneticfields.jl:11-15
  SymmetricGauge          src/zoo/magneticfields.jl:27-31
  PointFlux / PointFluxes src/zoo/magneticfields.jl:46-104,119-187
  adapt_field (PBC)       src/zoo/magneticfields.jl:237-266
"""
from __future__ import annotations

import itertools
import warnings

import numpy as np

from . import _lib


class AbstractField:
    def terms(self):
        """Flat list of closed-form terms: [(kind, (p0, p1, p2)), ...]."""
        raise NotImplementedError

    def __add__(self, other):
        return FieldSum(_flatten(self) + _flatten(other))

    def adapt(self, lat):
        return self

    def descriptor(self):
        t = self.terms()
        kinds = np.array([k for k, _ in t], np.int32)
        params = np.array([p for _, p in t], np.float64).reshape(len(t), 3)
        return kinds, np.ascontiguousarray(params)

    def line_integral(self, r1, r2):
        """Vectorised over bonds: r1, r2 of shape (nb, 2)."""
        r1 = np.asarray(r1, float).reshape(-1, 2)
        r2 = np.asarray(r2, float).reshape(-1, 2)
        out = np.zeros(len(r1))
        for kind, p in self.terms():
            out += _line_integral(kind, p, r1, r2)
        return out


def _flatten(f):
    return tuple(f.fields) if isinstance(f, FieldSum) else (f,)


def _line_integral(kind, p, r1, r2):
    x1, y1, x2, y2 = r1[:, 0], r1[:, 1], r2[:, 0], r2[:, 1]
    if kind == _lib.FIELD_LANDAU:
        return (x1 + x2) * (y2 - y1) * p[0] / 2
    if kind == _lib.FIELD_SYMMETRIC:
        return (x1 * y2 - x2 * y1) / 2 * p[0]
    ax, ay, bx, by = x1 - p[1], y1 - p[2], x2 - p[1], y2 - p[2]
    if kind == _lib.FIELD_POINTFLUX_AXIAL:
        n1, n2 = np.sqrt(ax * ax + ay * ay), np.sqrt(bx * bx + by * by)
        bad = (n1 < 1e-11) | (n2 < 1e-11)
        nn = np.where(bad, 1.0, n1 * n2)
        c = (ax * bx + ay * by) / nn / (1 + 1e-11)
        ang = np.arccos(np.clip(c, -1, 1)) * np.sign(ax * by - ay * bx)
        return np.where(bad, 0.0, ang * p[0] / (2 * np.pi))
    if kind == _lib.FIELD_POINTFLUX_SINGULAR:
        sg = bx - ax
        zero = (np.abs(sg) < 1e-11) | (ax * bx > 0) | (np.maximum(ax, bx) == 0)
        den = np.where(np.abs(ax - bx) < 1e-300, 1.0, ax - bx)
        yint = (-ay * bx + by * ax) / den
        return np.where(zero | (yint > 0), 0.0, p[0] * np.sign(sg))
    raise ValueError("unknown field kind %r" % kind)


class NoField(AbstractField):
    def terms(self):
        return []

    def __eq__(self, other):
        return isinstance(other, NoField)

    def __hash__(self):
        return hash("NoField")


class LandauGauge(AbstractField):
    def __init__(self, B):
        self.B = float(B)

    def terms(self):
        return [(_lib.FIELD_LANDAU, (self.B, 0.0, 0.0))]


class SymmetricGauge(AbstractField):
    def __init__(self, B):
        self.B = float(B)

    def terms(self):
        return [(_lib.FIELD_SYMMETRIC, (self.B, 0.0, 0.0))]


_GAUGES = {"axial": _lib.FIELD_POINTFLUX_AXIAL, "singular": _lib.FIELD_POINTFLUX_SINGULAR}


class PointFlux(AbstractField):
    def __init__(self, flux, point=(0, 0), gauge="axial"):
        if gauge not in _GAUGES:
            raise _lib.ArgumentError("Invalid gauge: %s; expected one of %s" % (gauge, tuple(_GAUGES)))
        self.flux, self.point, self.gauge = float(flux), (float(point[0]), float(point[1])), gauge

    def terms(self):
        return [(_GAUGES[self.gauge], (self.flux, self.point[0], self.point[1]))]

    def adapt(self, lat):
        if not lat.boundaries:
            return self
        return PointFluxes([self.flux], [self.point], gauge=self.gauge).adapt(lat)


class PointFluxes(AbstractField):
    def __init__(self, fluxes=(), points=(), gauge="axial"):
        if gauge not in _GAUGES:
            raise _lib.ArgumentError("Invalid gauge: %s; expected one of %s" % (gauge, tuple(_GAUGES)))
        points = [(float(p[0]), float(p[1])) for p in points]
        if np.isscalar(fluxes):
            fluxes = [float(fluxes)] * len(points)
        fluxes = [float(f) for f in fluxes]
        if len(fluxes) != len(points):
            raise _lib.ArgumentError("Length of fluxes and points should be the same")
        self.fluxes, self.points, self.gauge = fluxes, points, gauge

    def terms(self):
        k = _GAUGES[self.gauge]
        return [(k, (f, p[0], p[1])) for f, p in zip(self.fluxes, self.points)]

    def adapt(self, lat):
        """Replicate over the 3^nb image cells in the singular gauge
        (src/zoo/magneticfields.jl:237-261)."""
        if not lat.boundaries:
            return self
        if self.gauge != "singular":
            warnings.warn("Setting flux gauge to singular for a lattice with periodic boundary conditions")
        nb = len(lat.boundaries)
        fl, pts = [], []
        for tup_rev in itertools.product(*([range(-1, 2)] * nb)):
            tup = tuple(reversed(tup_rev))
            shift = np.zeros(2)
            for (tr, _), n in zip(lat.boundaries, tup):
                shift -= n * (lat.a @ np.array(tr[:2], float))
            for f, p in zip(self.fluxes, self.points):
                fl.append(f)
                pts.append((p[0] + shift[0], p[1] + shift[1]))
        return PointFluxes(fl, pts, gauge="singular")


class FieldSum(AbstractField):
    def __init__(self, fields):
        self.fields = tuple(fields)

    def terms(self):
        return [t for f in self.fields for t in f.terms()]

    def adapt(self, lat):
        return FieldSum(tuple(f.adapt(lat) for f in self.fields))
