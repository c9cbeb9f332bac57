"""Oracle: magnetic fields and Peierls line integrals (scalar restatement).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows (paths relative to /root/reference):
  LandauGauge          src/zoo/magneticfields.jl:11-15
  SymmetricGauge       src/zoo/magneticfields.jl:27-31
  PointFlux{:axial}    src/zoo/magneticfields.jl:64-85
  PointFlux{:singular} src/zoo/magneticfields.jl:86-104
  PointFluxes          src/zoo/magneticfields.jl:181-184
  FieldSum             src/operators/magneticfield.jl:100-104
  generic quadrature   src/operators/magneticfield.jl:24-35
  adapt_field (PBC)    src/zoo/magneticfields.jl:237-266
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass

import numpy as np


class Field:
    def line_integral(self, p1, p2):
        # generic fallback = one-step midpoint rule (src/operators/magneticfield.jl:24-25)
        p1 = np.asarray(p1, float)
        p2 = np.asarray(p2, float)
        a = np.asarray(self.vector_potential((p1 + p2) / 2), float)
        d = p2 - p1
        n = min(len(a), len(d))
        return float(np.dot(d[:n], a[:n]))

    def vector_potential(self, p):
        raise ValueError("no vector potential function defined for field type %s" % type(self).__name__)

    def __add__(self, other):
        return FieldSum(_terms(self) + _terms(other))

    def adapt(self, lat):
        return self


def line_integral_quadrature(field, p1, p2, n_steps):
    """src/operators/magneticfield.jl:26-35 (n-step midpoint rule)."""
    p1 = np.asarray(p1, float)
    p2 = np.asarray(p2, float)
    integral = 0.0
    dp = (p2 - p1) / n_steps
    p = p1 + 0.5 * dp
    for _ in range(n_steps):
        a = np.asarray(field.vector_potential(p), float)
        n = min(len(a), len(dp))
        integral += float(np.dot(dp[:n], a[:n]))
        p = p + dp
    return integral


@dataclass
class NoField(Field):
    def line_integral(self, p1, p2):
        return 0.0


@dataclass
class LandauGauge(Field):
    B: float

    def vector_potential(self, p):
        return (0.0, p[0] * self.B)

    def line_integral(self, p1, p2):
        return (p1[0] + p2[0]) * (p2[1] - p1[1]) * self.B / 2


@dataclass
class SymmetricGauge(Field):
    B: float

    def vector_potential(self, p):
        return (-p[1] * self.B / 2, p[0] * self.B / 2)

    def line_integral(self, p1, p2):
        return (p1[0] * p2[1] - p2[0] * p1[1]) / 2 * self.B


def _sign(x):
    return float(np.sign(x))


@dataclass
class PointFlux(Field):
    flux: float
    point: tuple = (0.0, 0.0)
    gauge: str = "axial"

    def vector_potential(self, p):
        if self.gauge != "axial":
            raise ValueError("Vector potential for a point flux is singular")
        # NOTE the reference's vector_potential ignores `point` (magneticfields.jl:66-70).
        x, y = p[0], p[1]
        normsq = x * x + y * y
        return (-y / normsq / (2 * math.pi) * self.flux, x / normsq / (2 * math.pi) * self.flux)

    def line_integral(self, p1, p2):
        px, py = self.point
        if self.gauge == "axial":
            # src/zoo/magneticfields.jl:72-85
            x1, y1 = p1[0] - px, p1[1] - py
            x2, y2 = p2[0] - px, p2[1] - py
            n1 = math.hypot(x1, y1)
            n2 = math.hypot(x2, y2)
            if n1 < 1e-11 or n2 < 1e-11:
                return 0.0
            nnorm = n1 * n2
            anglesinsign = x1 * y2 - y1 * x2
            anglecos = (x1 * x2 + y1 * y2) / nnorm / (1 + 1e-11)
            angle = math.acos(anglecos) * _sign(anglesinsign)
            return angle * self.flux / (2 * math.pi)
        elif self.gauge == "singular":
            # src/zoo/magneticfields.jl:89-104
            x1, y1 = p1[0] - px, p1[1] - py
            x2, y2 = p2[0] - px, p2[1] - py
            sg = x2 - x1
            if abs(sg) < 1e-11:
                return 0.0
            if x1 * x2 > 0 or max(x1, x2) == 0:
                return 0.0
            yintercept = (-y1 * x2 + y2 * x1) / (x1 - x2)
            return 0.0 if yintercept > 0 else self.flux * _sign(sg)
        raise ValueError("Invalid gauge: %s" % self.gauge)

    def adapt(self, lat):
        if not lat.boundaries:
            return self
        return PointFluxes([self.flux], [self.point], self.gauge).adapt(lat)


@dataclass
class PointFluxes(Field):
    fluxes: list
    points: list
    gauge: str = "axial"

    def line_integral(self, p1, p2):
        return sum(PointFlux(f, pt, self.gauge).line_integral(p1, p2)
                   for f, pt in zip(self.fluxes, self.points))

    def vector_potential(self, p):
        ax = ay = 0.0
        for f, pt in zip(self.fluxes, self.points):
            a = PointFlux(f, pt, self.gauge).vector_potential(p)
            ax += a[0]
            ay += a[1]
        return (ax, ay)

    def adapt(self, lat):
        """adapt_field(::PointFluxes, lat) (src/zoo/magneticfields.jl:237-261): replicate the
        fluxes over the 3^nb image cells in the singular gauge."""
        if not lat.boundaries:
            return self
        nb = len(lat.boundaries)
        new_f, new_p = [], []
        rng = range(-lat.depth, lat.depth + 1)
        for tup_rev in itertools.product(*([rng] * nb)):
            tup = tuple(reversed(tup_rev))
            # site = lat[1] shifted by nshifts(site, tr_i, tup_i) = site - tup_i * tr_i
            shift = np.zeros(lat.unitcell.translations.shape[0])
            for i in range(nb):
                shift = shift - tup[i] * (lat.unitcell.translations @ np.array(lat.boundaries[i].translate_uc, float))
            for f, pt in zip(self.fluxes, self.points):
                new_f.append(f)
                new_p.append((pt[0] + shift[0], pt[1] + shift[1]))
        return PointFluxes(new_f, new_p, "singular")


@dataclass
class FieldSum(Field):
    fields: tuple

    def line_integral(self, p1, p2):
        return sum(f.line_integral(p1, p2) for f in self.fields)

    def vector_potential(self, p):
        acc = np.zeros(2)
        for f in self.fields:
            a = np.asarray(f.vector_potential(p), float)
            acc[: len(a[:2])] += a[:2]
        return tuple(acc)

    def adapt(self, lat):
        return FieldSum(tuple(f.adapt(lat) for f in self.fields))


def _terms(f):
    return tuple(f.fields) if isinstance(f, FieldSum) else (f,)


@dataclass
class GaugeField(Field):
    """GaugeField(func; n) (src/operators/magneticfield.jl:66-74)."""
    func: object
    n: int

    def vector_potential(self, p):
        return self.func(p)

    def line_integral(self, p1, p2):
        return line_integral_quadrature(self, p1, p2, self.n)
