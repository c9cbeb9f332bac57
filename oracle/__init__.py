"""CPU oracle for the unitary-evolution hot path of LatticeModels.jl (v1.0.7).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import or execute it, and there only as the
checker (or the timed CPU baseline), never as the thing shipped.

It is a plain numpy/scipy *restatement* of the reference's Julia algorithms
(the reference is 100 % Julia and Julia is not installed in the build container,
so the reference itself cannot be executed: SURVEY.md section 8c).  Each function
cites the reference ``file:line`` it follows (paths relative to
``/root/reference``).

Parity pin status: "golden-vector parity unpinned; identity- and exp-parity
pinned".  The reference's test-suite holds NO stored golden vectors for
``localdensity`` / ``DensityCurrents`` time series.  What it does hold - doctest
known answers, closed-form-vs-quadrature field checks, the dense ``exp(-i dt H)``
propagator check (test/test_timedeps.jl:42-68), the Heisenberg / von Neumann
continuity identities (test/test_currents.jl:17-26, test/test_operators.jl:36-42)
and the assembly equivalences (test/test_operators.jl:78-88) - is all re-expressed
in ``tests/test_oracle_pins.py`` and must pass before the oracle is trusted.

Third-party arithmetic not vendored in the reference tree:
``KrylovKit.exponentiate`` (KrylovKit.jl, compat "0.4 - 0.9", Project.toml:10,26;
no Manifest so the exact version is unpinned) - its published algorithm (Lanczos
exp-action with residual-based stopping, defaults krylovdim=30, tol=1e-12,
maxiter=100) is restated in ``oracle/evolution.py::krylov_exponentiate``.
"""

from . import lattice, fields, operators, spectrum, evolution, observables  # noqa: F401
