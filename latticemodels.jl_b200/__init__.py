"""B200-native unitary-evolution backend for LatticeModels.jl (Python host mirror).

The product is the C-ABI shared library ``lib/liblm_b200.so`` (include/lm_b200.h) of
hand-written sm_100a kernels; this package mirrors the reference's operator / solver interface
for the hot path (same names, argument meaning and error behaviour) on top of it via ctypes.
Nothing here falls back to a CPU implementation of the hot path.
"""
from ._lib import ArgumentError, BackendUnavailable, LM_C64, LM_C128, library_path, load  # noqa: F401
from .build import build  # noqa: F401
from .context import Context, default_context, shard_range  # noqa: F401
from .fields import (FieldSum, LandauGauge, NoField, PointFlux, PointFluxes,  # noqa: F401
                     SymmetricGauge)
from .lattices import (Bravais, BravaisLattice, BravaisTranslation, HoneycombLattice, KagomeLattice,  # noqa: F401
                       NearestNeighbor, SquareLattice, TriangularLattice, honeycomb_2nn)
from .hamiltonian import (DeviceHam, Hamiltonian, construct_hamiltonian, construct_operator,  # noqa: F401
                          haldane, kanemele, qwz, tightbinding_hamiltonian)
from .states import (DeviceState, PsiProjector, densitymatrix, diagonalize, eigs_lowest, groundstate,  # noqa: F401
                     groundstate_device)
from .evolution import B200Exp, Evolution, EvolutionSolver, EvolutionTimestamp  # noqa: F401
from .observables import (Currents, DensityCurrents, LatticeValue, LocalOperatorCurrents, SubCurrents, SubLattice,  # noqa: F401
                          currentsfrom, currentsfromto, findnz, localdensity, localexpect)
from .timesequence import AsyncFrameSink, TimeSequence  # noqa: F401
