"""numpy restatement of the device generator of synthetic Psi blocks (csrc/kernels.cuh
k_synth_block / lm_state_create_psi_synth): element (i, c) is a SplitMix64 hash of (seed, i, global
column c), uniform in [-1, 1]^2 and scaled by sqrt(3 / (2 N)).  Used by the tests and by bench.py's
parity checks to know, on the host, what a device-generated block holds (checker only)."""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z):
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_block(N, M, col0=0, seed=1234, dtype=np.complex128):
    """The N x M block lm_state_create_psi_synth(ctx, N, M, col0, seed) generates (Fortran order)."""
    with np.errstate(over="ignore"):
        i = np.arange(N, dtype=np.uint64)[:, None]
        c = (np.arange(M, dtype=np.uint64) + np.uint64(col0))[None, :]
        k = i * np.uint64(4294967311) + c
        h1 = _splitmix64(k ^ (np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)))
        h2 = _splitmix64(h1)
    scale = np.sqrt(1.5 / N)
    re = (h1 >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0
    im = (h2 >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0
    out = np.empty((N, M), dtype=np.complex128, order="F")
    out.real = re * scale
    out.imag = im * scale
    return out.astype(dtype, order="F") if dtype != np.complex128 else out
