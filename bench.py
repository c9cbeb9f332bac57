#!/usr/bin/env python
"""Benchmark of the Psi-block unitary-evolution hot path (BASELINE.json metric:
"evolution steps/sec (N x Nocc Psi block)"; SpMM HBM GB/s vs peak).

    python bench.py --gpus N --steps K --warmup W [--workload c4|c3|c2] [--impl reference]

The default workload is the north-star headline: config 4, the Haldane model on a 500 x 500
honeycomb lattice (N = 5e5), a 4096-column Psi block, complex128.

One rank per GPU (torchrun sets RANK / LOCAL_RANK / WORLD_SIZE); rank 0 prints ONE JSON line.
A "step" is one application of exp(-i H dt) to the whole Psi block (all ranks' column shards).
Strong scaling: the block is fixed, its columns are sharded over the ranks, H is replicated.
  value     : steps/s with Psi resident in HBM (K lm_step calls, CUDA events, max over ranks)
  e2e       : steps/s through the C ABI with HOST buffers every step: H values uploaded from
              pinned host memory (host-assembled time-dependent H path; N > 1: rank 0 uploads, the
              other ranks receive them over NVLink, lm_ham_update_values_bcast), step, fused
              localdensity + bond currents reduced (all-reduced for N > 1) and copied back
  secondary : the default run (c4) also carries configs 3 and 2 as sub-objects (same legs and checks)
  roofline  : the dominant kernel (k_apply_stencil_tma = lattice-stencil SpMM fused with one
              product-form propagator factor): algorithmic bytes per launch / average launch
              duration vs the measured HBM copy bandwidth
  parity_check : in-run self-checks on the bench's own state - trace identity on the reduced frame,
              sharded (all ranks) vs unsharded (rank 0 alone) evolution + observables of a check
              block, and 8 columns against a host Taylor series of exp(-i H dt)
  cpu_baseline / --impl reference : the reference's CPU algorithm (one KrylovKit-style Lanczos
              exponentiate per ket, oracle/cpu_ref.c) on the host cores, bounded column sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, torchrun notices)
# write to fd 1 behind Python's back, so fd 1 is pointed at stderr for the whole run and the JSON
# line goes to a private duplicate of the original stdout.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(obj):
    _REAL_STDOUT.write(json.dumps(obj) + "\n")
    _REAL_STDOUT.flush()

import numpy as np  # noqa: E402


# ------------------------------------------------------------------------------------ workloads
def workload(name):
    """Synthetic inputs of SURVEY.md section 8(d).  Returns dict(H builder, N, M, dt, label)."""
    import lm_b200 as lm
    if name == "c2":
        lat = lm.SquareLattice(100, 100)
        return dict(label="c2: SquareLattice(100,100) tight-binding, constant H, N=1e4, M=5000 (half filling), complex128, dt=0.1",
                    ham=lambda t: lm.tightbinding_hamiltonian(lat), M=5000, dt=0.1, time_dependent=False)
    if name == "c3":
        lat = lm.SquareLattice(300, 300)
        return dict(label="c3: QWZ m=1 on SquareLattice(300,300), Landau field ramp regenerated on device each step, N=1.8e5, M=4096, complex128, dt=0.1",
                    ham=lambda t: lm.qwz(lat, field=lm.LandauGauge(0.1 * min(t, 10.0) / 10.0)), M=4096, dt=0.1, time_dependent=True)
    if name == "c4":
        lat = lm.HoneycombLattice(500, 500)
        return dict(label="c4: Haldane t1=1 t2=0.2 m=0.1 on HoneycombLattice(500,500), N=5e5, M=4096, complex128, dt=0.1",
                    ham=lambda t: lm.haldane(lat, 1.0, 0.2, 0.1), M=4096, dt=0.1, time_dependent=False)
    if name == "c1":
        lat = lm.SquareLattice(10, 10)
        return dict(label="c1: README example - SquareLattice(10,10), PointFlux ramp 0.2 min(t,10)/10 at (5.5,5.5), dense density matrix densitymatrix(mu=0), N=100, complex128, dt=0.1",
                    ham=lambda t: lm.tightbinding_hamiltonian(lat, field=lm.PointFlux(0.2 * min(t, 10.0) / 10.0, (5.5, 5.5))), M=100, dt=0.1, time_dependent=True, dense=True)
    if name == "c1x":
        lat = lm.SquareLattice(64, 32)
        return dict(label="c1x: dense-P path at N=2048 - SquareLattice(64,32) tight-binding, constant H, dense density matrix of 1024 occupied synthetic orbitals, complex128, dt=0.1",
                    ham=lambda t: lm.tightbinding_hamiltonian(lat), M=2048, dt=0.1, time_dependent=False, dense=True)
    if name == "tiny":
        lat = lm.SquareLattice(20, 20)
        return dict(label="tiny: SquareLattice(20,20), M=64 (harness self-test)",
                    ham=lambda t: lm.tightbinding_hamiltonian(lat), M=64, dt=0.1, time_dependent=False)
    raise SystemExit("unknown workload %r" % name)


def synth_block(N, M, seed):
    """Random complex block with unit-norm columns (seeded; values do not affect timing)."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((N, M)) + 1j * rng.standard_normal((N, M))
    a /= np.linalg.norm(a, axis=0, keepdims=True)
    return np.asfortranarray(a)


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except Exception:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU baseline
def cpu_reference_rate(wl, n_sample_cols, n_steps, seed=99):
    """Times the restated reference CPU algorithm (KrylovKitExp semantics: one Lanczos
    exponentiate per ket, src/evolution.jl:150-154,245-247, + localdensity per frame) on a
    bounded column sample with all host threads; returns (steps/s of the FULL block, info)."""
    from oracle import cpu_ref
    H = wl["ham"](0.0).data
    ham = cpu_ref.CsrHam(H)
    N, M = H.shape[0], wl["M"]
    cores = cpu_ref.max_threads()
    ns = min(n_sample_cols if n_sample_cols > 0 else (-n_sample_cols or 4) * cores, M)
    psi = synth_block(N, ns, seed)
    cpu_ref.krylov_block_step(ham, psi[:, :min(ns, 2 * cores)].copy(order="F"), wl["dt"])    # warm-up
    per_step = []
    for _ in range(n_steps):
        t0 = time.perf_counter()
        cpu_ref.krylov_block_step(ham, psi, wl["dt"])
        cpu_ref.localdensity(psi)
        per_step.append(time.perf_counter() - t0)
    t = float(np.median(per_step)) * (M / ns)           # scale the sample to the full block
    return 1.0 / t, per_step, dict(cores=cores, ns=ns, sample="%d of %d columns x %d steps (Lanczos krylovdim=30 tol=1e-12 per ket + localdensity), scaled by M/sample" % (ns, M, n_steps))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(args.workload)
    if args.M:
        wl["M"] = args.M
    rate, per_step, info = cpu_reference_rate(wl, args.ref_cols, args.steps + args.warmup)
    per_step = per_step[args.warmup:] or per_step
    M = wl["M"]
    t_sample = float(np.median(per_step))
    t = t_sample * (M / info["ns"])
    out = {"impl": "reference", "metric": "evolution steps/sec (N x Nocc Psi block)", "value": 1.0 / t, "unit": "steps/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "complex128", "data": "synthetic",
           "config": {"workload": wl["label"], "M_total": M, "sample_columns": info["ns"], "ms_per_step_sample": 1e3 * t_sample,
                      "note": "reference CPU algorithm restated in C (oracle/cpu_ref.c, kind=port): Julia is not installed, the reference itself cannot run; "
                              "each timed step evolves a bounded column sample (+ localdensity) and value / ms_per_step are scaled by M / sample_columns"},
           "cpu_baseline": {"value": 1.0 / t, "unit": "steps/s", "cores": info["cores"], "kind": "port", "sample": info["sample"]},
           "e2e": {"value": 1.0 / t, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


# ------------------------------------------------------------------------------------ in-run parity checks
def parity_check(lm, _lib, lib, torch, dist, ctx, dev, state, Hmat, H0, rho_frame, N, M, dt, args, rank, world, local, cdt):
    """Self-checks of THIS run's state (complex128 tolerances; none of it touches oracle/):
    (a) trace identity on the last reduced frame of the timed e2e leg:  sum_i rho_i == sum_c ||psi_c||^2
        (column norms from an independent kernel, summed over the ranks through torch.distributed);
    (b) N > 1: a check block of 32 columns per rank, sharded over all ranks, evolved 2 steps and reduced
        through the same exchange as the timed frames, against the SAME block evolved unsharded by rank 0
        alone on a second context without a communicator - rho and J;
    (c) 8 columns of that block after 2 steps against a host Taylor series of exp(-i H dt) (scipy sparse
        matrix-vector products on the host-assembled H)."""
    import ctypes as C
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from synth import synth_block as synth_host
    out = {}
    n2 = float(state.column_norms2().sum())
    if world > 1:
        t = torch.tensor([n2], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        n2 = float(t.item())
    out["trace_rel"] = abs(float(rho_frame.sum()) - n2) / n2
    nmv = C.c_int32()
    method = {"auto": 0, "chebyshev": 1, "taylor": 2, "taylor_horner": 4, "chebyshev_clenshaw": 5}[args.method]
    Mc = 32 * world
    n_sites, npairs = N // H0.n_int, len(dev.pairs()[0])
    chk = lm.DeviceState.synthetic(N, Mc, ctx=ctx, seed=4321, lattice=H0.lattice, n_int=H0.n_int)
    for _ in range(2):
        _lib.check(lib.lm_step(dev.handle, chk.handle, dt, args.tol, method, C.byref(nmv)))
    rho, J = np.empty(n_sites), np.empty(max(npairs, 1))
    _lib.check(lib.lm_observables(dev.handle, chk.handle, _lib.ptr(rho), _lib.ptr(J)))
    if rank == 0:
        lat = H0.lattice
        dims = lat.sizes if len(lat) == lat.sizes[0] * lat.sizes[1] * lat.nb else None
        if world > 1:
            solo = lm.Context(device=local, precision=args.precision)          # no communicator: nothing is reduced
            dev1 = lm.DeviceHam.from_csc(solo, Hmat, H0.n_int, coords=lat.coords, lattice_dims=dims)
            full = lm.DeviceState.synthetic(N, Mc, ctx=solo, seed=4321, lattice=lat, n_int=H0.n_int)
            for _ in range(2):
                _lib.check(lib.lm_step(dev1.handle, full.handle, dt, args.tol, method, C.byref(nmv)))
            rho1, J1 = np.empty(n_sites), np.empty(max(npairs, 1))
            _lib.check(lib.lm_observables(dev1.handle, full.handle, _lib.ptr(rho1), _lib.ptr(J1)))
            out["sharded_vs_unsharded_rho_rel"] = float(np.abs(rho - rho1).max() / np.abs(rho1).max())
            out["sharded_vs_unsharded_J_rel"] = float(np.abs(J - J1).max() / max(np.abs(J1).max(), 1e-300)) if npairs else 0.0
            small_ctx, small_dev = solo, dev1
        else:
            small_ctx, small_dev = ctx, dev
        # (c) host Taylor series on 8 columns (first 8 of a 32-column device block: the stencil kernels need >= 32)
        s8 = lm.DeviceState.synthetic(N, 32, ctx=small_ctx, seed=4321, lattice=lat, n_int=H0.n_int, shard=False)
        for _ in range(2):
            _lib.check(lib.lm_step(small_dev.handle, s8.handle, dt, args.tol, method, C.byref(nmv)))
        got = s8.download()[:, :8].astype(np.complex128)
        X = synth_host(N, 8, 0, seed=4321).astype(cdt).astype(np.complex128)
        A = (-1j * dt) * Hmat.tocsr()
        for _ in range(2):
            term, acc, k = X.copy(), X.copy(), 0
            while k < 200:
                k += 1
                term = (A @ term) / k
                acc += term
                if np.abs(term).max() <= 1e-18 * max(np.abs(acc).max(), 1e-300):
                    break
            X = acc
        out["psi_vs_host_taylor_rel"] = float(np.abs(got - X).max() / np.abs(X).max())
        out["checked"] = "trace identity on the timed frame; %d-column check block, 2 steps%s; 8 columns vs host Taylor series at N = %d" % (
            Mc, ", sharded over %d ranks vs unsharded on rank 0 (rho, J)" % world if world > 1 else "", N)
        out["max_rel"] = max(v for k, v in out.items() if k.endswith("_rel"))
    if world > 1:
        dist.barrier()
    return out


# ------------------------------------------------------------------------------------ the reference's published workload
def disc_hamiltonian(n):
    """benchmarks/models.jl:3-12: graphene disc of ~n sites (HoneycombLattice filtered by a circle centred at
    the origin), t1 = -2.8, p-n junction potential 0.2 inside r/2, SymmetricGauge(B 1.519e-3).  Returns
    (lattice, t -> Hamiltonian) with the ramp of benchmarks/benchmark_evolution_dynamic.jl:7: B(t) = 3 min(t/5, 1)."""
    import math
    import lm_b200 as lm
    from importlib import import_module
    LT = import_module("lm_b200.lattices")
    r = math.sqrt(n * (math.sqrt(3) / 4) / math.pi)            # area per site = sqrt(3)/4
    n2 = int(math.ceil(2 * r / (math.sqrt(3) / 2))) + 3
    n1 = int(math.ceil(2 * r + n2 / 2)) + 3
    a = np.array([[1.0, 0.5], [0.0, math.sqrt(3) / 2]])
    basis = np.array([[0.0, 0.5], [0.0, math.sqrt(3) / 6]])
    centre = a[:, 0] * (n1 + 1) / 2 + a[:, 1] * (n2 + 1) / 2 + basis.mean(axis=1)
    lat = LT.BravaisLattice(a, basis - centre[:, None], (n1, n2), predicate=lambda x, y: x * x + y * y < r * r, kind="HoneycombLattice")
    pot = np.where(lat.x ** 2 + lat.y ** 2 < r * r / 4, 0.2, 0.0)
    return lat, (lambda t: lm.construct_hamiltonian(lat, (-2.8, lm.NearestNeighbor(1)), (1, pot),
                                                    field=lm.SymmetricGauge(3.0 * min(t / 5.0, 1.0) * 1.519e-3)))


def main_published(args):
    """BASELINE.md section 1, the only first-party number of the reference: a single Ket on a graphene disc,
    Evolution over 0:0.1:10 with a ramped SymmetricGauge (H changes every step), localdensity per frame through
    TimeSequence; total wall seconds of the whole loop, as benchmarks/benchmark_evolution_dynamic.jl:3-13 times it.
    Through the public host API (Evolution / TimeSequence / localdensity); the filtered lattice runs on the ELL kernels."""
    import torch
    import lm_b200 as lm
    n = 10 ** int(args.workload[-1])
    published = {4: 1.43, 5: 12.7}[int(args.workload[-1])]
    torch.cuda.set_device(0)
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    ctx = lm.Context(device=0, precision=args.precision, stream=tstream.cuda_stream)
    lat, h = disc_hamiltonian(n)
    N = len(lat)
    rng = np.random.default_rng(1234)
    psi0 = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    psi0 /= np.linalg.norm(psi0)
    ts = np.arange(0, 101) * 0.1

    def run_once():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev = lm.Evolution(lm.B200Exp(tol=args.tol, method=args.method, ctx=ctx), h, psi0)
        dens = lm.TimeSequence(lambda m: lm.localdensity(m.state).values, ev, ts)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, dens
    for _ in range(args.warmup):
        run_once()
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.2)
    l0 = ctx.launch_count()
    t_a = time.time()
    runs = [run_once() for _ in range(args.steps)]
    t_b = time.time()
    launches = ctx.launch_count() - l0
    clocks = sampler.stop(t_a, t_b)
    secs = np.array([r[0] for r in runs])
    dens = runs[-1][1]
    # parity of the last frame: the ket evolved on the host by a Taylor series of exp(-i H(t_k) dt)
    import scipy.sparse as sp
    x = psi0.copy()
    for k in range(100):
        A = (-0.1j) * sp.csr_matrix(h(ts[k]).data)
        term, acc, q = x.copy(), x.copy(), 0
        while q < 200:
            q += 1
            term = (A @ term) / q
            acc += term
            if np.abs(term).max() <= 1e-18 * np.abs(acc).max():
                break
        x = acc
    rho_ref = np.abs(x) ** 2
    rho = dens[ts[-1]]
    total = float(np.median(secs))
    H0 = h(0.0)
    dev = H0.device(ctx)
    esz = 16 if args.precision == "c128" else 8
    out = {"metric": "total seconds, Evolution over 0:0.1:10 (100 steps, H re-evaluated every step) + localdensity per frame, single Ket, graphene disc (the reference's published workload)",
           "value": total, "unit": "s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / 100, "higher_is_better": False,
           "scaling": "replicas only", "vs_baseline": total / published, "dtype": "complex128" if esz == 16 else "complex64", "data": "synthetic",
           "config": {"workload": "%s: graphene disc, %d sites (target %d), t1 = -2.8, p-n potential, SymmetricGauge ramp, one Ket" % (args.workload, N, n),
                      "published_seconds": published, "published_source": "BASELINE.md section 1 (benchmark_evolution_dynamic.jl, KrylovKitExp, Xeon 8259CL, 2 threads)",
                      "best_seconds": float(secs.min()), "runs_timed": int(args.steps), "steps_per_second": 100.0 / total, "nnz": int(dev.nnz), "tol": args.tol,
                      "l2": "single ket: everything is L2 resident; the loop is launch / host-latency bound, not bandwidth bound"},
           "clocks": clocks, "gpu_launches": int(launches),
           "e2e": {"value": total, "unit": "s", "h2d_bytes_per_step": 24, "d2h_bytes_per_step": int(8 * N),
                   "what": "the loop IS end to end: per step 3 field parameters to the device (phases regenerated there), lm_step, lm_local_density -> host"},
           "roofline": {"bound": "hbm", "kernel": "lm::k_apply (ELL gather, single ket)", "achieved": None, "peak": None, "unit": "GB/s", "frac": None, "traffic": None,
                        "note": "latency-bound regime (N = %d, one column): no roofline claim" % N},
           "parity_check": {"rho_last_frame_rel": float(np.abs(rho - rho_ref).max() / rho_ref.max()), "norm": float(rho.sum()),
                            "max_rel": float(np.abs(rho - rho_ref).max() / rho_ref.max()), "checked": "localdensity of the last frame vs a host Taylor-series evolution of the ket"}}
    emit(out)


# ------------------------------------------------------------------------------------ dense-P path (config 1)
def dense_inputs(wl, H0):
    """Initial dense density matrix: c1 = the README's densitymatrix(H(0), mu = 0) (host eigh, N = 100);
    c1x = P = X X' of 1024 seeded synthetic orbitals (tests/synth.py)."""
    import lm_b200 as lm
    N = H0.structure.dim
    if N <= 512:
        proj = lm.densitymatrix(H0, mu=0.0)
        return proj.dense(), proj.psi * np.sqrt(proj.weights if proj.weights is not None else 1.0)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from synth import synth_block as synth_host
    X = synth_host(N, N // 2, 0, seed=1234)
    return X @ X.conj().T, X


def cpu_dense_rate(wl, P0, n_steps):
    """CachedExp semantics on the host (src/evolution.jl:62-128, restated in oracle/evolution.py): U = myexp!(-i H dt)
    cached while H is unchanged, P <- U P U' as two dense products (numpy / BLAS threads) + localdensity."""
    from oracle import evolution as EV
    import scipy.sparse as sp
    per_step, t = [], 0.0
    P = P0.copy()
    U, Hprev = None, None
    for _ in range(n_steps):
        t0 = time.perf_counter()
        H = sp.csc_matrix(wl["ham"](t).data)
        if Hprev is None or (H != Hprev).nnz:
            U, _ = EV.myexp(H, -1j * wl["dt"], threshold=1e-12)
            U = U.toarray() if sp.issparse(U) else np.asarray(U)
            Hprev = H
        P = U @ P @ U.conj().T
        np.real(np.diag(P)).copy()
        per_step.append(time.perf_counter() - t0)
        t += wl["dt"]
    return 1.0 / float(np.median(per_step)), per_step


def main_dense(args):
    """U P U' on the FP64 tensor cores: value = device-resident steps/s, e2e = the README loop through the C ABI
    (H values from the host every step, localdensity + currents back), roofline bound = tensor."""
    import torch
    import ctypes as C
    import lm_b200 as lm
    from importlib import import_module
    _lib = import_module("lm_b200._lib")
    D = import_module("lm_b200.distributed")
    rank, world, local = D.env_rank()
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    ctx = lm.Context(device=local, precision=args.precision, stream=tstream.cuda_stream)
    lib = _lib.load()
    wl = workload(args.workload)
    H0 = wl["ham"](0.0)
    N, dt = H0.structure.dim, wl["dt"]
    P0, X0 = dense_inputs(wl, H0)
    esz = 16 if args.precision == "c128" else 8
    cdt = np.complex128 if esz == 16 else np.complex64
    state = lm.DeviceState.from_dense(P0, ctx=ctx, lattice=H0.lattice, n_int=H0.n_int)
    sol = lm.B200Exp(tol=args.tol, method=args.method, precision=args.precision, ctx=ctx)

    def timed(fn, n):
        torch.cuda.synchronize()
        t0 = time.time()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for k in range(n):
            fn(k)
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1), t0, time.time()

    tcur = [0.0]

    def dev_step(k):
        sol.update_solver(wl["ham"](tcur[0]), dt)
        sol.step(state)
        tcur[0] += dt
    for k in range(args.warmup):
        dev_step(k)
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    l0 = ctx.launch_count()
    ms, t0, t1 = timed(dev_step, args.steps)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop(t0, t1)
    ms_per_step = ms / args.steps
    # roofline: two complex GEMMs per step (8 N^3 real flops each in the 4-multiplication count) on the DMMA pipe
    flops = 2 * 8.0 * float(N) ** 3
    peak = C.c_double(0.0)
    lib.lm_dbg_dmma_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.lm_dbg_dmma_peak.restype = C.c_int32
    if os.environ.get("LM_EMUL_LIB"):
        peak.value = 40.0                     # CPU dry run of the harness (tools/bench_dryrun_cpu.py): no probe
    else:
        _lib.check(lib.lm_dbg_dmma_peak(ctx.handle, C.byref(peak)))
    achieved = flops / (ms_per_step * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "lm::k_zgemm_dmma_3m (complex GEMM on the FP64 tensor cores, mma.sync.m8n8k4.f64, 3M product)" if N > 256 else "lm::k_zgemm_dmma (complex GEMM on the FP64 tensor cores, 32 x 32 tiles)",
                "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s", "frac": achieved / peak.value, "traffic": None,
                "peak_source": "measured in this run: register-resident DMMA chains (lm_dbg_dmma_peak); nominal B200 FP64 tensor peak 40 TFLOP/s",
                "flops_per_step": flops, "note": "flops counted as 2 GEMMs x 8 N^3 (4-multiplication complex product); the 3M kernel executes 3/4 of them; the whole step (incl. rebuilding U when H changed) is in the time"}

    # e2e: README loop through the C ABI with host buffers
    Hmat = H0.data
    lat = H0.lattice
    dims = lat.sizes if len(lat) == lat.sizes[0] * lat.sizes[1] * lat.nb else None
    dev = lm.DeviceHam.from_csc(ctx, Hmat, H0.n_int, coords=lat.coords, lattice_dims=dims)
    nnz = dev.nnz
    npairs, n_sites = len(dev.pairs()[0]), N // H0.n_int
    nz_all = [np.ascontiguousarray(wl["ham"](k * dt).data.data.astype(cdt)) for k in range(args.warmup + args.steps)] if wl["time_dependent"] else None
    nz_pinned = torch.from_numpy(np.ascontiguousarray(Hmat.data.astype(cdt)).view(np.float64 if esz == 16 else np.float32).copy()).pin_memory()
    nz_np = nz_pinned.numpy().view(cdt)
    rho_np, j_np = np.empty(n_sites), np.empty(max(npairs, 1))
    st2 = lm.DeviceState.from_dense(P0, ctx=ctx, lattice=lat, n_int=H0.n_int)
    nmv = C.c_int32()
    method = {"auto": 0, "chebyshev": 1, "taylor": 2, "taylor_horner": 4, "chebyshev_clenshaw": 5}[args.method]
    kk = [0]

    def e2e_step(k):
        if nz_all is not None:
            nz_np[:] = nz_all[kk[0]]
        kk[0] += 1
        _lib.check(lib.lm_ham_update_values(dev.handle, _lib.ptr(nz_np)))
        _lib.check(lib.lm_step(dev.handle, st2.handle, dt, args.tol, method, C.byref(nmv)))
        _lib.check(lib.lm_observables(dev.handle, st2.handle, _lib.ptr(rho_np), _lib.ptr(j_np)))
    for k in range(args.warmup):
        e2e_step(k)
    ms_e2e, _, _ = timed(e2e_step, args.steps)
    e2e = {"value": 1e3 / (ms_e2e / args.steps), "unit": "steps/s", "h2d_bytes_per_step": int(nnz * esz), "d2h_bytes_per_step": int(8 * (n_sites + npairs)),
           "what": "lm_ham_update_values(host nzval) + lm_step (U P U') + lm_observables (rho, J of the dense P -> host) per step"}
    # parity: the e2e state against the host: orbitals evolved by a Taylor series of exp(-i H(t_k) dt), P = X W X'
    import scipy.sparse as sp
    w = np.ones(X0.shape[1])
    X = X0.astype(np.complex128)
    for k in range(args.warmup + args.steps):
        A = (-1j * dt) * sp.csr_matrix(wl["ham"](k * dt).data)
        term, acc, q = X.copy(), X.copy(), 0
        while q < 200:
            q += 1
            term = (A @ term) / q
            acc += term
            if np.abs(term).max() <= 1e-18 * np.abs(acc).max():
                break
        X = acc
    Pref = (X * w) @ X.conj().T
    got = st2.download().astype(np.complex128)
    rho_ref = np.real(np.diag(Pref)).reshape(n_sites, H0.n_int).sum(1)
    parity = {"P_vs_host_taylor_rel": float(np.abs(got - Pref).max() / np.abs(Pref).max()),
              "rho_rel": float(np.abs(rho_np - rho_ref).max() / np.abs(rho_ref).max()),
              "checked": "dense P after %d e2e steps vs orbitals evolved by a host Taylor series (P = X X'); localdensity of the last frame" % (args.warmup + args.steps)}
    parity["max_rel"] = max(parity["P_vs_host_taylor_rel"], parity["rho_rel"])
    out = {"metric": "evolution steps/sec (dense N x N density matrix, U P U')", "value": 1e3 / ms_per_step, "unit": "steps/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "replicas only", "vs_baseline": None,
           "dtype": "complex128" if esz == 16 else "complex64", "data": "synthetic",
           "config": {"workload": wl["label"], "N": N, "nnz": int(nnz), "dt": dt, "tol": args.tol, "method": args.method,
                      "l2": "N = %d: the three N x N buffers (%.1f MB) %s" % (N, 3 * N * N * esz / 1e6, "exceed" if 3 * N * N * esz > 126e6 else "fit in the 126 MB L2 (compute-bound path; no flush)")},
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "parity_check": parity}
    if rank == 0 and not args.no_cpu_baseline:
        rate, per_step = cpu_dense_rate(wl, P0, 3)
        out["cpu_baseline"] = {"value": rate, "unit": "steps/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "3 steps of the CachedExp algorithm (myexp! + two dense products, numpy BLAS) on the full N = %d matrix" % N}
    if rank == 0:
        emit(out)


# ------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--tol", type=float, default=1e-12)
    ap.add_argument("--method", default="auto")
    ap.add_argument("--precision", default="c128")
    ap.add_argument("--ref-cols", type=int, default=0, help="columns of the reference arm's sample (0 = 4 per host thread)")
    ap.add_argument("--cpu-cols", type=int, default=0, help="columns of the cpu_baseline sample (0 = 8 per host thread)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--M", type=int, default=0, help="override the block width")
    ap.add_argument("--no-secondary", action="store_true", help="default workload only: skip the c3 / c2 sub-objects")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.workload in ("pub4", "pub5"):
        return main_published(args)
    if workload(args.workload).get("dense"):
        return main_dense(args)

    import torch
    import torch.distributed as dist
    import ctypes as C
    import lm_b200 as lm
    from importlib import import_module
    _lib = import_module("lm_b200._lib")
    D = import_module("lm_b200.distributed")

    rank, world, local = D.env_rank()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # a dedicated (non-default) stream: the default stream's handle is 0, which the C ABI reads as
    # "create your own"; torch events must be recorded on the stream the kernels launch on
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    ctx = lm.Context(device=local, precision=args.precision, stream=stream)
    if world > 1:
        D.attach_communicator(ctx)

    lib = _lib.load()

    def run_block(wname, steps, warmup):
        """One Psi-block workload on the already initialised context: the `value` leg, the roofline of the dominant
        kernel, the `e2e` leg and the in-run parity checks.  Returns (bench line, workload dict)."""
        wl = workload(wname)
        if args.M:
            wl["M"] = args.M
        H0 = wl["ham"](0.0)
        N, M, dt = H0.structure.dim, wl["M"], wl["dt"]
        b, e = lm.shard_range(M, rank, world)
        Ml = e - b
        esz = 16 if args.precision == "c128" else 8
        cdt = np.complex128 if args.precision == "c128" else np.complex64
        # the block is generated on the device (a 32.8 GB host block would take minutes): this rank's
        # column shard [b, e) of the seeded N x M block (tests/synth.py restates the generator)
        state = lm.DeviceState.synthetic(N, M, ctx=ctx, seed=1234, lattice=H0.lattice, n_int=H0.n_int)
        assert state.col_range == (b, e) and state.M == Ml
        sol = lm.B200Exp(tol=args.tol, method=args.method, precision=args.precision, ctx=ctx)
        lib = _lib.load()

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def timed(fn, n, tail=None):
            barrier()
            t0 = time.time()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for k in range(n):
                fn(k)
            if tail is not None:
                tail()
            ev1.record()
            barrier()
            ms = ev0.elapsed_time(ev1)
            if world > 1:
                t = torch.tensor([ms], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms, t0, time.time()

        # ---------------- device-resident steps (the headline `value`) ----------------
        tcur = [0.0]

        def dev_step(k):
            sol.update_solver(wl["ham"](tcur[0]), dt)      # device phase regeneration if time dependent
            sol.step(state)
            tcur[0] += dt
        for k in range(warmup):
            dev_step(k)
        sampler = ClockSampler(local)
        sampler.start()
        time.sleep(0.3)
        l0 = ctx.launch_count()
        ms, t0, t1 = timed(dev_step, steps)
        launches = ctx.launch_count() - l0
        clocks = sampler.stop(t0, t1)
        K = sol.n_matvec
        ms_per_step = ms / steps
        value = 1e3 / ms_per_step

        # ---------------- roofline of the dominant kernel ----------------
        dev = sol.dev
        nnz = dev.nnz
        bytes_spmm = 2.0 * N * Ml * esz + nnz * (esz + 4) + 4.0 * (N + 1)
        n_apply = steps * K
        avg_launch_ms = ms / max(n_apply, 1)            # the step is K back-to-back k_apply launches
        achieved = bytes_spmm / (avg_launch_ms * 1e-3) / 1e9
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_file):
            peak, peak_src = float(json.load(open(peaks_file))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        traffic = None                                   # DRAM bytes are not measured inside this run (ncu captures: profiles/)
        sid = C.c_int32(-1)
        lib.lm_dbg_stencil_info.argtypes = [C.c_void_p] * 5
        lib.lm_dbg_stencil_info(dev.handle, C.byref(sid), None, None, None)
        in_class = C.c_int32(0)
        if sid.value >= 0:
            lib.lm_dbg_stencil_ri_state.argtypes = [C.c_void_p, C.c_void_p]
            lib.lm_dbg_stencil_ri_state(dev.handle, C.byref(in_class))
        ri_on = bool(in_class.value) and int(os.environ.get("LM_STENCIL_RI", "1") or 0) != 0
        kernel = ("lm::k_apply_stencil_tma (fused lattice-stencil SpMM + one product-form propagator factor; TMA-staged patch, register-tiled unit cells, shared value loads for Hermitian H, %s)"
                  % ("real / imaginary value class: scalar values, two FMAs per element" if ri_on else "complex values")
                  if (sid.value >= 0 and state.M >= 32) else "lm::k_apply / k_apply_rows (ELL gather SpMM fused with one product-form propagator factor)")
        roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "bytes_per_launch": bytes_spmm, "launches_timed": n_apply, "avg_launch_ms": avg_launch_ms,
                    "K_matvec_per_step": K}

        # ---------------- end to end through the C ABI with host buffers ----------------
        Hmat = H0.data                                    # host-assembled CSC (what t -> H(t) returns)
        lat = H0.lattice
        dims = lat.sizes if len(lat) == lat.sizes[0] * lat.sizes[1] * lat.nb else None   # unfiltered: rows are cell-major
        csc_dev = lm.DeviceHam.from_csc(ctx, Hmat, H0.n_int, coords=lat.coords, lattice_dims=dims)
        nz_pinned = torch.empty(nnz * (2 if esz == 16 else 1), dtype=torch.float64 if esz == 16 else torch.complex64).pin_memory()
        nz_np = nz_pinned.numpy().view(cdt)
        nz_np[:] = Hmat.data.astype(cdt)
        npairs = len(csc_dev.pairs()[0])
        n_sites = N // H0.n_int
        rho_pinned = torch.empty(n_sites, dtype=torch.float64).pin_memory()
        j_pinned = torch.empty(max(npairs, 1), dtype=torch.float64).pin_memory()
        rho_np, j_np = rho_pinned.numpy(), j_pinned.numpy()
        nmv = C.c_int32()
        method = {"auto": 0, "chebyshev": 1, "taylor": 2, "taylor_horner": 4, "chebyshev_clenshaw": 5}[args.method]

        # every step: nzval host -> device, one propagation step, the frame [rho | J] device -> host through
        # the asynchronous frame sink (double-buffered: the copy of frame k overlaps step k + 1; every
        # frame is read back inside the timed region, the last one by `drain`)
        pending = []

        def drain():
            while pending:
                _lib.check(lib.lm_frame_wait(ctx.handle, pending.pop(0), _lib.ptr(rho_np), _lib.ptr(j_np)))

        slot = [0]
        bcast_values = int(os.environ.get("LM_BENCH_BCAST", "1") or 0) != 0

        def e2e_step(k):
            # H2D (pinned buffer; enclosure checked on the device).  N > 1: H(t) is the same on every rank, so rank 0 uploads it and the
            # others receive it over NVLink (lm_ham_update_values_bcast) instead of eight PCIe uploads through host memory
            if world > 1 and bcast_values:
                _lib.check(lib.lm_ham_update_values_bcast(csc_dev.handle, _lib.ptr(nz_np) if rank == 0 else None, 0))
            else:
                _lib.check(lib.lm_ham_update_values_async(csc_dev.handle, _lib.ptr(nz_np)))
            _lib.check(lib.lm_step(csc_dev.handle, state.handle, dt, args.tol, method, C.byref(nmv)))
            if len(pending) == 2:
                _lib.check(lib.lm_frame_wait(ctx.handle, pending.pop(0), _lib.ptr(rho_np), _lib.ptr(j_np)))   # D2H of frame k - 2 lands
            _lib.check(lib.lm_observables_async(csc_dev.handle, state.handle, slot[0], 1))
            pending.append(slot[0])
            slot[0] ^= 1
        for k in range(warmup):
            e2e_step(k)
        drain()
        ms_e2e, _, _ = timed(e2e_step, steps, tail=drain)
        e2e = {"value": 1e3 / (ms_e2e / steps), "unit": "steps/s", "h2d_bytes_per_step": int(nnz * esz),
               "d2h_bytes_per_step": int(8 * (n_sites + npairs)),
               "what": ("lm_ham_update_values_bcast(pinned nzval on rank 0, NVLink broadcast)" if (world > 1 and bcast_values) else "lm_ham_update_values_async(pinned nzval)")
                       + " + lm_step + lm_observables_async / lm_frame_wait (rho, J -> host, double-buffered) per step"}
        parity = parity_check(lm, _lib, lib, torch, dist, ctx, csc_dev, state, Hmat, H0, rho_np.copy(), N, M, dt, args, rank, world, local, cdt)

        out = {"metric": "evolution steps/sec (N x Nocc Psi block)", "value": value, "unit": "steps/s", "n_gpus": world,
               "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "complex128" if esz == 16 else "complex64", "data": "synthetic",
               "config": {"workload": wl["label"], "N": N, "M_total": M, "M_per_gpu": Ml, "nnz": int(nnz), "dt": dt, "tol": args.tol,
                          "method": args.method, "sharding": "Psi columns over %d GPU(s), H replicated" % world,
                          "schedule": {"pdl": int(os.environ.get("LM_STEP_PDL", "0") or 0), "launches_per_step": launches / max(steps, 1),
                                       "stencil_shared_value_loads": int(os.environ.get("LM_STENCIL_HERM", "1") or 0), "stencil_tensor_map_boxes": int(os.environ.get("LM_STENCIL_TMAP", "1") or 0),
                                       "stencil_value_class_scalars": int(ri_on), "stencil_l2_lookahead": os.environ.get("LM_STENCIL_PF", "auto (half a resident wave on multi-row stencils)")},
                          "l2": "inputs larger than L2 (3 x %.0f MB Psi buffers per GPU); no flush" % (N * Ml * esz / 1e6)},
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "parity_check": parity}

        import gc
        del state, csc_dev, sol
        gc.collect()
        torch.cuda.synchronize()
        return out, wl

    out, wl = run_block(args.workload, args.steps, args.warmup)
    # continuity with round 1 and driver-run evidence for the other Psi-block configs: the default run (c4) also
    # carries config 3 (QWZ 300 x 300, Landau ramp regenerated on the device every step) and config 2 as sub-objects
    if args.workload == "c4" and not args.M and not args.no_secondary:
        out["secondary"] = {}
        for w in ("c3", "c2"):
            try:
                o, _ = run_block(w, args.steps, args.warmup)
                out["secondary"][w] = {"workload": o["config"]["workload"], "value": o["value"], "unit": o["unit"], "ms_per_step": o["ms_per_step"],
                                       "e2e": o["e2e"]["value"], "roofline_frac": o["roofline"]["frac"], "roofline_achieved_gbs": o["roofline"]["achieved"],
                                       "K_matvec_per_step": o["roofline"]["K_matvec_per_step"], "gpu_launches": o["gpu_launches"],
                                       "clocks": o["clocks"], "parity_max_rel": o["parity_check"].get("max_rel")}
            except Exception as ex:        # the headline line must not depend on the sub-objects
                out["secondary"][w] = {"error": repr(ex)[:300]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, per_step, info = cpu_reference_rate(wl, args.cpu_cols if args.cpu_cols > 0 else -8, 3)
        out["cpu_baseline"] = {"value": rate, "unit": "steps/s", "cores": info["cores"], "kind": "port", "sample": info["sample"]}
    if rank == 0:
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
