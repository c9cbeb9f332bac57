#!/usr/bin/env python
"""Benchmark of the Psi-block unitary-evolution hot path (BASELINE.json metric:
"evolution steps/sec (N x Nocc Psi block)"; SpMM HBM GB/s vs peak).

    python bench.py --gpus N --steps K --warmup W [--workload c2|c3|c4] [--impl reference]

One rank per GPU (torchrun sets RANK / LOCAL_RANK / WORLD_SIZE); rank 0 prints ONE JSON line.
A "step" is one application of exp(-i H dt) to the whole Psi block (all ranks' column shards).
Strong scaling: the block is fixed, its columns are sharded over the ranks, H is replicated.
  value     : steps/s with Psi resident in HBM (K lm_step calls, CUDA events, max over ranks)
  e2e       : steps/s through the C ABI with HOST buffers every step: H values uploaded from
              pinned host memory (host-assembled time-dependent H path), step, fused
              localdensity + bond currents reduced (all-reduced for N > 1) and copied back
  roofline  : the dominant kernel (k_apply = fused ELL SpMM + polynomial term): algorithmic
              bytes per launch / average launch duration vs the measured HBM copy bandwidth
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
    ns = min(n_sample_cols, M)
    psi = synth_block(N, ns, seed)
    cores = cpu_ref.max_threads()
    cpu_ref.krylov_block_step(ham, psi[:, :min(ns, 2 * cores)].copy(order="F"), wl["dt"])    # warm-up
    per_step = []
    for _ in range(n_steps):
        t0 = time.perf_counter()
        cpu_ref.krylov_block_step(ham, psi, wl["dt"])
        cpu_ref.localdensity(psi)
        per_step.append(time.perf_counter() - t0)
    t = float(np.median(per_step)) * (M / ns)           # scale the sample to the full block
    return 1.0 / t, per_step, dict(cores=cores, sample="%d of %d columns x %d steps (Lanczos krylovdim=30 tol=1e-12 per ket + localdensity), scaled by M/sample" % (ns, M, n_steps))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(args.workload)
    ns = args.ref_cols
    rate, per_step, info = cpu_reference_rate(wl, ns, args.steps + args.warmup)
    per_step = per_step[args.warmup:] or per_step
    M = wl["M"]
    t = float(np.median(per_step)) * (M / min(ns, M))
    out = {"impl": "reference", "metric": "evolution steps/sec (N x Nocc Psi block)", "value": 1.0 / t, "unit": "steps/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "complex128", "data": "synthetic",
           "config": {"workload": wl["label"], "note": "reference CPU algorithm restated in C (oracle/cpu_ref.c, kind=port): Julia is not installed, the reference itself cannot run"},
           "cpu_baseline": {"value": 1.0 / t, "unit": "steps/s", "cores": info["cores"], "kind": "port", "sample": info["sample"]},
           "e2e": {"value": 1.0 / t, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


# ------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--tol", type=float, default=1e-12)
    ap.add_argument("--method", default="auto")
    ap.add_argument("--precision", default="c128")
    ap.add_argument("--ref-cols", type=int, default=512)
    ap.add_argument("--cpu-cols", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--M", type=int, default=0, help="override the block width")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

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

    wl = workload(args.workload)
    if args.M:
        wl["M"] = args.M
    H0 = wl["ham"](0.0)
    N, M, dt = H0.structure.dim, wl["M"], wl["dt"]
    b, e = lm.shard_range(M, rank, world)
    Ml = e - b
    esz = 16 if args.precision == "c128" else 8
    cdt = np.complex128 if args.precision == "c128" else np.complex64
    psi = synth_block(N, Ml, 1234 + rank).astype(cdt, order="F")
    state = lm.DeviceState.from_psi(psi, None, ctx=ctx, lattice=H0.lattice, n_int=H0.n_int, shard=False)
    del psi
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
    for k in range(args.warmup):
        dev_step(k)
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    l0 = ctx.launch_count()
    ms, t0, t1 = timed(dev_step, args.steps)
    launches = ctx.launch_count() - l0
    # schedule of the timed steps: strip width in columns (0 = plain, -1 = fixed by LM_STEP_L2_MB)
    lib.lm_dbg_step_schedule.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    sched_cols, sched_cal = C.c_int64(0), C.c_int32(0)
    lib.lm_dbg_step_schedule(state.handle, C.byref(sched_cols), C.byref(sched_cal))
    clocks = sampler.stop(t0, t1)
    K = sol.n_matvec
    ms_per_step = ms / args.steps
    value = 1e3 / ms_per_step

    # ---------------- roofline of the dominant kernel ----------------
    dev = sol.dev
    nnz = dev.nnz
    bytes_spmm = 2.0 * N * Ml * esz + nnz * (esz + 4) + 4.0 * (N + 1)
    n_apply = args.steps * K
    avg_launch_ms = ms / max(n_apply, 1)            # the step is K back-to-back k_apply launches
    achieved = bytes_spmm / (avg_launch_ms * 1e-3) / 1e9
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        peak, peak_src = float(json.load(open(peaks_file))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic = None
    tf = os.path.join(ROOT, "profiles", "traffic.json")
    plain_schedule = launches <= args.steps * (K + 2)
    if os.path.exists(tf) and plain_schedule:        # the stored ncu traffic is that of the plain schedule
        traffic = json.load(open(tf)).get("%s_n%d" % (args.workload, world))
    roofline = {"bound": "hbm", "kernel": "lm::k_apply_stencil_tma (fused lattice-stencil SpMM + one product-form propagator factor; TMA-staged patch, register-tiled unit cells)", "achieved": achieved, "peak": peak,
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

    def e2e_step(k):
        _lib.check(lib.lm_ham_update_values(csc_dev.handle, _lib.ptr(nz_np)))                  # H2D
        _lib.check(lib.lm_step(csc_dev.handle, state.handle, dt, args.tol, method, C.byref(nmv)))
        if len(pending) == 2:
            _lib.check(lib.lm_frame_wait(ctx.handle, pending.pop(0), _lib.ptr(rho_np), _lib.ptr(j_np)))   # D2H of frame k - 2 lands
        _lib.check(lib.lm_observables_async(csc_dev.handle, state.handle, slot[0], 1))
        pending.append(slot[0])
        slot[0] ^= 1
    for k in range(args.warmup):
        e2e_step(k)
    drain()
    ms_e2e, _, _ = timed(e2e_step, args.steps, tail=drain)
    e2e = {"value": 1e3 / (ms_e2e / args.steps), "unit": "steps/s", "h2d_bytes_per_step": int(nnz * esz),
           "d2h_bytes_per_step": int(8 * (n_sites + npairs)),
           "what": "lm_ham_update_values(pinned nzval) + lm_step + lm_observables_async / lm_frame_wait (rho, J -> host, double-buffered) per step"}

    out = {"metric": "evolution steps/sec (N x Nocc Psi block)", "value": value, "unit": "steps/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "complex128" if esz == 16 else "complex64", "data": "synthetic",
           "config": {"workload": wl["label"], "N": N, "M_total": M, "M_per_gpu": Ml, "nnz": int(nnz), "dt": dt, "tol": args.tol,
                      "method": args.method, "sharding": "Psi columns over %d GPU(s), H replicated" % world,
                      "schedule": {"LM_STEP_L2_MB": os.environ.get("LM_STEP_L2_MB", "unset (plain)"), "online_choice_strip_cols": int(sched_cols.value),
                                   "pdl": int(os.environ.get("LM_STEP_PDL", "0") or 0), "launches_per_step": launches / max(args.steps, 1)},
                      "l2": "inputs larger than L2 (3 x %.0f MB Psi buffers per GPU); no flush" % (N * Ml * esz / 1e6)},
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, per_step, info = cpu_reference_rate(wl, args.cpu_cols, 3)
        out["cpu_baseline"] = {"value": rate, "unit": "steps/s", "cores": info["cores"], "kind": "port", "sample": info["sample"]}
    if rank == 0:
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
