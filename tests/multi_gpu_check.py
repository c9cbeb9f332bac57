#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29520 tests/multi_gpu_check.py

Every rank evolves ITS column shard of a Psi block (H replicated, device Peierls phases) and the
per-frame [rho | J] is combined across ranks - by the fused finalize + NVLink peer-memory
all-gather (default) or by NCCL (LM_OBS_P2P=0).  Each rank compares the combined frame with the
same computation done UNSHARDED on its own GPU (a second, communicator-less context) and with the
CPU oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import lm_b200 as lm  # noqa: E402
from importlib import import_module  # noqa: E402

D = import_module("lm_b200.distributed")
rank, world, local = D.env_rank()
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = lm.Context(device=local)
D.attach_communicator(ctx, peer_slot_doubles=1 << 16)
ref = lm.Context(device=local)                      # no communicator: unsharded reference

l = lm.HoneycombLattice(12, 10)
rng = np.random.default_rng(42)                      # identical on every rank
N, M = 2 * 12 * 10, 37
Psi = np.linalg.qr(rng.standard_normal((N, M)) + 1j * rng.standard_normal((N, M)))[0]
w = rng.random(M)
h = lambda t: lm.haldane(l, 1.0, 0.2, 0.1, field=lm.LandauGauge(0.02 * t))
sharded = lm.DeviceState.from_psi(Psi, w, ctx=ctx, lattice=l)            # shards columns by rank
full = lm.DeviceState.from_psi(Psi, w, ctx=ref, lattice=l, shard=False)
assert sharded.M == lm.shard_range(M, rank, world)[1] - lm.shard_range(M, rank, world)[0]
sol_s, sol_f = lm.B200Exp(tol=1e-13, ctx=ctx), lm.B200Exp(tol=1e-13, ctx=ref)
worst = 0.0
for k in range(12):
    H = h(0.1 * k)
    for sol, st in ((sol_s, sharded), (sol_f, full)):
        sol.update_solver(H, 0.1)
        sol.step(st)
    rho_s = lm.localdensity(sharded).values
    rho_f = lm.localdensity(full).values
    Js = lm.DensityCurrents(H, sharded).pair_values()[2]
    Jf = lm.DensityCurrents(H, full).pair_values()[2]
    worst = max(worst, np.abs(rho_s - rho_f).max() / np.abs(rho_f).max(), np.abs(Js - Jf).max() / max(np.abs(Jf).max(), 1e-300))
assert worst < 1e-12, worst
# oracle cross-check of the last frame on rank 0
if rank == 0:
    from oracle import evolution as EV, fields as F, lattice as L, observables as OB, operators as OP
    lo = L.honeycomb_lattice(12, 10)
    ho = lambda t: OP.haldane(lo, 1.0, 0.2, 0.1, field=F.LandauGauge(0.02 * t))
    X = Psi.copy()
    for k in range(12):
        X = EV.exact_propagator(ho(0.1 * k), 0.1) @ X
    want = OB.localdensity(OB.State(X, w, block=True), 1)
    assert np.abs(rho_s - want).max() < 1e-10 * np.abs(want).max()
import ctypes as C  # noqa: E402
_lib = import_module("lm_b200._lib").load()
_lib.lm_dbg_p2p_frames.argtypes = [C.c_void_p, C.c_void_p]
nfr = C.c_int64()
_lib.lm_dbg_p2p_frames(ctx.handle, C.byref(nfr))
if os.environ.get("LM_OBS_P2P", "1") != "0":
    assert nfr.value >= 24, "peer-memory path was not taken (%d frames)" % nfr.value     # 12 densities + 12 currents
else:
    assert nfr.value == 0
# a single ket (every rank holds the whole vector) must NOT be summed across ranks
ket = Psi[:, 0].copy()
ks, kf = lm.DeviceState.from_any(ket, ctx, l, 1), lm.DeviceState.from_any(ket, ref, l, 1)
assert ks.replicated and not kf.replicated
Hk = h(0.3)
for sol, st in ((sol_s, ks), (sol_f, kf)):
    sol.update_solver(Hk, 0.1)
    sol.step(st)
rk_s, rk_f = lm.localdensity(ks).values, lm.localdensity(kf).values
Jk_s, Jk_f = lm.DensityCurrents(Hk, ks).pair_values()[2], lm.DensityCurrents(Hk, kf).pair_values()[2]
assert abs(rk_f.sum() - 1.0) < 1e-12
worst = max(worst, np.abs(rk_s - rk_f).max() / np.abs(rk_f).max(), np.abs(Jk_s - Jk_f).max() / max(np.abs(Jk_f).max(), 1e-300))
assert worst < 1e-12, ("replicated ket", worst)
# lm_ham_update_values_bcast: only rank 0 supplies the host values of H(t) (the others pass NULL and hold garbage), every rank
# receives them over NVLink; same frames as synchronous per-rank updates on the unsharded reference
import scipy.sparse as sp  # noqa: E402
_L = import_module("lm_b200._lib")
Hs = [sp.csc_matrix(h(0.1 * k).data).astype(np.complex128) for k in range(5)]
dev_s = lm.DeviceHam.from_csc(ctx, Hs[0], 1, coords=l.coords, lattice_dims=l.sizes)
dev_f = lm.DeviceHam.from_csc(ref, Hs[0], 1, coords=l.coords, lattice_dims=l.sizes)
sb = lm.DeviceState.from_psi(Psi, w, ctx=ctx, lattice=l)
fb = lm.DeviceState.from_psi(Psi, w, ctx=ref, lattice=l, shard=False)
nmv = C.c_int32()
keep = []
for k in range(1, 5):
    nz = np.ascontiguousarray(Hs[k].data)
    keep.append(nz)
    _L.check(_lib.lm_ham_update_values_bcast(dev_s.handle, _L.ptr(nz) if rank == 0 else None, 0))
    _L.check(_lib.lm_step(dev_s.handle, sb.handle, 0.1, 1e-13, 0, C.byref(nmv)))
    _L.check(_lib.lm_ham_update_values(dev_f.handle, _L.ptr(nz)))
    _L.check(_lib.lm_step(dev_f.handle, fb.handle, 0.1, 1e-13, 0, C.byref(nmv)))
_L.check(_lib.lm_ctx_synchronize(ctx.handle))
b0, b1 = lm.shard_range(M, rank, world)
dev_b = np.abs(sb.download() - fb.download()[:, b0:b1]).max()
assert dev_b < 1e-13, ("broadcast value update", dev_b)
worst = max(worst, dev_b)
t = torch.tensor([worst], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("multi-GPU check OK: world=%d p2p=%s (%d peer-memory frames) worst rel deviation sharded vs unsharded = %.2e" % (world, os.environ.get("LM_OBS_P2P", "1"), nfr.value, t.item()))
dist.destroy_process_group()
