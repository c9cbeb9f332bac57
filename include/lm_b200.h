/*
 * lm_b200.h - C ABI of the B200-native unitary-evolution backend for LatticeModels.jl.
 *
 * The reference (aryavorskiy/LatticeModels.jl v1.0.7, pure Julia) has no FFI; its plug-in
 * surface for this path is Julia multiple dispatch on the `EvolutionSolver` interface
 * (src/evolution.jl:25-33).  The entry points below are exactly what a Julia `ccall` glue
 * (julia/B200Backend.jl) binds; every function cites the reference
 * interface it replaces (paths relative to the reference repository).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes only.  One host thread per context.
 *  - every call returns int32 status (LM_OK = 0); lm_last_error() gives the message of the
 *    last failing call on the calling thread (Julia side: throw(ArgumentError(msg)),
 *    mirroring src/evolution.jl:152,239).
 *  - complex numbers are interleaved (re, im): double[2] for precision LM_C128 (Julia
 *    ComplexF64 / numpy complex128), float[2] for LM_C64.
 *  - matrices handed over by the host are COLUMN-major (Julia / Fortran order).
 *  - `index_base` is 1 for Julia arrays, 0 for C/numpy; it applies to every index array of
 *    the call (colptr, rowval, site indices).
 *  - host pointers are borrowed for the duration of the call; the library owns all device
 *    memory.  Calls are synchronous with respect to host buffers (results are complete on
 *    return); device work that produces no host result is only enqueued on the context's
 *    stream.
 *  - one process drives one GPU.  Multi-GPU = one process per GPU, Psi columns sharded by
 *    the host with lm_shard_range(), H replicated; the only collective is the per-frame
 *    all-reduce inside lm_local_density / lm_observables (NCCL, opened lazily with dlopen).
 */
#ifndef LM_B200_H
#define LM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LM_OK                0
#define LM_ERR_INVALID       1   /* bad argument (Julia: ArgumentError) */
#define LM_ERR_CUDA          2   /* CUDA runtime failure */
#define LM_ERR_NOT_CONVERGED 3   /* propagator did not reach `tol` (src/evolution.jl:152) */
#define LM_ERR_NCCL          4
#define LM_ERR_UNSUPPORTED   5

#define LM_C128 0                /* complex128 arithmetic (default, parity 1e-10) */
#define LM_C64  1                /* complex64 arithmetic (optional mode, parity 1e-5) */

/* lm_step `method` */
#define LM_METHOD_AUTO      0    /* cheapest plan by HBM-stream count (normally product-form Chebyshev) */
#define LM_METHOD_CHEBYSHEV 1    /* Chebyshev approximant of exp(-iH dt) on the Gershgorin interval, applied in
                                    product form prod_j (I - Ht/x_j) (Leja-ordered roots): 2 HBM streams per term */
#define LM_METHOD_TAYLOR    2    /* truncated Taylor series (the polynomial myexp! sums, src/evolution.jl:93-128)
                                    applied in product form prod_j (I - A/r_j): 2 HBM streams per term */
#define LM_METHOD_LANCZOS   3    /* per-column Lanczos (KrylovKit.exponentiate semantics, src/evolution.jl:150-154) */
#define LM_METHOD_TAYLOR_HORNER 4 /* same polynomial in Horner form (3 streams per term; cross-check) */
#define LM_METHOD_CHEBYSHEV_CLENSHAW 5 /* same Chebyshev approximant by the Clenshaw recurrence (4 streams; fallback) */

/* gauge-field kinds for lm_ham_set_fields; each field owns 3 doubles of `params` */
#define LM_FIELD_LANDAU            1   /* (B, -, -)         src/zoo/magneticfields.jl:15  */
#define LM_FIELD_SYMMETRIC         2   /* (B, -, -)         src/zoo/magneticfields.jl:31  */
#define LM_FIELD_POINTFLUX_AXIAL   3   /* (flux, px, py)    src/zoo/magneticfields.jl:72-85 */
#define LM_FIELD_POINTFLUX_SINGULAR 4  /* (flux, px, py)    src/zoo/magneticfields.jl:89-104 */

typedef struct lm_ctx   lm_ctx;
typedef struct lm_ham   lm_ham;
typedef struct lm_state lm_state;

/* ------------------------------------------------------------------ library / errors */
int32_t     lm_version(void);
const char* lm_last_error(void);

/* ------------------------------------------------------------------ context
 * Replaces: nothing in the reference (it is single-process CPU).  `stream` may be NULL (the
 * library creates its own non-blocking stream) or an existing cudaStream_t handle so that a
 * host framework can time / order the work on its own stream. */
int32_t lm_ctx_create(int32_t device, int32_t precision, void* stream, lm_ctx** out);
int32_t lm_ctx_destroy(lm_ctx* ctx);
int32_t lm_ctx_synchronize(lm_ctx* ctx);
int32_t lm_ctx_stream(lm_ctx* ctx, void** stream_out);
/* number of kernels this context has launched so far (bench.py "gpu_launches") */
int32_t lm_ctx_launch_count(lm_ctx* ctx, int64_t* count_out);
/* CUDA-event timer on the context's stream */
int32_t lm_timer_start(lm_ctx* ctx);
int32_t lm_timer_stop(lm_ctx* ctx, double* elapsed_ms_out);

/* multi-GPU plumbing: rank 0 calls lm_comm_unique_id, the host broadcasts the 128 bytes by
 * any means (torch.distributed, MPI, Julia Distributed), every rank calls lm_ctx_comm_init. */
int32_t lm_comm_unique_id(void* id128_out);
int32_t lm_ctx_comm_init(lm_ctx* ctx, const void* id128, int32_t rank, int32_t nranks);
/* Optional NVLink peer-memory path for the per-frame reduction (2..8 ranks of ONE node): every
 * rank allocates a symmetric exchange buffer (slot_doubles >= n_sites + n_pairs of the largest
 * Hamiltonian) and exports its 64-byte CUDA IPC handle; the host all-gathers the handles (rank
 * order) and every rank attaches.  Afterwards lm_observables / lm_local_density push their
 * partial [rho | J] straight into the peers' slots from the finalize kernel and reduce locally;
 * without it (or if a frame does not fit the slot) the NCCL all-reduce is used. */
int32_t lm_ctx_peer_handle(lm_ctx* ctx, int64_t slot_doubles, void* handle64_out);
int32_t lm_ctx_peer_attach(lm_ctx* ctx, const void* all_handles /* nranks x 64 bytes */);
/* contiguous column range [begin, end) of rank `rank` out of `nranks` for M columns */
int32_t lm_shard_range(int64_t M, int32_t rank, int32_t nranks, int64_t* begin, int64_t* end);

/* ------------------------------------------------------------------ Hamiltonian
 * Replaces the host-resident `Hamiltonian.data::SparseMatrixCSC{ComplexF64,Int}`
 * (src/operators/system.jl:380-387) as the operand of update_solver! (src/evolution.jl:83-92,
 * 146-149).  N = Hilbert dimension = n_sites * n_int, composite index = internal index
 * fastest (src/operators/system.jl:16). */
int32_t lm_ham_create_csc(lm_ctx* ctx, int64_t N, int32_t n_int,
                          const int64_t* colptr, const int64_t* rowval, const void* nzval,
                          int32_t index_base, lm_ham** out);
/* same sparsity pattern, new values (time-dependent H assembled on the host: the parity
 * fallback for arbitrary closures t -> H(t), src/evolution.jl:43).  The values need not be
 * Hermitian for lm_spmm, but lm_step and the currents refuse a non-Hermitian operator
 * (max |H_ij - conj(H_ji)| > 1e-13 ||H||; 1e-6 in LM_C64): the propagators assume H = H'. */
int32_t lm_ham_update_values(lm_ham* ham, const void* nzval);
/* Same update without a host synchronisation (update_solver! inside a frame loop whose closure only
 * moves Peierls phases, src/evolution.jl:83-92,243): the copy, the scatter and a device-side check
 * are enqueued on the context's stream and the call returns.  `nzval` must stay valid and unchanged
 * until the next synchronising call on the context (lm_frame_wait, lm_ctx_synchronize,
 * lm_observables, a download).  The propagator keeps the spectral enclosure of the last
 * lm_ham_update_values; if the new values leave it, or are not Hermitian, a sticky error is raised
 * by that next synchronising call (LM_ERR_NOT_CONVERGED / LM_ERR_INVALID): re-do the frame with
 * lm_ham_update_values.  Peierls phases never change the Gershgorin bounds, so a gauge-field ramp
 * never trips it. */
int32_t lm_ham_update_values_async(lm_ham* ham, const void* nzval);
/* One process per GPU, H replicated: only rank `root` supplies host values (nzval may be NULL on the other ranks); the root uploads
 * them and every rank receives them over NVLink (ncclBroadcast), instead of one PCIe upload per rank.  Same asynchronous semantics as
 * lm_ham_update_values_async (sticky enclosure / Hermiticity flag).  Must be called by every rank of the communicator; without a
 * communicator it is lm_ham_update_values_async (root must be 0).  Replaces, for sharded runs, the per-process re-upload of
 * `update_solver!(solver, mat, dt, force)` (src/evolution.jl:83-92). */
int32_t lm_ham_update_values_bcast(lm_ham* ham, const void* nzval, int32_t root);

/* Device-resident time-dependent Hamiltonian (the AbstractTimeDependentOperator branch,
 * src/evolution.jl:44-47,243).  Restates OperatorBuilder.setindex! + expand_bond
 * (src/operators/builder.jl:282-309) on the device: for every directed bond b
 *     H[src_b block, dst_b block] += amp_b * f_b ;  H[dst_b block, src_b block] += amp_b' * conj(f_b)
 *     f_b = bfac_b * exp(-2 pi i * line_integral(field, r_src_b, r_dst_b))
 * with r_dst the UNWRAPPED destination coordinates and bfac the boundary phase
 * (src/core/boundaries.jl:280-289).  `amp` holds nb column-major n_int x n_int blocks,
 * `onsite` (nullable) n_sites column-major blocks added on the diagonal, r_* are (x, y)
 * pairs.  The Hermitian partner is NOT added for self-bonds (src == dst). */
int32_t lm_ham_create_bonds(lm_ctx* ctx, int64_t n_sites, int32_t n_int, int64_t nb,
                            const int32_t* src, const int32_t* dst,
                            const double* r_src, const double* r_dst,
                            const void* amp, const void* bfac, const void* onsite,
                            int32_t index_base, lm_ham** out);
/* FieldSum of nfields closed-form fields (src/operators/magneticfield.jl:100-104); params
 * has 3 doubles per field.  nfields = 0 is NoField.  Regenerates the stored values. */
int32_t lm_ham_set_fields(lm_ham* ham, int32_t nfields, const int32_t* kinds, const double* params);
/* same kinds, new parameters (a few doubles per step instead of a host re-assembly) */
int32_t lm_ham_set_field_params(lm_ham* ham, const double* params);

/* Site coordinates (n_sites (x, y) pairs, the lattice's own site_coords,
 * src/lattices/bravais/unitcell.jl:119-122): lets the library group rows into compact 2-D
 * patches for the TMA-staged SpMM.  Optional - without it the register-gather kernel is used. */
int32_t lm_ham_set_site_coords(lm_ham* ham, const double* xy);
/* Optional hint, to be given BEFORE lm_ham_set_site_coords: `rows` consecutive Hilbert rows form
 * one block that is kept together (default n_int = the orbitals of a site; a Bravais lattice
 * with an NB-site basis passes NB * n_int = all rows of a unit cell,
 * src/lattices/bravais/unitcell.jl:126-132 orders the basis index innermost).  Blocks of 2 rows
 * are processed by the site-blocked kernel. */
int32_t lm_ham_set_row_block(lm_ham* ham, int32_t rows);
/* Optional: the Hilbert rows are CELL-MAJOR on an n1 x n2 grid of Bravais unit cells - row =
 * ((j1 * n2) + j2) * RC + r with RC = N / (n1 n2) rows (basis sites x orbitals) per cell - which
 * is the reference's own order for an unfiltered lattice (src/lattices/bravais/lattice.jl:101-111:
 * last lattice axis fastest; src/lattices/bravais/unitcell.jl:126-132: basis index innermost).
 * If every entry couples cells at most one apart (periodic images included) and the pattern is
 * covered by a compiled stencil, H x runs on the register-tiled stencil kernel: one thread owns a
 * block of unit cells of one Psi column and loads every element of the haloed block once.
 * Anything else keeps the ELL kernels; the call never fails for an unsupported pattern. */
int32_t lm_ham_set_lattice_dims(lm_ham* ham, int32_t n1, int32_t n2);

int32_t lm_ham_dims(lm_ham* ham, int64_t* N, int32_t* n_int, int64_t* nnz, int32_t* ell_width);
/* CSC view of the current H (pattern + values), index_base as given at creation */
int32_t lm_ham_get_csc(lm_ham* ham, int64_t* colptr, int64_t* rowval, void* nzval);
/* Gershgorin bounds [emin, emax] of the spectrum used by the polynomial propagators */
int32_t lm_ham_spectral_bounds(lm_ham* ham, double* emin, double* emax);
/* Optional, NOT rigorous: tighten [emin, emax] with `iters` Lanczos steps on one random vector,
 * widened by margin x width and clamped to the Gershgorin interval (fewer polynomial terms when
 * phases/signs cancel, e.g. QWZ).  Reset by lm_ham_update_values. */
int32_t lm_ham_refine_bounds(lm_ham* ham, int32_t iters, double margin);
int32_t lm_ham_destroy(lm_ham* ham);

/* ------------------------------------------------------------------ states
 * Replaces the `.data` of the evolved `Ket` / density `Operator` (EvolutionStateType,
 * src/evolution.jl:36).  A Psi state is the block reformulation P = Psi diag(w) Psi'
 * (SURVEY.md section 8, Appendix B); M = 1 with w = NULL is a plain Ket.  Under multi-GPU
 * each rank passes ITS shard of the columns (lm_shard_range). */
int32_t lm_state_create_psi(lm_ctx* ctx, int64_t N, int64_t M, const void* psi_colmajor,
                            const double* weights, lm_state** out);
/* Synthetic block generated on the device (SURVEY.md section 8d inputs: uniform complex in
 * [-1, 1]^2, columns scaled to norm ~ 1): element (i, col0 + c) is a SplitMix64 hash of (seed, i,
 * col0 + c), so a column shard [col0, col0 + M) equals the same columns of the unsharded block.
 * Replaces nothing in the reference; bench.py and the full-size tests use it instead of a
 * multi-GB host upload. */
int32_t lm_state_create_psi_synth(lm_ctx* ctx, int64_t N, int64_t M, int64_t col0, uint64_t seed, lm_state** out);
/* squared norms of the local columns, M doubles (norm(ket)^2; for P = Psi diag(w) Psi':
 * tr P = sum_c w_c out[c] = sum_i localdensity_i) */
int32_t lm_state_column_norms2(lm_state* state, double* out);
/* dense density matrix P (N x N column-major): stepped as U P U' (src/evolution.jl:73-78) */
int32_t lm_state_create_dense(lm_ctx* ctx, int64_t N, const void* P_colmajor, lm_state** out);
int32_t lm_state_copy(lm_state* state, lm_state** out);           /* copy(state), src/evolution.jl:193 */
/* Multi-GPU: mark a state whose columns are the SAME on every rank (a single Ket evolved by
 * Evolution(solver, H, psi), an unsharded block): its densities / currents are complete on each
 * rank and are NOT summed across ranks.  Default 0 = the columns are this rank's shard. */
int32_t lm_state_set_replicated(lm_state* state, int32_t replicated);
int32_t lm_state_dims(lm_state* state, int64_t* N, int64_t* M, int32_t* is_dense);
int32_t lm_state_download_psi(lm_state* state, void* psi_colmajor_out);
/* dense P (N x N column-major); for a Psi state materialises Psi diag(w) Psi' (local columns
 * only) - the escape hatch for arbitrary operator algebra on the yielded state */
int32_t lm_state_download_dense(lm_state* state, void* P_colmajor_out);
int32_t lm_state_destroy(lm_state* state);

/* Lowest `nev` (<= 64) eigenpairs of H on the device: replaces
 * diagonalize(ham, :krylovkit; n) (src/spectrum.jl:56-64) / groundstate (src/spectrum.jl:205) for sizes
 * where the host LAPACK route is impossible.  Chebyshev-filtered subspace iteration on the SpMM kernels
 * of the propagator; converged when every residual ||H x_j - theta_j x_j|| <= tol * max|Gershgorin bound|.
 * evals_out: nev ascending eigenvalues; resid_out (nullable): their residual norms; vecs_out
 * (nullable): a Psi state with nev orthonormal columns (weights NULL), ready for lm_step /
 * lm_observables - the Fermi sphere of the nev lowest levels; iters_out (nullable): outer iterations.
 * max_iter <= 0: 300; degree <= 0: filter degree 40.  LM_ERR_NOT_CONVERGED if max_iter is exhausted
 * (evals_out / resid_out then hold the last Ritz values). */
int32_t lm_eigs_lowest(lm_ham* ham, int32_t nev, double tol, int32_t max_iter, int32_t degree,
                       double* evals_out, double* resid_out, lm_state** vecs_out, int32_t* iters_out);

/* ------------------------------------------------------------------ hot path
 * lm_step replaces step!(solver, state.data, cache) (src/evolution.jl:69-78,150-154):
 * Psi <- exp(-i H dt) Psi, or P <- U P U' for a dense state.  dt < 0 is allowed here (the
 * negative-time-step guard lives in step!(evol, dt), src/evolution.jl:239, host side).
 * n_matvec_out (nullable) receives the number of H applications executed (the K of
 * SURVEY.md section 8d). */
int32_t lm_step(lm_ham* ham, lm_state* state, double dt, double tol, int32_t method,
                int32_t* n_matvec_out);
/* Y = H X on device-resident states (bandwidth sweep / unit parity) and on host buffers */
int32_t lm_spmm_state(lm_ham* ham, lm_state* x, lm_state* y);
int32_t lm_spmm(lm_ham* ham, const void* X_colmajor, void* Y_colmajor, int64_t N, int64_t M);

/* ------------------------------------------------------------------ observables
 * localdensity (src/operators/latticeutils.jl:41-45): rho_i = sum_alpha Re P[i', i'].
 * All-reduced over ranks when a communicator is attached. */
int32_t lm_local_density(lm_state* state, int32_t n_int, double* rho_out /* n_sites */);
/* DensityCurrents over H's own site-level sparsity, pairs i < j ordered like findnz of a CSC
 * matrix filtered by I < J (src/currents.jl:173-177); curr[i,j] = sum_ab 2 Im(H[i',j'] P[j',i'])
 * (src/zoo/currents.jl:92-102).  lm_currents_pairs returns the static pair list. */
int32_t lm_currents_npairs(lm_ham* ham, int64_t* npairs);
int32_t lm_currents_pairs(lm_ham* ham, int32_t* I, int32_t* J);
/* fused pass: density (nullable) and all pair currents (nullable) from ONE read of Psi */
int32_t lm_observables(lm_ham* ham, lm_state* state, double* rho_out, double* J_out);
/* Asynchronous frame sink (TimeSequence collection, src/timesequence.jl:41-43, without a host
 * stall per frame): lm_observables_async enqueues the fused reductions of the CURRENT state and
 * the device->host copy of the frame [rho | J] into slot 0 or 1 and returns at once - the copy
 * runs on a second stream, so the following lm_step calls overlap it.  lm_frame_wait blocks until
 * that slot's frame has landed and copies it out (either pointer may be NULL).  A slot must be
 * waited for before it is enqueued again.  Multi-GPU: the frame is the rank-reduced one. */
int32_t lm_observables_async(lm_ham* ham, lm_state* state, int32_t slot, int32_t want_currents);
int32_t lm_frame_wait(lm_ctx* ctx, int32_t slot, double* rho_out, double* J_out);
/* Region sums on the device (src/currents.jl:85-109) so that only one number / one LatticeValue
 * crosses PCIe: masks are n_sites bytes (non-zero = in the region); dst_mask NULL = every site
 * outside src.  reuse_frame != 0 sums over the most recent currents frame of `ham`
 * (lm_observables / lm_observables_async with currents) instead of recomputing it - the caller
 * guarantees the state has not changed; `state` may then be NULL.
 *   lm_currents_fromto: sum_{i in src, j in dst} curr[i, j]
 *   lm_currents_from  : out[j] = (j in src) ? 0 : sum_{i in src} curr[i, j]   (n_sites doubles) */
int32_t lm_currents_fromto(lm_ham* ham, lm_state* state, const uint8_t* src_mask, const uint8_t* dst_mask,
                           int32_t reuse_frame, double* out);
int32_t lm_currents_from(lm_ham* ham, lm_state* state, const uint8_t* src_mask, int32_t reuse_frame, double* out);
/* Currents(curr, bonds) (src/currents.jl:238-255) on a host-given bond list */
int32_t lm_bond_currents(lm_ham* ham, lm_state* state, int64_t nb, const int32_t* I,
                         const int32_t* J, double* J_out);

/* ------------------------------------------------------------------ local operators
 * localexpect(op, state) (src/operators/latticeutils.jl:13-20):
 *   out_i = sum_{j,k} op[j,k] P[(i,k),(i,j)];  op = n_int x n_int column-major complex128,
 *   out = n_sites complex128.  All-reduced over ranks. */
int32_t lm_local_expect(lm_state* state, int32_t n_int, const void* op, void* out);
/* LocalOperatorCurrents(ham, state, op)[i,j] (src/zoo/currents.jl:150-184) for every site pair of
 * lm_currents_pairs:  J_p = sum_{a,b} 2 Im( (op T_ij)_{ab} P[j_b, i_a] ),  T_ij the H block. */
int32_t lm_operator_currents(lm_ham* ham, lm_state* state, const void* op, double* J_out);

#ifdef __cplusplus
}
#endif
#endif /* LM_B200_H */
