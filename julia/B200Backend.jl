# B200Backend.jl - Julia glue that plugs liblm_b200.so into LatticeModels.jl (v1.0.7).
#
# NOT EXECUTED in the build container (Julia is not installed there): every method below is a
# short forwarder to the C ABI of include/lm_b200.h, written against the reference sources cited on
# each definition (paths relative to the reference repository).  No reference file is edited:
# everything is new methods on the reference's own generic functions - multiple dispatch on the
# solver type and on the array type stored in `.data` is the reference's plug-in API
# (src/evolution.jl:25-33).  julia/runtests_b200.jl re-runs the reference's own testsets through it.
#
#   using LatticeModels, B200Backend
#   h(t) = tightbinding_hamiltonian(l, field = PointFlux(0.2 * min(t, 10) / 10, (5.5, 5.5)))
#   P0 = densitymatrix(h(0), mu = 0)
#   # (1) exactly the reference loop, dense P on the FP64 tensor cores (U P U'):
#   for (P, H, t) in Evolution(B200Exp(tol = 1e-12), h, P0)(0:0.1:20)
#       localdensity(P); Currents(DensityCurrents(H, P))
#   end
#   # (2) the same P as an occupied-orbital block that never leaves the device:
#   Pd = psi_densitymatrix(diagonalize(h(0)), mu = 0)            # Operator(basis, PsiProjector)
#   for (P, H, t) in Evolution(B200Exp(tol = 1e-12), h, Pd)(0:0.1:20)
#       localdensity(P); Currents(DensityCurrents(H, P))         # same calls, fused device reductions
#   end
module B200Backend

using LatticeModels, SparseArrays, LinearAlgebra
import LatticeModels: EvolutionSolver, EvolutionIterator, update_solver!, step!, evolution_cache,
                      localdensity, localexpect, DensityCurrents, LocalOperatorCurrents, Currents,
                      currentsfrom, currentsfromto, lattice, internal_length, internal_basis,
                      hasinternal, LatticeValue, TimeSequence, AbstractBonds, adapt_bonds, to_inds,
                      OneParticleBasis, CompositeLatticeBasis, AbstractEigensystem, densfun, FermiDirac
import QuantumOpticsBase
import QuantumOpticsBase: Operator, DataOperator, Ket, basis, check_samebases

const LIB = get(ENV, "LM_B200_LIB", joinpath(@__DIR__, "..", "latticemodels.jl_b200", "lib", "liblm_b200.so"))
const CURRENTS_EPS = 1e-10                                   # src/currents.jl:4

# ---- status -> ArgumentError (mirrors src/evolution.jl:152,239) ---------------------------------
lasterror() = unsafe_string(ccall((:lm_last_error, LIB), Cstring, ()))
check(status::Int32) = status == 0 || throw(ArgumentError(lasterror()))

# ---- context: one process drives one GPU --------------------------------------------------------
mutable struct Context
    handle::Ptr{Cvoid}
    rank::Int
    nranks::Int
    obsdev::Any                       # DeviceHam cache of the observables (see device_ham)
    function Context(; device::Integer = 0, precision::Symbol = :c128)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:lm_ctx_create, LIB), Int32, (Int32, Int32, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                    device, precision === :c64 ? 1 : 0, C_NULL, h))
        finalizer(c -> ccall((:lm_ctx_destroy, LIB), Int32, (Ptr{Cvoid},), c.handle), new(h[], 0, 1, nothing))
    end
end
const DEFAULT_CTX = Ref{Union{Nothing,Context}}(nothing)
default_context() = something(DEFAULT_CTX[], (DEFAULT_CTX[] = Context()))
synchronize(ctx::Context) = check(ccall((:lm_ctx_synchronize, LIB), Int32, (Ptr{Cvoid},), ctx.handle))

# Multi-GPU = one Julia process per GPU (Distributed / MPI.jl / torchrun-style launchers all work):
# rank 0 makes the 128-byte id, the host framework broadcasts it, every rank attaches.  Psi columns
# are sharded by the caller with shard_range; H is replicated; the only exchange is the per-frame
# [rho | J] reduction inside the observables (SURVEY.md section 8e).
function unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:lm_comm_unique_id, LIB), Int32, (Ptr{UInt8},), id)); id
end
function comm_init!(ctx::Context, id::Vector{UInt8}, rank::Integer, nranks::Integer)
    check(ccall((:lm_ctx_comm_init, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), ctx.handle, id, rank, nranks))
    ctx.rank, ctx.nranks = rank, nranks; ctx
end
# optional NVLink peer-memory exchange: handle = peer_handle(ctx, slot); all-gather the 64-byte
# handles in rank order; peer_attach!(ctx, vcat(handles...))
function peer_handle(ctx::Context, slot_doubles::Integer)
    h = Vector{UInt8}(undef, 64)
    check(ccall((:lm_ctx_peer_handle, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{UInt8}), ctx.handle, slot_doubles, h)); h
end
peer_attach!(ctx::Context, all_handles::Vector{UInt8}) =
    (check(ccall((:lm_ctx_peer_attach, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}), ctx.handle, all_handles)); ctx)
"1-based column range of `rank` (0-based) out of `nranks` for an M-column block"
function shard_range(M::Integer, rank::Integer, nranks::Integer)
    b = Ref{Int64}(0); e = Ref{Int64}(0)
    check(ccall((:lm_shard_range, LIB), Int32, (Int64, Int32, Int32, Ref{Int64}, Ref{Int64}), M, rank, nranks, b, e))
    (b[] + 1):e[]
end

# ---- device Hamiltonian: upload of a Julia SparseMatrixCSC (1-based indices accepted as is) ----
mutable struct DeviceHam
    handle::Ptr{Cvoid}
    colptr::Vector{Int64}
    rowval::Vector{Int64}
    n_int::Int
    pairs::Union{Nothing,Tuple{Vector{Int32},Vector{Int32}}}     # site pairs I < J in findnz order (static)
    src::Any                                                     # the host matrix whose values it holds (identity check)
end
# (n1, n2) if `l` is an UNFILTERED 2-D Bravais lattice spanned over n1 x n2 unit cells - its rows
# are then cell-major (src/lattices/bravais/lattice.jl:101-111: last lattice axis fastest, basis
# index innermost, unitcell.jl:126-132), which is what lm_ham_set_lattice_dims declares and what
# the register-tiled stencil kernel needs.  `nothing` for any other lattice (ELL kernels).
function lattice_dims(l)
    l = LatticeModels.stripmeta(l)
    l isa LatticeModels.BravaisLattice || return nothing
    isempty(l.pointers) && return nothing
    length(first(l.pointers).latcoords) == 2 || return nothing
    lo1 = minimum(p -> p.latcoords[1], l.pointers); hi1 = maximum(p -> p.latcoords[1], l.pointers)
    lo2 = minimum(p -> p.latcoords[2], l.pointers); hi2 = maximum(p -> p.latcoords[2], l.pointers)
    n1, n2, nb = hi1 - lo1 + 1, hi2 - lo2 + 1, length(l.unitcell)
    length(l.pointers) == n1 * n2 * nb ? (n1, n2) : nothing
end
site_coords2(l) = Float64[site.coords[k] for k in 1:2, site in l]       # 2 x n_sites, column-major = (x, y) pairs
function set_geometry!(dev::DeviceHam, l)
    l === nothing && return dev
    check(ccall((:lm_ham_set_site_coords, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), dev.handle, site_coords2(l)))
    dims = lattice_dims(l)
    dims === nothing ||
        check(ccall((:lm_ham_set_lattice_dims, LIB), Int32, (Ptr{Cvoid}, Int32, Int32), dev.handle, dims[1], dims[2]))
    dev
end
function DeviceHam(ctx::Context, mat::SparseMatrixCSC{ComplexF64,Int64}, n_int::Integer; lat = nothing)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:lm_ham_create_csc, LIB), Int32,
                (Ptr{Cvoid}, Int64, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{ComplexF64}, Int32, Ref{Ptr{Cvoid}}),
                ctx.handle, size(mat, 1), n_int, mat.colptr, mat.rowval, mat.nzval, 1, h))
    dev = DeviceHam(h[], copy(mat.colptr), copy(mat.rowval), n_int, nothing, mat)
    finalizer(d -> ccall((:lm_ham_destroy, LIB), Int32, (Ptr{Cvoid},), d.handle), dev)
    set_geometry!(dev, lat)
end
samepattern(d::DeviceHam, m::SparseMatrixCSC) = d.colptr == m.colptr && d.rowval == m.rowval
function update_values!(d::DeviceHam, m::SparseMatrixCSC{ComplexF64,Int64})
    check(ccall((:lm_ham_update_values, LIB), Int32, (Ptr{Cvoid}, Ptr{ComplexF64}), d.handle, m.nzval))
    d.src = m; d
end
# Multi-process runs (one Julia process per GPU, the same H(t) in every process): only `root` hands its host values over, the other
# ranks receive them over NVLink (ncclBroadcast) instead of one PCIe upload per process.  Asynchronous like lm_ham_update_values_async:
# `m.nzval` must stay alive until the next synchronising call (lm_frame_wait / lm_ctx_synchronize / an observable); every rank of the
# communicator must make the call.
function update_values_bcast!(d::DeviceHam, m::Union{Nothing,SparseMatrixCSC{ComplexF64,Int64}}, root::Integer = 0)
    check(ccall((:lm_ham_update_values_bcast, LIB), Int32, (Ptr{Cvoid}, Ptr{ComplexF64}, Int32),
                d.handle, m === nothing ? Ptr{ComplexF64}(C_NULL) : pointer(m.nzval), root))
    m === nothing || (d.src = m)
    d
end
function currents_pairs(d::DeviceHam)
    d.pairs === nothing || return d.pairs
    np = Ref{Int64}(0)
    check(ccall((:lm_currents_npairs, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}), d.handle, np))
    I = Vector{Int32}(undef, np[]); J = Vector{Int32}(undef, np[])
    check(ccall((:lm_currents_pairs, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), d.handle, I, J))
    d.pairs = (I, J)
end

# ---- device state: P = Psi diag(w) Psi' (Psi block), a Ket (M = 1), or a dense density matrix ----
# An AbstractMatrix so that Operator(basis, PsiProjector) IS a DataOperator of the reference
# (EvolutionStateType, src/evolution.jl:36; StateType, src/operators/bases.jl:35).
mutable struct PsiProjector <: AbstractMatrix{ComplexF64}
    handle::Ptr{Cvoid}
    n::Int
    ctx::Context
    solver::Any                       # the B200Exp that stepped it last (its device H is the one yielded with the frame)
    function PsiProjector(handle::Ptr{Cvoid}, n::Integer, ctx::Context)
        finalizer(s -> ccall((:lm_state_destroy, LIB), Int32, (Ptr{Cvoid},), s.handle), new(handle, n, ctx, nothing))
    end
end
"`PsiProjector(psi[, w]; ctx)`: the columns of `psi` (N x M, this rank's shard under multi-GPU) with occupation weights `w`"
function PsiProjector(psi::AbstractMatrix{<:Number}, w::Union{Nothing,AbstractVector{<:Real}} = nothing;
                      ctx::Context = default_context())
    h = Ref{Ptr{Cvoid}}(C_NULL)
    p = Matrix{ComplexF64}(psi)                          # column-major N x M
    wv = w === nothing ? nothing : Vector{Float64}(w)
    check(ccall((:lm_state_create_psi, LIB), Int32,
                (Ptr{Cvoid}, Int64, Int64, Ptr{ComplexF64}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                ctx.handle, size(p, 1), size(p, 2), p, wv === nothing ? C_NULL : wv, h))
    PsiProjector(h[], size(p, 1), ctx)
end
"`dense_state(P; ctx)`: a dense density matrix kept on the device (stepped as U P U' on the DMMA path)"
function dense_state(P::AbstractMatrix{<:Number}; ctx::Context = default_context())
    h = Ref{Ptr{Cvoid}}(C_NULL)
    Pm = Matrix{ComplexF64}(P)
    check(ccall((:lm_state_create_dense, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{ComplexF64}, Ref{Ptr{Cvoid}}),
                ctx.handle, size(Pm, 1), Pm, h))
    PsiProjector(h[], size(Pm, 1), ctx)
end
Base.size(s::PsiProjector) = (s.n, s.n)
function Base.copy(s::PsiProjector)                      # copy(state), src/evolution.jl:193
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:lm_state_copy, LIB), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), s.handle, h))
    PsiProjector(h[], s.n, s.ctx)
end
function Base.Matrix(s::PsiProjector)                    # escape hatch: materialise Psi W Psi' (or download dense P)
    P = Matrix{ComplexF64}(undef, s.n, s.n)
    check(ccall((:lm_state_download_dense, LIB), Int32, (Ptr{Cvoid}, Ptr{ComplexF64}), s.handle, P))
    P
end
Base.getindex(s::PsiProjector, i::Int, j::Int) = Matrix(s)[i, j]     # slow, debugging only
Base.show(io::IO, s::PsiProjector) = print(io, "PsiProjector(", s.n, "x", s.n, " on device)")
Base.show(io::IO, ::MIME"text/plain", s::PsiProjector) = show(io, s)
"columns of the device block, N x M (this rank's shard)"
function psi_columns(s::PsiProjector)
    N = Ref{Int64}(0); M = Ref{Int64}(0); d = Ref{Int32}(0)
    check(ccall((:lm_state_dims, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int32}), s.handle, N, M, d))
    out = Matrix{ComplexF64}(undef, N[], M[])
    check(ccall((:lm_state_download_psi, LIB), Int32, (Ptr{Cvoid}, Ptr{ComplexF64}), s.handle, out))
    out
end
"mark a block whose columns are the SAME on every rank (no cross-rank sum of its observables)"
set_replicated!(s::PsiProjector, flag::Bool = true) =
    (check(ccall((:lm_state_set_replicated, LIB), Int32, (Ptr{Cvoid}, Int32), s.handle, flag ? 1 : 0)); s)

# N1 (SURVEY.md section 8f): the occupied eigenvectors and their weights go to the device as a Psi
# block instead of being multiplied out to a dense N x N matrix (projector(f, eig), src/spectrum.jl:226-232).
"`psi_projector(f, eig)`: `projector(f, eig)` in Psi-block form - Operator(eig.basis, PsiProjector)"
function psi_projector(f, eig::AbstractEigensystem; ctx::Context = default_context(), cols = nothing)
    w = Float64[f(E) for E in eig.values]
    keep = findall(!iszero, w)
    cols === nothing || (keep = keep[cols])              # multi-GPU: keep[shard_range(length(keep), rank, nranks)]
    Operator(eig.basis, PsiProjector(eig.states[:, keep], w[keep]; ctx = ctx))
end
"`psi_densitymatrix(eig; T = 0, mu = 0, statistics = FermiDirac)`: `ensemble_densitymatrix` (src/spectrum.jl:253-257) in Psi-block form"
psi_densitymatrix(eig::AbstractEigensystem; T::Real = 0, μ::Real = 0, mu::Real = μ, statistics = FermiDirac, kw...) =
    psi_projector(densfun(T, mu, statistics), eig; kw...)

const DevOp{B} = Operator{B,B,<:PsiProjector}

# N1, second half: the lowest eigenpairs on the device (Chebyshev-filtered subspace iteration on the SpMM
# kernels of the propagator) for sizes where the LAPACK route of `diagonalize` is impossible.  Plugs into the
# reference's routine dispatch (src/spectrum.jl:48-67):  diagonalize(H, :b200; n = 10)
"`eigs_lowest(ham; n, tol)` -> (eigenvalues, Operator(basis, PsiProjector)): the n lowest levels as a device block (Fermi sphere)"
function eigs_lowest(ham::DataOperator; n::Integer = 10, tol::Real = 1e-10, maxiter::Integer = 0, degree::Integer = 0,
                     ctx::Context = default_context())
    mat = ham.data isa SparseMatrixCSC{ComplexF64,Int64} ? ham.data : SparseMatrixCSC{ComplexF64,Int64}(sparse(ham.data))
    isl = basis(ham) isa LatticeModels.AbstractLatticeBasis
    dev = DeviceHam(ctx, mat, isl ? internal_length(ham) : 1; lat = isl ? lattice(ham) : nothing)
    vals = Vector{Float64}(undef, n); res = Vector{Float64}(undef, n)
    h = Ref{Ptr{Cvoid}}(C_NULL); it = Ref{Int32}(0)
    check(ccall((:lm_eigs_lowest, LIB), Int32,
                (Ptr{Cvoid}, Int32, Float64, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}, Ref{Int32}),
                dev.handle, n, tol, maxiter, degree, vals, res, h, it))
    vals, Operator(basis(ham), PsiProjector(h[], size(mat, 1), ctx))
end
function LatticeModels.diagonalize_routine(op::DataOperator, ::Val{:b200}; n = 10, tol = 1e-10, kw...)
    vals, P = eigs_lowest(op; n = n, tol = tol)
    LatticeModels.Eigensystem(basis(op), psi_columns(P.data), vals)
end

# ---- the solver ---------------------------------------------------------------------------------
mutable struct B200Exp <: EvolutionSolver
    ctx::Context
    dev::Union{Nothing,DeviceHam}
    mat::Any                      # the matrix the device H currently holds (identity check, src/evolution.jl:86-88)
    dt::Float64
    tol::Float64
    method::Int32                 # LM_METHOD_*: 0 auto, 1 Chebyshev, 2 Taylor, 3 Lanczos (KrylovKitExp semantics)
    n_int::Int
    lat::Any                      # lattice of the Hamiltonian (site coordinates / cell-major dims for the stencil kernels)
end
# B200Exp(; kw...) without a Hamiltonian comes for free via IncompleteSolver (src/evolution.jl:219-231);
# Evolution then calls B200Exp(hamiltonian; kw...) with whatever the user passed (operator, function of t, matrix)
function B200Exp(ham; tol = 1e-12, method = 0, precision::Symbol = :c128,
                 ctx = precision === :c128 ? default_context() : Context(precision = precision))
    h0 = LatticeModels._eval_ham(ham, 0.0)
    if h0 isa DataOperator && basis(h0) isa LatticeModels.AbstractLatticeBasis
        B200Exp(ctx, nothing, nothing, 0.0, tol, Int32(method), internal_length(h0), lattice(h0))
    else
        B200Exp(ctx, nothing, nothing, 0.0, tol, Int32(method), 1, nothing)
    end
end

function update_solver!(s::B200Exp, mat::SparseMatrixCSC{ComplexF64,Int64}, dt, force = false)   # src/evolution.jl:83-92
    s.dt = dt
    !force && s.mat === mat && s.dev !== nothing && return
    if s.dev !== nothing && !isempty(s.dev.colptr) && samepattern(s.dev, mat)
        update_values!(s.dev, mat)                                  # same sparsity pattern: nzval only
    else
        s.dev = DeviceHam(s.ctx, mat, s.n_int; lat = s.lat)
    end
    s.mat = mat
    return
end
update_solver!(s::B200Exp, mat::Base.RefValue, dt) = update_solver!(s, mat[], dt, true)           # src/evolution.jl:82
update_solver!(s::B200Exp, mat::AbstractMatrix, dt, force = false) =
    update_solver!(s, SparseMatrixCSC{ComplexF64,Int64}(sparse(mat)), dt, force)

function lm_step!(s::B200Exp, handle::Ptr{Cvoid})
    s.dev === nothing && throw(ArgumentError("B200Exp: update_solver! has not been called"))
    check(ccall((:lm_step, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64, Int32, Ptr{Int32}),
                s.dev.handle, handle, s.dt, s.tol, s.method, C_NULL))
end
# (a) device-resident states
evolution_cache(::B200Exp, ::PsiProjector) = nothing
function step!(s::B200Exp, state::PsiProjector, _cache)                            # src/evolution.jl:69-78
    lm_step!(s, state.handle)
    state.solver = s
    state
end
# (b) plain host arrays, exactly as with CachedExp: a Matrix is a density matrix (U P U', src/evolution.jl:73-78),
# a Vector a ket.  The cache is the device twin, uploaded ONCE when the Evolution is built
# (evolution_cache runs on the original state, src/evolution.jl:190-194); every step runs on the
# device and lands in the host array, which the reference's own localdensity / DensityCurrents read.
struct DeviceTwin
    state::PsiProjector
end
evolution_cache(s::B200Exp, state::Matrix{ComplexF64}) = DeviceTwin(dense_state(state; ctx = s.ctx))
evolution_cache(s::B200Exp, state::Vector{ComplexF64}) = DeviceTwin(PsiProjector(reshape(state, :, 1); ctx = s.ctx))
function step!(s::B200Exp, state::Matrix{ComplexF64}, cache::DeviceTwin)
    lm_step!(s, cache.state.handle)
    check(ccall((:lm_state_download_dense, LIB), Int32, (Ptr{Cvoid}, Ptr{ComplexF64}), cache.state.handle, state))
    state
end
function step!(s::B200Exp, state::Vector{ComplexF64}, cache::DeviceTwin)
    lm_step!(s, cache.state.handle)
    check(ccall((:lm_state_download_psi, LIB), Int32, (Ptr{Cvoid}, Ptr{ComplexF64}), cache.state.handle, state))
    state
end

# ---- the device Hamiltonian that goes with a currents object -----------------------------------
# DensityCurrents(H, P) carries the HOST Hamiltonian the evolution yielded with the frame.  The
# solver that stepped P holds exactly that matrix on the device (same object: `===`); any other H
# (first frame before a step, a user-built operator) is uploaded - values only if the pattern is known.
function device_ham(ham::DataOperator, st::PsiProjector)
    ham.data isa DeviceMatrix && return ham.data.dev
    mat = ham.data isa SparseMatrixCSC{ComplexF64,Int64} ? ham.data : SparseMatrixCSC{ComplexF64,Int64}(sparse(ham.data))
    s = st.solver
    s isa B200Exp && s.dev !== nothing && s.mat === ham.data && return s.dev
    d = st.ctx.obsdev
    if d isa DeviceHam && d.src === ham.data
        return d
    elseif d isa DeviceHam && samepattern(d, mat)
        return update_values!(d, mat)
    end
    st.ctx.obsdev = DeviceHam(st.ctx, mat, internal_length(ham); lat = lattice(ham))
end

# ---- observables --------------------------------------------------------------------------------
function localdensity(state::DevOp{<:OneParticleBasis})                           # latticeutils.jl:41-45
    l = lattice(state); n = internal_length(state)
    rho = Vector{Float64}(undef, length(l))
    check(ccall((:lm_local_density, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}), state.data.handle, n, rho))
    LatticeValue(l, rho)
end
function localexpect(op::DataOperator, state::DevOp{<:CompositeLatticeBasis})     # latticeutils.jl:13-20
    check_samebases(internal_basis(state), basis(op))
    l = lattice(state); n = internal_length(state)
    out = Vector{ComplexF64}(undef, length(l))
    check(ccall((:lm_local_expect, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{ComplexF64}, Ptr{ComplexF64}),
                state.data.handle, n, Matrix{ComplexF64}(op.data), out))
    LatticeValue(l, out)
end

const DevDensityCurrents = DensityCurrents{<:Any,<:DevOp}
const DevOperatorCurrents = LocalOperatorCurrents{<:Any,<:DevOp}
const DevCurrents = Union{DevDensityCurrents,DevOperatorCurrents}

"(I, J, V): every site pair I < J of H's own sparsity in findnz order, unfiltered - ONE fused pass over Psi"
function pair_values(curr::DevDensityCurrents)
    dev = device_ham(curr.hamiltonian, curr.state.data)
    I, J = currents_pairs(dev)
    V = Vector{Float64}(undef, max(length(I), 1))
    check(ccall((:lm_observables, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
                dev.handle, curr.state.data.handle, C_NULL, V))
    I, J, V[1:length(I)]
end
function pair_values(curr::DevOperatorCurrents)                                    # src/zoo/currents.jl:150-184
    dev = device_ham(curr.hamiltonian, curr.state.data)
    I, J = currents_pairs(dev)
    V = Vector{Float64}(undef, max(length(I), 1))
    check(ccall((:lm_operator_currents, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Float64}),
                dev.handle, curr.state.data.handle, Matrix{ComplexF64}(curr.op.data), V))
    I, J, V[1:length(I)]
end
function antisymmetric(l, I, J, V)
    keep = abs.(V) .>= CURRENTS_EPS                                                # src/currents.jl:230
    n = length(l)
    Currents(l, sparse(vcat(Int.(I[keep]), Int.(J[keep])), vcat(Int.(J[keep]), Int.(I[keep])), vcat(V[keep], -V[keep]), n, n))
end
# Currents(curr): only pairs with H_ij != 0 can carry a current, so the O(n^2) getindex loop of
# src/currents.jl:223-237 becomes one pass over H's bond list
Currents(curr::DevCurrents) = antisymmetric(lattice(curr), pair_values(curr)...)
# curr[i, j] (src/zoo/currents.jl:92-102)
function Base.getindex(curr::DevDensityCurrents, i::Int, j::Int)
    dev = device_ham(curr.hamiltonian, curr.state.data)
    out = Ref{Float64}(0.0)
    check(ccall((:lm_bond_currents, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ref{Int32}, Ref{Int32}, Ref{Float64}),
                dev.handle, curr.state.data.handle, 1, Ref(Int32(i)), Ref(Int32(j)), out))
    out[]
end
function Base.getindex(curr::DevOperatorCurrents, i::Int, j::Int)
    i == j && return 0.0
    I, J, V = pair_values(curr)
    a, b, sgn = i < j ? (i, j, 1.0) : (j, i, -1.0)
    k = findfirst(q -> I[q] == a && J[q] == b, eachindex(I))
    k === nothing ? 0.0 : sgn * V[k]
end
# Currents(curr, bonds) (src/currents.jl:238-255)
function Currents(curr::DevDensityCurrents, bonds::AbstractBonds)
    l = lattice(curr)
    Is = Int32[]; Js = Int32[]
    for (s1, s2) in adapt_bonds(bonds, l)
        push!(Is, s1.index); push!(Js, s2.index)
    end
    dev = device_ham(curr.hamiltonian, curr.state.data)
    V = Vector{Float64}(undef, max(length(Is), 1))
    check(ccall((:lm_bond_currents, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}),
                dev.handle, curr.state.data.handle, length(Is), Is, Js, V))
    V = V[1:length(Is)]
    keep = abs.(V) .>= CURRENTS_EPS
    n = length(l)
    I2 = Int.(Is[keep]); J2 = Int.(Js[keep])
    Currents(l, sparse(vcat(I2, J2), vcat(J2, I2), vcat(V[keep], -V[keep]), n, n, (a, b) -> a))   # a bond listed twice counts once
end
# findnz (src/currents.jl:159-177): pairs come back in the order of findnz of a CSC matrix filtered by I < J
function SparseArrays.findnz(curr::DevCurrents)
    I, J, V = pair_values(curr)
    keep = abs.(V) .>= CURRENTS_EPS
    Int.(I[keep]), Int.(J[keep]), V[keep]
end
# currentsfromto / currentsfrom (src/currents.jl:85-109) summed on the device: only one number / one
# LatticeValue crosses PCIe.  Regions go through the reference's own `to_inds`.
function region_mask(l, region)
    m = zeros(UInt8, length(l))
    for i in to_inds(l, region)          # a single site gives one Int, a collection / mask a vector of them
        m[i] = 1
    end
    m
end
function currentsfromto(curr::DevDensityCurrents, src, dst = nothing)
    l = lattice(curr); out = Ref{Float64}(0.0)
    dev = device_ham(curr.hamiltonian, curr.state.data)
    ms = region_mask(l, src)
    md = dst === nothing ? UInt8[] : region_mask(l, dst)          # NULL = every site outside src
    GC.@preserve md begin
        check(ccall((:lm_currents_fromto, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt8}, Ptr{UInt8}, Int32, Ref{Float64}),
                    dev.handle, curr.state.data.handle, ms, dst === nothing ? Ptr{UInt8}(C_NULL) : pointer(md), 0, out))
    end
    out[]
end
function currentsfrom(curr::DevDensityCurrents, src)
    l = lattice(curr); out = Vector{Float64}(undef, length(l))
    dev = device_ham(curr.hamiltonian, curr.state.data)
    check(ccall((:lm_currents_from, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt8}, Int32, Ptr{Float64}),
                dev.handle, curr.state.data.handle, region_mask(l, src), 0, out))
    LatticeValue(l, out)
end

# ---- asynchronous frame sink for TimeSequence collection (src/timesequence.jl:41-43) ------------
# frame k is reduced and copied to the host on a second stream while step k + 1 runs; two slots alternate.
mutable struct FrameSink
    ctx::Context
    pending::Vector{Tuple{Int32,Int,Int}}
    next::Int32
    rho::Vector{Vector{Float64}}
    J::Vector{Vector{Float64}}
end
FrameSink(ctx::Context = default_context()) = FrameSink(ctx, Tuple{Int32,Int,Int}[], Int32(0), Vector{Float64}[], Vector{Float64}[])
function collect_frame!(k::FrameSink)
    slot, ns, np = popfirst!(k.pending)
    rho = Vector{Float64}(undef, ns); J = Vector{Float64}(undef, max(np, 1))
    check(ccall((:lm_frame_wait, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}), k.ctx.handle, slot, rho, J))
    push!(k.rho, rho); push!(k.J, J[1:np]); k
end
function Base.push!(k::FrameSink, H::DataOperator, state::DevOp)
    length(k.pending) == 2 && collect_frame!(k)
    dev = device_ham(H, state.data)
    np = length(currents_pairs(dev)[1])
    check(ccall((:lm_observables_async, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32), dev.handle, state.data.handle, k.next, 1))
    push!(k.pending, (k.next, length(lattice(state)), np)); k.next = xor(k.next, Int32(1)); k
end
finish!(k::FrameSink) = (while !isempty(k.pending) collect_frame!(k) end; k)
"""
    TimeSequence(sink::FrameSink, ev_iter)

`TimeSequence(moment -> (localdensity(moment.state), Currents(DensityCurrents(moment.H, moment.state))), ev_iter)`
without a host stall per frame: returns a `TimeSequence` of `(density = LatticeValue, currents = Currents)`
named tuples keyed by the iterator's times (src/timesequence.jl:41-43).
"""
function TimeSequence(sink::FrameSink, iter::EvolutionIterator)
    lat = nothing; pairs = nothing
    for moment in iter
        st = moment.state
        push!(sink, moment.H, st)
        if lat === nothing
            lat = lattice(st); pairs = currents_pairs(device_ham(moment.H, st.data))
        end
    end
    finish!(sink)
    vals = [(density = LatticeValue(lat, sink.rho[k]), currents = antisymmetric(lat, pairs[1], pairs[2], sink.J[k])) for k in eachindex(sink.rho)]
    TimeSequence(collect(Float64, iter.times), vals)
end

# ---- device-resident time-dependent Hamiltonian (N2) --------------------------------------------
# Holds the directed bond table once; settime! only ships the field parameters, the Peierls phases
# are regenerated on the device (builder.jl:282-309 restated in lm_ham_create_bonds).  It IS a
# DataOperator on the lattice basis - so DensityCurrents(H, P) passes check_samebases
# (src/zoo/currents.jl:85) - whose `.data` is a handle instead of a host matrix; use it through the
# Function branch of Evolution (src/evolution.jl:43):   Evolution(B200Exp(), t -> settime!(Hdev, t), P0)
struct DeviceMatrix <: AbstractMatrix{ComplexF64}
    dev::DeviceHam
    n::Int
end
Base.size(m::DeviceMatrix) = (m.n, m.n)
function SparseArrays.sparse(m::DeviceMatrix)                 # the current H as a host CSC matrix (1-based)
    N = Ref{Int64}(0); ni = Ref{Int32}(0); nnz = Ref{Int64}(0); W = Ref{Int32}(0)
    check(ccall((:lm_ham_dims, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int32}, Ref{Int64}, Ref{Int32}), m.dev.handle, N, ni, nnz, W))
    cp = Vector{Int64}(undef, N[] + 1); rv = Vector{Int64}(undef, nnz[]); nz = Vector{ComplexF64}(undef, nnz[])
    check(ccall((:lm_ham_get_csc, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{ComplexF64}), m.dev.handle, cp, rv, nz))
    SparseMatrixCSC(Int(N[]), Int(N[]), cp, rv, nz)
end
Base.getindex(m::DeviceMatrix, i::Int, j::Int) = sparse(m)[i, j]                   # slow, debugging only
Base.show(io::IO, m::DeviceMatrix) = print(io, "DeviceMatrix(", m.n, "x", m.n, " on device)")
Base.show(io::IO, ::MIME"text/plain", m::DeviceMatrix) = show(io, m)

mutable struct B200Hamiltonian{SystemT,BasisT} <: DataOperator{BasisT,BasisT}
    sys::SystemT
    basis_l::BasisT
    basis_r::BasisT
    data::DeviceMatrix
    field::Function               # t -> a closed-form field: LandauGauge / SymmetricGauge / PointFlux / FieldSum of those / NoField
end
# field descriptors of include/lm_b200.h (LM_FIELD_*; 3 doubles per field)
field_descr(::LatticeModels.NoField) = (Int32[], Float64[])
field_descr(f::LatticeModels.LandauGauge) = (Int32[1], Float64[f.B, 0, 0])
field_descr(f::LatticeModels.SymmetricGauge) = (Int32[2], Float64[f.B, 0, 0])
field_descr(f::LatticeModels.PointFlux{:axial}) = (Int32[3], Float64[f.flux, f.point[1], f.point[2]])
field_descr(f::LatticeModels.PointFlux{:singular}) = (Int32[4], Float64[f.flux, f.point[1], f.point[2]])
function field_descr(f::LatticeModels.PointFluxes{G}) where G
    k = G === :axial ? Int32(3) : Int32(4)
    (fill(k, length(f.fluxes)), reduce(vcat, [Float64[f.fluxes[q], f.points[q][1], f.points[q][2]] for q in eachindex(f.fluxes)]; init = Float64[]))
end
function field_descr(f::LatticeModels.FieldSum)
    parts = map(field_descr, f.fields)
    (reduce(vcat, first.(parts); init = Int32[]), reduce(vcat, last.(parts); init = Float64[]))
end
# every (site1, site2) pair of a bonds description, as add_term! enumerates them (src/operators/constructoperator.jl:14-38):
# a NearestNeighbor / BravaisSiteMapping adapts to a set of translations, each of which iterates its site pairs
function foreach_bond(f, what, l)
    b = adapt_bonds(what, l)
    if b isa LatticeModels.BravaisSiteMapping
        for tr in b.translations
            foreach_bond(f, tr, l)
        end
    else
        for (a1, a2) in b
            f(a1, a2)
        end
    end
end
"""
    B200Hamiltonian(template, terms...; field = t -> NoField())

`template` is the reference Hamiltonian of the same system at any time (it supplies system and basis);
`terms` are the `op => bonds` / `op => LatticeValue | number` pairs of `construct_hamiltonian`
(src/operators/constructoperator.jl:4-47) WITHOUT the field.  The bond table is built once from the
reference's own bond iteration and site resolution (`adapt_bonds`, `resolve_site`: pre-wrap
coordinates and boundary factors, src/operators/builder.jl:282-285).
"""
function B200Hamiltonian(template::LatticeModels.Hamiltonian, terms::Pair...; field::Function = t -> LatticeModels.NoField(),
                         ctx::Context = default_context())
    l = lattice(template); n = internal_length(template); ns = length(l)
    sample = LatticeModels.sample(template)
    src = Int32[]; dst = Int32[]; rs = Float64[]; rd = Float64[]; amp = ComplexF64[]; bf = ComplexF64[]
    onsite = zeros(ComplexF64, n, n, ns); has_onsite = false
    for (op, what) in terms
        iszero(op) && continue                                                    # add_pair_terms!, constructoperator.jl:49-52
        B = Matrix{ComplexF64}(LatticeModels.op_to_matrix(sample, op))
        if what isa LatticeValue
            for i in 1:ns; onsite[:, :, i] .+= what.values[i] .* B; end; has_onsite = true
        elseif what isa Number
            for i in 1:ns; onsite[:, :, i] .+= what .* B; end; has_onsite = true
        else
            foreach_bond(what, l) do a1, a2
                s1 = LatticeModels.resolve_site(l, a1); s2 = LatticeModels.resolve_site(l, a2)
                (s1 === nothing || s2 === nothing) && return
                push!(src, s1.index); push!(dst, s2.index)
                append!(rs, s1.old_site.coords[1:2]); append!(rd, s2.old_site.coords[1:2])
                append!(amp, vec(B)); push!(bf, s2.factor * s1.factor')
            end
        end
    end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:lm_ham_create_bonds, LIB), Int32,
                (Ptr{Cvoid}, Int64, Int32, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64},
                 Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Int32, Ref{Ptr{Cvoid}}),
                ctx.handle, ns, n, length(src), src, dst, rs, rd, amp, bf, has_onsite ? onsite : C_NULL, 1, h))
    dev = DeviceHam(h[], Int64[], Int64[], n, nothing, nothing)
    finalizer(d -> ccall((:lm_ham_destroy, LIB), Int32, (Ptr{Cvoid},), d.handle), dev)
    set_geometry!(dev, l)
    H = B200Hamiltonian(template.sys, template.basis_l, template.basis_r, DeviceMatrix(dev, ns * n), field)
    kinds, params = field_descr(field(0.0))
    check(ccall((:lm_ham_set_fields, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}), dev.handle, length(kinds), kinds, params))
    H
end
"`settime!(H, t)`: a few doubles to the device, phases regenerated there; returns H (use as `t -> settime!(H, t)`)"
function settime!(H::B200Hamiltonian, t)
    kinds, params = field_descr(H.field(t))
    check(ccall((:lm_ham_set_fields, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}), H.data.dev.handle, length(kinds), kinds, params))
    H
end
update_solver!(s::B200Exp, mat::DeviceMatrix, dt, force = false) = (s.dt = dt; s.dev = mat.dev; s.mat = mat; nothing)

export B200Exp, PsiProjector, Context, B200Hamiltonian, settime!, FrameSink, psi_projector, psi_densitymatrix, eigs_lowest,
       dense_state, psi_columns, shard_range, unique_id, comm_init!, peer_handle, peer_attach!, set_replicated!, update_values_bcast!

end # module
