# B200Backend.jl - thin Julia glue that plugs liblm_b200.so into LatticeModels.jl.
#
# NOT EXECUTED in the build container (Julia is not installed there); it is kept minimal and
# line-for-line reviewable.  All logic lives in the shared library behind include/lm_b200.h;
# this file only (a) declares the new `EvolutionSolver` subtype and the device state type,
# (b) forwards the reference's extension points to `ccall`s.  No reference file is edited:
# everything below is new methods on existing generic functions (multiple dispatch is the
# reference's plug-in API, src/evolution.jl:25-33).
#
#   using LatticeModels, B200Backend
#   P0 = densitymatrix(h(0), mu = 0)                        # host, as before
#   ev = Evolution(B200Exp(tol = 1e-12), t -> h(t), PsiProjector(P0_eig...))   # or P0 itself
#   for (P, H, t) in ev(0:0.1:20)
#       localdensity(P); Currents(DensityCurrents(H, P))
#   end
module B200Backend

using LatticeModels, SparseArrays, LinearAlgebra
import LatticeModels: EvolutionSolver, update_solver!, step!, evolution_cache, localdensity,
                      DensityCurrents, Currents, lattice, internal_length
import QuantumOpticsBase: Operator, basis

const LIB = get(ENV, "LM_B200_LIB", joinpath(@__DIR__, "..", "latticemodels.jl_b200", "lib", "liblm_b200.so"))

# ---- status -> ArgumentError (mirrors src/evolution.jl:152,239) ---------------------------------
lasterror() = unsafe_string(ccall((:lm_last_error, LIB), Cstring, ()))
check(status::Int32) = status == 0 || throw(ArgumentError(lasterror()))

# ---- context (one process drives one GPU) ------------------------------------------------------
mutable struct Context
    handle::Ptr{Cvoid}
    function Context(; device::Integer = 0, precision::Symbol = :c128)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:lm_ctx_create, LIB), Int32, (Int32, Int32, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                    device, precision === :c64 ? 1 : 0, C_NULL, h))
        finalizer(c -> ccall((:lm_ctx_destroy, LIB), Int32, (Ptr{Cvoid},), c.handle), new(h[]))
    end
end
const DEFAULT_CTX = Ref{Union{Nothing,Context}}(nothing)
default_context() = something(DEFAULT_CTX[], (DEFAULT_CTX[] = Context()))

# ---- device Hamiltonian: upload of a Julia SparseMatrixCSC (1-based indices accepted as is) ----
mutable struct DeviceHam
    handle::Ptr{Cvoid}
    colptr::Vector{Int64}
    rowval::Vector{Int64}
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
function set_lattice_dims!(dev, dims)
    dims === nothing && return dev
    check(ccall((:lm_ham_set_lattice_dims, LIB), Int32, (Ptr{Cvoid}, Int32, Int32), dev.handle, dims[1], dims[2]))
    dev
end

function DeviceHam(ctx::Context, mat::SparseMatrixCSC{ComplexF64,Int64}, n_int::Integer; coords = nothing, dims = nothing)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:lm_ham_create_csc, LIB), Int32,
                (Ptr{Cvoid}, Int64, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{ComplexF64}, Int32, Ref{Ptr{Cvoid}}),
                ctx.handle, size(mat, 1), n_int, mat.colptr, mat.rowval, mat.nzval, 1, h))
    dev = DeviceHam(h[], copy(mat.colptr), copy(mat.rowval))
    if coords !== nothing       # 2 x n_sites Float64 matrix of site coordinates (site.coords[1:2])
        check(ccall((:lm_ham_set_site_coords, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), dev.handle, coords))
    end
    set_lattice_dims!(dev, dims)                         # lattice_dims(lattice(H)) of the caller's Hamiltonian
    finalizer(d -> ccall((:lm_ham_destroy, LIB), Int32, (Ptr{Cvoid},), d.handle), dev)
end
samepattern(d::DeviceHam, m::SparseMatrixCSC) = d.colptr == m.colptr && d.rowval == m.rowval

# ---- device state: P = Psi diag(w) Psi' (Psi-block) or a dense density matrix -----------------
# An AbstractMatrix so that Operator(basis, PsiProjector) IS a DataOperator of the reference
# (EvolutionStateType, src/evolution.jl:36; StateType, src/operators/bases.jl:35).
mutable struct PsiProjector <: AbstractMatrix{ComplexF64}
    handle::Ptr{Cvoid}
    n::Int
    ctx::Context
end
function PsiProjector(psi::AbstractMatrix{ComplexF64}, w::Union{Nothing,Vector{Float64}} = nothing;
                      ctx::Context = default_context())
    h = Ref{Ptr{Cvoid}}(C_NULL)
    p = Matrix(psi)                                      # column-major N x M
    check(ccall((:lm_state_create_psi, LIB), Int32,
                (Ptr{Cvoid}, Int64, Int64, Ptr{ComplexF64}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                ctx.handle, size(p, 1), size(p, 2), p, w === nothing ? C_NULL : w, h))
    finalizer(s -> ccall((:lm_state_destroy, LIB), Int32, (Ptr{Cvoid},), s.handle), PsiProjector(h[], size(p, 1), ctx))
end
Base.size(s::PsiProjector) = (s.n, s.n)
function Base.copy(s::PsiProjector)                      # copy(state), src/evolution.jl:193
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:lm_state_copy, LIB), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), s.handle, h))
    finalizer(t -> ccall((:lm_state_destroy, LIB), Int32, (Ptr{Cvoid},), t.handle), PsiProjector(h[], s.n, s.ctx))
end
function Base.Matrix(s::PsiProjector)                    # escape hatch: materialise Psi W Psi'
    P = Matrix{ComplexF64}(undef, s.n, s.n)
    check(ccall((:lm_state_download_dense, LIB), Int32, (Ptr{Cvoid}, Ptr{ComplexF64}), s.handle, P))
    P
end
Base.getindex(s::PsiProjector, i::Int, j::Int) = Matrix(s)[i, j]     # slow, debugging only

# ---- the solver ---------------------------------------------------------------------------------
mutable struct B200Exp <: EvolutionSolver
    ctx::Context
    dev::Union{Nothing,DeviceHam}
    mat::Any
    dt::Float64
    tol::Float64
    method::Int32
    n_int::Int
    coords::Any
    dims::Any                     # (n1, n2) of an unfiltered Bravais lattice, or nothing
end
# B200Exp(; kw...) without a Hamiltonian comes for free via IncompleteSolver (src/evolution.jl:219-231)
function B200Exp(ham; tol = 1e-12, method = 0, precision = :c128, ctx = default_context(), coords = nothing)
    n_int = ham isa LatticeModels.Hamiltonian ? internal_length(ham) : 1
    dims = nothing
    if ham isa LatticeModels.Hamiltonian
        l = lattice(ham)
        coords === nothing && (coords = Float64[site.coords[k] for k in 1:2, site in l])
        dims = lattice_dims(l)
    end
    B200Exp(ctx, nothing, nothing, 0.0, tol, Int32(method), n_int, coords, dims)
end

function update_solver!(s::B200Exp, mat::SparseMatrixCSC, dt, force = false)      # src/evolution.jl:83-92
    s.dt = dt
    !force && s.mat === mat && s.dev !== nothing && return
    if s.dev !== nothing && samepattern(s.dev, mat)
        check(ccall((:lm_ham_update_values, LIB), Int32, (Ptr{Cvoid}, Ptr{ComplexF64}), s.dev.handle, mat.nzval))
    else
        s.dev = DeviceHam(s.ctx, mat, s.n_int; coords = s.coords, dims = s.dims)
    end
    s.mat = mat
    return
end
update_solver!(s::B200Exp, mat::AbstractMatrix, dt, force = false) =
    update_solver!(s, sparse(ComplexF64.(mat)), dt, force)

evolution_cache(::B200Exp, ::PsiProjector) = nothing
function step!(s::B200Exp, state::PsiProjector, _cache)                            # src/evolution.jl:69-78
    check(ccall((:lm_step, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64, Int32, Ptr{Int32}),
                s.dev.handle, state.handle, s.dt, s.tol, s.method, C_NULL))
    state
end
# A plain dense Matrix state means a density matrix: upload once, keep it on the device.
# (Evolution copies the state at construction; convert there.)
PsiProjector(P::Matrix{ComplexF64}; ctx::Context = default_context()) = begin
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:lm_state_create_dense, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{ComplexF64}, Ref{Ptr{Cvoid}}),
                ctx.handle, size(P, 1), P, h))
    finalizer(s -> ccall((:lm_state_destroy, LIB), Int32, (Ptr{Cvoid},), s.handle), PsiProjector(h[], size(P, 1), ctx))
end

# ---- observables --------------------------------------------------------------------------------
const DevOp = Operator{B,B,<:PsiProjector} where {B}
function localdensity(state::DevOp)                                               # latticeutils.jl:41-45
    l = lattice(state); n = internal_length(state)
    rho = Vector{Float64}(undef, length(l))
    check(ccall((:lm_local_density, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}), state.data.handle, n, rho))
    LatticeValue(l, rho)
end

# Currents(DensityCurrents(H, P)): all bonds of H in one fused pass (src/currents.jl:223-237).
# The solver's device Hamiltonian currently holds exactly the H yielded with this frame.
function Currents(curr::DensityCurrents{<:Any,<:DevOp}, solver::B200Exp)
    dev = solver.dev
    np = Ref{Int64}(0)
    check(ccall((:lm_currents_npairs, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}), dev.handle, np))
    I = Vector{Int32}(undef, np[]); J = similar(I); V = Vector{Float64}(undef, np[])
    check(ccall((:lm_currents_pairs, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), dev.handle, I, J))
    check(ccall((:lm_observables, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
                dev.handle, curr.state.data.handle, C_NULL, V))
    keep = abs.(V) .>= 1e-10                                                       # CURRENTS_EPS, src/currents.jl:4
    l = lattice(curr); n = length(l)
    Currents(l, sparse(vcat(I[keep], J[keep]), vcat(J[keep], I[keep]), vcat(V[keep], -V[keep]), n, n))
end

# currentsfromto / currentsfrom (src/currents.jl:85-109) summed on the device: only one number / one
# LatticeValue crosses PCIe.  Regions go through the reference's own `to_inds`.
function region_mask(l, region)
    m = zeros(UInt8, length(l)); m[LatticeModels.to_inds(l, region)] .= 1; m
end
function currentsfromto(curr::DensityCurrents{<:Any,<:DevOp}, solver::B200Exp, src, dst = nothing)
    l = lattice(curr); out = Ref{Float64}(0.0)
    ms = region_mask(l, src)
    md = dst === nothing ? C_NULL : region_mask(l, dst)
    check(ccall((:lm_currents_fromto, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt8}, Ptr{UInt8}, Int32, Ref{Float64}),
                solver.dev.handle, curr.state.data.handle, ms, md, 0, out))
    out[]
end
function currentsfrom(curr::DensityCurrents{<:Any,<:DevOp}, solver::B200Exp, src)
    l = lattice(curr); out = Vector{Float64}(undef, length(l))
    check(ccall((:lm_currents_from, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt8}, Int32, Ptr{Float64}),
                solver.dev.handle, curr.state.data.handle, region_mask(l, src), 0, out))
    LatticeValue(l, out)
end

# Asynchronous frame sink for TimeSequence-style collection (src/timesequence.jl:41-43): frame k is
# reduced and copied to the host on a second stream while step k + 1 runs; two slots alternate.
# push!(sink, solver, state, t) after every step, then finish!(sink) -> (times, rho frames, J frames).
mutable struct FrameSink
    ctx::Context; pending::Vector{Tuple{Int32,Float64,Int,Int}}; next::Int32
    times::Vector{Float64}; rho::Vector{Vector{Float64}}; J::Vector{Vector{Float64}}
end
FrameSink(ctx::Context = default_context()) = FrameSink(ctx, Tuple{Int32,Float64,Int,Int}[], Int32(0), Float64[], Vector{Float64}[], Vector{Float64}[])
function collect_frame!(k::FrameSink)
    slot, t, ns, np = popfirst!(k.pending)
    rho = Vector{Float64}(undef, ns); J = Vector{Float64}(undef, np)
    check(ccall((:lm_frame_wait, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}), k.ctx.handle, slot, rho, J))
    push!(k.times, t); push!(k.rho, rho); push!(k.J, J); k
end
function Base.push!(k::FrameSink, solver::B200Exp, state::DevOp, t)
    length(k.pending) == 2 && collect_frame!(k)
    np = Ref{Int64}(0)
    check(ccall((:lm_currents_npairs, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}), solver.dev.handle, np))
    check(ccall((:lm_observables_async, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32),
                solver.dev.handle, state.data.handle, k.next, 1))
    push!(k.pending, (k.next, Float64(t), length(lattice(state)), Int(np[]))); k.next = xor(k.next, Int32(1)); k
end
finish!(k::FrameSink) = (while !isempty(k.pending) collect_frame!(k) end; (k.times, k.rho, k.J))

# ---- device-resident time-dependent Hamiltonian (AbstractTimeDependentOperator branch) --------
# Holds the directed bond table once; set_time! only ships the field parameters, the Peierls
# phases are regenerated on the device (src/evolution.jl:44-47,243; builder.jl:282-309 restated
# in lm_ham_create_bonds).  `fieldparams(t)` returns the 3-doubles-per-field parameter matrix.
mutable struct B200Hamiltonian <: QuantumOpticsBase.AbstractTimeDependentOperator
    dev::DeviceHam
    kinds::Vector{Int32}
    fieldparams::Function
    template::Any                 # a reference Hamiltonian (basis, system) for DensityCurrents(H, P)
end
function B200Hamiltonian(ctx::Context, l, n_int::Integer, src::Vector{Int32}, dst::Vector{Int32},
                         r_src::Matrix{Float64}, r_dst::Matrix{Float64}, amp::Array{ComplexF64,3},
                         bfac::Vector{ComplexF64}, onsite::Union{Nothing,Array{ComplexF64,3}},
                         kinds::Vector{Int32}, fieldparams::Function, template)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:lm_ham_create_bonds, LIB), Int32,
                (Ptr{Cvoid}, Int64, Int32, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64},
                 Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Int32, Ref{Ptr{Cvoid}}),
                ctx.handle, length(l), n_int, length(src), src, dst, r_src, r_dst, amp, bfac,
                onsite === nothing ? C_NULL : onsite, 1, h))
    dev = DeviceHam(h[], Int64[], Int64[])
    coords = Float64[site.coords[k] for k in 1:2, site in l]
    check(ccall((:lm_ham_set_site_coords, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), dev.handle, coords))
    set_lattice_dims!(dev, lattice_dims(l))
    p0 = fieldparams(0.0)
    check(ccall((:lm_ham_set_fields, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}),
                dev.handle, length(kinds), kinds, p0))
    finalizer(x -> ccall((:lm_ham_destroy, LIB), Int32, (Ptr{Cvoid},), x.dev.handle),
              B200Hamiltonian(dev, kinds, fieldparams, template))
end
function QuantumOpticsBase.set_time!(H::B200Hamiltonian, t)
    check(ccall((:lm_ham_set_field_params, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), H.dev.handle, H.fieldparams(t)))
    H
end
LatticeModels._data(H::B200Hamiltonian) = H          # update_solver! receives the operator itself
update_solver!(s::B200Exp, H::B200Hamiltonian, dt, force = false) = (s.dt = dt; s.dev = H.dev; nothing)

export B200Exp, PsiProjector, Context, B200Hamiltonian

end # module
