# Re-runs the reference's own testsets for the evolution hot path with the B200 backend plugged in:
#   test/test_timedeps.jl:42-68   (Evolution against the dense exponential, atol 1e-10)
#   test/test_currents.jl:1-63    (DensityCurrents / Currents / Subcurrents)
#   test/test_workflows.jl:29-62  (README loop on a dense P)
# plus device-vs-reference comparisons of every bound method.  Needs Julia >= 1.9 with LatticeModels
# installed and liblm_b200.so built (python -c "import __graft_entry__ as g; g.build()") on a B200:
#     julia --project=<env with LatticeModels> julia/runtests_b200.jl
# NOT run in the build container (no Julia there); kept as the acceptance script of the binding.
using Test, LinearAlgebra, SparseArrays
using LatticeModels
include(joinpath(@__DIR__, "B200Backend.jl"))
using .B200Backend

@testset "Evolution (test/test_timedeps.jl:42-68) with B200Exp" begin
    l = SquareLattice(10, 10)
    Hs = qwz(l)
    Hd = dense(Hs)
    psi = groundstate(Hd)
    ts = 0:0.1:10
    correct_val = ComplexF64[]
    correct_ev = exp(-im * step(ts) * Hd.data)
    psidata = psi.data
    for _ in ts
        push!(correct_val, psidata[2])
        psidata = correct_ev * psidata
    end
    for method in (0, 1, 2, 3)                      # auto, Chebyshev, Taylor, Lanczos (KrylovKitExp semantics)
        val = ComplexF64[]
        for st in Evolution(B200Exp(tol = 1e-12, method = method), Hs, psi, timedomain = ts)
            push!(val, st.state.data[2])
        end
        @test val ≈ correct_val atol = 1e-10
    end
    # solver given as a type, host Ket state, iterator form (the KrylovKitExp line of the reference test)
    val1 = ComplexF64[]
    for state in Evolution(B200Exp, Hs, psi)(ts)
        push!(val1, state[1].data[2])
    end
    @test val1 ≈ correct_val atol = 1e-10
    # dense density matrix: U P U' on the device, read back into the host matrix every step
    P = psi ⊗ psi'
    Pref = copy(P.data)
    for st in Evolution(B200Exp(tol = 1e-13), Hs, P, timedomain = ts[1:21])
        @test st.state.data ≈ Pref atol = 1e-10
        Pref = correct_ev * Pref * correct_ev'
    end
    @test_throws ArgumentError for _ in Evolution(B200Exp(), Hs, psi)([0.0, -1.0]) end       # negative time step
end

@testset "Currents (test/test_currents.jl:1-63) on device states" begin
    l = SquareLattice(4, 4)
    x, y = coordvalues(l)
    H_0 = qwz(l)
    H_1 = qwz(l, field = LandauGauge(0.1))
    dg = diagonalize(H_0)
    P = densitymatrix(dg, statistics = FermiDirac, info = false)
    Pd = psi_densitymatrix(dg)                                     # the same P as an occupied-orbital block
    @test Matrix(Pd.data) ≈ P.data atol = 1e-12
    dc = DensityCurrents(H_1, P)
    dd = DensityCurrents(H_1, Pd)                                  # same constructor: check_samebases passes

    s1 = l[6]; s2 = l[11]
    @test dd[s1, s2] == -dd[s2, s1]
    @test dd[s1, s1] ≈ 0 atol = eps()
    @test dd[s1, s2] ≈ dc[s1, s2] atol = 1e-12
    @test currentsfromto(dd, s1) ≈ currentsfromto(dc, s1) atol = 1e-12
    @test currentsfrom(dd, s1).values ≈ currentsfrom(dc, s1).values atol = 1e-12
    @test Currents(dd) ≈ Currents(dc)
    @test all(isapprox.(findnz(dd)[3], findnz(Currents(dc))[3]; atol = 1e-12))
    @test findnz(dd)[1] == findnz(Currents(dc))[1] && findnz(dd)[2] == findnz(Currents(dc))[2]

    bs = AdjacencyMatrix(H_1)
    @test Currents(dd, bs) ≈ Currents(dc, bs)
    m1 = Currents(dd)[x .< y]
    m2 = Currents(dd[x .< y])                                      # SubCurrents through the reference's generic path
    m3 = Currents(dd, bs)[x .< y]
    @test m1 ≈ m2
    @test m1 ≈ m3

    # localdensity / localexpect / LocalOperatorCurrents
    @test localdensity(Pd).values ≈ localdensity(P).values atol = 1e-12
    sz = Operator(SpinBasis(1 // 2), ComplexF64[1 0; 0 -1])
    @test localexpect(sz, Pd).values ≈ localexpect(sz, P).values atol = 1e-12
    @test Currents(LocalOperatorCurrents(H_1, Pd, sz)) ≈ Currents(LocalOperatorCurrents(H_1, P, sz))
end

@testset "README loop (test/test_workflows.jl:29-62): dense P, Psi block, device-resident H" begin
    l = SquareLattice(10, 10)
    h(t) = tightbinding_hamiltonian(l, field = PointFlux(0.2 * min(t, 10) / 10, (5.5, 5.5)))
    dg = diagonalize(h(0))
    P0 = densitymatrix(dg, mu = 0, statistics = FermiDirac, info = false)
    Pd0 = psi_densitymatrix(dg, mu = 0)
    ts = 0:0.1:2
    ref = [(localdensity(P).values, Currents(DensityCurrents(H, P))) for (P, H, t) in Evolution(CachedExp(threshold = 1e-14), h, P0)(ts)]
    # (1) dense P through B200Exp: reference observables on the host copy
    k = 0
    for (P, H, t) in Evolution(B200Exp(tol = 1e-13), h, P0)(ts)
        k += 1
        @test localdensity(P).values ≈ ref[k][1] atol = 1e-10
        @test Currents(DensityCurrents(H, P)) ≈ ref[k][2]
    end
    # (2) Psi block on the device, same calls
    k = 0
    for (P, H, t) in Evolution(B200Exp(tol = 1e-13), h, Pd0)(ts)
        k += 1
        @test localdensity(P).values ≈ ref[k][1] atol = 1e-10
        @test Currents(DensityCurrents(H, P)) ≈ ref[k][2]
    end
    # (3) device-resident time-dependent Hamiltonian + asynchronous frame sink
    Hdev = B200Hamiltonian(h(0), 1 => NearestNeighbor(1); field = t -> PointFlux(0.2 * min(t, 10) / 10, (5.5, 5.5)))
    @test sparse(Hdev.data) ≈ h(0).data
    seq = TimeSequence(FrameSink(), Evolution(B200Exp(tol = 1e-13), t -> settime!(Hdev, t), Pd0)(ts))
    for (k, t) in enumerate(ts)
        @test seq[t].density.values ≈ ref[k][1] atol = 1e-10
        @test seq[t].currents ≈ ref[k][2]
    end
end
@testset "device eigensolver (diagonalize(H, :b200; n))" begin
    l = SquareLattice(12, 11)
    H = qwz(l)
    ref = diagonalize(dense(H))
    eig = diagonalize(H, :b200, n = 10)
    @test eig.values ≈ ref.values[1:10] atol = 1e-9
    @test norm(H.data * eig.states - eig.states * Diagonal(eig.values)) < 1e-8
    vals, P = eigs_lowest(H; n = 10)
    @test localdensity(P).values ≈ localdensity(projector(ref[1:10])).values atol = 1e-7
end
println("B200 backend: all reference testsets passed")
