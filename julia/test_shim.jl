# test_shim.jl -- what a maintainer runs ONCE on a box with Julia, PowerSpectra.jl and a B200 to accept the shim:
#
#   julia --project=<env with PowerSpectra, Healpix> julia/test_shim.jl /path/to/libpsb200.so [ngpus]
#
# Every inner loop the shim overrides (PowerSpectraB200.jl) is evaluated twice on the same synthetic inputs -- by the
# stock Julia methods, then, after `enable!`, by libpsb200 -- and compared entry by entry with the north-star criterion
# (relative error <= 1e-10 above 1e-30 of the row maximum, plus the condition-aware floor for cancelling sums).  No data
# files are needed.  Afterwards the reference's own test-suite can be run with the shim enabled
# (`PowerSpectraB200.enable!(lib); include("test/runtests.jl")`).
#
# STATUS: UNTESTED, like the shim itself -- there is no Julia toolchain in the build image.  The same comparisons run
# against the C/OpenMP restatement of these loops (oracle/) in tests/test_gpu_parity.py, and against multiprecision
# known answers in tests/test_highl_golden.py.

using PowerSpectra
using Healpix
using Test
using Random
using LinearAlgebra

include(joinpath(@__DIR__, "PowerSpectraB200.jl"))

lib = length(ARGS) >= 1 ? ARGS[1] : "libpsb200.so"
ngpus = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 1

# ---- synthetic inputs: smooth random window spectra and CMB-like spectra (no maps, no SHTs) -------------------
Random.seed!(20261017)
lmax = 767
ℓ = 0:lmax
window() = SpectralVector([(1 + 0.3 * randn()) / (1 + (l / 40)^2) * (isodd(l ÷ 7) ? -0.2 : 1.0) for l in ℓ])
tt() = SpectralVector([6000 * 2π / (max(l, 1) * (max(l, 1) + 1)) * exp(-(l / 1500)^2) * (1 + 0.05 * rand()) + 1e-5 for l in ℓ])
ee() = SpectralVector(0.02 .* parent(tt()))
te() = SpectralVector(0.1 .* parent(tt()) .* cos.(ℓ ./ 90))
ratio() = SpectralVector([sqrt(1 + (l / 2000)^2) * (1 + 0.1 * rand()) for l in ℓ])

"north-star check of B (libpsb200) against A (stock Julia): strict where the two agree that nothing cancels"
function close_enough(A, B; lo = 0)
    a, b = parent(A)[(lo + 1):end, (lo + 1):end], parent(B)[(lo + 1):end, (lo + 1):end]
    rowmax = maximum(abs, a; dims = 2)
    sel = abs.(a) .> 1e-30 .* rowmax
    # Float64 evaluations of a cancelling sum differ by ~1e-13 of the row scale (DESIGN.md section 3)
    return all(abs.(a[sel] .- b[sel]) .<= 1e-10 .* abs.(a[sel]) .+ 1e-13 .* (rowmax .* ones(size(a)))[sel])
end

# ---- 1. stock results ------------------------------------------------------------------------------------------
V = window()
stock = Dict{Symbol,Any}()
for (name, f!) in ((:M00, PowerSpectra.inner_mcm⁰⁰!), (:M02, PowerSpectra.inner_mcm⁰²!),
                   (:Mpp, PowerSpectra.inner_mcm⁺⁺!), (:Mmm, PowerSpectra.inner_mcm⁻⁻!))
    M = spectralzeros(0:lmax, 0:lmax)
    f!(M, V)
    stock[name] = M
end
sp4, rt4, W8 = [tt() for _ in 1:4], [ratio() for _ in 1:4], [window() for _ in 1:8]
eesp = [ee() for _ in 1:4]
tesp = [te() for _ in 1:4]
covcalls = Dict(
    :TTTT => C -> PowerSpectra.loop_covTTTT!(C, sp4..., rt4..., W8...),
    :EEEE => C -> PowerSpectra.loop_covEEEE!(C, eesp..., rt4..., W8...),
    :TTTE => C -> PowerSpectra.loop_covTTTE!(C, sp4[1], sp4[2], tesp[1], tesp[2], rt4[1], rt4[2], W8[1:4]...),
    :TETE => C -> PowerSpectra.loop_covTETE!(C, sp4[1], eesp[2], tesp[3], tesp[4], rt4[1], rt4[2], W8[1:5]...),
    :TEEE_planck => C -> PowerSpectra.loop_covTEEE_planck!(C, eesp[1], eesp[2], tesp[1], tesp[2], rt4[1], rt4[2], W8[1:4]...),
    :TEEE => C -> PowerSpectra.loop_covTEEE!(C, eesp[1], eesp[2], tesp[1], tesp[2], rt4[1], rt4[2], W8[1:4]...),
    :TTEE => C -> PowerSpectra.loop_covTTEE!(C, tesp[1], tesp[2], tesp[3], tesp[4], W8[1], W8[2]))
for (name, call) in covcalls
    C = spectralzeros(0:lmax, 0:lmax)
    call(C)
    stock[name] = C
end

# ---- 2. the same calls through libpsb200 ---------------------------------------------------------------------------
PowerSpectraB200.enable!(lib; ngpus = ngpus)

@testset "inner_mcm overrides" begin
    for (name, f!, lo) in ((:M00, PowerSpectra.inner_mcm⁰⁰!, 0), (:M02, PowerSpectra.inner_mcm⁰²!, 2),
                           (:Mpp, PowerSpectra.inner_mcm⁺⁺!, 2), (:Mmm, PowerSpectra.inner_mcm⁻⁻!, 2))
        M = spectralzeros(0:lmax, 0:lmax)
        f!(M, V)                                   # dispatches to the override (V is a SpectralVector{Float64})
        @test close_enough(stock[name], M; lo = lo)
        @test all(isfinite, parent(M))             # rows l < 2 of the spin-2 kinds: what the family routine yields
    end
end

@testset "loop_cov overrides" begin
    for (name, call) in covcalls
        C = spectralzeros(0:lmax, 0:lmax)
        call(C)                                    # an ambiguity MethodError here means the override signatures drifted
        lo = name in (:TTTT, :TTTE, :TTEE) ? 0 : 2
        @test close_enough(stock[name], C; lo = lo)
        @test parent(C) == parent(C)'              # symmetric by copy (src/covariance.jl:119)
    end
end

@testset "errors and extras" begin
    @test_throws ArgumentError PowerSpectraB200.mcm_call!(9, spectralzeros(0:7, 0:7), SpectralVector(ones(8)))
    Mpp, Mmm = spectralzeros(0:lmax, 0:lmax), spectralzeros(0:lmax, 0:lmax)
    PowerSpectraB200.mcm_call!(4, Mpp, V; 𝐌2 = Mmm)                      # fused M⁺⁺ / M⁻⁻
    @test close_enough(stock[:Mpp], Mpp; lo = 2) && close_enough(stock[:Mmm], Mmm; lo = 2)
    P = PowerSpectraB200.pinned_spectralzeros(0:lmax)
    PowerSpectra.inner_mcm⁰⁰!(P, V)
    @test close_enough(stock[:M00], P)
    PowerSpectraB200.free_pinned!(P)
end
