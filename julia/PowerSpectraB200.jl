# PowerSpectraB200.jl -- drop-in shim: routes PowerSpectra.jl's Wigner-3j inner loops to libpsb200.so.
#
#   using PowerSpectra
#   include("julia/PowerSpectraB200.jl")      # after PowerSpectra is loaded
#   PowerSpectraB200.enable!("/path/to/libpsb200.so"; ngpus = 1)
#   M = mcm(:TT, alm1, alm2)                   # unchanged user code; the (l1,l2) loops now run on the GPU
#
# What is overridden (method re-definition, same signatures as the reference):
#   inner_mcm⁰⁰!  inner_mcm⁰²!  inner_mcm⁺⁺!  inner_mcm⁻⁻!          src/modecoupling.jl:78,99,123,143
#   loop_covTTTT! loop_covEEEE! loop_covTTTE! loop_covTETE!
#   loop_covTEEE! loop_covTEEE_planck! loop_covTTEE!                 src/covariance.jl:92,153,208,261,337,376,422
#   quickpolΞ!(𝚵, ν₁, ν₂, s₁, s₂, ω₁, ω₂, buf1, buf2)                src/beam.jl:72-101
# Everything else -- mcm, coupledcov, CovarianceWorkspace, window_function_W!, SpectralArray, `\`,
# decouple_covmat, master -- is the reference's own code and keeps running on the host.
#
# STATUS: UNTESTED.  No Julia toolchain exists in the build image or on the GPU boxes, so this file
# has never been executed.  The C ABI it binds is exercised (same argument order, same memory
# layout) by the ctypes mirror in powerspectra.jl_b200/ and its tests.

module PowerSpectraB200

using PowerSpectra
using LinearAlgebra
import PowerSpectra: SpectralArray, SpectralVector

const LIB = Ref{String}("libpsb200.so")
const NGPUS = Ref{Cint}(1)

"0-based contiguous x[l], l = 0..lmax (entries below the first stored multipole are never read)."
function zero_based(x::SpectralVector{Float64}, lmax::Int)
    out = zeros(Float64, lmax + 1)
    lo = max(firstindex(x), 0)
    lastindex(x) >= lmax || throw(ArgumentError("vector ends at l=$(lastindex(x)), need lmax=$lmax"))
    @inbounds for l in lo:lmax
        out[l + 1] = x[l]
    end
    return out
end

function check(rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:psb200_last_error, LIB[]), Cstring, ()))
    rc == 1 && throw(ArgumentError(msg))
    rc == 6 && throw(LinearAlgebra.SingularException(0))      # exactly singular system in a device-side solve
    error("libpsb200 error $rc: $msg")
end

function mcm_call!(kind::Int, 𝐌::SpectralArray{Float64,2}, V::SpectralVector{Float64};
                   𝐌2::Union{Nothing,SpectralArray{Float64,2}} = nothing)
    @assert axes(𝐌, 1) == axes(𝐌, 2)
    lmin, lmax = first(axes(𝐌, 1)), last(axes(𝐌, 1))
    firstindex(V) == 0 || throw(ArgumentError("window spectrum must start at l = 0 (got $(firstindex(V)))"))
    v = collect(parent(V))                       # V is 0-indexed: SpectralVector(alm2cl(...)[1:lmax+1])
    P = parent(𝐌)                                # dense column-major N x N
    p2 = 𝐌2 === nothing ? Ptr{Cdouble}(C_NULL) : pointer(parent(𝐌2))
    GC.@preserve v P 𝐌2 begin
        rc = ccall((:psb200_mcm, LIB[]), Cint,
                   (Cint, Cint, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Clong, Ptr{Cdouble}, Cint),
                   kind, lmin, lmax, v, length(v), P, stride(P, 2), p2, NGPUS[])
    end
    check(rc)
    return 𝐌
end

function cov_call!(block::Int, 𝐂::SpectralArray{Float64,2}, spectra, ratios, Ws)
    @assert axes(𝐂, 1) == axes(𝐂, 2)
    lmin, lmax = first(axes(𝐂, 1)), last(axes(𝐂, 1))
    sp = [zero_based(s, lmax) for s in spectra]
    rt = [zero_based(r, lmax) for r in ratios]
    all(w -> firstindex(w) == 0, Ws) || throw(ArgumentError("window spectra must start at l = 0"))
    ws = [collect(parent(w)) for w in Ws]        # 0-indexed, length workspace.lmax + 1
    lenW = minimum(length, ws)
    psp, prt, pws = pointer.(sp), pointer.(rt), pointer.(ws)
    P = parent(𝐂)
    GC.@preserve sp rt ws psp prt pws P begin
        rc = ccall((:psb200_cov, LIB[]), Cint,
                   (Cint, Cint, Cint, Ptr{Ptr{Cdouble}}, Cint, Ptr{Ptr{Cdouble}}, Cint,
                    Ptr{Ptr{Cdouble}}, Cint, Cint, Ptr{Cdouble}, Clong, Cint),
                   block, lmin, lmax, psp, length(sp), prt, length(rt), pws, length(ws), lenW,
                   P, stride(P, 2), NGPUS[])
    end
    check(rc)
    return 𝐂
end

"""
quickpolΞ! on the GPU.  𝚵 wraps a BandedMatrix over 0:lmax (docs/src/beams.md); its `data` field is the
(l+u+1) x n band storage the C ABI addresses directly.  quickpolW (src/beam.jl:43-56) stays on the host.
"""
function quickpol_call!(𝚵::SpectralArray{Float64,2}, ν₁, ν₂, s₁, s₂, ω₁, ω₂)
    size(𝚵, 1) != size(𝚵, 2) && throw(ArgumentError("𝚵 is not square."))
    lmax = lastindex(𝚵, 1)
    B = parent(𝚵)                                # the BandedMatrix (Base.parent(::SpectralArray) unwraps the OffsetArray, src/spectralarray.jl:45)
    W = collect(parent(PowerSpectra.quickpolW(ω₁, ω₂)))
    bl, bu = PowerSpectra.BandedMatrices.bandwidths(B)
    data = PowerSpectra.BandedMatrices.bandeddata(B)      # data[u + 1 + i - j, j] = B[i, j]
    # `𝚵 .*= sgn` (:98-99): the library overwrites every entry the loop visits (sign included), so scaling the
    # whole band storage first leaves exactly the unvisited entries (rows / columns below 2) multiplied by sgn
    flip = isodd(s₁ + s₂ + ν₁ + ν₂)
    flip && (data .*= -1)
    rc = GC.@preserve W data begin
        ccall((:psb200_quickpol_xi, LIB[]), Cint,
              (Cint, Cint, Cint, Cint, Cint, Ptr{Cdouble}, Cint, Cint, Cint, Ptr{Cdouble}, Clong, Cint),
              ν₁, ν₂, s₁, s₂, lmax, W, length(W), bl, bu, data, stride(data, 2), NGPUS[])
    end
    rc != 0 && flip && (data .*= -1)             # a failed call leaves the matrix as it was handed in
    check(rc)
    return 𝚵
end

"""
    pinned_spectralzeros(lmin:lmax; interleave = false) -> SpectralArray{Float64,2}

A zero N x N `SpectralArray` over `lmin:lmax` in page-locked memory of the library (psb200_host_alloc): the result of a
host call lands in it at the full PCIe rate (about 55 GB/s against 20 GB/s into a pageable `spectralzeros`).  With
`interleave = true` the 2 MB pieces of the array alternate between the NUMA nodes of the host, which is what a call on
several GPUs of a two-socket box wants.  Pass it to the in-place methods (`PowerSpectra.inner_mcm⁰⁰!(𝐌, V)` and the
`loop_cov*!` family); release it with `free_pinned!(𝐌)` when done (the memory is not garbage collected).
"""
function pinned_spectralzeros(r::AbstractUnitRange; interleave::Bool = false)
    n = length(r)
    p = ccall((:psb200_host_alloc, LIB[]), Ptr{Cdouble}, (Csize_t, Cint), n * n * sizeof(Float64), interleave ? 1 : 0)
    p == C_NULL && error("libpsb200: " * unsafe_string(ccall((:psb200_last_error, LIB[]), Cstring, ())))
    A = unsafe_wrap(Array, p, (n, n); own = false)            # zero-filled by the library
    return SpectralArray(A, (first(r) - 1, first(r) - 1))
end

"Release an array made by `pinned_spectralzeros`; it must not be touched afterwards."
free_pinned!(𝐌::SpectralArray{Float64,2}) =
    (ccall((:psb200_host_free, LIB[]), Cint, (Ptr{Cvoid},), pointer(parent(𝐌))) == 0 || throw(ArgumentError("not a pinned_spectralzeros array")); nothing)

"""
    enable!(libpath = "libpsb200.so"; ngpus = 1)

Re-define the inner loops of PowerSpectra to call the B200 library.  `ngpus = 0` uses every
visible GPU of the box (work-balanced l1 row bands, gathered on GPU 0).
"""
function enable!(libpath::AbstractString = "libpsb200.so"; ngpus::Integer = 1)
    LIB[] = String(libpath)
    NGPUS[] = Cint(ngpus)
    ccall((:psb200_device_count, LIB[]), Cint, ()) > 0 ||
        error("libpsb200: no CUDA device visible (there is no CPU fallback; keep the stock PowerSpectra loops instead)")
    @eval PowerSpectra begin
        inner_mcm⁰⁰!(𝐌::SpectralArray{Float64,2}, V::SpectralVector{Float64}) = $(mcm_call!)(0, 𝐌, V)
        inner_mcm⁰²!(𝐌::SpectralArray{Float64,2}, V::SpectralVector{Float64}) = $(mcm_call!)(1, 𝐌, V)
        inner_mcm⁺⁺!(𝐌::SpectralArray{Float64,2}, V::SpectralVector{Float64}) = $(mcm_call!)(2, 𝐌, V)
        inner_mcm⁻⁻!(𝐌::SpectralArray{Float64,2}, V::SpectralVector{Float64}) = $(mcm_call!)(3, 𝐌, V)

        # The spectra / ratio arguments carry the reference's own types (src/covariance.jl:92-97 etc.) with T = Float64,
        # so every override is strictly more specific than the method it shadows (typing only 𝐂 would make the
        # two methods ambiguous: ours wins on argument 1, the reference's on arguments 2 onward).
        loop_covTTTT!(𝐂::SpectralArray{Float64,2}, TTip::SpectralVector{Float64}, TTjq::SpectralVector{Float64}, TTiq::SpectralVector{Float64}, TTjp::SpectralVector{Float64},
                      r_ip::SpectralVector{Float64}, r_jq::SpectralVector{Float64}, r_iq::SpectralVector{Float64}, r_jp::SpectralVector{Float64},
                      W1, W2, W3, W4, W5, W6, W7, W8) =
            $(cov_call!)(0, 𝐂, (TTip, TTjq, TTiq, TTjp), (r_ip, r_jq, r_iq, r_jp), (W1, W2, W3, W4, W5, W6, W7, W8))
        loop_covEEEE!(𝐂::SpectralArray{Float64,2}, EEip::SpectralVector{Float64}, EEjq::SpectralVector{Float64}, EEiq::SpectralVector{Float64}, EEjp::SpectralVector{Float64},
                      r_ip::SpectralVector{Float64}, r_jq::SpectralVector{Float64}, r_iq::SpectralVector{Float64}, r_jp::SpectralVector{Float64},
                      W1, W2, W3, W4, W5, W6, W7, W8) =
            $(cov_call!)(1, 𝐂, (EEip, EEjq, EEiq, EEjp), (r_ip, r_jq, r_iq, r_jp), (W1, W2, W3, W4, W5, W6, W7, W8))
        loop_covTTTE!(𝐂::SpectralArray{Float64,2}, TTip::SpectralVector{Float64}, TTjp::SpectralVector{Float64}, TEiq::SpectralVector{Float64}, TEjq::SpectralVector{Float64},
                      r_ip::SpectralVector{Float64}, r_jp::SpectralVector{Float64}, W1, W2, W3, W4) =
            $(cov_call!)(2, 𝐂, (TTip, TTjp, TEiq, TEjq), (r_ip, r_jp), (W1, W2, W3, W4))
        loop_covTETE!(𝐂::SpectralArray{Float64,2}, TTip::SpectralVector{Float64}, EEjq::SpectralVector{Float64}, TEiq::SpectralVector{Float64}, TEjp::SpectralVector{Float64},
                      r_TT_ip::SpectralVector{Float64}, r_PP_jq::SpectralVector{Float64}, W1, W2, W3, W4, W5) =
            $(cov_call!)(3, 𝐂, (TTip, EEjq, TEiq, TEjp), (r_TT_ip, r_PP_jq), (W1, W2, W3, W4, W5))
        loop_covTEEE_planck!(𝐂::SpectralArray{Float64,2}, EEjq::SpectralVector{Float64}, EEjp::SpectralVector{Float64}, TEip::SpectralVector{Float64}, TEiq::SpectralVector{Float64},
                             r_EE_jq::SpectralVector{Float64}, r_EE_jp::SpectralVector{Float64}, W1, W2, W3, W4) =
            $(cov_call!)(4, 𝐂, (EEjq, EEjp, TEip, TEiq), (r_EE_jq, r_EE_jp), (W1, W2, W3, W4))
        loop_covTEEE!(𝐂::SpectralArray{Float64,2}, EEjq::SpectralVector{Float64}, EEjp::SpectralVector{Float64}, TEip::SpectralVector{Float64}, TEiq::SpectralVector{Float64},
                      r_EE_jq::SpectralVector{Float64}, r_EE_jp::SpectralVector{Float64}, W1, W2, W3, W4) =
            $(cov_call!)(5, 𝐂, (EEjq, EEjp, TEip, TEiq), (r_EE_jq, r_EE_jp), (W1, W2, W3, W4))
        loop_covTTEE!(𝐂::SpectralArray{Float64,2}, TEip::SpectralVector{Float64}, TEiq::SpectralVector{Float64}, TEjq::SpectralVector{Float64}, TEjp::SpectralVector{Float64}, W1, W2) =
            $(cov_call!)(6, 𝐂, (TEip, TEiq, TEjq, TEjp), (), (W1, W2))

        quickpolΞ!(𝚵::SpectralArray{Float64,2}, ν₁, ν₂, s₁, s₂, ω₁::Alm, ω₂::Alm,
                   buf1::Array{Array{Float64,1},1}, buf2::Array{Array{Float64,1},1}) =
            $(quickpol_call!)(𝚵, ν₁, ν₂, s₁, s₂, ω₁, ω₂)
    end
    return nothing
end

"""
    mcm_EE_BB_fused(alm₁, alm₂; lmin = 0, lmax = nothing) -> (𝐌⁺⁺, 𝐌⁻⁻)

Optional extra entry point: both spin-2 blocks from ONE evaluation of the (0,-2,2) family
(`mcm(:EE_BB, ...)` in the reference evaluates it twice, src/modecoupling.jl:213-214).
"""
function mcm_EE_BB_fused(alm₁, alm₂; lmin = 0, lmax = nothing)
    lmax = isnothing(lmax) ? min(alm₁.lmax, alm₂.lmax) : lmax
    V = SpectralVector(PowerSpectra.alm2cl(alm₁, alm₂)[1:(lmax + 1)])
    𝐌⁺⁺ = PowerSpectra.spectralzeros(lmin:lmax, lmin:lmax)
    𝐌⁻⁻ = PowerSpectra.spectralzeros(lmin:lmax, lmin:lmax)
    mcm_call!(4, 𝐌⁺⁺, V; 𝐌2 = 𝐌⁻⁻)
    return 𝐌⁺⁺, 𝐌⁻⁻
end

"""
    mcm_master(maskT₁, maskP₁, maskT₂, maskP₂; lmin = 0, lmax = nothing)

Optional extra entry point: every mode-coupling matrix `maskedalm2spectra` asks for
(src/modecoupling.jl:348-362: TT, TE(=TB), ET(=BT) and the (EE_BB, EB_BE) blocks) from ONE fused
GPU pass over both 3j families (the reference makes 5 `mcm` calls plus the tuple call, i.e. 11
family evaluations per pair).  Returns (𝐌⁰⁰, 𝐌⁰²_TP, 𝐌⁰²_PT, 𝐌⁺⁺, 𝐌⁻⁻).
"""
function mcm_master(maskT₁, maskP₁, maskT₂, maskP₂; lmin = 0, lmax = nothing)
    lmax = isnothing(lmax) ? minimum(a.lmax for a in (maskT₁, maskP₁, maskT₂, maskP₂)) : lmax
    V = [collect(PowerSpectra.alm2cl(a, b)[1:(lmax + 1)]) for (a, b) in
         ((maskT₁, maskT₂), (maskT₁, maskP₂), (maskP₁, maskT₂), (maskP₁, maskP₂))]
    out = [PowerSpectra.spectralzeros(lmin:lmax, lmin:lmax) for _ in 1:5]
    P = parent.(out)
    GC.@preserve V P begin
        rc = ccall((:psb200_mcm_master, LIB[]), Cint,
                   (Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cint,
                    Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Clong, Cint),
                   lmin, lmax, V[1], V[2], V[3], V[4], lmax + 1,
                   P[1], P[2], P[3], P[4], P[5], stride(P[1], 2), NGPUS[])
    end
    check(rc)
    return Tuple(out)
end

"""
    mcm_solve(spec, alm₁, alm₂, pCl; lmin = 0, lmax = nothing)

Optional extra entry point: `mcm(spec, alm₁, alm₂; lmin) \\ pCl` (src/modecoupling.jl:359-362,
src/blockspectralmatrix.jl:124-129) with the LU solve on the GPU that holds the matrix -- only the spectrum
crosses PCIe.  `spec` ∈ (:TT, :TE, :ET, :TB, :BT, :M⁺⁺); for the block systems pass `:EE_BB` / `:EB_BE` and
`pCl = (pCl_1, pCl_2)` (stacked like `[pCl_EE; pCl_BB]`, src/modecoupling.jl:365-379); returns a SpectralVector
(a tuple of two for the block systems).
"""
function mcm_solve(spec::Symbol, alm₁, alm₂, pCl; lmin = 0, lmax = nothing)
    lmax = isnothing(lmax) ? min(alm₁.lmax, alm₂.lmax) : lmax
    sys = spec in (:TT, :M⁰⁰) ? 0 : spec in (:TE, :ET, :TB, :BT, :M⁰², :M²⁰) ? 1 : spec == :M⁺⁺ ? 2 :
          spec == :M⁻⁻ ? 3 : spec == :EE_BB ? 4 : spec == :EB_BE ? 5 : throw(ArgumentError("$(spec) not a valid spectrum."))
    v = collect(PowerSpectra.alm2cl(alm₁, alm₂)[1:(lmax + 1)])
    N = lmax - lmin + 1
    rhs = sys >= 4 ? vcat((Float64[p[l] for l in lmin:lmax] for p in pCl)...) : Float64[pCl[l] for l in lmin:lmax]
    out = similar(rhs)
    GC.@preserve v rhs out begin
        rc = ccall((:psb200_mcm_solve, LIB[]), Cint,
                   (Cint, Cint, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Clong, Cint, Ptr{Cdouble}, Clong, Cint),
                   sys, lmin, lmax, v, length(v), rhs, length(rhs), 1, out, length(out), NGPUS[])
    end
    check(rc)
    wrap(x) = SpectralVector(x, lmin:lmax)      # same axes as `M \ pCl` returns (B.parent.offsets)
    return sys >= 4 ? (wrap(out[1:N]), wrap(out[(N + 1):end])) : wrap(out)
end

"""
    maskedalm2spectra_device(maskedmap₁vec, maskT₁, maskP₁, maskedmap₂vec, maskT₂, maskP₂; lmin = 0)

`maskedalm2spectra` (src/modecoupling.jl:341-377) with the five mode-coupling matrices built in one fused GPU pass and
ALL its solves done on the device (psb200_master_solve): returns the same Dict of nine decoupled spectra; no matrix is
ever copied to the host.
"""
function maskedalm2spectra_device(m₁::Vector, maskT₁, maskP₁, m₂::Vector, maskT₂, maskP₂; lmin = 0)
    lmax = minimum(a.lmax for a in (maskT₁, maskP₁, maskT₂, maskP₂))
    V = [collect(PowerSpectra.alm2cl(a, b)[1:(lmax + 1)]) for (a, b) in
         ((maskT₁, maskT₂), (maskT₁, maskP₂), (maskP₁, maskT₂), (maskP₁, maskP₂))]
    names = (:TT, :TE, :ET, :TB, :BT, :EE, :BB, :EB, :BE)
    idx = Dict('T' => 1, 'E' => 2, 'B' => 3)
    N = lmax - lmin + 1
    pcl = Matrix{Float64}(undef, N, 9)
    for (k, n) in enumerate(names)
        x, y = String(n)
        pcl[:, k] = PowerSpectra.alm2cl(m₁[idx[x]], m₂[idx[y]])[(lmin + 1):(lmax + 1)]
    end
    cl = similar(pcl)
    GC.@preserve V pcl cl begin
        rc = ccall((:psb200_master_solve, LIB[]), Cint,
                   (Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cint,
                    Ptr{Cdouble}, Clong, Ptr{Cdouble}, Clong, Cint),
                   lmin, lmax, V[1], V[2], V[3], V[4], lmax + 1, pcl, N, cl, N, NGPUS[])
    end
    check(rc)
    return Dict{Symbol,SpectralVector}(n => SpectralVector(cl[:, k], lmin:lmax) for (k, n) in enumerate(names))
end

"""
    decouple_covmat_device(Y, B1, B2)

`decouple_covmat` (src/covariance.jl:8-14) on the GPU: lu(B1'), lu(B2') and both n-right-hand-side solves in
libpsb200 (cuSOLVER getrf/getrs).
"""
function decouple_covmat_device(Y::SpectralArray{Float64,2}, B1::SpectralArray{Float64,2}, B2::SpectralArray{Float64,2})
    M = deepcopy(Y)
    C, P1, P2 = parent(M), parent(B1), parent(B2)
    n = size(C, 1)
    (size(C) == size(P1) == size(P2) == (n, n)) || throw(ArgumentError("decouple_covmat needs three square matrices of one size"))
    GC.@preserve C P1 P2 begin
        rc = ccall((:psb200_decouple_covmat, LIB[]), Cint,
                   (Cint, Ptr{Cdouble}, Clong, Ptr{Cdouble}, Clong, Ptr{Cdouble}, Clong, Ptr{Cdouble}, Clong),
                   n, C, stride(C, 2), P1, stride(P1, 2), P2, stride(P2, 2), C, stride(C, 2))
    end
    check(rc)
    return M
end

"""
    map2alm_device(maps...; lmax, niter = 3, scale = 1.0) -> Alm

`Healpix.map2alm(scale .* maps[1] .* maps[2] .* ...; lmax, niter)` on the GPU (psb200_map2alm: ring FFTs, Legendre
recurrences and the Jacobi iterations all on the device; 1 to 3 RING-ordered Float64 maps of one nside).
"""
function map2alm_device(maps::PowerSpectra.HealpixMap{Float64,PowerSpectra.RingOrder}...; lmax::Integer, niter::Integer = 3,
                        scale::Real = 1.0)
    1 <= length(maps) <= 3 || throw(ArgumentError("one to three maps"))
    nside = maps[1].resolution.nside
    all(m -> m.resolution.nside == nside, maps) || throw(ArgumentError("maps of different resolution"))
    alm = PowerSpectra.Alm(lmax, lmax, zeros(ComplexF64, PowerSpectra.numberOfAlms(lmax, lmax)))
    pix = [parent(m) for m in maps]
    ptrs = [pointer(p) for p in pix]
    GC.@preserve pix ptrs alm begin
        rc = ccall((:psb200_map2alm, LIB[]), Cint,
                   (Cint, Cint, Cint, Cint, Ptr{Ptr{Cdouble}}, Cdouble, Ptr{Cdouble}),
                   nside, lmax, niter, length(pix), ptrs, Float64(scale), alm.alm)      # ComplexF64 = interleaved (re, im)
    end
    check(rc)
    return alm
end

"""
    map2alm_many(maps, products; lmax, niter = 3, scales = ones(length(products)), ngpus = NGPUS[]) -> Vector{Alm}

`[map2alm(scales[k] .* prod(maps[i] for i in products[k]); lmax, niter) for k in eachindex(products)]` in one library call
(psb200_map2alm_many): every distinct map crosses PCIe once per GPU instead of once per product, and the products are dealt
to `ngpus` devices.  `products[k]` holds one to three 1-based indices into `maps`.  What a workspace needs: all the
`effective_weight_alm!` products of its masks and variance maps (src/workspace.jl:141-171).
"""
function map2alm_many(maps::Vector{<:PowerSpectra.HealpixMap{Float64,PowerSpectra.RingOrder}}, products::Vector{<:AbstractVector{<:Integer}};
                      lmax::Integer, niter::Integer = 3, scales::Vector{Float64} = ones(length(products)), ngpus::Integer = NGPUS[])
    nside = maps[1].resolution.nside
    all(m -> m.resolution.nside == nside, maps) || throw(ArgumentError("maps of different resolution"))
    all(p -> 1 <= length(p) <= 3, products) || throw(ArgumentError("one to three factors per product"))
    idx = fill(Cint(-1), 3 * length(products))
    for (k, p) in enumerate(products), (f, i) in enumerate(p)
        idx[3 * (k - 1) + f] = Cint(i - 1)
    end
    out = [PowerSpectra.Alm(lmax, lmax, zeros(ComplexF64, PowerSpectra.numberOfAlms(lmax, lmax))) for _ in products]
    pix = [parent(m) for m in maps]
    mp = [pointer(p) for p in pix]
    op = [Ptr{Cdouble}(pointer(a.alm)) for a in out]
    GC.@preserve pix mp out op idx scales begin
        rc = ccall((:psb200_map2alm_many, LIB[]), Cint,
                   (Cint, Cint, Cint, Cint, Ptr{Ptr{Cdouble}}, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Ptr{Cdouble}}, Cint),
                   nside, lmax, niter, length(pix), mp, length(products), idx, scales, op, ngpus)
    end
    check(rc)
    return out
end

"""
    alm2cl_device(a, b) -> Vector{Float64}

`Healpix.alm2cl(a, b)` on the GPU for two full alm (mmax == lmax) of one lmax.
"""
function alm2cl_device(a::PowerSpectra.Alm, b::PowerSpectra.Alm)
    (a.lmax == b.lmax && a.mmax == a.lmax && b.mmax == b.lmax) || throw(ArgumentError("two full alm of one lmax"))
    cl = Vector{Float64}(undef, a.lmax + 1)
    GC.@preserve a b cl begin
        rc = ccall((:psb200_alm2cl, LIB[]), Cint, (Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), a.lmax, a.alm, b.alm, cl)
    end
    check(rc)
    return cl
end

"""
    enable_workspace!()

Optional: re-define `effective_weight_alm!` (src/workspace.jl:141-171) so that the mask products and their `map2alm`
run on the GPU; `window_function_W!` (:174-213) then works unchanged on the alm it caches.  HealpixMap workspaces only.
"""
function enable_workspace!()
    @eval PowerSpectra begin
        function effective_weight_alm!(workspace::CovarianceWorkspace{Float64,Dict{Tuple{String,Symbol},HealpixMap{Float64,RingOrder,Vector{Float64}}}},
                                       A, i, j, α)
            haskey(workspace.effective_weights, (A, i, j, α)) && return workspace.effective_weights[(A, i, j, α)]
            X, Y = split_maptype(α)
            lmax = workspace.lmax
            m_iX, m_jY = workspace.mask_p[i, X], workspace.mask_p[j, Y]
            if A == :∅∅
                w = $(map2alm_device)(m_iX, m_jY; lmax = lmax)
            elseif (A in (:II, :QQ, :UU)) && i == j
                w = $(map2alm_device)(m_iX, m_jY, workspace.weight_p[i, A]; lmax = lmax, scale = pixsize(m_iX))
            else
                return Alm(lmax, lmax, zeros(ComplexF64, numberOfAlms(lmax, lmax)))
            end
            workspace.effective_weights[A, i, j, α] = w
            return w
        end
    end
    return nothing
end

end # module
