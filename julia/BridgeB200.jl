# BridgeB200.jl -- thin ccall shim that puts libbridge_b200.so (hand-written CUDA, sm_100a) behind Bridge.jl's
# own generic functions for the data-parallel hot path.  Host code stays Julia; see include/bridge_b200.h for the ABI
# and INTEGRATION.md for how a maintainer wires it in.
#
# STATUS: written against the header; NOT executed (no Julia runtime exists in the build environment).  The Python
# binding bridge.jl_b200/_cabi.py makes exactly the same calls and IS tested on the GPU (tests/test_gpu_parity.py).
#
# What it adds (methods only; no Bridge.jl source is modified):
#   PathEnsemble            device-resident container of P chains x S segments (replaces P*S (W, X) SamplePath pairs)
#   Bridge.sample!(E, Wiener())                         -> bb_wiener_sample          (src/wiener.jl:50-58)
#   Bridge.solve!(EulerMaruyama(), E, u, P)             -> bb_euler                  (src/euler.jl:135-152)
#   Bridge.solve!(Euler(), E, u, Po::Vector{<:Guide})   -> bb_guided_euler_ll        (src/euler.jl:247-268)
#   Bridge.llikelihood(LeftRule(), E, Po; skip)         -> bb_llikelihood / fused    (src/partialbridgenuH.jl:171-189)
#   pcn!(E, P, Po, rho, seed, iter)                     -> bb_pcn_step               (test/partialbridgenuH.jl:176-191)
#   solve!/llikelihood on plain SamplePath (P = 1 plumbing) go through a one-chain ensemble.
module BridgeB200

using Bridge, StaticArrays, LinearAlgebra
import Bridge: sample!, solve!, llikelihood, EulerMaruyama, Euler, LeftRule, SamplePath, Wiener, ContinuousTimeProcess

const lib = get(ENV, "BRIDGE_B200_LIB", joinpath(@__DIR__, "..", "bridge.jl_b200", "lib", "libbridge_b200.so"))

# ---- status codes -> the reference's own errors (include/bridge_b200.h)
function check(st::Cint)
    st == 0 && return nothing
    msg = unsafe_string(ccall((:bb_strerror, lib), Cstring, (Cint,), st))
    st == -4 && throw(DimensionMismatch("length(tt) != size(yy, 2)"))        # src/types.jl:127
    st == -5 && throw(AssertionError("m == length(v)"))                      # src/partialbridgenuH.jl:3
    if st == -8 || st == -9
        msg *= ": " * unsafe_string(ccall((:bb_last_cuda_error, lib), Cstring, ()))
    end
    error(msg)   # "Y and W differ in length." / "Time axis mismatch ..." / "Starting point has wrong length."
end

# ---- registry models (bb_model): the device cannot call Julia closures
struct BBModel
    id::Int32; d::Int32; dprime::Int32; reserved::Int32
    par::NTuple{32,Float64}
end
pad32(v) = ntuple(i -> i <= length(v) ? Float64(v[i]) : 0.0, 32)
bbmodel(::Wiener{Float64}) = BBModel(0, 1, 1, 0, pad32(()))
bbmodel(::Wiener{SVector{d,Float64}}) where {d} = BBModel(0, d, d, 0, pad32(()))
bbmodel(P::Bridge.LinPro) = (d = size(P.B, 1); BBModel(2, d, d, 0, pad32(vcat(vec(P.B'), P.μ, vec(P.σ')))))   # row-major
bbmodel(P::Bridge.Models.FitzHughNagumo) = BBModel(3, 2, 2, 0, pad32((P.ϵ, P.s, P.γ, P.β, P.σ1, P.σ2)))        # src/Models.jl:9-20
# user structs opt in by defining bbmodel(P), e.g. for project_partialbridge/partialbridge_fitzhugh.jl:36-46
#   BridgeB200.bbmodel(P::FitzhughDiffusion) = BridgeB200.BBModel(4, 2, 1, 0, BridgeB200.pad32((P.ϵ, P.s, P.γ, P.β, P.σ)))

# ---- context / ensemble handles
mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:bb_ctx_create, lib), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r))
        c = new(r[]); finalizer(c -> ccall((:bb_ctx_destroy, lib), Cint, (Ptr{Cvoid},), c.h), c); c
    end
end

mutable struct PathEnsemble
    h::Ptr{Cvoid}; ctx::Context
    P::Int; S::Int; N::Int; d::Int; dprime::Int
    function PathEnsemble(ctx::Context, P, S, N, d, dprime; double_buffer = true, store_x = true, chain_offset = 0)
        flags = UInt32((double_buffer ? 1 : 0) | (store_x ? 0 : 2))
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:bb_ens_create, lib), Cint, (Ptr{Cvoid}, Int64, Int32, Int32, Int32, Int32, UInt32, Ref{Ptr{Cvoid}}),
                    ctx.h, P, S, N, d, dprime, flags, r))
        E = new(r[], ctx, P, S, N, d, dprime)
        chain_offset != 0 && check(ccall((:bb_ens_set_chain_offset, lib), Cint, (Ptr{Cvoid}, Int64), E.h, chain_offset))
        finalizer(E -> ccall((:bb_ens_destroy, lib), Cint, (Ptr{Cvoid},), E.h), E); E
    end
end
setgrid!(E::PathEnsemble, seg, tt::Vector{Float64}) =
    check(ccall((:bb_ens_set_grid, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32), E.h, seg - 1, tt, length(tt)))
setstart!(E::PathEnsemble, u::SVector) = (v = collect(u);
    check(ccall((:bb_ens_set_start, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32), E.h, v, length(v), 1)))

# per-chain starting points: X0 is d x P (one column per chain), the ABI wants [P][d] -- the same bytes
setstart!(E::PathEnsemble, X0::Matrix{Float64}) =
    check(ccall((:bb_ens_set_start, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32), E.h, X0, length(X0), 0))

# ---- guiding tables: the constructors of Bridge.jl have already run the backward ODE on the host, or use
#      bb_backward_nuH / bb_backward_HV / bb_backward_LMmu to run it on the device (same R3 / Lyapunov schemes)
mutable struct Guide
    h::Ptr{Cvoid}
end
rowmajor(A::AbstractMatrix) = collect(vec(permutedims(A)))
function Guide(ctx::Context, Po::Bridge.PartialBridgeνH)           # fields Target, Pt, tt, ν, H, C  (src/partialbridgenuH.jl:122-130)
    d = length(Po.ν[1]); N = length(Po.tt)
    H = reduce(vcat, rowmajor.(Po.H)); ν = reduce(vcat, collect.(Po.ν))
    Bt = reduce(vcat, [rowmajor(Bridge.B(t, Po.Pt)) for t in Po.tt]); βt = reduce(vcat, [collect(Bridge.β(t, Po.Pt)) for t in Po.tt])
    r = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve H ν Bt βt check(ccall((:bb_guide_create, lib), Cint,
        (Ptr{Cvoid}, Int32, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Int32, Ref{Ptr{Cvoid}}),
        ctx.h, 1, N, d, 0, Po.tt, H, ν, C_NULL, C_NULL, Bt, βt, 0, r))
    g = Guide(r[]); finalizer(g -> ccall((:bb_guide_destroy, lib), Cint, (Ptr{Cvoid},), g.h), g); g
end
# GuidedBridge (kind 2: A = H♢, b = V) and PartialBridge (kind 3: A = L, b = μ, Mm = M, v) are built the same way.

# ---- the methods added to Bridge's generic functions
function sample!(E::PathEnsemble, ::Wiener; seed::UInt64 = UInt64(0), stream::UInt32 = UInt32(0))
    check(ccall((:bb_wiener_sample, lib), Cint, (Ptr{Cvoid}, UInt64, UInt32), E.h, seed, stream)); E
end
function solve!(::EulerMaruyama, E::PathEnsemble, u, P::ContinuousTimeProcess)
    setstart!(E, u); m = Ref(bbmodel(P))
    check(ccall((:bb_euler, lib), Cint, (Ptr{Cvoid}, Ref{BBModel}), E.h, m)); E
end
function solve!(::EulerMaruyama, E::PathEnsemble, u, P::ContinuousTimeProcess, guides::Vector{Guide}; skip = 0, store_x = true)
    setstart!(E, u); m = Ref(bbmodel(P)); hs = [g.h for g in guides]
    GC.@preserve guides check(ccall((:bb_guided_euler_ll, lib), Cint, (Ptr{Cvoid}, Ref{BBModel}, Ptr{Ptr{Cvoid}}, Int32, UInt32),
                                    E.h, m, hs, skip, store_x ? 1 : 0))
    E   # end points: bb_ens_get_f64(E, BB_F_XEND, ...)
end
function llikelihood(::LeftRule, E::PathEnsemble, P::ContinuousTimeProcess, guides::Vector{Guide}; skip = 0)
    m = Ref(bbmodel(P)); hs = [g.h for g in guides]
    GC.@preserve guides check(ccall((:bb_llikelihood, lib), Cint, (Ptr{Cvoid}, Ref{BBModel}, Ptr{Ptr{Cvoid}}, Int32), E.h, m, hs, skip))
    ll = Vector{Float64}(undef, E.P)
    check(ccall((:bb_ens_get_f64, lib), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Float64}), E.h, 0, 0, E.P, ll)); ll
end
"""One pCN / MH update of every chain: the body of `for iter in 1:iterations` in test/partialbridgenuH.jl:176-191."""
function pcn!(E::PathEnsemble, P::ContinuousTimeProcess, guides::Vector{Guide}, ρ, seed::UInt64, iter::Integer; skip = 0, store_x = true)
    m = Ref(bbmodel(P)); hs = [g.h for g in guides]
    GC.@preserve guides check(ccall((:bb_pcn_step, lib), Cint,
        (Ptr{Cvoid}, Ref{BBModel}, Ptr{Ptr{Cvoid}}, Float64, UInt64, UInt32, Int32, UInt32), E.h, m, hs, ρ, seed, iter, skip, store_x ? 1 : 0))
    acc = Ref{Int64}(0); check(ccall((:bb_ens_get_acc, lib), Cint, (Ptr{Cvoid}, Ref{Int64}), E.h, acc)); acc[]
end

"""Current paths of all chains as an array [d, N, S, P] (refreshes the chains whose last proposal was rejected)."""
function download_x(E::PathEnsemble, P::ContinuousTimeProcess, guides::Vector{Guide})
    m = Ref(bbmodel(P)); hs = [g.h for g in guides]
    GC.@preserve guides check(ccall((:bb_ens_refresh_x, lib), Cint, (Ptr{Cvoid}, Ref{BBModel}, Ptr{Ptr{Cvoid}}), E.h, m, hs))
    X = Array{Float64}(undef, E.d, E.N, E.S, E.P)   # column-major == the ABI's [P][S][N][d]
    check(ccall((:bb_ens_download, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Float64}), E.h, 1, 0, 0, E.P, X)); X
end

# ---- P = 1 plumbing: the reference's own signatures on SamplePath (a one-chain ensemble per call)
function solve!(::EulerMaruyama, Y::SamplePath{T}, u::T, W::SamplePath, P::ContinuousTimeProcess{T}, ctx::Context) where {T}
    N = length(W); N != length(Y) && error("Y and W differ in length.")
    d = length(u); dp = length(W.yy[1])
    E = PathEnsemble(ctx, 1, 1, N, d, dp; double_buffer = false)
    setgrid!(E, 1, W.tt); setstart!(E, SVector{d}(u...))
    w = reinterpret(Float64, W.yy)
    check(ccall((:bb_ens_upload, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Float64}), E.h, 0, 0, 0, 1, w))
    m = Ref(bbmodel(P)); check(ccall((:bb_euler, lib), Cint, (Ptr{Cvoid}, Ref{BBModel}), E.h, m))
    x = reinterpret(Float64, Y.yy)
    check(ccall((:bb_ens_download, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Float64}), E.h, 1, 0, 0, 1, x))
    Y.tt .= W.tt; Y
end

# ---- per-chain parameters: the `updateparams` branch of partialbridge_bolus3.jl:248-365 (bb_theta_* of the header)
struct BBThetaSpec   # == bb_theta_spec
    m::Int32; aux_kind::Int32
    L::NTuple{16,Float64}; Sigma::NTuple{16,Float64}; eps::Float64
    v::NTuple{64,Float64}                      # v[s][0..3], row per segment
    prior_kind::NTuple{8,Int32}; prior_a::NTuple{8,Float64}; prior_b::NTuple{8,Float64}
    start_sd::Float64; start_dir::NTuple{4,Float64}   # joint start-point random walk, bolus3.jl:311-318 (0 = off)
end
const AUXKIND = Dict(:fhn_matching => Int32(1), :fhn_linearised_end => Int32(2), :bolus => Int32(3))
function theta_attach!(E::PathEnsemble, P::ContinuousTimeProcess, L, Σ, ϵ, obs; aux = :fhn_matching, priors = Dict(),
                       start_sd = 0.0, start_dir = (0.0, 0.0, 0.0, 0.0))
    m, d = size(L)
    Lr = zeros(16); Sr = zeros(16); vr = zeros(64); pk = zeros(Int32, 8); pa = zeros(8); pb = zeros(8)
    for i in 1:m, j in 1:d; Lr[(i-1)*d + j] = L[i, j]; end        # row-major in the ABI
    for i in 1:m, j in 1:m; Sr[(i-1)*m + j] = Σ[i, j]; end
    for (s, v) in enumerate(obs), i in 1:m; vr[(s-1)*4 + i] = v[i]; end
    for (k, (kind, a, b)) in priors; pk[k+1] = 1; pa[k+1] = a; pb[k+1] = b; end   # k: 0-based parameter index
    spec = Ref(BBThetaSpec(m, AUXKIND[aux], Tuple(Lr), Tuple(Sr), ϵ, Tuple(vr), Tuple(pk), Tuple(pa), Tuple(pb), start_sd, Tuple(Float64.(start_dir))))
    mdl = Ref(bbmodel(P))
    check(ccall((:bb_theta_attach, lib), Cint, (Ptr{Cvoid}, Ref{BBModel}, Ref{BBThetaSpec}), E.h, mdl, spec)); E
end
theta_solve!(E::PathEnsemble; skip = 0, store_x = true) =   # backward ODEs of every chain + solve! + llikelihood
    check(ccall((:bb_theta_guided_euler_ll, lib), Cint, (Ptr{Cvoid}, Int32, UInt32), E.h, skip, store_x ? 1 : 0))
theta_pcn!(E::PathEnsemble, ρ, seed::UInt64, iter::Integer; skip = 0, store_x = true) =
    check(ccall((:bb_theta_pcn_step, lib), Cint, (Ptr{Cvoid}, Float64, UInt64, UInt32, Int32, UInt32),
                E.h, ρ, seed, iter, skip, store_x ? 1 : 0))
function theta_param_step!(E::PathEnsemble, rw_sd, seed::UInt64, iter::Integer; skip = 0, store_x = true)
    sd = zeros(8); sd[1:length(rw_sd)] .= rw_sd
    check(ccall((:bb_theta_param_step, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, UInt64, UInt32, Int32, UInt32),
                E.h, sd, seed, iter, skip, store_x ? 1 : 0))
    acc = Ref{Int64}(0); check(ccall((:bb_theta_get_acc, lib), Cint, (Ptr{Cvoid}, Ref{Int64}), E.h, acc)); acc[]
end
function theta(E::PathEnsemble)   # param(P) of every chain: 8 x P (column per chain)
    θ = Matrix{Float64}(undef, 8, E.P)
    check(ccall((:bb_theta_get, lib), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Float64}), E.h, 0, 0, E.P, θ)); θ
end

end # module
