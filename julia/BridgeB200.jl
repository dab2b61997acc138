# BridgeB200.jl -- thin ccall shim that puts libbridge_b200.so (hand-written CUDA, sm_100a) behind Bridge.jl's
# own generic functions for the data-parallel hot path.  Host code stays Julia; see include/bridge_b200.h for the ABI
# and INTEGRATION.md for how a maintainer wires it in.
#
# STATUS: written against the header; NOT executed (no Julia runtime exists in the build environment).  Every `ccall`
# below is checked mechanically against include/bridge_b200.h by tools/check_shim.py (function name, arity, argument
# classes and widths, and the byte layout of the mirrored structs BBModel / BBAux / BBThetaSpec against gcc's
# sizeof/offsetof); tests/test_cabi.py runs that check.  The Python binding bridge.jl_b200/_cabi.py makes the same
# calls and IS tested on the GPU (tests/test_gpu_*.py).
#
# Two layers:
#  (1) the reference's OWN signatures on SamplePath (one path per call), so that example/ and project_partialbridge/
#      scripts run unchanged after `using BridgeB200`:
#        sample!(W, Wiener{T}())                                   src/wiener.jl:24-58
#        solve!(EulerMaruyama(), Y, u, W, P)                        src/euler.jl:135-152   (registry targets)
#        solve!(Euler(), Y, u, W, P°) -> Y.yy[end]                  src/euler.jl:246-268   (P° ∈ GuidedBridge | PartialBridge | PartialBridgeνH)
#        llikelihood(LeftRule(), X, P°; skip = 0)                   src/partialbridgenuH.jl:171, guip.jl:429, partialbridge.jl:67
#        bridge!(Y, W, P°), bridge!(X, x0, W, P°)                   src/deprecated.jl:16-17, project/partialbridge.jl:63
#        innovations!(EulerMaruyama(), W, Y, P)                     src/euler.jl:357-376
#        lptilde(x, P°::PartialBridgeνH), lptilde(P°::GuidedBridge, u)
#      They are more specific than Bridge's methods, so dispatch prefers them; `BridgeB200.enable!(false)` hands every
#      call back to Bridge's CPU methods (`invoke`).  A process P is on the device iff `bbmodel(P)` is defined for it.
#  (2) PathEnsemble: P chains x S segments resident in HBM -- what a many-chain sampler uses instead of P*S
#      SamplePath pairs: sample!, solve!, llikelihood, pcn!, theta_* (per-chain parameters), Communicator (multi-GPU).
module BridgeB200

using Bridge, StaticArrays, LinearAlgebra
import Bridge: sample!, solve!, llikelihood, bridge!, innovations!, lptilde, EulerMaruyama, Euler, LeftRule, SamplePath,
               Wiener, ContinuousTimeProcess, GuidedBridge, PartialBridge, PartialBridgeνH

const lib = get(ENV, "BRIDGE_B200_LIB", joinpath(@__DIR__, "..", "bridge.jl_b200", "lib", "libbridge_b200.so"))

# ---- status codes -> the reference's own errors (include/bridge_b200.h)
function check(st::Cint)
    st == 0 && return nothing
    msg = unsafe_string(ccall((:bb_strerror, lib), Cstring, (Cint,), st))
    st == -4 && throw(DimensionMismatch("length(tt) != size(yy, 2)"))        # src/types.jl:127
    st == -5 && throw(AssertionError("m == length(v)"))                      # src/partialbridgenuH.jl:3
    if st == -8 || st == -9
        msg *= ": " * unsafe_string(ccall((:bb_last_cuda_error, lib), Cstring, ()))
    elseif st == -14
        msg *= ": " * unsafe_string(ccall((:bb_comm_last_error, lib), Cstring, ()))
    end
    error(msg)   # "Y and W differ in length." / "Time axis mismatch ..." / "Starting point has wrong length."
end

const ENABLED = Ref(true)
"""`enable!(false)`: every overloaded Bridge call falls back to Bridge's own CPU method."""
enable!(on::Bool = true) = (ENABLED[] = on)

# ---- registry models (bb_model): the device cannot call Julia closures
struct BBModel
    id::Int32; d::Int32; dprime::Int32; reserved::Int32
    par::NTuple{32,Float64}
end
pad32(v) = ntuple(i -> i <= length(v) ? Float64(v[i]) : 0.0, 32)
rowmajor(A::AbstractMatrix) = collect(vec(permutedims(A)))
bbmodel(::Wiener{Float64}) = BBModel(0, 1, 1, 0, pad32(()))
bbmodel(::Wiener{SVector{d,Float64}}) where {d} = BBModel(0, d, d, 0, pad32(()))
bbmodel(P::Bridge.LinPro) = (d = size(P.B, 1); BBModel(2, d, d, 0, pad32(vcat(rowmajor(P.B), P.μ, rowmajor(P.σ)))))
bbmodel(P::Bridge.Models.FitzHughNagumo) = BBModel(3, 2, 2, 0, pad32((P.ϵ, P.s, P.γ, P.β, P.σ1, P.σ2)))        # src/Models.jl:9-20
bbmodel(P::Bridge.Models.Lorenz) = BBModel(7, 3, 3, 0, pad32(vcat(collect(P.θ), diag(Matrix(P.σ)))))           # src/Models.jl:38-55
# user structs opt in by defining bbmodel(P), e.g. for project_partialbridge/partialbridge_fitzhugh.jl:36-46
#   BridgeB200.bbmodel(P::FitzhughDiffusion) = BridgeB200.BBModel(4, 2, 1, 0, BridgeB200.pad32((P.ϵ, P.s, P.γ, P.β, P.σ)))
# ids: include/bridge_b200.h bb_model_id (OU 1, INTDIFF 5, NCLAR3 6, LANDMARKS 8, BOLUS 9)
ondevice(P) = ENABLED[] && hasmethod(bbmodel, Tuple{typeof(P)})

# ---- context: one per process by default (device = LOCAL_RANK), no extra argument anywhere
mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = parse(Int, get(ENV, "LOCAL_RANK", "0")))
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:bb_ctx_create, lib), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r))
        new(r[])   # never destroyed before its ensembles/guides: bb_ctx_destroy refuses while any is alive
    end
end
const DEFAULT_CTX = Ref{Union{Nothing,Context}}(nothing)
default_context() = (DEFAULT_CTX[] === nothing && (DEFAULT_CTX[] = Context()); DEFAULT_CTX[]::Context)
synchronize(ctx::Context = default_context()) = check(ccall((:bb_ctx_synchronize, lib), Cint, (Ptr{Cvoid},), ctx.h))
"""0: reference arithmetic in the shared-table constructors (default), 1: fused multiply-adds"""
set_arith!(a::Integer, ctx::Context = default_context()) = check(ccall((:bb_ctx_set_arith, lib), Cint, (Ptr{Cvoid}, Cint), ctx.h, a))

# a model that is not in the registry: CUDA C source compiled at run time into the same kernels (bb_user_model_create)
mutable struct UserModel
    h::Ptr{Cvoid}; handle::Int32; d::Int; dprime::Int
end
"""`UserModel(d, d′, drift, col, sigma)`: `drift` = C statements assigning o[0..d-1] from x[] and par[]; col[i] = column of W
entering component i (0-based, -1: none); sigma[i] = C expression in par[] for that entry of σ.  Then
`BridgeB200.bbmodel(P::MyProcess) = BridgeB200.BBModel(10, d, d′, um.handle, BridgeB200.pad32(params(P)))`."""
function UserModel(d::Integer, dprime::Integer, drift::String, col::Vector{Int32}, sigma::Vector{String}; ctx::Context = default_context())
    r = Ref{Ptr{Cvoid}}(C_NULL)
    sp = [col[i] >= 0 ? Base.unsafe_convert(Cstring, sigma[i]) : Cstring(C_NULL) for i in 1:d]
    st = GC.@preserve sigma ccall((:bb_user_model_create, lib), Cint, (Ptr{Cvoid}, Int32, Int32, Cstring, Ptr{Int32}, Ptr{Cstring}, Ref{Ptr{Cvoid}}),
                                  ctx.h, d, dprime, drift, col, sp, r)
    if st != 0
        log = r[] == C_NULL ? "" : unsafe_string(ccall((:bb_user_model_log, lib), Cstring, (Ptr{Cvoid},), r[]))
        r[] == C_NULL || ccall((:bb_user_model_destroy, lib), Cint, (Ptr{Cvoid},), r[])
        error(unsafe_string(ccall((:bb_strerror, lib), Cstring, (Cint,), st)) * "\n" * log)
    end
    um = UserModel(r[], ccall((:bb_user_model_handle, lib), Int32, (Ptr{Cvoid},), r[]), d, dprime)
    finalizer(u -> ccall((:bb_user_model_destroy, lib), Cint, (Ptr{Cvoid},), u.h), um); um
end

mutable struct PathEnsemble
    h::Ptr{Cvoid}; ctx::Context
    P::Int; S::Int; N::Int; d::Int; dprime::Int
    function PathEnsemble(P, S, N, d, dprime; ctx::Context = default_context(), double_buffer = true, store_x = true, chain_offset = 0)
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
setstart!(E::PathEnsemble, u::Number) = setstart!(E, SVector{1,Float64}(u))
setstart!(E::PathEnsemble, u::AbstractVector) = (v = collect(Float64, u);
    check(ccall((:bb_ens_set_start, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32), E.h, v, length(v), 1)))
# per-chain starting points: X0 is d x P (one column per chain), the ABI wants [P][d] -- the same bytes
setstart!(E::PathEnsemble, X0::Matrix{Float64}) =
    check(ccall((:bb_ens_set_start, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32), E.h, X0, length(X0), 0))
upload!(E::PathEnsemble, what::Integer, A::Array{Float64}; which = 0, p0 = 0, np = E.P) =   # A: [k, N, S, np] == [np][S][N][k]
    check(ccall((:bb_ens_upload, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Float64}), E.h, what, which, p0, np, A))
download!(A::Array{Float64}, E::PathEnsemble, what::Integer; which = 0, p0 = 0, np = E.P) =
    (check(ccall((:bb_ens_download, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Float64}), E.h, what, which, p0, np, A)); A)
function getf64(E::PathEnsemble, field::Integer, width = 1)     # BB_F_LL 0, LL_PROP 1, LOGU 2, XEND 3, XEND_PROP 4
    out = Array{Float64}(undef, width, E.P)
    check(ccall((:bb_ens_get_f64, lib), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Float64}), E.h, field, 0, E.P, out)); out
end
function accepted(E::PathEnsemble)
    a = Vector{UInt8}(undef, E.P)
    check(ccall((:bb_ens_get_accepted, lib), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{UInt8}), E.h, 0, E.P, a)); a
end
acc(E::PathEnsemble) = (r = Ref{Int64}(0); check(ccall((:bb_ens_get_acc, lib), Cint, (Ptr{Cvoid}, Ref{Int64}), E.h, r)); r[])
reset_acc!(E::PathEnsemble) = check(ccall((:bb_ens_reset_acc, lib), Cint, (Ptr{Cvoid},), E.h))
"""`set_ll!(E, ll)`: log-likelihoods of the chains' current paths supplied by the caller (e.g. after `upload!` of W and X)."""
set_ll!(E::PathEnsemble, ll::Vector{Float64}; p0 = 0) =
    check(ccall((:bb_ens_set_ll, lib), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}), E.h, p0, length(ll), ll))
device_bytes(E::PathEnsemble) = ccall((:bb_ens_bytes, lib), Int64, (Ptr{Cvoid},), E.h)
function grid(E::PathEnsemble, seg::Integer)   # the time grid of segment `seg` (1-based) as set by set_grid!
    tt = Vector{Float64}(undef, E.N)
    check(ccall((:bb_ens_get_grid, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32), E.h, seg - 1, tt, E.N)); tt
end
abi_version() = Int(ccall((:bb_abi_version, lib), Cint, ()))
launch_count(ctx::Context = default_context()) = ccall((:bb_ctx_launch_count, lib), Int64, (Ptr{Cvoid},), ctx.h)

# ---- auxiliary process as values (bb_aux): constants, or values at the Ralston stage times of every interval
struct BBAux
    d::Int32; is_const::Int32
    B::Ptr{Float64}; beta::Ptr{Float64}; a::Ptr{Float64}; a_left::Ptr{Float64}
end
struct AuxValues   # keeps the arrays alive
    B::Vector{Float64}; β::Vector{Float64}; a::Vector{Float64}; al::Vector{Float64}; isconst::Bool; d::Int
end
isconstaux(Pt) = Pt isa Bridge.LinPro   # every other auxiliary process is sampled on the grid
function AuxValues(Pt, tt)
    d = size(Bridge.B(tt[1], Pt), 1)
    if isconstaux(Pt)
        return AuxValues(rowmajor(Bridge.B(tt[1], Pt)), collect(Float64, Bridge.β(tt[1], Pt)), rowmajor(Matrix(Bridge.a(tt[1], Pt))), Float64[], true, d)
    end
    Bs = Float64[]; βs = Float64[]; as = Float64[]; al = Float64[]
    for i in 1:length(tt)-1                     # backward step over [tt[i], tt[i+1]]: t = tt[i+1], h = tt[i]-tt[i+1]  (src/ode.jl:44-49,92-95)
        t, h = tt[i+1], tt[i] - tt[i+1]
        for c in (0.0, 1/2, 3/4)
            s = t + c*h
            append!(Bs, rowmajor(Bridge.B(s, Pt))); append!(βs, collect(Float64, Bridge.β(s, Pt))); append!(as, rowmajor(Matrix(Bridge.a(s, Pt))))
        end
        append!(al, rowmajor(Matrix(Bridge.a(tt[i], Pt))))
    end
    AuxValues(Bs, βs, as, al, false, d)
end
bbaux(A::AuxValues) = BBAux(A.d, A.isconst ? 1 : 0, pointer(A.B), pointer(A.β), pointer(A.a), A.isconst ? Ptr{Float64}(C_NULL) : pointer(A.al))

# ---- guiding tables of one segment on the device (bb_guide), built from the reference's own proposal structs
mutable struct Guide
    h::Ptr{Cvoid}
end
function newguide(ctx, kind, N, d, m, tt, A, b, Mm, v, Bt, βt, auxconst)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve tt A b Mm v Bt βt check(ccall((:bb_guide_create, lib), Cint,
        (Ptr{Cvoid}, Int32, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Int32, Ref{Ptr{Cvoid}}),
        ctx.h, kind, N, d, m, tt, A, b, Mm === nothing ? C_NULL : Mm, v === nothing ? C_NULL : v, Bt, βt, auxconst ? 1 : 0, r))
    g = Guide(r[]); finalizer(g -> ccall((:bb_guide_destroy, lib), Cint, (Ptr{Cvoid},), g.h), g); g
end
function auxongrid(Pt, tt)   # B̃(tt[i]), β̃(tt[i]) for llikelihood (b̃ = B̃x + β̃, src/partialbridgenuH.jl:176)
    isconstaux(Pt) && return rowmajor(Bridge.B(tt[1], Pt)), collect(Float64, Bridge.β(tt[1], Pt)), true
    reduce(vcat, [rowmajor(Bridge.B(t, Pt)) for t in tt]), reduce(vcat, [collect(Float64, Bridge.β(t, Pt)) for t in tt]), false
end
flat(v::Vector{<:SVector}) = collect(reinterpret(Float64, v))
flat(v::Vector{<:SMatrix}) = reduce(vcat, rowmajor.(v))
flat(v::Vector{Float64}) = v
function Guide(Po::PartialBridgeνH; ctx::Context = default_context())   # fields Target, Pt, tt, ν, H, C  (src/partialbridgenuH.jl:122-130)
    Bt, βt, c = auxongrid(Po.Pt, Po.tt)
    newguide(ctx, 1, length(Po.tt), length(Po.ν[1]), 0, Po.tt, flat(Po.H), flat(Po.ν), nothing, nothing, Bt, βt, c)
end
function Guide(Po::GuidedBridge; ctx::Context = default_context())      # fields Target, Pt, tt, H♢, V  (src/guip.jl:165-170)
    Bt, βt, c = auxongrid(Po.Pt, Po.tt)
    newguide(ctx, 2, length(Po.tt), length(Po.V[1]), 0, Po.tt, flat(Po.H♢), flat(Po.V), nothing, nothing, Bt, βt, c)
end
function Guide(Po::PartialBridge; ctx::Context = default_context())     # fields Target, Pt, tt, v, L, M, μ  (src/partialbridge.jl:33-41)
    Bt, βt, c = auxongrid(Po.Pt, Po.tt)
    m, d = size(Po.L[1])
    newguide(ctx, 3, length(Po.tt), d, m, Po.tt, flat(Po.L), flat(Po.μ), flat(Po.M), collect(Float64, Po.v), Bt, βt, c)
end
const Proposal = Union{GuidedBridge,PartialBridge,PartialBridgeνH}
const GUIDES = IdDict{Any,Guide}()    # one device table per proposal object (uploaded on first use)
guide_for(Po::Proposal) = get!(() -> Guide(Po), GUIDES, Po)
forget!(Po::Proposal) = delete!(GUIDES, Po)

# ---- backward ODEs on the device (the constructors' work; same R3 / Lyapunov schemes, reference arithmetic)
"""`partialbridgeνH(tt, Pt, νend, Hend⁺)` on the device -> (ν [d, N], H [d, d, N] row-major per matrix, ν(tt[1]), H⁺(tt[1]), C)
(src/partialbridgenuH.jl:86-103,148-155; `method` 0 = R3, 1 = Lyap)."""
function backward_nuH(tt::Vector{Float64}, Pt, νend, Hendp; method = 1, C0 = 0.0, ctx::Context = default_context())
    A = AuxValues(Pt, tt); N, d = length(tt), A.d
    ν = Matrix{Float64}(undef, d, N); H = Array{Float64}(undef, d, d, N)
    νl = Vector{Float64}(undef, d); Hl = Matrix{Float64}(undef, d, d); C = Ref{Float64}(0.0)
    νe = collect(Float64, νend); He = rowmajor(Matrix(Hendp)); aux = Ref(bbaux(A))
    GC.@preserve A νe He check(ccall((:bb_backward_nuH, lib), Cint,
        (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Float64}, Ref{BBAux}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}),
        ctx.h, method, N, d, tt, aux, νe, He, C0, ν, H, νl, Hl, C))
    ν, H, νl, permutedims(Hl), C[]
end
"""Observation update of (ν, H⁺) between segments (partialbridge_bolus3.jl:128-137), in place."""
function gpupdate_nuH!(ν::Vector{Float64}, Hp::Matrix{Float64}, L, Σ, v; ctx::Context = default_context())
    m, d = size(L); Hr = rowmajor(Hp); Lr = rowmajor(Matrix(L)); Sr = rowmajor(Matrix(Σ)); vv = collect(Float64, v)
    check(ccall((:bb_gpupdate_nuH, lib), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                ctx.h, d, m, ν, Hr, Lr, Sr, vv))
    Hp .= permutedims(reshape(Hr, d, d)); ν, Hp
end

# ---- (2) ensemble methods added to Bridge's generic functions
function sample!(E::PathEnsemble, ::Wiener; seed::UInt64 = UInt64(0), stream::UInt32 = UInt32(0))
    check(ccall((:bb_wiener_sample, lib), Cint, (Ptr{Cvoid}, UInt64, UInt32), E.h, seed, stream)); E
end
function solve!(::EulerMaruyama, E::PathEnsemble, u, P::ContinuousTimeProcess)
    setstart!(E, u); m = Ref(bbmodel(P))
    check(ccall((:bb_euler, lib), Cint, (Ptr{Cvoid}, Ref{BBModel}), E.h, m)); E
end
function solve!(::EulerMaruyama, E::PathEnsemble, u, P::ContinuousTimeProcess, guides::Vector{Guide}; skip = 0, store_x = true, ll = true)
    setstart!(E, u); m = Ref(bbmodel(P)); hs = [g.h for g in guides]
    GC.@preserve guides check(ccall((:bb_guided_euler_ll, lib), Cint, (Ptr{Cvoid}, Ref{BBModel}, Ptr{Ptr{Cvoid}}, Int32, UInt32),
                                    E.h, m, hs, skip, (store_x ? 1 : 0) | (ll ? 0 : 2)))
    E   # end points: getf64(E, 3, E.d)
end
function llikelihood(::LeftRule, E::PathEnsemble, P::ContinuousTimeProcess, guides::Vector{Guide}; skip = 0)
    m = Ref(bbmodel(P)); hs = [g.h for g in guides]
    GC.@preserve guides check(ccall((:bb_llikelihood, lib), Cint, (Ptr{Cvoid}, Ref{BBModel}, Ptr{Ptr{Cvoid}}, Int32), E.h, m, hs, skip))
    vec(getf64(E, 0))
end
# solve! of an ensemble with the reference's other one-step schemes (src/euler.jl:68-88,178-198,330-356)
scheme_id(::EulerMaruyama) = 0
scheme_id(::Bridge.StratonovichEuler) = 1
scheme_id(::Bridge.StochasticHeun) = 2
scheme_id(::Bridge.StochasticRungeKutta) = 3
function solve!(s::Union{Bridge.StratonovichEuler,Bridge.StochasticHeun,Bridge.StochasticRungeKutta}, E::PathEnsemble, u,
                P::ContinuousTimeProcess)
    setstart!(E, u); m = Ref(bbmodel(P))
    check(ccall((:bb_solve_scheme, lib), Cint, (Ptr{Cvoid}, Ref{BBModel}, Int32), E.h, m, scheme_id(s))); E
end
"""sample!(W, Wiener()) fused with solve!(EulerMaruyama(), X, u, W, P): W and X are both written, W is never read."""
function sample_solve!(E::PathEnsemble, u, P::ContinuousTimeProcess; seed::UInt64 = UInt64(0), stream::UInt32 = UInt32(0))
    setstart!(E, u); m = Ref(bbmodel(P))
    check(ccall((:bb_sample_euler, lib), Cint, (Ptr{Cvoid}, Ref{BBModel}, UInt64, UInt32), E.h, m, seed, stream)); E
end
# kernel time of the last call on the context (CUDA events), for benchmarking a script
set_timing!(on::Bool, ctx::Context = default_context()) = check(ccall((:bb_ctx_set_timing, lib), Cint, (Ptr{Cvoid}, Cint), ctx.h, on ? 1 : 0))
last_kernel_ms(ctx::Context = default_context()) = ccall((:bb_ctx_last_kernel_ms, lib), Float64, (Ptr{Cvoid},), ctx.h)

"""One pCN / MH update of every chain: the body of `for iter in 1:iterations` in test/partialbridgenuH.jl:176-191."""
function pcn!(E::PathEnsemble, P::ContinuousTimeProcess, guides::Vector{Guide}, ρ, seed::UInt64, iter::Integer; skip = 0, store_x = true)
    m = Ref(bbmodel(P)); hs = [g.h for g in guides]
    GC.@preserve guides check(ccall((:bb_pcn_step, lib), Cint,
        (Ptr{Cvoid}, Ref{BBModel}, Ptr{Ptr{Cvoid}}, Float64, UInt64, UInt32, Int32, UInt32), E.h, m, hs, ρ, seed, iter, skip, store_x ? 1 : 0))
    acc(E)
end
"""Current paths of all chains as an array [d, N, S, P] (refreshes the chains whose last proposal was rejected)."""
function download_x(E::PathEnsemble, P::ContinuousTimeProcess, guides::Vector{Guide})
    m = Ref(bbmodel(P)); hs = [g.h for g in guides]
    GC.@preserve guides check(ccall((:bb_ens_refresh_x, lib), Cint, (Ptr{Cvoid}, Ref{BBModel}, Ptr{Ptr{Cvoid}}), E.h, m, hs))
    download!(Array{Float64}(undef, E.d, E.N, E.S, E.P), E, 1)   # column-major == the ABI's [P][S][N][d]
end

"""The same iteration with W, X kept in HOST arrays as the reference loop does (test/partialbridgenuH.jl:168-191):
`W` [d′, N, S, P] goes up; `Wo`, `Xo` ([d, N, S, P]), `llo`, `accepted` come back.  `skip_rejected = true`: only the chains that
accept write their rows; with page-locked arrays (`CUDA.Mem.pin` / `cudaHostRegister`) `Wo === W` is allowed and `W` is then
updated in place -- the loop's `W, Wo = Wo, W`."""
function pcn_host!(E::PathEnsemble, P::ContinuousTimeProcess, guides::Vector{Guide}, ρ, seed::UInt64, iter::Integer,
                   W::Array{Float64}, Wo::Array{Float64}, Xo::Array{Float64}, llo::Vector{Float64}, accepted::Vector{UInt8};
                   skip = 0, skip_rejected = false)
    m = Ref(bbmodel(P)); hs = [g.h for g in guides]
    GC.@preserve guides check(ccall((:bb_pcn_step_host, lib), Cint,
        (Ptr{Cvoid}, Ref{BBModel}, Ptr{Ptr{Cvoid}}, Float64, UInt64, UInt32, Int32, UInt32, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}),
        E.h, m, hs, ρ, seed, iter, skip, 1 | (skip_rejected ? 4 : 0), W, Wo, Xo, llo, accepted))
    acc(E)
end
# which kernel runs pcn!: 0 automatic, 1 one thread per chain, 2 / 3 warp-specialised (one / two chains per dynamics thread)
set_pcn_kernel!(mode::Integer, ctx::Context = default_context()) =
    check(ccall((:bb_ctx_set_pcn_kernel, lib), Cint, (Ptr{Cvoid}, Cint), ctx.h, mode))

"""`solve!(Mdb(), Y, u, W, P°)` (src/euler.jl:308-327) for every chain of an ensemble."""
function solve!(::Bridge.Mdb, E::PathEnsemble, u, P::ContinuousTimeProcess, guides::Vector{Guide})
    setstart!(E, u); m = Ref(bbmodel(P)); hs = [g.h for g in guides]
    GC.@preserve guides check(ccall((:bb_guided_mdb, lib), Cint, (Ptr{Cvoid}, Ref{BBModel}, Ptr{Ptr{Cvoid}}), E.h, m, hs)); E
end

"""The backward pass of a chain of S `PartialBridgeνH` segments in ONE launch, tables written in place on the device:
the loop of partialbridge_bolus3.jl:162-180 (νend = 0, Hend⁺ = I/ϵ, gpupdate with V.yy[end]; for i = S:-1:1
`partialbridgeνH(tt_i, P, Pt_i, νend, Hend⁺)` and, for i > 1, `gpupdate(νend, Hend⁺, Σ, L, v_{i-1})`).
`tts[i]`: grid of segment i; `Pts[i]`: its (constant) auxiliary process; `vs[i]`: observation at its right end.
`guides` = the vector returned by an earlier call, to update the same device tables.  Returns (guides, ν(0), H⁺(0), C)."""
function guides_chain_nuH(tts::Vector{Vector{Float64}}, Pts, L, Σ, vs, ϵ; guides::Vector{Guide} = Guide[], method = 1,   # 1 = BB_ODE_LYAP (as the script), 0 = BB_ODE_R3
                          ctx::Context = default_context())
    S = length(tts); N = length(tts[1]); m, d = size(L)
    auxv = [AuxValues(Pts[i], tts[i]) for i in 1:S]; auxs = [bbaux(A) for A in auxv]
    tt = reduce(vcat, tts); v = reduce(vcat, [collect(Float64, vs[i]) for i in 1:S])
    hs = isempty(guides) ? fill(Ptr{Cvoid}(C_NULL), S) : [g.h for g in guides]
    ν0 = zeros(d); H0 = zeros(d * d); C = Ref(0.0)
    GC.@preserve auxv guides check(ccall((:bb_guides_chain_nuH, lib), Cint,
        (Ptr{Cvoid}, Int32, Int32, Int32, Int32, Int32, Ptr{Float64}, Ptr{BBAux}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Float64, Ptr{Ptr{Cvoid}}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}),
        ctx.h, method, S, N, d, m, tt, auxs, rowmajor(Matrix{Float64}(L)), rowmajor(Matrix{Float64}(Σ)), v, ϵ, hs, ν0, H0, C))
    if isempty(guides)
        guides = [Guide(h) for h in hs]
        foreach(g -> finalizer(x -> ccall((:bb_guide_destroy, lib), Cint, (Ptr{Cvoid},), x.h), g), guides)
    end
    guides, ν0, permutedims(reshape(H0, d, d)), C[]
end

# ---- mcstart / mcnext! / mcstats of src/mclog.jl:22-93 for an ensemble: moments of the current paths per grid point, pooled
# over the chains (the reference's mcnext is per chain over iterations) and over calls
mc_reset!(E::PathEnsemble) = check(ccall((:bb_ens_mc_reset, lib), Cint, (Ptr{Cvoid},), E.h))
mc_update!(E::PathEnsemble) = check(ccall((:bb_ens_mc_update, lib), Cint, (Ptr{Cvoid},), E.h))   # after download_x / a refresh
function mc_stats(E::PathEnsemble)
    mean = Array{Float64}(undef, E.d, E.N, E.S); cov = Array{Float64}(undef, E.d, E.d, E.N, E.S); n = Ref{Int64}(0)
    check(ccall((:bb_ens_mc_stats, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{Int64}), E.h, mean, cov, n))
    mean, cov, n[]
end
# The reference's own semantics: one (m, m2, k) per chain over the recorded iterations, `mcnext!(mcstate[i], XX[i].yy)` of
# partialbridge_fitzhugh.jl:169-189 for every chain at once (Welford in the reference's operation order, bit-identical).
# Arrays are column-major views of the C layout: mean[a, j, s, p], cov[b, a, j, s, p] = m2[a, b]/(k - 1) of chain p.
chain_mc_reset!(E::PathEnsemble) = check(ccall((:bb_ens_chain_mc_reset, lib), Cint, (Ptr{Cvoid},), E.h))          # mcstart
chain_mc_update!(E::PathEnsemble) = check(ccall((:bb_ens_chain_mc_update, lib), Cint, (Ptr{Cvoid},), E.h))        # mcnext!
function chain_mc_stats(E::PathEnsemble, chains::UnitRange{Int} = 1:E.P)                                           # mcstats
    np = length(chains)
    mean = Array{Float64}(undef, E.d, E.N, E.S, np); cov = Array{Float64}(undef, E.d, E.d, E.N, E.S, np); k = Ref{Int64}(0)
    check(ccall((:bb_ens_chain_mc_stats, lib), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Int64}),
                E.h, first(chains) - 1, np, mean, cov, k))
    mean, cov, k[]
end
function chain_mc_band(E::PathEnsemble, chains::UnitRange{Int} = 1:E.P)                                            # mcband
    np = length(chains)
    lower = Array{Float64}(undef, E.d, E.N, E.S, np); upper = similar(lower)
    check(ccall((:bb_ens_chain_mc_band, lib), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}),
                E.h, first(chains) - 1, np, lower, upper))
    lower, upper
end

# ---- multi-GPU: one Julia process (or task) per GPU; the acceptance counter is the only exchange (SURVEY 8e)
mutable struct Communicator
    h::Ptr{Cvoid}
end
function unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:bb_comm_unique_id, lib), Cint, (Ptr{UInt8},), id)); id
end
"""All ranks call this with the 128 bytes rank 0 got from `unique_id()` (ship them with MPI, a file, Distributed, ...)."""
function Communicator(nranks::Integer, rank::Integer, id::Vector{UInt8}; ctx::Context = default_context())
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:bb_comm_create, lib), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}, Ref{Ptr{Cvoid}}), ctx.h, nranks, rank, id, r))
    c = Communicator(r[]); finalizer(c -> ccall((:bb_comm_destroy, lib), Cint, (Ptr{Cvoid},), c.h), c); c
end
"""Start the all-reduce of the acceptance counter (asynchronous, NCCL on its own stream); `acc(comm)` fetches the sum."""
allreduce_acc!(E::PathEnsemble, c::Communicator) = check(ccall((:bb_allreduce_acc, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), E.h, c.h))
acc(c::Communicator) = (r = Ref{Int64}(0); check(ccall((:bb_comm_get_acc, lib), Cint, (Ptr{Cvoid}, Ref{Int64}), c.h, r)); r[])
"""The same for the acceptance counter of the parameter updates (`theta_param!`)."""
allreduce_theta_acc!(E::PathEnsemble, c::Communicator) = check(ccall((:bb_allreduce_theta_acc, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), E.h, c.h))
synchronize(c::Communicator) = check(ccall((:bb_comm_synchronize, lib), Cint, (Ptr{Cvoid},), c.h))   # all issued all-reduces done
Base.size(c::Communicator) = Int(ccall((:bb_comm_size, lib), Cint, (Ptr{Cvoid},), c.h))
rank(c::Communicator) = Int(ccall((:bb_comm_rank, lib), Cint, (Ptr{Cvoid},), c.h))

# ---- (1) the reference's own signatures on SamplePath: a cached one-chain ensemble per (N, d, d')
const SMALL = Dict{NTuple{3,Int},PathEnsemble}()
small(N, d, dp) = get!(() -> PathEnsemble(1, 1, N, d, dp; double_buffer = false), SMALL, (N, d, dp))
dimof(::Type{Float64}) = 1
dimof(::Type{SVector{d,Float64}}) where {d} = d
asf64(yy::Vector{Float64}) = yy
asf64(yy::Vector{SVector{d,Float64}}) where {d} = reinterpret(Float64, yy)   # same bytes as the ABI's [N][d]
const RNG = Ref((UInt64(0), UInt32(0)))
"""`seed!(s)`: seeds the Philox streams `sample!` draws from (the device cannot reproduce Julia's own RNG streams)."""
seed!(s::Integer) = (RNG[] = (UInt64(s), UInt32(0)))

function sample!(W::SamplePath{T}, P::Wiener{T}, y1 = W.yy[1]) where {T<:Union{Float64,SVector}}
    ENABLED[] || return invoke(sample!, Tuple{SamplePath{T},Wiener{T},Any}, W, P, y1)
    dp = dimof(T); E = small(length(W.tt), dp, dp)
    W.yy[1] = y1
    setgrid!(E, 1, W.tt)
    w = asf64(W.yy)
    GC.@preserve W begin
        check(ccall((:bb_ens_upload, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Float64}), E.h, 0, 0, 0, 1, w))
        seed, stream = RNG[]; RNG[] = (seed, stream + UInt32(1))
        check(ccall((:bb_wiener_sample, lib), Cint, (Ptr{Cvoid}, UInt64, UInt32), E.h, seed, stream))
        check(ccall((:bb_ens_download, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Float64}), E.h, 0, 0, 0, 1, w))
    end
    W
end

function solve!(s::EulerMaruyama, Y::SamplePath{T}, u::T, W::SamplePath, P::ContinuousTimeProcess{T}) where {T}
    ondevice(P) || return invoke(solve!, Tuple{EulerMaruyama,Any,T,SamplePath,Bridge.ProcessOrCoefficients}, s, Y, u, W, P)
    N = length(W); N != length(Y) && error("Y and W differ in length.")
    m = bbmodel(P); E = small(N, Int(m.d), Int(m.dprime))
    setgrid!(E, 1, W.tt); setstart!(E, u)
    w = asf64(W.yy); x = asf64(Y.yy); mr = Ref(m)
    GC.@preserve W Y begin
        check(ccall((:bb_ens_upload, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Float64}), E.h, 0, 0, 0, 1, w))
        check(ccall((:bb_euler, lib), Cint, (Ptr{Cvoid}, Ref{BBModel}), E.h, mr))
        check(ccall((:bb_ens_download, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Float64}), E.h, 1, 0, 0, 1, x))
    end
    Y.tt[:] = W.tt
    Y
end

"""solve!(EulerMaruyama!(), Y::VSamplePath, u, W, P)   src/sde!.jl:21-53: d x N matrices (column = time) have the byte
layout of the ABI's [N][d]; same recurrence, same kernel"""
function solve!(s::Bridge.EulerMaruyama!, Y::Bridge.VSamplePath, u, W::Bridge.VSamplePath, P::ContinuousTimeProcess)
    ondevice(P) || return invoke(solve!, Tuple{Bridge.EulerMaruyama!,Bridge.VSamplePath,Any,Any,Bridge.ProcessOrCoefficients}, s, Y, u, W, P)
    N = length(W); N != length(Y) && error("Y and W differ in length.")
    size(Y.yy) != (length(u), N) && error("Starting point has wrong length.")          # src/sde!.jl:30
    m = bbmodel(P); E = small(N, Int(m.d), Int(m.dprime))
    setgrid!(E, 1, W.tt); setstart!(E, u)
    mr = Ref(m)
    GC.@preserve W Y begin
        check(ccall((:bb_ens_upload, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Float64}), E.h, 0, 0, 0, 1, W.yy))
        check(ccall((:bb_euler, lib), Cint, (Ptr{Cvoid}, Ref{BBModel}), E.h, mr))
        check(ccall((:bb_ens_download, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Float64}), E.h, 1, 0, 0, 1, Y.yy))
    end
    Y.tt[:] = W.tt
    Y
end

"""solve!(Euler(), Y, u, W, P°) -> Y.yy[end]   (src/euler.jl:247-268; tt[:] = P.tt, end point rule of GuidedBridge)"""
function solve!(s::EulerMaruyama, Y::SamplePath, u, W::SamplePath, Po::Proposal)
    ondevice(Po.Target) || return invoke(solve!, Tuple{EulerMaruyama,Any,Any,SamplePath,Proposal}, s, Y, u, W, Po)
    W.tt === Po.tt && error("Time axis mismatch between bridge P and driving W.")          # src/euler.jl:248
    N = length(W); (N != length(Y) || N != length(Po.tt)) && error("Y and W differ in length.")   # :251
    m = bbmodel(Po.Target); E = small(N, Int(m.d), Int(m.dprime)); g = guide_for(Po)
    setstart!(E, u)
    w = asf64(W.yy); x = asf64(Y.yy); mr = Ref(m); hs = [g.h]
    GC.@preserve W Y g begin
        check(ccall((:bb_ens_upload, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Float64}), E.h, 0, 0, 0, 1, w))
        check(ccall((:bb_guided_euler_ll, lib), Cint, (Ptr{Cvoid}, Ref{BBModel}, Ptr{Ptr{Cvoid}}, Int32, UInt32), E.h, mr, hs, 0, 1 | 2))
        check(ccall((:bb_ens_download, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Float64}), E.h, 1, 0, 0, 1, x))
    end
    Y.tt[:] = Po.tt
    Y.yy[end]
end

"""llikelihood(LeftRule(), X, P°; skip = 0) -> Float64"""
function llikelihood(r::LeftRule, X::SamplePath, Po::Proposal; skip = 0)
    ondevice(Po.Target) || return invoke(llikelihood, Tuple{LeftRule,SamplePath,typeof(Po)}, r, X, Po; skip = skip)
    N = length(X); N != length(Po.tt) && error("Y and W differ in length.")
    m = bbmodel(Po.Target); E = small(N, Int(m.d), Int(m.dprime)); g = guide_for(Po)
    x = asf64(X.yy); mr = Ref(m); hs = [g.h]; ll = Ref{Float64}(0.0)
    GC.@preserve X g begin
        check(ccall((:bb_ens_upload, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Float64}), E.h, 1, 0, 0, 1, x))
        check(ccall((:bb_llikelihood, lib), Cint, (Ptr{Cvoid}, Ref{BBModel}, Ptr{Ptr{Cvoid}}, Int32), E.h, mr, hs, skip))
        check(ccall((:bb_ens_get_f64, lib), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Ref{Float64}), E.h, 0, 0, 1, ll))
    end
    ll[]
end

# bridge!(Y, W, P°) (src/deprecated.jl:16-17) and the 4-argument form older scripts use (project/partialbridge.jl:63,65,87)
bridge!(Y::SamplePath, W::SamplePath, Po::Proposal) = (solve!(Euler(), Y, Y.yy[1], W, Po); Y)
bridge!(X::SamplePath, x0, W::SamplePath, Po::Proposal) = (solve!(Euler(), X, x0, W, Po); X)

"""innovations!(EulerMaruyama(), W, Y, P) -> W   (src/euler.jl:357-376; P a registry target or a proposal; d' = d)"""
function innovations!(s::EulerMaruyama, W::SamplePath, Y::SamplePath, P)
    target = P isa Proposal ? P.Target : P
    ondevice(target) || return invoke(innovations!, Tuple{EulerMaruyama,Any,Any,Any}, s, W, Y, P)
    N = length(Y); N != length(W) && error("Y and W differ in length.")
    m = bbmodel(target); E = small(N, Int(m.d), Int(m.dprime))
    x = asf64(Y.yy); w = asf64(W.yy); mr = Ref(m)
    hs = P isa Proposal ? [guide_for(P).h] : Ptr{Cvoid}[]
    P isa Proposal || setgrid!(E, 1, Y.tt)
    GC.@preserve W Y begin
        check(ccall((:bb_ens_upload, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Float64}), E.h, 1, 0, 0, 1, x))
        check(ccall((:bb_innovations, lib), Cint, (Ptr{Cvoid}, Ref{BBModel}, Ptr{Ptr{Cvoid}}), E.h, mr, isempty(hs) ? C_NULL : hs))
        check(ccall((:bb_ens_download, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Float64}), E.h, 0, 0, 0, 1, w))
    end
    W.tt[:] = Y.tt
    W
end

"""lptilde(x, P°::PartialBridgeνH) in the form the reference tests (test/partialbridgenuH.jl:124; the method in
src/partialbridgenuH.jl:169 has a typo and does not run)"""
function lptilde(x, Po::PartialBridgeνH; ctx::Context = default_context())
    d = length(Po.ν[1]); out = Ref{Float64}(0.0)
    ν0 = collect(Float64, Po.ν[1]); H0 = rowmajor(Matrix(Po.H[1])); xx = collect(Float64, x)
    check(ccall((:bb_lptilde_nuH, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Ref{Float64}),
                ctx.h, d, ν0, H0, Po.C, xx, out))
    out[]
end
"""lptilde(P°::GuidedBridge, u) = logpdfnormal(V[1] - u, H♢[1]) - traceB(tt, Pt)   (src/guip.jl:203-206)"""
function lptilde(Po::GuidedBridge, u; ctx::Context = default_context())
    tt = Po.tt; N = length(tt); d = length(Po.V[1]); out = Ref{Float64}(0.0)
    tr = Float64[]
    if isconstaux(Po.Pt)
        push!(tr, LinearAlgebra.tr(Bridge.B(tt[1], Po.Pt)))
    else
        for i in 1:N-1, c in (0.0, 1/2, 3/4)      # forward Ralston stages of src/ode.jl:44-49,178-184
            push!(tr, LinearAlgebra.tr(Bridge.B(tt[i] + c*(tt[i+1] - tt[i]), Po.Pt)))
        end
    end
    V0 = collect(Float64, Po.V[1]); K0 = rowmajor(Matrix(Po.H♢[1])); uu = collect(Float64, u)
    check(ccall((:bb_lptilde_HV, lib), Cint,
                (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}),
                ctx.h, N, d, tt, tr, isconstaux(Po.Pt) ? 1 : 0, V0, K0, uu, out))
    out[]
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
    a = Ref{Int64}(0); check(ccall((:bb_theta_get_acc, lib), Cint, (Ptr{Cvoid}, Ref{Int64}), E.h, a)); a[]
end
# blocked path update of segments klow .. kup-1 (1-based, as ind = (kup-1):-1:klow of partialbridge_bolus3.jl:258-275);
# Hzero⁺ = hzero*I (:234) conditions a block that does not end the chain on the chain's own path at its right end
theta_block!(E::PathEnsemble, klow::Integer, kup::Integer, ρ, seed::UInt64, iter::Integer; hzero = 0.1, skip = 0) =
    check(ccall((:bb_theta_block_step, lib), Cint, (Ptr{Cvoid}, Int32, Int32, Float64, Float64, UInt64, UInt32, Int32),
                E.h, klow - 1, kup - 1, ρ, hzero, seed, iter, skip))
# one sweep klow = 1 .. obsnum of the `while !finished` loop (:258-362, updateparams == false); returns the blocks
function theta_blocked_sweep!(E::PathEnsemble, ρ, seed::UInt64, iter0::Integer; hzero = 0.1, skip = 0)
    obsnum = E.S + 1; klow = 1; blocks = Tuple{Int,Int}[]
    while klow != obsnum
        kup = klow + rand(1:obsnum-klow)
        theta_block!(E, klow, kup, ρ, seed, iter0 + length(blocks); hzero = hzero, skip = skip)
        push!(blocks, (klow, kup)); klow = kup
    end
    blocks
end
function set_theta!(E::PathEnsemble, θ::Matrix{Float64})   # 8 x P, column per chain
    check(ccall((:bb_theta_set, lib), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}), E.h, 0, E.P, θ)); E
end
theta_refresh_x!(E::PathEnsemble) = check(ccall((:bb_theta_refresh_x, lib), Cint, (Ptr{Cvoid},), E.h))
function theta_block(E::PathEnsemble)   # numbers of the last block update: (5 + 2S) x P
    out = Matrix{Float64}(undef, 5 + 2 * E.S, E.P)
    check(ccall((:bb_theta_get_block, lib), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}), E.h, 0, E.P, out)); out
end
function theta(E::PathEnsemble)   # param(P) of every chain: 8 x P (column per chain)
    θ = Matrix{Float64}(undef, 8, E.P)
    check(ccall((:bb_theta_get, lib), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Float64}), E.h, 0, 0, E.P, θ)); θ
end

end # module
