# baseline/julia_cpu.jl -- times the TRUE reference (Bridge.jl, Julia >= 1.5) on the workload bench.py measures.
# STATUS: written, NOT run (no Julia runtime in the build environment or on the GPU box).  Uses only reference API:
# sample!, solve!(Euler(), ...), llikelihood(LeftRule(), ...), partialbridgeνH, and the loop of
# test/partialbridgenuH.jl:155-198 / project_partialbridge/partialbridge_bolus3.jl:162-192.
# usage: julia --project=/path/to/Bridge.jl baseline/julia_cpu.jl [chains] [iterations]
using Bridge, StaticArrays, LinearAlgebra, Random
const ℝ = SVector{N,Float64} where {N}

struct FitzhughDiffusion <: ContinuousTimeProcess{ℝ{2}}
    ϵ::Float64; s::Float64; γ::Float64; β::Float64; σ::Float64
end
Bridge.b(t, x, P::FitzhughDiffusion) = ℝ{2}((x[1] - x[2] - x[1]^3 + P.s) / P.ϵ, P.γ * x[1] - x[2] + P.β)
Bridge.σ(t, x, P::FitzhughDiffusion) = ℝ{2}(0.0, P.σ)
Bridge.constdiff(::FitzhughDiffusion) = true
struct FitzhughDiffusionAux <: ContinuousTimeProcess{ℝ{2}}
    ϵ::Float64; s::Float64; γ::Float64; β::Float64; σ::Float64; v::Float64
end
Bridge.B(t, P::FitzhughDiffusionAux) = @SMatrix [1/P.ϵ -1/P.ϵ; P.γ -1.0]
Bridge.β(t, P::FitzhughDiffusionAux) = ℝ{2}(P.s / P.ϵ - (P.v^3) / P.ϵ, P.β)
Bridge.σ(t, P::FitzhughDiffusionAux) = ℝ{2}(0.0, P.σ)
Bridge.constdiff(::FitzhughDiffusionAux) = true
Bridge.b(t, x, P::FitzhughDiffusionAux) = Bridge.B(t, P) * x + Bridge.β(t, P)
Bridge.a(t, P::FitzhughDiffusionAux) = Bridge.σ(t, P) * Bridge.σ(t, P)'

function gpupdate(ν, H⁺, Σ, L, v)                       # partialbridge_bolus3.jl:128-137
    Z = I - H⁺ * L' * inv(Σ + L * H⁺ * L') * L
    SVector(Z * H⁺ * L' * inv(Σ) * v + Z * ν), Z * H⁺
end

function main(chains = 64, iterations = 4)
    P = FitzhughDiffusion(0.1, 0.0, 1.5, 0.8, 0.3); x0 = ℝ{2}(-0.5, -0.6); ρ = 0.99
    L = @SMatrix [1.0 0.0]; Σ = @SMatrix [1e-10]; ϵ = 1e-3
    obs_t = (0.0, 0.5, 1.0, 1.5, 2.0); obs_v = (-1.0, -0.5, 0.5, 1.1); n = 1001
    τ(t, T0, T1) = T0 + (t - T0) * (2 - (t - T0) / (T1 - T0))
    grids = [τ.(range(obs_t[k], obs_t[k+1], length = n), obs_t[k], obs_t[k+1]) for k in 1:4]
    ν = ℝ{2}(0.0, 0.0); H⁺ = SMatrix{2,2}(I / ϵ)
    ν, H⁺ = gpupdate(ν, H⁺, Σ, L, ℝ{1}(obs_v[4]))
    Q = Vector{Any}(undef, 4)
    for i in 4:-1:1
        Q[i], ν, H⁺ = Bridge.partialbridgeνH(grids[i], P, FitzhughDiffusionAux(P.ϵ, P.s, P.γ, P.β, P.σ, obs_v[i]), ν, H⁺)
        i > 1 && ((ν, H⁺) = gpupdate(ν, H⁺, Σ, L, ℝ{1}(obs_v[i-1])))
    end
    WW = [[sample(grids[i], Wiener()) for i in 1:4] for c in 1:chains]
    XX = [[Bridge.samplepath(grids[i], zero(x0)) for i in 1:4] for c in 1:chains]
    WWo = deepcopy(WW); XXo = deepcopy(XX); W2 = deepcopy(WW[1]); ll = zeros(chains)
    for c in 1:chains
        xs = x0
        for i in 1:4
            xs = solve!(Euler(), XX[c][i], xs, WW[c][i], Q[i]); ll[c] += llikelihood(LeftRule(), XX[c][i], Q[i])
        end
    end
    acc = 0
    t = @elapsed for iter in 1:iterations, c in 1:chains
        xs = x0; llo = 0.0
        for i in 1:4
            sample!(W2[i], Wiener())
            WWo[c][i].yy .= ρ * WW[c][i].yy + sqrt(1 - ρ^2) * W2[i].yy
            xs = solve!(Euler(), XXo[c][i], xs, WWo[c][i], Q[i])
            llo += llikelihood(LeftRule(), XXo[c][i], Q[i])
        end
        if log(rand()) <= llo - ll[c]
            XX[c], XXo[c] = XXo[c], XX[c]; WW[c], WWo[c] = WWo[c], WW[c]; ll[c] = llo; acc += 1
        end
    end
    steps = chains * iterations * 4 * (n - 1)
    println("path-steps/s = ", steps / t, "  (1 thread; ", chains, " chains x ", iterations, " iterations; acc = ", acc, ")")
end
main(parse.(Int, ARGS)...)
