"""The BASELINE.json workloads, written against the public API of this package (no oracle here).

config 2  Wiener + EulerMaruyama ensemble, N = 1001 on [0, 1]                       (mode A / B)
config 3  LinPro d = 3 GuidedBridge, N = 1001 on [0, 1], v = (0.5, 0, -0.5)           (mode G)
config 4  FitzHugh-Nagumo (hypoelliptic) PartialBridgeνH pCN, 4 segments x N = 1001   (mode M)
          project_partialbridge/partialbridge_fitzhugh.jl:13-14,22,31-33,44-50,107-109 for the model,
          partialbridge_bolus3.jl:162-192 for the multi-segment backward/forward chaining
"""
from __future__ import annotations

import numpy as np

from . import api as B

# ---- config 4
FHN_PAR = (0.1, 0.0, 1.5, 0.8, 0.3)  # ϵ, s, γ, β, σ   (Ditlevsen-Samson; partialbridge_fitzhugh.jl:48)
FHN_X0 = np.array([-0.5, -0.6])      # :50
FHN_L = np.array([[1.0, 0.0]])       # :31
FHN_SIGMA = np.array([[1e-10]])      # :32-33
FHN_EPS = 1e-3                       # :22
FHN_RHO = 0.99                       # :109
FHN_OBS_T = (0.5, 1.0, 1.5, 2.0)
FHN_OBS_V = (-1.0, -0.5, 0.5, 1.1)


def tau_grid(t0: float, t1: float, n: int) -> np.ndarray:
    """τ(T)(x) = x (2 - x/T) on a uniform grid (partialbridge_fitzhugh.jl:13-14), shifted to [t0, t1]."""
    s = np.linspace(0.0, t1 - t0, n)
    return t0 + s * (2.0 - s / (t1 - t0))


def fhn_matching_aux(v: float, par=FHN_PAR):
    """The "matching" auxiliary process of partialbridge_fitzhugh.jl:106-108 (constant in t)."""
    ϵ, s, γ, β, σ = par
    Bt = np.array([[1 / ϵ, -1 / ϵ], [γ, -1.0]])
    bt = np.array([s / ϵ - (v * v * v) / ϵ, β])  # P.v^3 is a product in Julia (literal_pow)
    at = np.array([[0.0, 0.0], [0.0, σ * σ]])
    return Bt, bt, at


def fhn_segment_grids(n: int = 1001, obs_t=FHN_OBS_T):
    t = (0.0,) + tuple(obs_t)
    return [tau_grid(t[k], t[k + 1], n) for k in range(len(obs_t))]


def fhn_config4(n: int = 1001, ctx=None, obs_t=FHN_OBS_T, obs_v=FHN_OBS_V):
    """Target, one PartialBridgeνH per segment (backward chain right to left), start point, ρ."""
    P = B.FitzhughDiffusion(*FHN_PAR)
    grids = fhn_segment_grids(n, obs_t)
    S = len(grids)
    ν = np.zeros(2)
    Hp = np.eye(2) / FHN_EPS
    ν, Hp = B.gpupdate_νH(ν, Hp, FHN_L, FHN_SIGMA, [obs_v[-1]], ctx=ctx)
    guides = [None] * S
    for i in range(S - 1, -1, -1):
        Pt = B.LinearAux(*fhn_matching_aux(obs_v[i]))
        guides[i], ν, Hp, _ = B.partialbridgeνH(grids[i], P, Pt, ν, Hp, ctx=ctx)
        if i > 0:
            ν, Hp = B.gpupdate_νH(ν, Hp, FHN_L, FHN_SIGMA, [obs_v[i - 1]], ctx=ctx)
    return P, guides, FHN_X0.copy(), FHN_RHO


def fhn_config4_chain(n: int = 1001, ctx=None, obs_t=FHN_OBS_T, obs_v=FHN_OBS_V, chain=None):
    """The same proposals as fhn_config4 through the single-launch backward chain (tables stay on the device);
    `chain` = an existing PartialBridgeνHChain to update in place."""
    P = B.FitzhughDiffusion(*FHN_PAR)
    Pts = [B.LinearAux(*fhn_matching_aux(v)) for v in obs_v]
    if chain is None:
        chain = B.PartialBridgeνHChain(fhn_segment_grids(n, obs_t), P, Pts, FHN_L, FHN_SIGMA, obs_v, FHN_EPS, ctx=ctx)
    else:
        chain.update_(Pts, obs_v)
    return P, chain, FHN_X0.copy(), FHN_RHO


# ---- config 3
LIN3_B1 = -np.array([[1.0, 0.1, 0.0], [-0.2, 1.0, 0.1], [0.0, -0.1, 1.0]])
LIN3_B2 = -np.eye(3)
LIN3_SIG = 0.5 * np.eye(3)
LIN3_V = np.array([0.5, 0.0, -0.5])


def linpro_config3(n: int = 1001, ctx=None):
    P = B.LinPro(LIN3_B1, np.zeros(3), LIN3_SIG)
    Pt = B.LinPro(LIN3_B2, np.zeros(3), LIN3_SIG)
    tt = np.linspace(0.0, 1.0, n)
    return P, B.GuidedBridge(tt, P, Pt, LIN3_V, ctx=ctx), np.zeros(3)


# ---- config 5: Landmarks d = 16 (n = 4 landmarks in the plane), noise on the momenta (d' = 8)
# project_partialbridge/partialbridge_landmarks.jl:47,86-101,111-146 is an unfinished draft: the numbers
# below are this repository's choice and are recorded in BASELINE.md
LM_A, LM_SIGMA, LM_LAMBDA = 0.5, 2.0, 0.5
LM_Q0 = np.array([[-1.0, -1.0], [1.0, -1.2], [1.1, 0.9], [-0.8, 1.0]])
LM_P0 = np.array([[0.5, 0.1], [-0.2, 0.4], [-0.3, -0.3], [0.2, -0.5]])
LM_QT = np.array([[-0.6, -1.4], [1.5, -0.9], [0.8, 1.4], [-1.2, 0.7]])
LM_EPS, LM_OBS_VAR, LM_RHO = 1e-3, 1e-4, 0.9


def landmarks_config5(n: int = 1001, ctx=None, T: float = 1.0):
    """Target, PartialBridgeνH towards the observed end positions qT (positions observed, momenta free), start."""
    x0 = np.concatenate([np.concatenate([LM_Q0[i], LM_P0[i]]) for i in range(4)])
    L = np.zeros((8, 16))
    for i in range(4):
        for c in range(2):
            L[2 * i + c, 4 * i + c] = 1.0
    tt = tau_grid(0.0, T, n)
    Pm = B.Landmarks(LM_A, LM_SIGMA, LM_LAMBDA)
    Pt = B.LandmarksTilde(LM_A, LM_SIGMA, LM_LAMBDA, LM_QT)
    Po = B.PartialBridgeνH(tt, Pm, Pt, L, LM_QT.ravel(), LM_EPS, LM_OBS_VAR * np.eye(8), ctx=ctx)
    return Pm, Po, x0
