#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/ (run from the repository root: python tests/golden/make_golden.py).

The reference (Bridge.jl) is Julia and cannot be run in this environment, so the fixtures come from two sources:
  * reference_docs_ou.npz  -- the doctest of the reference itself (docs/src/manual.md:59-77): W and the Euler path X of
                              OrnsteinUhlenbeck(20, 1), as printed there (6 significant digits);
  * oracle_*.npz           -- seeded inputs and outputs of the CPU restatement in REFERENCE arithmetic
                              (oracle/liboracle_ref.so), frozen here so that later edits of the oracle or of the
                              kernels are caught: tests/test_golden.py replays them on the CPU (bit-exact) and on the
                              GPU (tolerance of test_gpu_parity.py).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
FHN_PAR = [0.1, 0.0, 1.5, 0.8, 0.3]


def warped(t0, t1, n):
    s = np.linspace(0, t1 - t0, n)
    return t0 + s * (2 - s / (t1 - t0))


def main():
    O.build()
    orc = O.load("ref")
    # ---- the reference's own golden vector
    np.savez(os.path.join(OUT, "reference_docs_ou.npz"),
             tt=np.arange(11) * 0.1,
             W=np.array([0.0, 0.0940107, 0.214935, 0.0259463, 0.0226432, -0.24268, -0.144298, 0.581472, -0.135443,
                         0.0321464, 0.168574]),
             X=np.array([0.1, -0.00598928, 0.126914, -0.315902, 0.312599, -0.577923, 0.676305, 0.0494658, -0.766381,
                         0.933971, -0.797544]), beta=20.0, sigma=1.0, u=0.1)
    # ---- FHN hypoelliptic, PartialBridgeνH, two chained segments, pCN proposal
    N, S, seed, rho = 97, 2, 11, 0.9
    grids = [warped(0.0, 0.5, N), warped(0.5, 1.0, N)]
    obs_v = [-1.0, -0.5]
    L = np.array([[1.0, 0.0]]); Sg = np.array([[1e-10]])
    nu = np.zeros(2); Hp = np.eye(2) / 1e-3
    nu, Hp = orc.gpupdate_nuH(nu, Hp, L, Sg, [obs_v[-1]])
    tabs = [None] * S
    for i in range(S - 1, -1, -1):
        v = obs_v[i]
        Bt = np.array([[10.0, -10.0], [1.5, -1.0]]); bt = np.array([-v ** 3 / 0.1, 0.8]); at = np.array([[0.0, 0.0], [0.0, 0.09]])
        nut, Ht, nu, Hp, _ = orc.backward_nuH(O.ODE_LYAP, grids[i], O.const_aux(Bt, bt, at), nu, Hp, 0.0)
        tabs[i] = (nut, Ht, Bt, bt)
        if i > 0:
            nu, Hp = orc.gpupdate_nuH(nu, Hp, L, Sg, [obs_v[i - 1]])
    om = O.make_model(O.FHN_HYPO, 2, 1, FHN_PAR)
    og = [O.GuideHolder(O.GUIDE_NUH, grids[s], tabs[s][1], tabs[s][0], Bt=tabs[s][2], betat=tabs[s][3]) for s in range(S)]
    x0 = np.array([-0.5, -0.6])
    chains = [0, 1, 77]
    Wc = np.stack([np.stack([orc.wiener_sample(grids[s], 1, seed, 0xFFFFFFFE, c * S + s) for s in range(S)]) for c in chains])
    Xc = np.zeros((len(chains), S, N, 2)); llc = np.zeros(len(chains))
    Wo = np.zeros_like(Wc); Xo = np.zeros_like(Xc); llo = np.zeros(len(chains)); logu = np.zeros(len(chains))
    for k, c in enumerate(chains):
        start = x0
        for s in range(S):
            Xc[k, s], start = orc.guided_euler(om, og[s], start, Wc[k, s])
            llc[k] += orc.llikelihood(om, og[s], Xc[k, s])
        llo[k], logu[k], Wo[k], Xo[k], _ = orc.pcn_propose(om, og, x0, Wc[k], rho, seed, 3, c)
    np.savez(os.path.join(OUT, "oracle_fhn_nuH_pcn.npz"), grids=np.stack(grids), nu=np.stack([t[0] for t in tabs]),
             H=np.stack([t[1] for t in tabs]), Bt=np.stack([t[2] for t in tabs]), bt=np.stack([t[3] for t in tabs]),
             par=np.array(FHN_PAR), x0=x0, chains=np.array(chains), seed=seed, rho=rho, it=3, W=Wc, X=Xc, ll=llc,
             Wo=Wo, Xo=Xo, llo=llo, logu=logu)
    # ---- LinPro d = 3, GuidedBridge (H♢, V) with end-point rule
    N3 = 129
    tt = np.linspace(0, 1, N3)
    B1 = -np.array([[1.0, 0.1, 0.0], [-0.2, 1.0, 0.1], [0.0, -0.1, 1.0]])
    sig = 0.5 * np.eye(3) + 0.05 * np.array([[0, 1, 0], [0, 0, 1], [1, 0, 0]])
    v3 = np.array([0.5, 0.0, -0.5])
    Hd, V = orc.backward_HV(tt, O.const_aux(-np.eye(3), np.zeros(3), sig @ sig.T), v3)
    om3 = O.linpro_model(B1, np.zeros(3), sig)
    og3 = O.GuideHolder(O.GUIDE_HV, tt, Hd, V, Bt=-np.eye(3), betat=np.zeros(3))
    W3 = np.stack([orc.wiener_sample(tt, 3, 3, 0, c) for c in range(3)])
    X3 = np.zeros((3, N3, 3)); ll3 = np.zeros(3)
    for c in range(3):
        X3[c], _ = orc.guided_euler(om3, og3, np.zeros(3), W3[c])
        ll3[c] = orc.llikelihood(om3, og3, X3[c])
    np.savez(os.path.join(OUT, "oracle_linpro3_guidedbridge.npz"), tt=tt, B1=B1, sigma=sig, v=v3, Hdia=Hd, V=V, W=W3,
             X=X3, ll=ll3)
    # ---- IntegratedDiffusion, PartialBridge (L, M, mu) on the grid of test/partialparam.jl (first 201 points)
    ttp = np.arange(201) / 1000
    aux = dict(B=np.array([[0.0, 1.0], [0.0, -1.0]]), beta=np.array([0.0, 0.5]), a=np.array([[0.0, 0.0], [0.0, 0.49]]))
    Lt, Mt, mut = orc.backward_LMmu(ttp, O.const_aux(**aux), [[1.0, 0.0]], [[0.1]])
    omi = O.make_model(O.INTDIFF, 2, 1, [0.7])
    ogi = O.GuideHolder(O.GUIDE_LMMU, ttp, Lt, mut, Mm=Mt, v=[2.5], Bt=aux["B"], betat=aux["beta"], m=1)
    Wi = orc.wiener_sample(ttp, 1, 1, 0, 0)
    Xi, _ = orc.guided_euler(omi, ogi, [2.0, 1.0], Wi)
    np.savez(os.path.join(OUT, "oracle_intdiff_partialbridge.npz"), tt=ttp, L=Lt, M=Mt, mu=mut, v=np.array([2.5]),
             Bt=aux["B"], bt=aux["beta"], W=Wi, X=Xi, ll=orc.llikelihood(omi, ogi, Xi))
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
