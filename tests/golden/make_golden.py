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


def savez(path, **arrays):
    """np.savez, but an existing file with identical contents is left alone (zip timestamps would dirty the work tree)."""
    if os.path.exists(path):
        old = np.load(path)
        if set(old.files) == set(arrays) and all(np.array_equal(old[k], np.asarray(v)) for k, v in arrays.items()):
            return
    np.savez(path, **arrays)
FHN_PAR = [0.1, 0.0, 1.5, 0.8, 0.3]


def warped(t0, t1, n):
    s = np.linspace(0, t1 - t0, n)
    return t0 + s * (2 - s / (t1 - t0))


def main():
    O.build()
    orc = O.load("ref")
    # ---- the reference's own golden vector
    savez(os.path.join(OUT, "reference_docs_ou.npz"),
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
    savez(os.path.join(OUT, "oracle_fhn_nuH_pcn.npz"), grids=np.stack(grids), nu=np.stack([t[0] for t in tabs]),
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
    savez(os.path.join(OUT, "oracle_linpro3_guidedbridge.npz"), tt=tt, B1=B1, sigma=sig, v=v3, Hdia=Hd, V=V, W=W3,
             X=X3, ll=ll3)
    # ---- IntegratedDiffusion, PartialBridge (L, M, mu) on the grid of test/partialparam.jl (first 201 points)
    ttp = np.arange(201) / 1000
    aux = dict(B=np.array([[0.0, 1.0], [0.0, -1.0]]), beta=np.array([0.0, 0.5]), a=np.array([[0.0, 0.0], [0.0, 0.49]]))
    Lt, Mt, mut = orc.backward_LMmu(ttp, O.const_aux(**aux), [[1.0, 0.0]], [[0.1]])
    omi = O.make_model(O.INTDIFF, 2, 1, [0.7])
    ogi = O.GuideHolder(O.GUIDE_LMMU, ttp, Lt, mut, Mm=Mt, v=[2.5], Bt=aux["B"], betat=aux["beta"], m=1)
    Wi = orc.wiener_sample(ttp, 1, 1, 0, 0)
    Xi, _ = orc.guided_euler(omi, ogi, [2.0, 1.0], Wi)
    savez(os.path.join(OUT, "oracle_intdiff_partialbridge.npz"), tt=ttp, L=Lt, M=Mt, mu=mut, v=np.array([2.5]),
             Bt=aux["B"], bt=aux["beta"], W=Wi, X=Xi, ll=orc.llikelihood(omi, ogi, Xi))
    make_theta_and_landmarks(orc)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


BOLUS_PAR = [70.0 / 0.6, 8.0, 15.0 / (20 * 0.6), 1.5, 0.5, 0.2]
BOLUS_PRIORS = {1: ("gamma", 1.0, 100.0), 4: ("gamma", 1.0, 100.0)}
BOLUS_RW = [0.0, 0.02, 0.0, 0.0, 0.02, 0.0, 0.0, 0.0]
LM_PAR = [0.5, 2.0, 0.5]
LM_QT = np.array([[-0.6, -1.4], [1.5, -0.9], [0.8, 1.4], [-1.2, 0.7]])
LM_X0 = np.array([-1.0, -1.0, 0.5, 0.1, 1.0, -1.2, -0.2, 0.4, 1.1, 0.9, -0.3, -0.3, -0.8, 1.0, 0.2, -0.5])


def make_theta_and_landmarks(orc):
    """Per-chain parameter step of partialbridge_bolus3.jl with its own model, and the landmarks bridge (config 5)."""
    # ---- bolus model: two chains, one parameter proposal each (reference arithmetic)
    n, S, seed, it = 41, 2, 21, 5
    tcut = (0.0, 0.8, 1.7); obs_v = [4.0, 9.0]
    grids = [warped(tcut[k], tcut[k + 1], n) for k in range(S)]
    x0 = np.array([0.5, 0.2]); L = np.array([[0.5, 0.5]]); Sg = 1e-2 * np.eye(1)
    # (the script's own Σ = 1e-4, ϵ = 1e-3 make the backward recursion ill-conditioned in rounding: tables computed with and
    #  without fused multiply-adds agree to 4e-4 only; Σ = 1e-2, ϵ = 0.1 agree to 1e-8, which is what a frozen fixture needs)
    EPSB = 0.1
    chains = [3, 250]
    th = np.array([BOLUS_PAR + [0.0, 0.0], [BOLUS_PAR[0], 7.1, BOLUS_PAR[2], BOLUS_PAR[3], 0.62, 0.2, 0.0, 0.0]])
    out = dict(grids=np.stack(grids), x0=x0, L=L, Sigma=Sg, eps=EPSB, obs_v=np.array(obs_v), chains=np.array(chains),
               theta=th, rw=np.array(BOLUS_RW), seed=seed, it=it)
    W = []; thp = []; left_c = []; left_o = []; Xo = []; llo = []; llc = []; logu = []
    for k, c in enumerate(chains):
        Wk = np.stack([orc.wiener_sample(grids[s], 2, seed, 0xFFFFFFF0, c * S + s) for s in range(S)])
        gc, lc = O.theta_backward(orc, O.BOLUS, th[k, :6], grids, x0, L, Sg, EPSB, obs_v, O.AUX_BOLUS, BOLUS_PRIORS)
        _, ll0, _ = O.theta_forward(orc, O.BOLUS, 2, th[k, :6], gc, x0, Wk)
        tp = O.theta_propose(orc, th[k], BOLUS_RW, seed, it, c)
        go, lo = O.theta_backward(orc, O.BOLUS, tp[:6], grids, x0, L, Sg, EPSB, obs_v, O.AUX_BOLUS, BOLUS_PRIORS)
        X1, ll1, _ = O.theta_forward(orc, O.BOLUS, 2, tp[:6], go, x0, Wk)
        W.append(Wk); thp.append(tp); Xo.append(X1); llo.append(ll1); llc.append(ll0)
        left_c.append([lc["lpn"], lc["trsum"], lc["lpri"]]); left_o.append([lo["lpn"], lo["trsum"], lo["lpri"]])
        logu.append(orc.logu_q(seed, it, c, O.Q_THETA_LOGU))
    out.update(W=np.stack(W), theta_prop=np.stack(thp), Xo=np.stack(Xo), llo=np.array(llo), llc=np.array(llc),
               left_c=np.array(left_c), left_o=np.array(left_o), logu=np.array(logu))
    savez(os.path.join(OUT, "oracle_bolus_theta_step.npz"), **out)
    # ---- landmarks: PartialBridgeνH tables at d = 16, two guided paths with log-likelihood, one pCN proposal
    N = 33
    tt = warped(0.0, 1.0, N)
    Lm = np.zeros((8, 16))
    for i in range(4):
        for c2 in range(2):
            Lm[2 * i + c2, 4 * i + c2] = 1.0
    a, sig, lam = LM_PAR
    Bt = np.zeros((16, 16))
    for i in range(4):
        for j in range(4):
            dq = LM_QT[i] - LM_QT[j]
            kk = np.exp(-np.dot(dq, dq) / (2 * a)) / (2 * np.pi * a)
            for c2 in range(2):
                Bt[4 * i + c2, 4 * j + 2 + c2] += 0.5 * kk
                Bt[4 * i + 2 + c2, 4 * j + 2 + c2] += -0.5 * lam * kk
    at = np.zeros((16, 16))
    for k in (2, 3, 6, 7, 10, 11, 14, 15):
        at[k, k] = sig * sig
    nu, Hp, C0 = orc.update_nuHC(Lm, 1e-4 * np.eye(8), LM_QT.ravel(), 1e-3)
    nus, Hs, _, _, Cc = orc.backward_nuH(O.ODE_R3, tt, O.const_aux(Bt, np.zeros(16), at), nu, Hp, C0)
    om = O.make_model(O.LANDMARKS, 16, 8, LM_PAR)
    og = O.GuideHolder(O.GUIDE_NUH, tt, Hs, nus, Bt=Bt, betat=np.zeros(16))
    chains = [0, 41]
    Wl = np.stack([orc.wiener_sample(tt, 8, 9, 0xFFFFFFF0, c) for c in chains])
    Xl = []; lll = []; Wo = []; Xo = []; llo = []; lu = []
    for k, c in enumerate(chains):
        X, _ = orc.guided_euler(om, og, LM_X0, Wl[k])
        Xl.append(X); lll.append(orc.llikelihood(om, og, X))
        l1, u1, W1, X1, _ = orc.pcn_propose(om, [og], LM_X0, Wl[k][None], 0.9, 9, 2, c)
        Wo.append(W1[0]); Xo.append(X1[0]); llo.append(l1); lu.append(u1)
    savez(os.path.join(OUT, "oracle_landmarks_bridge.npz"), tt=tt, par=np.array(LM_PAR), qT=LM_QT, x0=LM_X0, L=Lm,
             Bt=Bt, at=at, nu=nus, H=Hs, C=Cc, chains=np.array(chains), W=Wl, X=np.stack(Xl), ll=np.array(lll),
             Wo=np.stack(Wo), Xo=np.stack(Xo), llo=np.array(llo), logu=np.array(lu), rho=0.9, seed=9, it=2)


if __name__ == "__main__":
    main()
