"""Replays the committed fixtures of tests/golden/ (see make_golden.py): on the CPU the oracle must reproduce them bit
for bit; on the GPU the CUDA path must reproduce them to the parity tolerance (the fixtures are in REFERENCE
arithmetic, the kernels use fused multiply-adds)."""
import os

import numpy as np
import pytest

from oracle import oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name))


def fhn_guides(f):
    S = f["grids"].shape[0]
    return [O.GuideHolder(O.GUIDE_NUH, f["grids"][s], f["H"][s], f["nu"][s], Bt=f["Bt"][s], betat=f["bt"][s])
            for s in range(S)]


def test_oracle_reproduces_reference_doc_vector(oracle_ref):
    f = load("reference_docs_ou.npz")
    X = oracle_ref.euler(O.make_model(O.OU, 1, 1, [float(f["beta"]), float(f["sigma"])]), f["tt"], [float(f["u"])], f["W"])
    assert np.max(np.abs(X[:, 0] - f["X"])) < 1e-5  # print precision of the doctest x stiffness (1 - 20*0.1 = -1)


def test_oracle_reproduces_golden_bit_exact(oracle_ref):
    f = load("oracle_fhn_nuH_pcn.npz")
    om = O.make_model(O.FHN_HYPO, 2, 1, f["par"]); og = fhn_guides(f)
    S = len(og)
    for k, c in enumerate(f["chains"]):
        for s in range(S):
            assert np.array_equal(oracle_ref.wiener_sample(f["grids"][s], 1, int(f["seed"]), 0xFFFFFFFE, int(c) * S + s), f["W"][k, s])
        llo, lu, Wo, Xo, _ = oracle_ref.pcn_propose(om, og, f["x0"], f["W"][k], float(f["rho"]), int(f["seed"]), int(f["it"]), int(c))
        assert np.array_equal(Wo, f["Wo"][k]) and np.array_equal(Xo, f["Xo"][k]) and llo == f["llo"][k] and lu == f["logu"][k]
    g = load("oracle_linpro3_guidedbridge.npz")
    om3 = O.linpro_model(g["B1"], np.zeros(3), g["sigma"])
    og3 = O.GuideHolder(O.GUIDE_HV, g["tt"], g["Hdia"], g["V"], Bt=-np.eye(3), betat=np.zeros(3))
    for c in range(3):
        X, _ = oracle_ref.guided_euler(om3, og3, np.zeros(3), g["W"][c])
        assert np.array_equal(X, g["X"][c]) and oracle_ref.llikelihood(om3, og3, X) == g["ll"][c]
    h = load("oracle_intdiff_partialbridge.npz")
    ogi = O.GuideHolder(O.GUIDE_LMMU, h["tt"], h["L"], h["mu"], Mm=h["M"], v=h["v"], Bt=h["Bt"], betat=h["bt"], m=1)
    X, _ = oracle_ref.guided_euler(O.make_model(O.INTDIFF, 2, 1, [0.7]), ogi, [2.0, 1.0], h["W"])
    assert np.array_equal(X, h["X"])


@pytest.mark.gpu
def test_gpu_reproduces_golden():
    import bridge_jl_b200 as B
    K = B.api.K
    xt = lambda a, b: np.max(np.abs(a - b)) <= 1e-10 * (1 + np.max(np.abs(b)))
    lt = lambda a, b: np.all(np.abs(a - b) <= 1e-6 * np.abs(b) + 1e-9)
    # reference doctest through the mirrored API
    f = load("reference_docs_ou.npz")
    X = B.solve(B.Euler(), float(f["u"]), B.SamplePath(f["tt"], f["W"]), B.OrnsteinUhlenbeck(float(f["beta"]), float(f["sigma"])))
    assert np.max(np.abs(X.yy - f["X"])) < 1e-5
    # FHN pCN
    f = load("oracle_fhn_nuH_pcn.npz")
    Pm = B.FitzhughDiffusion(*f["par"])
    S, N = f["grids"].shape
    guides = [B.GuideTables(K.GUIDE_NUH, f["grids"][s], Pm, f["H"][s], f["nu"][s], f["Bt"][s], f["bt"][s]) for s in range(S)]
    for k, c in enumerate(f["chains"]):
        ens = B.PathEnsemble(1, S, N, 2, 1, chain_offset=int(c))
        for s in range(S):
            ens.set_grid(s, f["grids"][s])
        ens.set_start(f["x0"])
        ens.sample_(int(f["seed"]), 0xFFFFFFFE)
        assert xt(ens.download(B.W)[0], f["W"][k])   # same normals; the running sum uses fma on the device
        ens.guided_euler_ll_(Pm, guides)
        assert xt(ens.download(B.X)[0], f["X"][k]) and lt(ens.ll[0], f["ll"][k])
        ens.pcn_step_(Pm, guides, float(f["rho"]), int(f["seed"]), int(f["it"]))
        assert xt(ens.download(B.W, which=B.PROP)[0], f["Wo"][k]) and xt(ens.download(B.X, which=B.PROP)[0], f["Xo"][k])
        assert lt(ens.ll_prop[0], f["llo"][k]) and ens.logu[0] == f["logu"][k]
        ens.close()
    g = load("oracle_linpro3_guidedbridge.npz")
    P3 = B.LinPro(g["B1"], np.zeros(3), g["sigma"])
    G3 = B.GuideTables(K.GUIDE_HV, g["tt"], P3, g["Hdia"], g["V"], -np.eye(3), np.zeros(3))
    ens = B.PathEnsemble(3, 1, len(g["tt"]), 3, 3, double_buffer=False)
    ens.set_start(np.zeros(3)); ens.upload(B.W, g["W"][:, None]); ens.guided_euler_ll_(P3, [G3])
    X = ens.download(B.X)[:, 0]
    assert xt(X[:, :-1], g["X"][:, :-1]) and np.array_equal(X[:, -1], g["X"][:, -1]) and lt(ens.ll, g["ll"])
    ens.close()
    h = load("oracle_intdiff_partialbridge.npz")
    Pi = B.IntegratedDiffusion(0.7)
    Gi = B.GuideTables(K.GUIDE_LMMU, h["tt"], Pi, h["L"], h["mu"], h["Bt"], h["bt"], Mm=h["M"], v=h["v"], m=1)
    ens = B.PathEnsemble(1, 1, len(h["tt"]), 2, 1, double_buffer=False)
    ens.set_start([2.0, 1.0]); ens.upload(B.W, h["W"][None, None]); ens.guided_euler_ll_(Pi, [Gi])
    assert xt(ens.download(B.X)[0, 0], h["X"]) and lt(ens.ll[0], float(h["ll"]))
    ens.close()


BOLUS_PRIORS = {1: ("gamma", 1.0, 100.0), 4: ("gamma", 1.0, 100.0)}


def test_oracle_reproduces_theta_and_landmarks_golden(oracle_ref):
    f = load("oracle_bolus_theta_step.npz")
    grids = list(f["grids"]); S = len(grids)
    for k, c in enumerate(f["chains"]):
        c = int(c)
        for s in range(S):
            assert np.array_equal(oracle_ref.wiener_sample(grids[s], 2, int(f["seed"]), 0xFFFFFFF0, c * S + s), f["W"][k, s])
        tp = O.theta_propose(oracle_ref, f["theta"][k], f["rw"], int(f["seed"]), int(f["it"]), c)
        assert np.array_equal(tp, f["theta_prop"][k])
        go, lo = O.theta_backward(oracle_ref, O.BOLUS, tp[:6], grids, f["x0"], f["L"], f["Sigma"], float(f["eps"]),
                                  list(f["obs_v"]), O.AUX_BOLUS, BOLUS_PRIORS)
        X, ll, _ = O.theta_forward(oracle_ref, O.BOLUS, 2, tp[:6], go, f["x0"], f["W"][k])
        assert np.array_equal(X, f["Xo"][k]) and ll == f["llo"][k]
        assert [lo["lpn"], lo["trsum"], lo["lpri"]] == list(f["left_o"][k])
        assert oracle_ref.logu_q(int(f["seed"]), int(f["it"]), c, O.Q_THETA_LOGU) == f["logu"][k]
    g = load("oracle_landmarks_bridge.npz")
    om = O.make_model(O.LANDMARKS, 16, 8, g["par"])
    nu, Hp, C0 = oracle_ref.update_nuHC(g["L"], 1e-4 * np.eye(8), g["qT"].ravel(), 1e-3)
    nus, Hs, _, _, Cc = oracle_ref.backward_nuH(O.ODE_R3, g["tt"], O.const_aux(g["Bt"], np.zeros(16), g["at"]), nu, Hp, C0)
    assert np.array_equal(nus, g["nu"]) and np.array_equal(Hs, g["H"]) and Cc == float(g["C"])
    og = O.GuideHolder(O.GUIDE_NUH, g["tt"], g["H"], g["nu"], Bt=g["Bt"], betat=np.zeros(16))
    for k, c in enumerate(g["chains"]):
        X, _ = oracle_ref.guided_euler(om, og, g["x0"], g["W"][k])
        assert np.array_equal(X, g["X"][k]) and oracle_ref.llikelihood(om, og, X) == g["ll"][k]
        llo, lu, Wo, Xo, _ = oracle_ref.pcn_propose(om, [og], g["x0"], g["W"][k][None], float(g["rho"]), int(g["seed"]),
                                                    int(g["it"]), int(c))
        assert np.array_equal(Wo[0], g["Wo"][k]) and np.array_equal(Xo[0], g["Xo"][k]) and llo == g["llo"][k]


@pytest.mark.gpu
def test_gpu_reproduces_theta_and_landmarks_golden():
    """The fixtures are in REFERENCE arithmetic (no fma, libm exp).  The bolus fixture uses Σ = 1e-2, ϵ = 0.1, for which
    the backward recursion is well conditioned (tables with and without fma agree to 1e-8; with the script's Σ = 1e-4,
    ϵ = 1e-3 only to 4e-4), and the device builds its own tables: paths to 1e-6, ll to 1e-5.  The landmarks run uses the
    fixture's tables and is compared to the parity tolerance."""
    import bridge_jl_b200 as B
    K = B.api.K
    f = load("oracle_bolus_theta_step.npz")
    grids = list(f["grids"]); S, n = len(grids), len(grids[0])
    Pm = B.BolusDiffusion(*f["theta"][0, :6])
    for k, c in enumerate(f["chains"]):
        ens = B.PathEnsemble(1, S, n, 2, 2, chain_offset=int(c))
        for s in range(S):
            ens.set_grid(s, grids[s])
        ens.set_start(f["x0"])
        ens.theta_attach_(Pm, f["L"], f["Sigma"], float(f["eps"]), f["obs_v"], aux_kind=K.AUX_BOLUS, priors=BOLUS_PRIORS)
        ens.set_theta(f["theta"][k][None])
        ens.sample_(int(f["seed"]), 0xFFFFFFF0)
        assert np.max(np.abs(ens.download(B.W)[0] - f["W"][k])) <= 1e-10 * (1 + np.max(np.abs(f["W"][k])))
        ens.theta_guided_euler_ll_()
        assert abs(ens.ll[0] - f["llc"][k]) <= 1e-5 * abs(f["llc"][k]) + 1e-6, (ens.ll[0], f["llc"][k])
        ens.theta_param_step_(f["rw"], int(f["seed"]), int(f["it"]))
        assert np.array_equal(ens.theta(B.PROP)[0], f["theta_prop"][k]) and ens.logu[0] == f["logu"][k]
        Xo = ens.download(B.X, which=B.PROP)[0]
        assert np.max(np.abs(Xo - f["Xo"][k])) <= 1e-6 * (1 + np.max(np.abs(f["Xo"][k])))
        assert abs(ens.ll_prop[0] - f["llo"][k]) <= 1e-5 * abs(f["llo"][k]) + 1e-6
        left = ens.theta_left(B.PROP)[0]
        assert np.allclose(left[7:10], f["left_o"][k], rtol=1e-5, atol=1e-9)  # logpdfnormal amplifies the table rounding
        ens.close()
    g = load("oracle_landmarks_bridge.npz")
    Pl = B.Landmarks(*g["par"])
    Gl = B.GuideTables(K.GUIDE_NUH, g["tt"], Pl, g["H"], g["nu"], g["Bt"], np.zeros(16))
    N = len(g["tt"])
    for k, c in enumerate(g["chains"]):
        ens = B.PathEnsemble(1, 1, N, 16, 8, chain_offset=int(c))
        ens.set_grid(0, g["tt"]); ens.set_start(g["x0"]); ens.sample_(int(g["seed"]), 0xFFFFFFF0)
        ens.guided_euler_ll_(Pl, [Gl])
        X = ens.download(B.X)[0, 0]
        assert np.max(np.abs(X - g["X"][k])) <= 1e-9 * (1 + np.max(np.abs(g["X"][k])))
        assert abs(ens.ll[0] - g["ll"][k]) <= 1e-6 * abs(g["ll"][k]) + 1e-9
        ens.pcn_step_(Pl, [Gl], float(g["rho"]), int(g["seed"]), int(g["it"]))
        assert np.max(np.abs(ens.download(B.W, which=B.PROP)[0, 0] - g["Wo"][k])) <= 1e-10 * (1 + np.max(np.abs(g["Wo"][k])))
        assert np.max(np.abs(ens.download(B.X, which=B.PROP)[0, 0] - g["Xo"][k])) <= 1e-9 * (1 + np.max(np.abs(g["Xo"][k])))
        assert abs(ens.ll_prop[0] - g["llo"][k]) <= 1e-6 * abs(g["llo"][k]) + 1e-9 and ens.logu[0] == g["logu"][k]
        ens.close()
