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
