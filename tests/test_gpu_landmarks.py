"""BASELINE config 5: Landmarks (d = 16, d' = 8) partial bridge on the wide path (csrc/bb_wide.cuh,
csrc/bb_backward_gen.cu) against the oracle restatement of project_partialbridge/partialbridge_landmarks.jl:47,86-101,
111-146.  The reference script is an unfinished draft that does not run, so this parity is oracle-only.

  * constructors (updateνH⁺C, partialbridgeodeνH! R3 / Lyap at d = 16, m = 8) vs liboracle_fma: BIT-EXACT;
  * sample! of the 8-dimensional Wiener process and the pCN proposal W°: BIT-EXACT;
  * paths / log-likelihoods vs liboracle_fma: BIT-EXACT (the kernel's exponential is built from +, *, fma, rint and is
    restated in the oracle's GPU-order build); vs liboracle_ref (libm exp, no fma): |dX| <= 1e-9 (1 + |X|),
    |dll| <= 1e-6 |ll| + 1e-9; accept decisions replayed exactly from the kernel's numbers.
"""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

A_K, SIG, LAM = 0.5, 2.0, 0.5  # a, σ, λ (the script's aa = 10 is for its own length scale; here |q_i - q_j| ~ 2)
Q0 = np.array([[-1.0, -1.0], [1.0, -1.2], [1.1, 0.9], [-0.8, 1.0]])
P0 = np.array([[0.5, 0.1], [-0.2, 0.4], [-0.3, -0.3], [0.2, -0.5]])
QT = np.array([[-0.6, -1.4], [1.5, -0.9], [0.8, 1.4], [-1.2, 0.7]])
EPS, SDIAG = 1e-3, 1e-4


@pytest.fixture(scope="module")
def B():
    import bridge_jl_b200 as B
    B.default_context()
    return B


@pytest.fixture
def no_dmma():
    """The four-lane kernel WITHOUT tensor-core mat-vecs: the rounding sequence of the oracle's GPU-order build, used
    for the bit-exact statements.  (The default guided kernel, bb_wide4m_kernel, sums the mat-vecs on the fp64 tensor
    cores and is compared at the contract tolerance: test_landmarks_tensor_core_kernel_vs_reference_arithmetic.)"""
    import os
    os.environ["BB_WIDE_MMA"] = "0"
    yield
    os.environ.pop("BB_WIDE_MMA", None)


def x0():
    return np.concatenate([np.concatenate([Q0[i], P0[i]]) for i in range(4)])


def obs():
    L = np.zeros((8, 16))
    for i in range(4):
        for c in range(2):
            L[2 * i + c, 4 * i + c] = 1.0
    return L, SDIAG * np.eye(8), QT.ravel()


def setup(B, orc, N, T=1.0):
    tt = np.linspace(0.0, T, N) * (2 - np.linspace(0.0, T, N) / T)
    Pm = B.Landmarks(A_K, SIG, LAM)
    Pt = B.LandmarksTilde(A_K, SIG, LAM, QT)
    L, Σ, v = obs()
    Po = B.PartialBridgeνH(tt, Pm, Pt, L, v, EPS, Σ)
    om = O.make_model(O.LANDMARKS, 16, 8, [A_K, SIG, LAM])
    og = O.GuideHolder(O.GUIDE_NUH, tt, Po.H, Po.ν, Bt=Pt.B(0.0), betat=Pt.β(0.0))
    return tt, Pm, Pt, Po, om, og


def test_landmarks_drift_and_aux_consistency(B, oracle_ref):
    """b(x) of the oracle equals the host mirror's restatement; with q frozen at qT the q-part of the drift is B~ x."""
    Pm = B.Landmarks(A_K, SIG, LAM)
    om = O.make_model(O.LANDMARKS, 16, 8, [A_K, SIG, LAM])
    rng = np.random.default_rng(0)
    for _ in range(5):
        x = rng.standard_normal(16)
        b = np.zeros(16)
        oracle_ref.lib.bbo_model_b(O.C.byref(om), O.C.c_double(0.0), O._p(x), O._p(b))
        assert np.allclose(b, Pm.b(0.0, x), rtol=1e-13, atol=1e-15)
    Pt = B.LandmarksTilde(A_K, SIG, LAM, QT)
    x = x0().reshape(4, 2, 2); x[:, 0] = QT
    bq = Pm.b(0.0, x.ravel()).reshape(4, 2, 2)[:, 0]
    assert np.allclose((Pt.B(0.0) @ x.ravel()).reshape(4, 2, 2)[:, 0], bq, rtol=1e-13)


def test_landmarks_constructors_bit_exact(B, oracle_fma):
    N = 41
    tt, Pm, Pt, Po, om, og = setup(B, oracle_fma, N)
    L, Σ, v = obs()
    nu, Hp, C0 = oracle_fma.update_nuHC(L, Σ, v, EPS)
    aux = O.const_aux(Pt.B(0.0), Pt.β(0.0), Pt.a(0.0))
    nus, Hs, _, _, Cc = oracle_fma.backward_nuH(O.ODE_R3, tt, aux, nu, Hp, C0)
    assert np.array_equal(Po.ν, nus) and np.array_equal(Po.H, Hs)
    assert abs(Po.C - Cc) <= 1e-12 * abs(Cc)  # C passes through the device log
    # Lyapunov variant (5-argument partialbridgeνH) at d = 16
    Po2, nul, Hpl, C2 = B.partialbridgeνH(tt, Pm, Pt, nu, Hp)
    nus2, Hs2, nul_o, Hpl_o, C2o = oracle_fma.backward_nuH(O.ODE_LYAP, tt, aux, nu, Hp, 0.0)
    assert np.array_equal(Po2.ν, nus2) and np.array_equal(Po2.H, Hs2)
    assert np.array_equal(nul, nul_o) and np.array_equal(Hpl, Hpl_o) and C2 == C2o


def test_landmarks_guided_euler_pcn_vs_oracle(B, oracle_fma, oracle_ref, no_dmma):
    N, P, seed, rho = 97, 150, 5, 0.9
    tt, Pm, Pt, Po, om, og = setup(B, oracle_fma, N)
    ens = B.PathEnsemble(P, 1, N, 16, 8, chain_offset=300)
    ens.set_grid(0, tt)
    ens.set_start(x0())
    ens.sample_(seed, 0xFFFFFFF0)
    W = ens.download(B.W)
    for p in (0, 63, 64, 149):
        assert np.array_equal(W[p, 0], oracle_fma.wiener_sample(tt, 8, seed, 0xFFFFFFF0, 300 + p))
    ens.guided_euler_ll_(Pm, [Po])
    X = ens.download(B.X); ll = ens.ll
    assert np.all(np.isfinite(X)) and np.all(np.isfinite(ll))
    for p in (0, 63, 64, 149):
        for o in (oracle_fma, oracle_ref):
            Xo, xend = o.guided_euler(om, og, x0(), W[p, 0])
            llo = o.llikelihood(om, og, Xo)
            if o is oracle_fma:
                assert np.array_equal(X[p, 0], Xo) and ll[p] == llo, p
            assert np.max(np.abs(X[p, 0] - Xo)) <= 1e-9 * (1 + np.max(np.abs(Xo))), (p, np.max(np.abs(X[p, 0] - Xo)))
            assert abs(ll[p] - llo) <= 1e-6 * abs(llo) + 1e-9
    # the bridge ends near the observed positions
    assert np.max(np.abs(X[:, 0, -1].reshape(P, 4, 2, 2)[:, :, 0] - QT)) < 0.15
    nacc = 0
    for it in range(3):
        llc = ens.ll; Wc = ens.download(B.W)
        ens.pcn_step_(Pm, [Po], rho, seed, it, store_x=(it != 1))
        llp, logu, flags = ens.ll_prop, ens.logu, ens.accepted.astype(bool)
        assert np.array_equal(flags, logu <= llp - llc)
        assert np.array_equal(ens.ll, np.where(flags, llp, llc))
        Wp = ens.download(B.W, which=B.PROP)
        for p in (0, 63, 64, 149):
            llo, lu, Wo, Xo, _ = oracle_fma.pcn_propose(om, [og], x0(), Wc[p], rho, seed, it, 300 + p)
            assert np.array_equal(Wp[p], Wo) and logu[p] == lu
            assert llp[p] == llo
            if it != 1:
                Xp = ens.download(B.X, which=B.PROP, p0=p, np_=1)
                assert np.array_equal(Xp[0], Xo)
        nacc += int(flags.sum())
    assert ens.acc == nacc and 0 < nacc < 3 * P
    # current paths: rejected proposals are recomputed from the current W
    Xc = ens.download(B.X); Wc = ens.download(B.W)
    for p in (0, 63, 64, 149):
        Xo, _ = oracle_fma.guided_euler(om, og, x0(), Wc[p, 0])
        assert np.array_equal(Xc[p, 0], Xo)
    ens.close()


def test_landmarks_plain_euler_and_unsupported(B, oracle_fma):
    N, P = 33, 70
    tt = np.linspace(0.0, 0.3, N)
    Pm = B.Landmarks(A_K, SIG, LAM)
    om = O.make_model(O.LANDMARKS, 16, 8, [A_K, SIG, LAM])
    ens = B.PathEnsemble(P, 1, N, 16, 8, double_buffer=False)
    ens.set_grid(0, tt); ens.set_start(x0()); ens.sample_(2, 1)
    W = ens.download(B.W)
    ens.euler_(Pm)
    X = ens.download(B.X)
    for p in (0, 69):
        Xo = oracle_fma.euler(om, tt, x0(), W[p, 0])
        assert np.array_equal(X[p, 0], Xo)
    # fused sample! + solve! gives the same W and X
    ens2 = B.PathEnsemble(P, 1, N, 16, 8, double_buffer=False)
    ens2.set_grid(0, tt); ens2.set_start(x0()); ens2.sample_euler_(Pm, 2, 1)
    assert np.array_equal(ens2.download(B.W), W) and np.array_equal(ens2.download(B.X), X)
    ens2.close()
    # not on the wide path: innovations! (sigma is 16 x 8), pooled statistics, second-pass llikelihood, other schemes
    with pytest.raises(B.BridgeError) as ei:
        ens.innovations_(Pm)
    assert ei.value.status in (-11, -12)
    with pytest.raises(B.BridgeError) as ei:
        ens.mc_update_()
    assert ei.value.status == -11
    with pytest.raises(B.BridgeError) as ei:
        ens.solve_scheme_(Pm, B.StochasticHeun.scheme)
    assert ei.value.status == -11
    ens.close()


def test_landmarks_four_lane_kernel_equals_one_thread_per_chain(B, no_dmma):
    """The production kernels split a chain over four lanes (one landmark each); BB_WIDE_LANES=1 selects the plain
    one-thread-per-chain kernels.  Same rounding sequence per component: results must agree bit for bit."""
    import os
    N, P, seed, rho = 65, 203, 9, 0.8
    tt, Pm, Pt, Po, om, og = setup(B, None, N)
    out = {}
    for lanes in ("4", "1"):
        os.environ["BB_WIDE_LANES"] = lanes
        try:
            ens = B.PathEnsemble(P, 1, N, 16, 8)
            ens.set_grid(0, tt); ens.set_start(x0()); ens.sample_(seed, 3)
            ens.euler_(Pm)
            Xe = ens.download(B.X)
            ens.guided_euler_ll_(Pm, [Po], skip=1)
            res = [Xe, ens.download(B.X), ens.ll, ens.xend]
            for it in range(3):
                ens.pcn_step_(Pm, [Po], rho, seed, it, skip=1, store_x=(it != 1))
                res += [ens.ll_prop, ens.logu, ens.accepted, ens.download(B.W, which=B.PROP), ens.xend_prop]
            res += [ens.download(B.X), ens.download(B.W), ens.ll, np.array([ens.acc])]
            out[lanes] = res
            ens.close()
        finally:
            os.environ.pop("BB_WIDE_LANES", None)
    assert len(out["4"]) == len(out["1"])
    for u, v in zip(out["4"], out["1"]):
        assert np.array_equal(u, v)


def test_landmarks_tensor_core_kernel_vs_reference_arithmetic(B, oracle_ref):
    """The default guided kernel of config 5 computes r = H (nu - x) and B~ x with fp64 tensor-core instructions
    (mma.sync m8n8k4.f64, 8 chains per warp) and reduces the Girsanov sum by a butterfly: its sums round in another
    order than the reference's, so it is held to the CONTRACT tolerance against the reference arithmetic (libm exp, no
    fma, sequential sums): |dX| <= 1e-9 (1 + |X|), |dll| <= 1e-6 |ll| + 1e-9, same noise, accept decisions replayed
    from the kernel's own numbers with flips against the oracle counted; and against the four-lane kernel without DMMA."""
    import os
    N, P, seed, rho = 129, 203, 5, 0.9
    tt, Pm, Pt, Po, om, og = setup(B, oracle_ref, N)
    res = {}
    for mma in ("1", "0"):
        os.environ["BB_WIDE_MMA"] = mma
        try:
            ens = B.PathEnsemble(P, 1, N, 16, 8, chain_offset=300)
            ens.set_grid(0, tt); ens.set_start(x0()); ens.sample_(seed, 0xFFFFFFF0)
            W = ens.download(B.W)
            ens.guided_euler_ll_(Pm, [Po], skip=1)
            X = ens.download(B.X); ll = ens.ll; xend = ens.xend
            rec = [W, X, ll, xend]
            flips = 0
            for it in range(3):
                llc = ens.ll; Wc = ens.download(B.W)
                ens.pcn_step_(Pm, [Po], rho, seed, it, skip=1, store_x=(it != 1))
                llp, logu, flags = ens.ll_prop, ens.logu, ens.accepted.astype(bool)
                assert np.array_equal(flags, logu <= llp - llc)
                rec += [ens.download(B.W, which=B.PROP), llp, logu]
                if mma == "1":
                    for p in (0, 7, 8, 100, 202):
                        llo, lu, Wo, Xo, _ = oracle_ref.pcn_propose(om, [og], x0(), Wc[p], rho, seed, it, 300 + p, skip=1)
                        assert lu == logu[p] and np.max(np.abs(ens.download(B.W, which=B.PROP, p0=p, np_=1)[0] - Wo)) <= 1e-12
                        assert abs(llp[p] - llo) <= 1e-6 * abs(llo) + 1e-9, (llp[p], llo)
                        flips += int(flags[p] != (lu <= llo - llc[p]))
                        if it != 1:
                            Xp = ens.download(B.X, which=B.PROP, p0=p, np_=1)[0]
                            assert np.max(np.abs(Xp - Xo)) <= 1e-9 * (1 + np.max(np.abs(Xo)))
            assert flips == 0
            rec += [ens.download(B.X), ens.ll, np.array([ens.acc])]
            res[mma] = rec
            if mma == "1":
                for p in (0, 7, 8, 100, 202):
                    Xo, _ = oracle_ref.guided_euler(om, og, x0(), W[p, 0])
                    llo = oracle_ref.llikelihood(om, og, Xo, 1)
                    assert np.max(np.abs(X[p, 0] - Xo)) <= 1e-9 * (1 + np.max(np.abs(Xo))), np.max(np.abs(X[p, 0] - Xo))
                    assert abs(ll[p] - llo) <= 1e-6 * abs(llo) + 1e-9
                assert np.max(np.abs(X[:, 0, -1].reshape(P, 4, 2, 2)[:, :, 0] - QT)) < 0.15
            ens.close()
        finally:
            os.environ.pop("BB_WIDE_MMA", None)
    a, b = res["1"], res["0"]
    assert np.array_equal(a[0], b[0])  # the noise is the same bits
    for u, v in zip(a[1:4], b[1:4]):   # paths / ll / end points: rounding-level differences only
        assert np.max(np.abs(u - v)) <= 1e-10 * (1 + np.max(np.abs(v)))
    print("tensor-core vs four-lane kernel: max |dX| %.2e, max rel dll %.2e" % (
        np.max(np.abs(a[1] - b[1])), np.max(np.abs(a[2] - b[2]) / np.abs(b[2]))))
