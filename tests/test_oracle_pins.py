"""Pins the CPU oracle (oracle/bridge_oracle.c) to the reference's own golden vector and
known-answer tests.  CPU only.

Each test names the reference test / doc it restates.  Seeded Julia streams cannot be
reproduced (no Julia here), so only seed-free assertions are restated; where the
reference test is statistical, the same statistic is computed with Philox noise.
"""
import numpy as np
import pytest
from scipy.linalg import expm, solve_continuous_lyapunov

from oracle import oracle as O


# --------------------------------------------------------------------------- RNG
def test_philox_known_answers(oracle_ref):
    """Random123 kat_vectors for philox4x32-10 (Salmon et al., SC'11)."""
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kat:
        got = oracle_ref.philox(ctr, key)
        assert [int(x) for x in got] == want


def test_normals_are_standard(oracle_ref):
    z = np.array([oracle_ref.normal(7, 3, 11, n) for n in range(40000)])
    assert abs(z.mean()) < 4 / np.sqrt(z.size)
    assert abs(z.var() - 1) < 0.03
    assert abs((z ** 3).mean()) < 0.05 and abs((z ** 4).mean() - 3) < 0.15
    # pairs are (sin, cos) of the same radius: uncorrelated
    assert abs(np.corrcoef(z[0::2], z[1::2])[0, 1]) < 0.03
    # distinct rows / streams give distinct numbers
    assert oracle_ref.normal(7, 3, 11, 0) != oracle_ref.normal(7, 3, 12, 0)
    assert oracle_ref.normal(7, 3, 11, 0) != oracle_ref.normal(7, 4, 11, 0)


def test_wiener_sample_recurrence(oracle_ref):
    """src/wiener.jl:50-58: yy[1] kept, yy[i] = yy[i-1] + sqrt(dt_i) * randn, component-minor."""
    tt = np.array([0.0, 0.1, 0.25, 0.7, 1.0])
    W = oracle_ref.wiener_sample(tt, 2, seed=5, stream=1, row=9, y1=[0.5, -1.0])
    assert np.array_equal(W[0], [0.5, -1.0])
    for j in range(1, 5):
        for k in range(2):
            xi = oracle_ref.normal(5, 1, 9, j * 2 + k)
            assert W[j, k] == W[j - 1, k] + np.sqrt(tt[j] - tt[j - 1]) * xi


# --------------------------------------------------------------------------- Euler-Maruyama
GOLD_TT = np.array([0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0])
GOLD_W = np.array([0.0, 0.0940107, 0.214935, 0.0259463, 0.0226432, -0.24268, -0.144298,
                   0.581472, -0.135443, 0.0321464, 0.168574])
GOLD_X = np.array([0.1, -0.00598928, 0.126914, -0.315902, 0.312599, -0.577923, 0.676305,
                   0.0494658, -0.766381, 0.933971, -0.797544])


def test_docs_golden_vector(oracle_ref, oracle_fma):
    """docs/src/manual.md:59-63,73-77: solve(Euler(), 0.1, W, OrnsteinUhlenbeck(20.0, 1.0)).
    W and X are printed with 6 significant digits; the stiff recurrence (1-20*0.1 = -1)
    carries the print error of W without damping, hence 1e-5 absolute."""
    P = O.make_model(O.OU, 1, 1, [20.0, 1.0])
    for orc in (oracle_ref, oracle_fma):
        X = orc.euler(P, GOLD_TT, [0.1], GOLD_W)[:, 0]
        assert np.max(np.abs(X - GOLD_X)) < 1e-5


def test_euler_matches_independent_loop(oracle_ref):
    """test/euler.jl:63-68 (three Euler code paths agree to eps()) restated as: the C oracle
    agrees to the last bit with a direct numpy transcription of src/euler.jl:144-151 on Lorenz d=3."""
    n = 500
    tt = np.linspace(0.0, 1.0, n + 1)
    P = O.make_model(O.LORENZ, 3, 3, [10.0, 28.0, 8 / 3, 3.0, 3.0, 3.0])
    W = oracle_ref.wiener_sample(tt, 3, 1, 0, 0)
    X = oracle_ref.euler(P, tt, [1.0, 0.0, 0.0], W)
    y = np.array([1.0, 0.0, 0.0])
    for i in range(n):
        assert np.array_equal(X[i], y)
        b = np.array([10.0 * (y[1] - y[0]), y[0] * (28.0 - y[2]) - y[1], y[0] * y[1] - (8 / 3) * y[2]])
        y = y + b * (tt[i + 1] - tt[i]) + 3.0 * (W[i + 1] - W[i])
    assert np.array_equal(X[n], y)


def test_wiener_as_process_reproduces_w(oracle_ref):
    """BASELINE config 2: P = Wiener() (b=0, sigma=I) gives X = u + W."""
    tt = np.linspace(0, 1, 101)
    W = oracle_ref.wiener_sample(tt, 1, 2, 0, 5)
    X = oracle_ref.euler(O.make_model(O.WIENER, 1, 1), tt, [0.0], W)
    assert np.allclose(X, W, rtol=0, atol=1e-14)


# --------------------------------------------------------------------------- LinPro closed forms
def linpro_closed(B, a, mu):
    lam = solve_continuous_lyapunov(B, -a)  # B lam + lam B' + a = 0   (src/linpro.jl:74)
    lam = 0.5 * (lam + lam.T)

    def Hinv(t, T):  # inverse of src/linpro.jl:122-125
        phim = expm(-(T - t) * B)
        return phim @ lam @ phim.T - lam

    def V(t, T, v):  # src/linpro.jl:127-130
        return expm(-(T - t) * B) @ (v - mu) + mu

    return lam, Hinv, V


B2 = np.array([[-1.0, 0.1], [-0.2, -1.0]])
SIG2 = 2 * np.array([[-0.212887, 0.0687025], [0.193157, 0.388997]])
A2 = SIG2 @ SIG2.T


def test_linprobridge_r3_vs_closed_form(oracle_ref):
    """test/linprobridge.jl:22-25: R3 backward solution of _dHinv and V on tt=0:2/10000:2 equals
    the matrix-exponential closed forms to 1e-8."""
    n, T = 10000, 2.0
    tt = np.arange(n + 1) * (T / n)
    mu = np.zeros(2)
    v = np.array([0.5, 0.0])
    aux = O.const_aux(B2, -B2 @ mu, A2)
    Hd, Vt = oracle_ref.backward_HV(tt, aux, v)
    _, Hinv, Vcf = linpro_closed(B2, A2, mu)
    assert np.linalg.norm(np.linalg.inv(Hd[0]) - np.linalg.inv(Hinv(0.0, T))) < 1e-8
    assert np.linalg.norm(Hd[0] - Hinv(0.0, T)) < 1e-8
    assert np.linalg.norm(Vt[0] - Vcf(0.0, T, v)) < 1e-8


def test_linpro_n150(oracle_ref):
    """test/linpro.jl:38-53: gpHinv!/gpV! at N=150 on [0.5, 2]: |K[1] H - I| and |V[1] - V_cf| < 10/150^3."""
    n2 = 150
    t, T = 0.5, 2.0
    tt = np.linspace(t, T, n2)
    mu = 0.1 * np.array([0.2, 0.3])
    v = np.array([0.5, 0.0])
    aux = O.const_aux(B2, -B2 @ mu, A2)
    K, Vt = oracle_ref.backward_HV(tt, aux, v)
    _, Hinv, Vcf = linpro_closed(B2, A2, mu)
    H_cf = np.linalg.inv(Hinv(t, T))
    assert np.linalg.norm(K[0] @ H_cf - np.eye(2)) < 10 / n2 ** 3
    assert np.linalg.norm(Vt[0] - Vcf(t, T, v)) < 10 / n2 ** 3


def test_vhk_1d(oracle_ref):
    """test/VHK.jl:29-30: GuidedBridge.H♢, V vs closed form to 1e-5 (n=200, T=2, LinPro(-0.8, 0.2, sqrt(0.7))).
    Session check values: max|H♢-1/H| = 5.87e-6, max|V-V_cf| = 1.7e-8."""
    n, T = 200, 2.0
    tt = np.linspace(0, T, n)
    beta, mu, a, v = 0.8, 0.2, 0.7, 0.1
    B = np.array([[-beta]])
    aux = O.const_aux(B, -B @ np.array([mu]), [[a]])
    Hd, Vt = oracle_ref.backward_HV(tt, aux, [v])
    lam = a / (2 * beta)
    Hcf = np.array([np.exp(2 * beta * (T - t)) * lam - lam for t in tt])  # phim*lam*phim' - lam
    Vcf = np.array([np.exp(beta * (T - t)) * (v - mu) + mu for t in tt])
    eH = np.max(np.abs(Hd[:, 0, 0] - Hcf))
    eV = np.max(np.abs(Vt[:, 0] - Vcf))
    assert eH < 1e-5 and eV < 1e-5
    assert abs(eH - 5.87e-6) < 0.05e-6
    assert abs(eV - 1.7e-8) < 0.1e-8


# --------------------------------------------------------------------------- partial bridges (test/partialparam.jl)
T_PB = 1.5
TT_PB = np.arange(1501) * (1 / 1000)
X0_PB = np.array([2.0, 1.0])
L_PB = np.array([[1.0, 0.0]])
SIG_PB = np.array([[0.1]])
V_PB = np.array([2.5])
GAMMA = 0.7
EPS_PB = 1e-5
AUX_PB = dict(B=np.array([[0.0, 1.0], [0.0, -1.0]]), beta=np.array([0.0, 0.5]),
              a=np.array([[0.0, 0.0], [0.0, GAMMA ** 2]]))


def test_partialbridge_finite_differences(oracle_ref):
    """test/partialbridge.jl:73-74 (j = 10, 1-based)."""
    aux = O.const_aux(**AUX_PB)
    Lt, Mt, mut = oracle_ref.backward_LMmu(TT_PB, aux, L_PB, SIG_PB)
    j, dt = 9, 1 / 1000  # 0-based index of Julia's j=10
    lhs = (mut[j + 1] - mut[j]) / dt
    assert np.linalg.norm(lhs - (-Lt[j + 1] @ AUX_PB["beta"])) < 0.01
    lhs = (np.linalg.inv(Mt[j + 1]) - np.linalg.inv(Mt[j])) / dt
    assert np.linalg.norm(lhs - (-Lt[j + 1] @ AUX_PB["a"] @ Lt[j + 1].T)) < 0.01
    assert np.array_equal(Lt[-1], L_PB) and np.allclose(Mt[-1], np.linalg.inv(SIG_PB))


def test_partialbridgenuH_consistency(oracle_ref):
    """test/partialbridgenuH.jl:108-127 (lines 243-262 of the file): ν/H vs F/H parametrisation and
    the LP ~ LP2 tie to the Gaussian transition density, with the session-derived check values."""
    aux = O.const_aux(**AUX_PB)
    nuT, HpT, C_ = oracle_ref.update_nuHC(L_PB, SIG_PB, V_PB, EPS_PB)
    F0, H0, C0 = oracle_ref.update_FHC(L_PB, SIG_PB, V_PB, np.zeros(2), np.zeros((2, 2)), EPS_PB)
    assert np.isclose(C0, C_)                       # @test C ≈ C_
    assert np.allclose(F0, H0 @ nuT)                # @test F ≈ H*ν
    assert np.allclose(HpT, np.linalg.inv(H0))      # @test H⁺ ≈ inv(H)

    nu, H, _, _, C = oracle_ref.backward_nuH(O.ODE_R3, TT_PB, aux, nuT, HpT, C_)
    Ft, Ht, C2 = oracle_ref.backward_FH(TT_PB, aux, F0, H0, C0)
    assert abs(C2 - C) < 0.03
    relH = max(np.linalg.norm(Ht[i] - H[i]) for i in range(len(TT_PB))) / np.linalg.norm(Ht.reshape(len(TT_PB), -1), axis=1).max()
    # reference: maximum(norm.(Ht .- Po2.H)./norm(Ht)) with norm(Ht) the norm of the vector of matrices
    normHt = np.sqrt(sum(np.linalg.norm(Ht[i]) ** 2 for i in range(len(TT_PB))))
    relH = max(np.linalg.norm(Ht[i] - H[i]) for i in range(len(TT_PB))) / normHt
    assert relH < 1e-5
    dF = max(np.linalg.norm(H[i] @ nu[i] - Ft[i]) for i in range(len(TT_PB)))
    assert dF < 0.015

    # LP from the (L, M, mu) backward solve: test/partialbridge.jl:87
    Lt, Mt, mut = oracle_ref.backward_LMmu(TT_PB, aux, L_PB, SIG_PB)
    mean = mut[0][0] + (Lt[0] @ X0_PB)[0]
    std = Mt[0][0, 0] ** -0.5
    LP = -0.5 * ((V_PB[0] - mean) / std) ** 2 - np.log(std) - 0.5 * np.log(2 * np.pi)
    LP2 = -0.5 * (X0_PB @ H[0] @ X0_PB - 2 * X0_PB @ H[0] @ nu[0]) - C
    assert abs(LP - LP2) < 0.01
    # session-derived anchors
    assert abs(LP - (-0.99293614)) < 1e-6
    assert abs(LP2 - (-0.98368522)) < 1e-6
    assert abs(C - 7.78123371) < 1e-6
    assert abs(abs(C - C2) - 0.02452) < 1e-4
    assert abs(dF - 0.01105) < 1e-4
    assert abs(np.linalg.cond(H[0]) - 1.6857e7) / 1.6857e7 < 1e-3


def test_lyap_step_preserves_psd(oracle_ref):
    """test/lyap.jl:9-24: the Lyapunov backward step keeps H⁺ positive definite (d=20, t=2:0.01:5)."""
    d = 20
    tt = np.arange(2.0, 5.0 + 1e-12, 0.01)
    rng = np.random.default_rng(4)
    # the reference draws a fresh random B, sigma at EVERY call of B(t,P), sigma(t,P)
    N = tt.size
    Bs = rng.random((N - 1, 3, d, d))
    sig = rng.random((N - 1, 3, d, 10))
    As = np.einsum("iskl,isml->iskm", sig, sig)
    sl = rng.random((N - 1, d, 10))
    al = np.einsum("ikl,iml->ikm", sl, sl)
    aux = O.AuxHolder(d, Bs, np.zeros((N - 1, 3, d)), As, al, is_const=False)
    nu, H, _, Hl, _ = oracle_ref.backward_nuH(O.ODE_LYAP, tt, aux, np.zeros(d), np.eye(d))
    for i in range(N):
        Hp = np.linalg.inv(H[i])
        Hp = 0.5 * (Hp + Hp.T)
        assert np.all(np.linalg.eigvalsh(Hp) > 0)


def test_lyap_agrees_with_r3(oracle_ref):
    """The two backward schemes of src/partialbridgenuH.jl (R3 :21-55, Lyap :86-103) solve the same ODE."""
    aux = O.const_aux(**AUX_PB)
    nuT, HpT, C_ = oracle_ref.update_nuHC(L_PB, SIG_PB, V_PB, 1e-3)
    nu1, H1, _, Hl1, C1 = oracle_ref.backward_nuH(O.ODE_R3, TT_PB, aux, nuT, HpT, C_)
    nu2, H2, _, Hl2, C2 = oracle_ref.backward_nuH(O.ODE_LYAP, TT_PB, aux, nuT, HpT, C_)
    assert np.max(np.abs(nu1 - nu2)) < 1e-12
    assert np.linalg.norm(Hl1 - Hl2) / np.linalg.norm(Hl1) < 1e-5
    assert abs(C1 - C2) < 5e-3


def test_gpupdate_matches_direct_formula(oracle_ref):
    """src/guip.jl:221-243 / bolus3.jl:128-137 against numpy."""
    rng = np.random.default_rng(0)
    A = rng.standard_normal((3, 3)); Hp = A @ A.T + np.eye(3)
    nu = rng.standard_normal(3)
    L = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.5]]); Sig = np.diag([0.1, 0.2]); v = np.array([0.3, -0.2])
    nu2, Hp2 = oracle_ref.gpupdate_nuH(nu, Hp, L, Sig, v)
    Z = np.eye(3) - Hp @ L.T @ np.linalg.inv(Sig + L @ Hp @ L.T) @ L
    assert np.allclose(Hp2, Z @ Hp, rtol=1e-12, atol=1e-14)
    assert np.allclose(nu2, Z @ Hp @ L.T @ np.linalg.inv(Sig) @ v + Z @ nu, rtol=1e-12, atol=1e-14)
    # Kalman form: posterior precision = prior precision + L' Sig^-1 L
    assert np.allclose(np.linalg.inv(Hp2), np.linalg.inv(Hp) + L.T @ np.linalg.inv(Sig) @ L)
    # infinite prior variance branch
    Hinf = np.diag([np.inf] * 2)
    nu3, Hp3 = oracle_ref.gpupdate_nuH(np.zeros(2), Hinf, np.eye(2), 0.5 * np.eye(2), np.array([1.0, 2.0]))
    assert np.allclose(Hp3, 0.5 * np.eye(2)) and np.allclose(nu3, [1.0, 2.0])


# --------------------------------------------------------------------------- forward guided path + ll
def nuH_guide(orc, tt, eps=EPS_PB):
    aux = O.const_aux(**AUX_PB)
    nuT, HpT, C_ = orc.update_nuHC(L_PB, SIG_PB, V_PB, eps)
    nu, H, _, _, C = orc.backward_nuH(O.ODE_R3, tt, aux, nuT, HpT, C_)
    return O.GuideHolder(O.GUIDE_NUH, tt, H, nu, Bt=AUX_PB["B"], betat=AUX_PB["beta"])


def test_guided_euler_and_ll_match_numpy(oracle_ref):
    """src/euler.jl:262-267 with the drift of src/partialbridgenuH.jl:157-159 and the sum of :171-189."""
    tt = TT_PB[:301]
    P = O.make_model(O.INTDIFF, 2, 1, [GAMMA])
    G = nuH_guide(oracle_ref, tt)
    W = oracle_ref.wiener_sample(tt, 1, 11, 0, 0)
    X, xend = oracle_ref.guided_euler(P, G, X0_PB, W)
    ll = oracle_ref.llikelihood(P, G, X)
    y = X0_PB.copy(); som = 0.0
    a = AUX_PB["a"]
    for i in range(len(tt) - 1):
        assert np.allclose(X[i], y, rtol=1e-13, atol=1e-13)
        b = np.array([y[1], -(y[1] + np.sin(y[1])) + 0.5])
        r = G.A[i] @ (G.b[i] - y)
        bt = AUX_PB["B"] @ y + AUX_PB["beta"]
        dt = tt[i + 1] - tt[i]
        som += np.dot(b - bt, r) * dt
        y = y + (b + a @ r) * dt + np.array([0.0, GAMMA]) * (W[i + 1, 0] - W[i, 0])
    assert np.allclose(xend, y, rtol=1e-12, atol=1e-12)
    assert abs(ll - som) < 1e-9 * max(1.0, abs(som))
    # skip drops the last terms
    assert oracle_ref.llikelihood(P, G, X, skip=5) != ll


def test_guided_bridge_hits_observation(oracle_ref):
    """With a small observation variance the guided path ends near L x = v (the point of the construction)."""
    tt = TT_PB
    P = O.make_model(O.INTDIFF, 2, 1, [GAMMA])
    aux = O.const_aux(**AUX_PB)
    nuT, HpT, C_ = oracle_ref.update_nuHC(L_PB, np.array([[1e-6]]), V_PB, 1e-3)
    nu, H, _, _, _ = oracle_ref.backward_nuH(O.ODE_R3, tt, aux, nuT, HpT, C_)
    G = O.GuideHolder(O.GUIDE_NUH, tt, H, nu, Bt=AUX_PB["B"], betat=AUX_PB["beta"])
    ends = []
    for row in range(20):
        W = oracle_ref.wiener_sample(tt, 1, 3, 0, row)
        ends.append(oracle_ref.guided_euler(P, G, X0_PB, W)[1][0])
    assert np.max(np.abs(np.array(ends) - V_PB[0])) < 0.05


def test_partialbridge_and_nuH_proposals_agree(oracle_ref):
    """PartialBridge (L,M,mu) and PartialBridgeνH describe the same guided drift when eps -> 0
    (test/partialbridgenuH.jl:276 `norm(Xo1.yy - Xo2.yy) < sqrt(eps())` is for the in-place twin;
    here the (L,M,mu) and (nu,H) forms are compared, tolerance set by eps=1e-5 regularisation)."""
    tt = TT_PB
    P = O.make_model(O.INTDIFF, 2, 1, [GAMMA])
    aux = O.const_aux(**AUX_PB)
    Lt, Mt, mut = oracle_ref.backward_LMmu(tt, aux, L_PB, SIG_PB)
    G1 = O.GuideHolder(O.GUIDE_LMMU, tt, Lt, mut, Mm=Mt, v=V_PB, Bt=AUX_PB["B"], betat=AUX_PB["beta"], m=1)
    G2 = nuH_guide(oracle_ref, tt)
    W = oracle_ref.wiener_sample(tt, 1, 1, 0, 0)
    X1, _ = oracle_ref.guided_euler(P, G1, X0_PB, W)
    X2, _ = oracle_ref.guided_euler(P, G2, X0_PB, W)
    assert np.max(np.abs(X1 - X2)) < 2e-3
    ll1 = oracle_ref.llikelihood(P, G1, X1); ll2 = oracle_ref.llikelihood(P, G2, X2)
    assert abs(ll1 - ll2) < 2e-3 * max(1.0, abs(ll1))


def test_guidedbridge_importance_weights_unbiased(oracle_ref):
    """test/guip.jl:245-274 (GuidedBridge z-score): E[exp(ll) p~/p] = 1.  Target and auxiliary are
    both LinPro so that p and p~ are closed-form Gaussians; a fine tau-warped grid keeps the
    discretisation bias (SURVEY 8c caveat) below the Monte-Carlo error."""
    T, n, m = 2.0, 1000, 3000
    s = np.linspace(0, T, n)
    tt = s * (2 - s / T)
    a, u, v = 0.7, 0.5, 0.1
    Bt_, mu_t = -0.8, 0.2      # auxiliary LinPro(-0.8, 0.2, sqrt(a))
    Bp_, mu_p = -0.3, -0.1     # target
    P = O.linpro_model([[Bp_]], [mu_p], [[np.sqrt(a)]])
    aux = O.const_aux([[Bt_]], [-Bt_ * mu_t], [[a]])
    Hd, Vt = oracle_ref.backward_HV(tt, aux, [v])
    G = O.GuideHolder(O.GUIDE_HV, tt, Hd, Vt, Bt=[[Bt_]], betat=[-Bt_ * mu_t])

    def logp(B, mu):
        mean = np.exp(B * T) * (u - mu) + mu
        var = a / (2 * -B) * (1 - np.exp(2 * B * T))
        return -0.5 * (v - mean) ** 2 / var - 0.5 * np.log(2 * np.pi * var)

    w = np.empty(m)
    for k in range(m):
        W = oracle_ref.wiener_sample(tt, 1, 99, 0, k)
        X, xend = oracle_ref.guided_euler(P, G, [u], W)
        assert xend[0] == v  # endpoint(y, P::GuidedBridge) = V[end] when H♢[end] = 0  (src/euler.jl:241-242)
        w[k] = np.exp(oracle_ref.llikelihood(P, G, X) + logp(Bt_, mu_t) - logp(Bp_, mu_p))
    z = abs(w.mean() - 1) * np.sqrt(m) / w.std()
    assert z < 3.5, (w.mean(), z)


def test_partialbridge_nuH_importance_weights_unbiased(oracle_ref):
    """E[exp(ll) p~/p] = 1 for PartialBridgeνH (the z-score check of test/guip.jl:245-274 carried over to the proposal
    north_star names), with p~ = exp(lptilde(x0, P°)) in the form of test/partialbridgenuH.jl:124.  Target and
    auxiliary are 2-d LinPro processes with a PARTIAL observation v = L X_T + N(0, Σ), L = [1 0], so that
    p = N(v; L m_T, L K_T L' + Σ) is a closed-form Gaussian (src/linpro.jl:98-134).  ϵ = 1e-5 (the reference tests'
    value, test/partialparam.jl) keeps the extra factor exp(-ϵ|x_T|²/2) of the regularised terminal condition
    (src/partialbridgenuH.jl:4) at 1 - O(1e-5); a much smaller ϵ makes H⁺ = O(1/ϵ) too stiff for the explicit R3 step
    (lptilde off by 0.1 at ϵ = 1e-9, N = 801 -- the conditioning issue behind test/partialbridgenuH.jl:121's @test_broken)."""
    from scipy.linalg import expm, solve_continuous_lyapunov
    T, n, m = 1.0, 801, 4000
    s = np.linspace(0, T, n)
    tt = s * (2 - s / T)
    sig = np.array([[0.6, 0.0], [0.2, 0.5]])
    a = sig @ sig.T
    Bp = np.array([[-0.5, 0.4], [-0.3, -0.8]]); mup = np.array([0.2, -0.1])    # target
    Bt = np.array([[-0.9, 0.1], [0.0, -0.6]]); mut = np.array([0.0, 0.1])      # auxiliary
    L = np.array([[1.0, 0.0]]); Sig = np.array([[0.05]]); v = np.array([0.4]); eps = 1e-5
    x0 = np.array([0.3, -0.2])
    P = O.linpro_model(Bp, mup, sig)
    aux = O.const_aux(Bt, -Bt @ mut, a)
    nuT, HpT, C0 = oracle_ref.update_nuHC(L, Sig, v, eps)
    nu, H, _, _, Cc = oracle_ref.backward_nuH(O.ODE_R3, tt, aux, nuT, HpT, C0)
    G = O.GuideHolder(O.GUIDE_NUH, tt, H, nu, Bt=Bt, betat=-Bt @ mut)

    def logp(B, mu):
        Phi = expm(B * T)
        Kinf = solve_continuous_lyapunov(B, -a)
        K = Kinf - Phi @ Kinf @ Phi.T
        mean = (L @ (Phi @ (x0 - mu) + mu))[0]
        var = (L @ K @ L.T + Sig)[0, 0]
        return -0.5 * (v[0] - mean) ** 2 / var - 0.5 * np.log(2 * np.pi * var)

    lpt = oracle_ref.lptilde_nuH(nu[0], H[0], Cc, x0)
    # lptilde equals the closed-form log-density of the auxiliary process up to the discretisation of C (a first-order
    # sum, src/partialbridgenuH.jl:46: 2.2e-3 here, 2.2e-4 at N = 8001); test/partialbridgenuH.jl:127 asks |LP - LP2| < 0.01
    assert abs(lpt - logp(Bt, mut)) < 5e-3, (lpt, logp(Bt, mut))
    w = np.empty(m)
    for k in range(m):
        W = oracle_ref.wiener_sample(tt, 2, 77, 0, k)
        X, _ = oracle_ref.guided_euler(P, G, x0, W)
        w[k] = np.exp(oracle_ref.llikelihood(P, G, X) + lpt - logp(Bp, mup))
    z = abs(w.mean() - 1) * np.sqrt(m) / w.std()
    assert z < 3.5, (w.mean(), w.std(), z)


def test_lptilde_guidedbridge_matches_closed_form(oracle_ref):
    """test/VHK.jl:63-65: lptilde(GP, u) equals lp(t, u, T, v, Pt) of the linear auxiliary process to 1e-5
    (1-d LinPro(-β, μ, σ), the setup of test/VHK.jl:12-27 with N = 200)."""
    β, μ, σ, T, u, v = 0.8, 0.2, np.sqrt(0.7), 2.0, 0.5, 0.1
    tt = np.linspace(0.0, T, 2001)
    Hd, V = oracle_ref.backward_HV(tt, O.const_aux([[-β]], [β * μ], [[σ * σ]]), [v])
    mean = np.exp(-β * T) * (u - μ) + μ
    var = σ * σ / (2 * β) * (1 - np.exp(-2 * β * T))
    lp = -0.5 * (v - mean) ** 2 / var - 0.5 * np.log(2 * np.pi * var)
    assert abs(oracle_ref.lptilde_HV(tt, -β, V[0], Hd[0], [u]) - lp) < 1e-5
    # time-dependent trace through the staged values: the same constant, tabulated
    tr = np.full((len(tt) - 1, 3), -β)
    assert abs(oracle_ref.lptilde_HV(tt, tr, V[0], Hd[0], [u]) - lp) < 1e-5


# --------------------------------------------------------------------------- pCN
def test_pcn_smoke_acceptance(oracle_ref):
    """test/partialbridge.jl:133 `1 < acc < iterations` on the partialparam setup (rho = 0.9)."""
    tt = TT_PB
    P = O.make_model(O.INTDIFF, 2, 1, [GAMMA])
    G = nuH_guide(oracle_ref, tt)
    iters = 300
    W = oracle_ref.wiener_sample(tt, 1, 1, 12345, 0)
    X, _ = oracle_ref.guided_euler(P, G, X0_PB, W)
    ll = oracle_ref.llikelihood(P, G, X)
    acc = 0
    for it in range(iters):
        llo, logu, Wo, Xo, _ = oracle_ref.pcn_propose(P, [G], X0_PB, W, 0.9, 1, it, 0)
        # the proposal is rho*W + sqrt(1-rho^2)*W2 with a fresh W2 (test/partialbridgenuH.jl:313-314)
        if it == 0:
            W2 = oracle_ref.wiener_sample(tt, 1, 1, 0, 0)
            assert np.allclose(Wo[0], 0.9 * W + np.sqrt(1 - 0.81) * W2, rtol=0, atol=1e-14)
        if logu <= llo - ll:
            W, ll, acc = Wo[0], llo, acc + 1
    assert 1 < acc < iters


def test_pcn_bench_driver_matches_stepwise(oracle_ref):
    """The OpenMP baseline driver is the same algorithm as the single-chain entry point."""
    tt = TT_PB[:201]
    P = O.make_model(O.INTDIFF, 2, 1, [GAMMA])
    G = nuH_guide(oracle_ref, tt)
    acc, secs, ll = oracle_ref.pcn_bench(P, [G], 4, X0_PB, 0.9, 7, 5, nthreads=2)
    # replay chain 2 by hand
    c = 2
    W = oracle_ref.wiener_sample(tt, 1, 7, 0xFFFFFFFE, c)
    X, _ = oracle_ref.guided_euler(P, G, X0_PB, W)
    l = oracle_ref.llikelihood(P, G, X)
    for it in range(5):
        llo, logu, Wo, _, _ = oracle_ref.pcn_propose(P, [G], X0_PB, W, 0.9, 7, it, c)
        if logu <= llo - l:
            W, l = Wo[0], llo
    assert l == ll[c]
    assert 0 <= acc <= 20 and secs >= 0


def test_mdb_step_properties(oracle_ref):
    """solve!(Mdb(), ...)  src/euler.jl:308-327: the noise of step i is scaled by sqrt((T - t[i+1]) / (T - t[i])), so the
    last step is deterministic, and with W = 0 the scheme is the guided Euler scheme."""
    tt = TT_PB[:201]
    P = O.make_model(O.INTDIFF, 2, 1, [GAMMA])
    G = nuH_guide(oracle_ref, tt)
    W = oracle_ref.wiener_sample(tt, 1, 5, 0, 0)
    X, xend = oracle_ref.guided_mdb(P, G, X0_PB, W)
    W2 = W.copy(); W2[-1] += 7.0                       # the last increment does not matter
    X2, _ = oracle_ref.guided_mdb(P, G, X0_PB, W2)
    assert np.array_equal(X, X2) and np.array_equal(xend, X[-1])
    Xe, _ = oracle_ref.guided_euler(P, G, X0_PB, W)
    assert not np.array_equal(X, Xe)
    Z = np.zeros_like(W)
    assert np.array_equal(oracle_ref.guided_mdb(P, G, X0_PB, Z)[0], oracle_ref.guided_euler(P, G, X0_PB, Z)[0])
    # one step by hand
    i = 57
    y = X[i]; r = G.A[i] @ (G.b[i] - y)
    b = np.array([y[1], -(y[1] + np.sin(y[1])) + 0.5]) + AUX_PB["a"] @ r
    s = np.sqrt((tt[-1] - tt[i + 1]) / (tt[-1] - tt[i]))
    want = y + b * (tt[i + 1] - tt[i]) + np.array([0.0, GAMMA * s]) * (W[i + 1, 0] - W[i, 0])
    assert np.allclose(X[i + 1], want, rtol=1e-13, atol=1e-15)


def test_tuned_baseline_driver_is_the_same_algorithm(oracle_ref, oracle_fma):
    """bench.py's CPU baseline runs a driver specialised for the FitzHugh-Nagumo / PartialBridgeνH workload (inlined
    2-d arithmetic, batched normals, every Philox call fully used).  In the contraction-free builds it must reproduce the
    generic restatement BIT FOR BIT: same acceptance count, same final log-likelihoods."""
    T = 0.5
    s = np.linspace(0, T, 97)
    P = O.make_model(O.FHN_HYPO, 2, 1, [0.1, 0.0, 1.5, 0.8, 0.3])
    for orc in (oracle_ref, oracle_fma):
        guides = []
        nu, Hp = np.zeros(2), np.eye(2) / 1e-3
        nu, Hp = orc.gpupdate_nuH(nu, Hp, [[1.0, 0.0]], [[1e-4]], [-0.5])
        for t0, v in ((0.5, -0.5), (0.0, -1.0)):
            tt = t0 + s * (2 - s / T)
            Bt = np.array([[10.0, -10.0], [1.5, -1.0]]); bet = np.array([0.0 - v ** 3 / 0.1, 0.8])
            nut, Ht, nu, Hp, _ = orc.backward_nuH(O.ODE_LYAP, tt, O.const_aux(Bt, bet, [[0.0, 0.0], [0.0, 0.09]]), nu, Hp)
            guides.insert(0, O.GuideHolder(O.GUIDE_NUH, tt, Ht, nut, Bt=Bt, betat=bet))
            nu, Hp = orc.gpupdate_nuH(nu, Hp, [[1.0, 0.0]], [[1e-4]], [-1.0])
        a1, _, l1 = orc.pcn_bench(P, guides, 24, [-0.5, -0.6], 0.95, 11, 5, nthreads=2)
        a2, _, l2 = orc.pcn_bench_fhn_tuned(P, guides, 24, [-0.5, -0.6], 0.95, 11, 5, nthreads=2)
        assert a1 == a2 and np.array_equal(l1, l2) and 0 < a1 < 24 * 5


# --------------------------------------------------------------------------- the two oracle builds
def test_ref_and_fma_builds_agree_to_rounding(oracle_ref, oracle_fma):
    """Contraction sensitivity of the FHN hypoelliptic bridge: reference arithmetic vs the kernels'
    FMA order on identical W.  This is the gap the GPU-vs-reference tolerance has to cover."""
    T = 0.5
    s = np.linspace(0, T, 501)
    tt = s * (2 - s / T)
    P = O.make_model(O.FHN_HYPO, 2, 1, [0.1, 0.0, 1.5, 0.8, 0.3])
    v = -1.0
    Bt = np.array([[10.0, -10.0], [1.5, -1.0]]); bet = np.array([0.0 - v ** 3 / 0.1, 0.8])
    at = np.array([[0.0, 0.0], [0.0, 0.09]])
    aux = O.const_aux(Bt, bet, at)
    out = []
    for orc in (oracle_ref, oracle_fma):
        nuT, HpT, C_ = orc.update_nuHC([[1.0, 0.0]], [[1e-10]], [v], 1e-3)
        nu, H, _, _, _ = orc.backward_nuH(O.ODE_R3, tt, aux, nuT, HpT, C_)
        G = O.GuideHolder(O.GUIDE_NUH, tt, H, nu, Bt=Bt, betat=bet)
        W = oracle_ref.wiener_sample(tt, 1, 4, 0, 0)
        X, _ = orc.guided_euler(P, G, [-0.5, -0.6], W)
        out.append((X, orc.llikelihood(P, G, X)))
    (X1, l1), (X2, l2) = out
    assert np.max(np.abs(X1 - X2)) < 1e-9 * (1 + np.max(np.abs(X1)))
    assert abs(l1 - l2) < 1e-6 * abs(l1) + 1e-9
    assert abs(X1[-1, 0] - v) < 1e-3


def test_nonconstdiff_llikelihood_terms(oracle_ref):
    """src/partialbridge.jl:79-84: for a pair with a != a~ the log-likelihood gains
    -0.5 tr((a-a~)H) dt + 0.5 r'(a-a~)r dt with H = L'ML; restated here in numpy on the PartialBridge tables."""
    tt = TT_PB[:151]
    P = O.make_model(O.INTDIFF, 2, 1, [GAMMA])
    at = np.array([[0.05, 0.0], [0.0, 0.6 * GAMMA ** 2]])      # auxiliary diffusion differs from the target's
    aux = O.const_aux(AUX_PB["B"], AUX_PB["beta"], at)
    Lt, Mt, mut = oracle_ref.backward_LMmu(tt, aux, L_PB, SIG_PB)
    Ad = AUX_PB["a"] - at
    G0 = O.GuideHolder(O.GUIDE_LMMU, tt, Lt, mut, Mm=Mt, v=V_PB, Bt=AUX_PB["B"], betat=AUX_PB["beta"], m=1)
    G1 = O.GuideHolder(O.GUIDE_LMMU, tt, Lt, mut, Mm=Mt, v=V_PB, Bt=AUX_PB["B"], betat=AUX_PB["beta"], m=1, Adiff=Ad)
    W = oracle_ref.wiener_sample(tt, 1, 2, 0, 0)
    X, _ = oracle_ref.guided_euler(P, G1, X0_PB, W)
    ll0, ll1 = oracle_ref.llikelihood(P, G0, X), oracle_ref.llikelihood(P, G1, X)
    extra = 0.0
    for i in range(len(tt) - 1):
        dt = tt[i + 1] - tt[i]
        H = Lt[i].T @ Mt[i] @ Lt[i]
        r = Lt[i].T @ Mt[i] @ (V_PB - mut[i] - Lt[i] @ X[i])
        extra += -0.5 * np.trace(Ad @ H) * dt + 0.5 * (r @ Ad @ r) * dt
    assert abs(extra) > 1e-3
    assert abs((ll1 - ll0) - extra) < 1e-10 * max(1.0, abs(extra))


def test_stochastic_heun_linear_closed_form(oracle_ref):
    """solve!(StochasticHeun(), ...) (src/euler.jl:178-198) for b = -βx, σ const: one step is
    y (1 - β dt + β² dt²/2) + σ dw; the loop stops at N-2 and the last grid point is not written."""
    beta, sig = 2.0, 0.7
    m = O.make_model(O.OU, 1, 1, [beta, sig])
    tt = np.linspace(0, 1, 41) ** 1.5
    W = oracle_ref.wiener_sample(tt, 1, 9, 0, 3)
    X0 = np.full((41, 1), 123.0)
    X = oracle_ref.heun(m, tt, [0.4], W, X0)
    y = 0.4
    want = np.empty(40)
    for i in range(39):
        want[i] = y
        dt = tt[i + 1] - tt[i]
        y = y * (1 - beta * dt + 0.5 * beta * beta * dt * dt) + sig * (W[i + 1, 0] - W[i, 0])
    want[39] = y
    assert np.allclose(X[:40, 0], want, rtol=1e-13, atol=1e-15)
    assert X[40, 0] == 123.0


def test_kernel_exponential_restatement_accuracy(oracle_fma, oracle_ref):
    """bb_exp of the GPU-order build (the landmarks kernel's exponential: +, *, fma, rint only) against libm:
    relative error < 3e-16 on the argument range of the Gaussian kernel; the reference build IS libm exp."""
    import math
    xs = -np.concatenate([np.linspace(0.0, 6.0, 4001), np.logspace(-14, 2.84, 4000)])
    err = max(abs(oracle_fma.lib.bbo_exp(float(x)) - math.exp(x)) / math.exp(x) for x in xs)
    assert err < 3e-16
    assert oracle_fma.lib.bbo_exp(-0.0) == 1.0 and oracle_fma.lib.bbo_exp(-750.0) == 0.0
    assert all(oracle_ref.lib.bbo_exp(float(x)) == math.exp(x) for x in xs[::97])


def test_landmarks_drift_restatement(oracle_ref, oracle_fma):
    """project_partialbridge/partialbridge_landmarks.jl:47,90-101: the oracle's drift against an independent NumPy
    transcription; with the positions frozen at qT the position part of the drift is B~ x of :129-138 (the auxiliary
    process is the drift's linearisation there)."""
    a, sig, lam = 0.5, 2.0, 0.5
    om = O.make_model(O.LANDMARKS, 16, 8, [a, sig, lam])

    def kern(x):
        return np.exp(-np.dot(x, x) / (2 * a)) / (2 * np.pi * a)

    def drift(x):
        x = x.reshape(4, 2, 2)
        out = np.zeros_like(x)
        for i in range(4):
            for j in range(4):
                k = kern(x[i, 0] - x[j, 0])
                out[i, 0] += 0.5 * x[j, 1] * k
                out[i, 1] += -lam * 0.5 * x[j, 1] * k + 1 / (2 * a) * np.dot(x[i, 1], x[j, 1]) * (x[i, 0] - x[j, 0]) * k
        return out.ravel()

    rng = np.random.default_rng(3)
    for orc, tol in ((oracle_ref, 1e-14), (oracle_fma, 1e-14)):
        for _ in range(4):
            x = rng.standard_normal(16)
            b = np.zeros(16)
            orc.lib.bbo_model_b(O.C.byref(om), O.C.c_double(0.0), O._p(x), O._p(b))
            assert np.allclose(b, drift(x), rtol=tol, atol=1e-15)
    qT = rng.standard_normal((4, 2))
    Bt = np.zeros((16, 16))
    for i in range(4):
        for j in range(4):
            k = kern(qT[i] - qT[j])
            for c in range(2):
                Bt[4 * i + c, 4 * j + 2 + c] += 0.5 * k
                Bt[4 * i + 2 + c, 4 * j + 2 + c] += -0.5 * lam * k
    x = rng.standard_normal((4, 2, 2)); x[:, 0] = qT
    assert np.allclose((Bt @ x.ravel()).reshape(4, 2, 2)[:, 0], drift(x.ravel()).reshape(4, 2, 2)[:, 0], rtol=1e-13)
    # noise acts on the momenta only: a = sigma^2 on the p components
    A = np.zeros((16, 16))
    oracle_ref.lib.bbo_model_a(O.C.byref(om), O._p(A))
    want = np.zeros(16); want[[2, 3, 6, 7, 10, 11, 14, 15]] = sig * sig
    assert np.array_equal(A, np.diag(want))


def test_theta_oracle_composition_matches_the_backward_chain(oracle_ref):
    """oracle.theta_backward (per-chain parameter path) is the chain of partialbridge_bolus3.jl:162-180 that the shared
    tables of config 4 are built with; with θ = the model's parameters both give the same tables."""
    eps_, sdiag = 1e-3, 1e-10
    par = (0.1, 0.0, 1.5, 0.8, 0.3)
    L = np.array([[1.0, 0.0]]); Sig = np.array([[sdiag]])
    obs_t, obs_v = (0.5, 1.0, 1.5), (-1.0, -0.5, 0.5)

    def tau(t0, t1, n):
        s = np.linspace(0.0, t1 - t0, n)
        return t0 + s * (2.0 - s / (t1 - t0))

    grids = [tau(a, b, 41) for a, b in zip((0.0,) + obs_t[:-1], obs_t)]
    guides, left = O.theta_backward(oracle_ref, O.FHN_HYPO, par, grids, [-0.5, -0.6], L, Sig, eps_, obs_v,
                                    O.AUX_FHN_MATCHING, {4: ("gamma", 1.0, 100.0)})
    nu = np.zeros(2); Hp = np.eye(2) / eps_
    nu, Hp = oracle_ref.gpupdate_nuH(nu, Hp, L, Sig, [obs_v[-1]])
    for s in range(2, -1, -1):
        Bt, bt = O.fhn_aux(O.AUX_FHN_MATCHING, par, obs_v[s])
        nus, Hs, nu, Hp, _ = oracle_ref.backward_nuH(O.ODE_LYAP, grids[s], O.const_aux(Bt, bt, O.fhn_a(O.FHN_HYPO, par)),
                                                     nu, Hp, 0.0)
        assert np.array_equal(guides[s].A, Hs) and np.array_equal(guides[s].b, nus)
        if s > 0:
            nu, Hp = oracle_ref.gpupdate_nuH(nu, Hp, L, Sig, [obs_v[s - 1]])
    assert np.array_equal(left["nu"], nu) and np.array_equal(left["Hp"], Hp)
    assert left["lpn"] == oracle_ref.logpdfnormal(np.array([-0.5, -0.6]) - nu, Hp)
    assert abs(left["lpri"] - (-np.log(100.0) - 0.3 / 100.0)) < 1e-15
    assert abs(left["trsum"] - 1.5 * (1 / 0.1 - 1.0)) < 1e-12


def test_logpdfnormal_and_gamma_prior_against_scipy(oracle_ref):
    """logpdfnormal(x, Σ) (src/gaussian.jl:66-75) against scipy's multivariate normal; the Gamma(shape, scale) log-density of
    logπ (bolus3.jl:237, Distributions.Gamma) as restated in oracle.theta_backward against scipy.stats.gamma."""
    from scipy.stats import gamma, multivariate_normal
    rng = np.random.default_rng(7)
    for d in (1, 2, 3):
        A = rng.standard_normal((d, d))
        S = A @ A.T + 0.3 * np.eye(d)
        x = rng.standard_normal(d)
        want = multivariate_normal(mean=np.zeros(d), cov=S).logpdf(x)
        assert abs(oracle_ref.logpdfnormal(x, S) - want) <= 1e-12 * (1 + abs(want))
    # symmetrisation: an asymmetric input is used as (Σ + Σ')/2  (Bridge.symmetrize, bolus3.jl:319)
    S = np.array([[2.0, 0.3], [0.1, 1.0]])
    x = np.array([0.4, -0.7])
    assert abs(oracle_ref.logpdfnormal(x, S) - multivariate_normal(cov=0.5 * (S + S.T)).logpdf(x)) < 1e-12
    grids = [np.linspace(0.0, 0.5, 9)]
    for (a, b, xval) in ((1.0, 100.0, 0.3), (2.0, 50.0, 0.7), (3.5, 2.0, 1.9)):
        par = [0.1, 0.0, 1.5, 0.8, xval]
        _, left = O.theta_backward(oracle_ref, O.FHN_HYPO, par, grids, [-0.5, -0.6], np.array([[1.0, 0.0]]),
                                   np.array([[1e-2]]), 0.1, [0.5], O.AUX_FHN_MATCHING, {4: ("gamma", a, b)})
        assert abs(left["lpri"] - gamma(a, scale=b).logpdf(xval)) < 1e-12
    _, left = O.theta_backward(oracle_ref, O.FHN_HYPO, [0.1, 0.0, 1.5, 0.8, -0.2], grids, [-0.5, -0.6],
                               np.array([[1.0, 0.0]]), np.array([[1e-2]]), 0.1, [0.5], O.AUX_FHN_MATCHING,
                               {4: ("gamma", 1.0, 100.0)})
    assert left["lpri"] == -np.inf  # outside the support: never accepted


def test_block_update_composition_properties(oracle_ref):
    """oracle.theta_block_step (blocked segment updates, partialbridge_bolus3.jl:258-355):
      * a block that spans the whole chain uses the tables of the full backward chain (theta_backward) and its start
        term is that chain's logpdfnormal;
      * with ρ = 1 and no start move the proposal IS the re-simulated current path: W° = W, diffll = 0 exactly;
      * a block that does not end the chain is conditioned on the chain's own path: ν at the block's right end is the
        path's value there and H⁺ = Hzero⁺, whatever the observations right of it are;
      * the proposal's noise is the pCN combination of one segment of bbo_pcn_propose (same rows, same normals)."""
    par = (0.1, 0.0, 1.5, 0.8, 0.3)
    L = np.array([[1.0, 0.0]]); Sig = np.array([[1e-2]]); eps_ = 0.1
    obs_t, obs_v = (0.5, 1.0, 1.5), (-1.0, -0.5, 0.5)

    def tau(t0, t1, n):
        s = np.linspace(0.0, t1 - t0, n)
        return t0 + s * (2.0 - s / (t1 - t0))

    S, n = 3, 31
    grids = [tau(a, b, n) for a, b in zip((0.0,) + obs_t[:-1], obs_t)]
    x0 = np.array([-0.5, -0.6])
    guides, left = O.theta_backward(oracle_ref, O.FHN_HYPO, par, grids, x0, L, Sig, eps_, obs_v, O.AUX_FHN_MATCHING)
    W = np.stack([oracle_ref.wiener_sample(g, 1, 3, 0, 10 + s) for s, g in enumerate(grids)])
    X, ll, _ = O.theta_forward(oracle_ref, O.FHN_HYPO, 1, par, guides, x0, W)
    args = (oracle_ref, O.FHN_HYPO, 1, par, grids, L, Sig, eps_, obs_v, O.AUX_FHN_MATCHING)
    # whole chain
    r = O.theta_block_step(*args, 0, S, 0.1, x0, X, W, 0.7, 5, 2, 77)
    for s in range(S):
        assert np.array_equal(r["guides"][s].A, guides[s].A) and np.array_equal(r["guides"][s].b, guides[s].b)
    assert r["lpn"] == left["lpn"] and r["lpno"] == r["lpn"]
    assert np.sum(r["llt"]) == pytest.approx(ll, rel=1e-14)  # XXtemp under the same guide is the current path
    mdl = O.make_model(O.FHN_HYPO, 2, 1, par)
    llo, lu, Wo, Xo, _ = oracle_ref.pcn_propose(mdl, guides, x0, W, 0.7, 5, 2, 77)
    for s in range(S):
        assert np.array_equal(r["Wo"][s], Wo[s]) and np.array_equal(r["Xo"][s], Xo[s])
    assert lu == r["logu"] and np.sum(r["llo"]) == pytest.approx(llo, rel=1e-14)
    # ρ = 1: nothing is proposed
    r1 = O.theta_block_step(*args, 1, 3, 0.1, x0, X, W, 1.0, 5, 3, 77)
    assert all(np.array_equal(r1["Wo"][s], W[s]) for s in (1, 2)) and r1["diff"] == 0.0
    # interior block: conditioned on the path's own value, observations to the right do not enter
    r2 = O.theta_block_step(*args, 0, 2, 0.1, x0, X, W, 0.7, 5, 4, 77, 0.1, [1.0, -1.0])
    g2, nuL, HpL = O.theta_block_backward(oracle_ref, O.FHN_HYPO, par, grids, L, Sig, eps_, (-1.0, -0.5, 123.0),
                                          O.AUX_FHN_MATCHING, 0, 2, X[1][-1], 0.1)
    assert np.array_equal(g2[1].b[-1], X[1][-1]) and np.array_equal(r2["guides"][1].b, g2[1].b)
    assert np.allclose(np.linalg.inv(g2[1].A[-1]), 0.1 * np.eye(2), rtol=1e-12)
    assert np.array_equal(r2["nuL"], nuL) and g2[2] is None
    # the start move: x0° = x0 + (0.1 u)(1, -1), and both start terms use the block's left-end (ν, H⁺)
    u = oracle_ref.normal(5, 4, 77, 4 * O.Q_THETA_NORMALS + 3)
    assert np.array_equal(r2["x0o"], x0 + (0.1 * u) * np.array([1.0, -1.0]))
    assert r2["lpno"] == oracle_ref.logpdfnormal(r2["x0o"] - nuL, HpL) and np.array_equal(r2["Xo"][0][0], r2["x0o"])


# --------------------------------------------------------------------------- online statistics (src/mclog.jl)
def test_online_statistics_known_answers():
    """test/onlinestat.jl:2-10: the running (mean, m2) of 10 random 5-vectors / 10 random scalars reproduce `mean` and
    `cov(x, corrected = true)` / `var` to eps(100.0).  `MeanCov`'s iteration (src/mclog.jl:157-168) and `mcnext!`
    (:47-56) are the same recurrence: delta = x - m; m += delta/(n+1); m2 += outer(delta, x - m).  Plus the worked
    example of the reference's own comment, src/mclog.jl:270-277 (1:10 repeated over 5 entries)."""
    rng = np.random.default_rng(0)
    x = rng.random((10, 5))
    mc = O.mcstart(x[0])
    for xi in x:
        mc = O.mcnext(mc, xi)
    m, cov = O.mcstats(mc)
    eps100 = np.spacing(100.0)
    assert mc[2] == 10
    assert np.linalg.norm(np.cov(x.T, ddof=1) - cov) < eps100
    assert np.linalg.norm(x.mean(axis=0) - m) < eps100
    y = rng.random(10)
    mc = O.mcstart(y[:1])
    for yi in y:
        mc = O.mcnext(mc, [yi])
    m, var = O.mcstats(mc)
    assert abs(np.var(y, ddof=1) - var[0, 0]) < eps100 and abs(y.mean() - m[0]) < eps100
    # S = OnlineStat(ones(5)); push!(S, i*ones(5)) for i in 2:10: mean 5.5, variance of 1:10
    mc = O.mcstart(np.ones((5, 1)))
    for i in range(1, 11):
        mc = O.mcnext(mc, i * np.ones((5, 1)))
    m, var = O.mcstats(mc)
    assert np.all(m == 5.5) and np.allclose(var, np.var(np.arange(1, 11), ddof=1), rtol=1e-15)
    lo, hi = O.mcband(mc)
    assert np.allclose(hi - m, O.MCBAND_Q * np.sqrt(var[..., 0]), rtol=1e-15) and np.allclose(m - lo, hi - m, rtol=1e-15)
    # Q = sqrt(2.) * erfinv(0.95) is the two-sided 95 % normal quantile
    from scipy.special import erfinv
    assert abs(O.MCBAND_Q - np.sqrt(2.0) * erfinv(0.95)) < 4e-16
    # one observation: m2/(k-1) = 0/0, as in the reference
    mc = O.mcnext(O.mcstart(np.ones(2)), np.ones(2))
    assert np.all(np.isnan(O.mcstats(mc)[1]))
