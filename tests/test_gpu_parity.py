"""Parity of the CUDA path (through the C ABI) with the CPU oracle on identical inputs.  GPU only.

Two comparisons per case:
  * against liboracle_fma.so (the oracle evaluated in the kernels' rounding order): BIT-EXACT
    (np.array_equal), except where libm's sin enters (INTDIFF / NCLAR3 drifts) or a device log;
  * against liboracle_ref.so (reference arithmetic: no fused multiply-add, true divisions):
    |dX| <= 1e-10 (1 + |X|),  |dll| <= 1e-6 |ll| + 1e-9   (north_star: ll within 1e-6 rel).
Accept/reject bookkeeping is compared exactly.
"""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

XTOL = 1e-10
LLREL, LLABS = 1e-6, 1e-9


@pytest.fixture(scope="module")
def B():
    import bridge_jl_b200 as B
    B.default_context()
    return B


def close_x(a, b):
    return np.max(np.abs(a - b)) <= XTOL * (1 + np.max(np.abs(b)))


def close_ll(a, b):
    return np.all(np.abs(a - b) <= LLREL * np.abs(b) + LLABS)


def warped(t0, t1, n):
    s = np.linspace(0, t1 - t0, n)
    return t0 + s * (2 - s / (t1 - t0))


# ----------------------------------------------------------------------------------------------- a4: sample!
@pytest.mark.parametrize("dp,N,S", [(1, 37, 2), (2, 64, 1), (3, 9, 3), (1, 5, 1)])
def test_wiener_sample_bit_exact(B, oracle_fma, dp, N, S):
    P = 70
    ens = B.PathEnsemble(P, S, N, dp, dp, double_buffer=False, chain_offset=1000)
    grids = [warped(0.3 * s, 0.3 * s + 1.0, N) for s in range(S)]
    for s, tt in enumerate(grids):
        ens.set_grid(s, tt)
    W0 = np.zeros((P, S, N, dp))
    W0[:, :, 0, :] = np.arange(P)[:, None, None] * 0.25  # y1 = W.yy[1] is kept (src/wiener.jl:50-51)
    ens.upload(B.W, W0)
    ens.sample_(seed=0x1234567890ABCDEF, stream=7)
    W = ens.download(B.W)
    for p in (0, 1, 31, 32, 69):
        for s in range(S):
            want = oracle_fma.wiener_sample(grids[s], dp, 0x1234567890ABCDEF, 7, (1000 + p) * S + s,
                                            y1=W0[p, s, 0])
            assert np.array_equal(W[p, s], want), (p, s)
    ens.close()


def test_upload_download_roundtrip(B):
    rng = np.random.default_rng(0)
    ens = B.PathEnsemble(33, 2, 21, 3, 2, double_buffer=True)
    Wh = rng.standard_normal((33, 2, 21, 2)); Xh = rng.standard_normal((33, 2, 21, 3))
    ens.upload(B.W, Wh); ens.upload(B.X, Xh)
    assert np.array_equal(ens.download(B.W), Wh) and np.array_equal(ens.download(B.X), Xh)
    assert np.array_equal(ens.download(B.W, p0=5, np_=7), Wh[5:12])
    ens.close()


# ----------------------------------------------------------------------------------------------- a5: Euler-Maruyama
GOLD_TT = np.arange(11) * 0.1
GOLD_W = np.array([0.0, 0.0940107, 0.214935, 0.0259463, 0.0226432, -0.24268, -0.144298, 0.581472, -0.135443,
                   0.0321464, 0.168574])
GOLD_X = np.array([0.1, -0.00598928, 0.126914, -0.315902, 0.312599, -0.577923, 0.676305, 0.0494658, -0.766381,
                   0.933971, -0.797544])


def test_docs_golden_vector_through_solve(B):
    """docs/src/manual.md:59-77: X = solve(Euler(), 0.1, W, OrnsteinUhlenbeck(20.0, 1.0))."""
    W = B.SamplePath(GOLD_TT, GOLD_W)
    X = B.solve(B.Euler(), 0.1, W, B.OrnsteinUhlenbeck(20.0, 1.0))
    assert X.yy.shape == (11,)
    assert np.max(np.abs(X.yy - GOLD_X)) < 1e-5
    assert np.array_equal(X.tt, GOLD_TT)


def test_config1_ou_single_path(B, oracle_ref, oracle_fma):
    """BASELINE config 1: 1-D OU, one path, n = 1001, tt = 0:0.01:10, u = 0.1 (plumbing of solve / solve!)."""
    tt = np.arange(1001) * 0.01
    for beta in (2.0, 20.0):
        Wy = oracle_ref.wiener_sample(tt, 1, 1, 0, 0)[:, 0]
        X = B.solve(B.EulerMaruyama(), 0.1, B.SamplePath(tt, Wy), B.OrnsteinUhlenbeck(beta, 1.0))
        m = O.make_model(O.OU, 1, 1, [beta, 1.0])
        assert np.array_equal(X.yy, oracle_fma.euler(m, tt, [0.1], Wy)[:, 0])
        assert close_x(X.yy, oracle_ref.euler(m, tt, [0.1], Wy)[:, 0])


MODELS = [
    ("wiener1", lambda B: B.Wiener(1), O.make_model(O.WIENER, 1, 1), True),
    ("wiener3", lambda B: B.Wiener(3), O.make_model(O.WIENER, 3, 3), True),
    ("ou", lambda B: B.OrnsteinUhlenbeck(2.0, 0.7), O.make_model(O.OU, 1, 1, [2.0, 0.7]), True),
    ("lorenz", lambda B: B.Lorenz([10.0, 28.0, 8 / 3], 3.0),
     O.make_model(O.LORENZ, 3, 3, [10.0, 28.0, 8 / 3, 3.0, 3.0, 3.0]), True),
    ("fhn_diag", lambda B: B.FitzHughNagumo(0.1, 0.0, 1.5, 0.8, 0.3, 0.2),
     O.make_model(O.FHN_DIAG, 2, 2, [0.1, 0.0, 1.5, 0.8, 0.3, 0.2]), True),
    ("fhn_hypo", lambda B: B.FitzhughDiffusion(0.1, 0.0, 1.5, 0.8, 0.3),
     O.make_model(O.FHN_HYPO, 2, 1, [0.1, 0.0, 1.5, 0.8, 0.3]), True),
    ("intdiff", lambda B: B.IntegratedDiffusion(0.7), O.make_model(O.INTDIFF, 2, 1, [0.7]), False),
    ("nclar3", lambda B: B.NclarDiffusion(1.5, 2.0, 0.4), O.make_model(O.NCLAR3, 3, 1, [1.5, 2.0, 0.4]), False),
    ("linpro2", lambda B: B.LinPro([[-1.0, 0.1], [-0.2, -1.0]], [0.1, -0.2], [[0.4, 0.1], [0.2, 0.8]]),
     O.linpro_model([[-1.0, 0.1], [-0.2, -1.0]], [0.1, -0.2], [[0.4, 0.1], [0.2, 0.8]]), True),
    ("linpro3", lambda B: B.LinPro(-np.array([[1.0, 0.1, 0.0], [-0.2, 1.0, 0.1], [0.0, -0.1, 1.0]]), np.zeros(3),
                                   0.5 * np.eye(3)),
     O.linpro_model(-np.array([[1.0, 0.1, 0.0], [-0.2, 1.0, 0.1], [0.0, -0.1, 1.0]]), np.zeros(3), 0.5 * np.eye(3)),
     True),
]


@pytest.mark.parametrize("name,mk,om,exact", MODELS, ids=[m[0] for m in MODELS])
def test_euler_ensemble_vs_oracle(B, oracle_ref, oracle_fma, name, mk, om, exact):
    """solve!(EulerMaruyama(), X, u, W, P) for an ensemble with per-chain starting points, two chained
    segments (the second starts at the first's end point) and a ragged N."""
    Pm = mk(B)
    d, dp = om.d, om.dprime
    P, S, N = 45, 2, 203
    grids = [warped(0.0, 0.5, N), warped(0.5, 1.2, N)]
    rng = np.random.default_rng(1)
    u = rng.standard_normal((P, d)) * 0.3
    ens = B.PathEnsemble(P, S, N, d, dp, double_buffer=False)
    for s in range(S):
        ens.set_grid(s, grids[s])
    ens.set_start(u)
    ens.sample_(seed=11, stream=0)
    W = ens.download(B.W)
    ens.euler_(Pm)
    X = ens.download(B.X)
    xend = ens.xend
    for p in (0, 7, 44):
        start = u[p]
        for s in range(S):
            Xf = oracle_fma.euler(om, grids[s], start, W[p, s])
            Xr = oracle_ref.euler(om, grids[s], start, W[p, s])
            if exact:
                assert np.array_equal(X[p, s], Xf), (name, p, s)
            assert close_x(X[p, s], Xr), (name, p, s, np.max(np.abs(X[p, s] - Xr)))
            start = X[p, s, -1]
        assert np.array_equal(xend[p], X[p, -1, -1])
    # fused sample! + solve! gives the same W and X
    ens2 = B.PathEnsemble(P, S, N, d, dp, double_buffer=False)
    for s in range(S):
        ens2.set_grid(s, grids[s])
    ens2.set_start(u)
    ens2.sample_euler_(Pm, seed=11, stream=0)
    assert np.array_equal(ens2.download(B.W), W)
    assert np.array_equal(ens2.download(B.X), X)
    ens.close(); ens2.close()


def test_config2_wiener_process_property(B):
    """BASELINE config 2 at reduced P: the Wiener process as target gives X = u + W for every chain (size-independent)."""
    P, N = 20000, 1001
    ens = B.PathEnsemble(P, 1, N, 1, 1, double_buffer=False)
    ens.set_grid(0, np.linspace(0, 1, N))
    ens.set_start([0.0])
    ens.sample_euler_(B.Wiener(1), seed=2, stream=0)
    W = ens.download(B.W); X = ens.download(B.X)
    assert np.allclose(X, W, rtol=0, atol=1e-13)
    inc = np.diff(W[:, 0, :, 0], axis=1) / np.sqrt(1 / 1000)
    assert abs(inc.mean()) < 5 / np.sqrt(inc.size) and abs(inc.var() - 1) < 5 * np.sqrt(2 / inc.size)
    ens.close()


# ----------------------------------------------------------------------------------------------- guided proposals
FHN_PAR = [0.1, 0.0, 1.5, 0.8, 0.3]


def fhn_aux(v):
    Bt = np.array([[10.0, -10.0], [1.5, -1.0]])
    bt = np.array([0.0 / 0.1 - v ** 3 / 0.1, 0.8])
    at = np.array([[0.0, 0.0], [0.0, 0.09]])
    return Bt, bt, at


def oracle_fhn_chain(orc, grids, obs_v, eps=1e-3, Sig=1e-10):
    """Backward chain of partialbridge_bolus3.jl:162-180 with the oracle (Lyapunov step + observation updates)."""
    S = len(grids)
    L = np.array([[1.0, 0.0]]); Sg = np.array([[Sig]])
    nu = np.zeros(2); Hp = np.eye(2) / eps
    nu, Hp = orc.gpupdate_nuH(nu, Hp, L, Sg, [obs_v[-1]])
    out = [None] * S
    for i in range(S - 1, -1, -1):
        Bt, bt, at = fhn_aux(obs_v[i])
        nut, Ht, nu, Hp, C = orc.backward_nuH(O.ODE_LYAP, grids[i], O.const_aux(Bt, bt, at), nu, Hp, 0.0)
        out[i] = (nut, Ht, Bt, bt)
        if i > 0:
            nu, Hp = orc.gpupdate_nuH(nu, Hp, L, Sg, [obs_v[i - 1]])
    return out


def test_guided_nuH_fhn_multisegment(B, oracle_ref, oracle_fma):
    """solve!(Euler(), X, u, W, P°) + llikelihood for PartialBridgeνH, 3 chained segments, FHN hypoelliptic."""
    N, P = 301, 40
    obs_t, obs_v = (0.5, 1.0, 1.5), (-1.0, -0.5, 0.5)
    grids = [warped(a, b, N) for a, b in zip((0.0,) + obs_t[:-1], obs_t)]
    S = len(grids)
    Pm = B.FitzhughDiffusion(*FHN_PAR)
    x0 = np.array([-0.5, -0.6])
    for orc, exact in ((oracle_fma, True), (oracle_ref, False)):
        om = O.make_model(O.FHN_HYPO, 2, 1, FHN_PAR)
        tabs = oracle_fhn_chain(orc, grids, obs_v)
        guides = [B.GuideTables(B.api.K.GUIDE_NUH, grids[s], Pm, tabs[s][1], tabs[s][0], tabs[s][2], tabs[s][3])
                  for s in range(S)]
        og = [O.GuideHolder(O.GUIDE_NUH, grids[s], tabs[s][1], tabs[s][0], Bt=tabs[s][2], betat=tabs[s][3])
              for s in range(S)]
        ens = B.PathEnsemble(P, S, N, 2, 1, double_buffer=False)
        for s in range(S):
            ens.set_grid(s, grids[s])
        ens.set_start(x0)
        ens.sample_(seed=5, stream=1)
        W = ens.download(B.W)
        for skip in (0, 3):
            ens.guided_euler_ll_(Pm, guides, skip=skip, store_x=True)
            X = ens.download(B.X); ll = ens.ll; xend = ens.xend
            for p in (0, 13, 39):
                start, llo = x0, 0.0
                for s in range(S):
                    Xo, xe = orc.guided_euler(om, og[s], start, W[p, s])
                    llo += orc.llikelihood(om, og[s], Xo, skip)
                    if exact:
                        assert np.array_equal(X[p, s], Xo), (p, s)
                    else:
                        assert close_x(X[p, s], Xo), (p, s)
                    start = xe
                if exact:
                    assert ll[p] == llo and np.array_equal(xend[p], start)
                else:
                    assert close_ll(ll[p], llo), (ll[p], llo)
            # second-pass llikelihood on the stored X equals the fused value
            ens.set_ll(np.zeros(P))
            ens.llikelihood_(Pm, guides, skip=skip)
            assert np.array_equal(ens.ll, ll)
            # the guided paths end near the observations
            assert np.max(np.abs(X[:, :, -1, 0] - np.array(obs_v))) < 2e-2
        # without storing X the log-likelihood is unchanged
        ens.guided_euler_ll_(Pm, guides, skip=0, store_x=False)
        ens.guided_euler_ll_(Pm, guides, skip=3, store_x=False)
        assert np.array_equal(ens.ll, ll)
        ens.close()


def test_guided_nuH_intdiff_partialparam(B, oracle_ref):
    """test/partialparam.jl setup (IntegratedDiffusion, N = 1501): tolerance parity (the drift contains sin)."""
    tt = np.arange(1501) / 1000
    Pm = B.IntegratedDiffusion(0.7)
    om = O.make_model(O.INTDIFF, 2, 1, [0.7])
    aux = dict(B=np.array([[0.0, 1.0], [0.0, -1.0]]), beta=np.array([0.0, 0.5]), a=np.array([[0.0, 0.0], [0.0, 0.49]]))
    nuT, HpT, C_ = oracle_ref.update_nuHC([[1.0, 0.0]], [[0.1]], [2.5], 1e-5)
    nu, H, _, _, _ = oracle_ref.backward_nuH(O.ODE_R3, tt, O.const_aux(**aux), nuT, HpT, C_)
    og = O.GuideHolder(O.GUIDE_NUH, tt, H, nu, Bt=aux["B"], betat=aux["beta"])
    g = B.GuideTables(B.api.K.GUIDE_NUH, tt, Pm, H, nu, aux["B"], aux["beta"])
    P = 16
    ens = B.PathEnsemble(P, 1, 1501, 2, 1, double_buffer=False)
    ens.set_grid(0, tt); ens.set_start([2.0, 1.0]); ens.sample_(3, 0)
    W = ens.download(B.W)
    ens.guided_euler_ll_(Pm, [g])
    X = ens.download(B.X); ll = ens.ll
    for p in range(P):
        Xo, _ = oracle_ref.guided_euler(om, og, [2.0, 1.0], W[p, 0])
        assert close_x(X[p, 0], Xo)
        assert close_ll(ll[p], oracle_ref.llikelihood(om, og, Xo))
    ens.close()


def test_guidedbridge_linpro3_config3(B, oracle_ref, oracle_fma):
    """BASELINE config 3 at reduced P: GuidedBridge (H♢, V), LinPro d = 3 with dense sigma, end point = v."""
    N, P = 1001, 24
    tt = np.linspace(0, 1, N)
    B1 = -np.array([[1.0, 0.1, 0.0], [-0.2, 1.0, 0.1], [0.0, -0.1, 1.0]])
    sig = 0.5 * np.eye(3) + 0.05 * np.array([[0, 1, 0], [0, 0, 1], [1, 0, 0]])
    v = np.array([0.5, 0.0, -0.5])
    Pm = B.LinPro(B1, np.zeros(3), sig)
    om = O.linpro_model(B1, np.zeros(3), sig)
    aux = O.const_aux(-np.eye(3), np.zeros(3), sig @ sig.T)
    for orc, exact in ((oracle_fma, True), (oracle_ref, False)):
        Hd, V = orc.backward_HV(tt, aux, v)
        og = O.GuideHolder(O.GUIDE_HV, tt, Hd, V, Bt=-np.eye(3), betat=np.zeros(3))
        g = B.GuideTables(B.api.K.GUIDE_HV, tt, Pm, Hd, V, -np.eye(3), np.zeros(3))
        ens = B.PathEnsemble(P, 1, N, 3, 3, double_buffer=False)
        ens.set_grid(0, tt); ens.set_start(np.zeros(3)); ens.sample_(3, 0)
        W = ens.download(B.W)
        ens.guided_euler_ll_(Pm, [g])
        X = ens.download(B.X); ll = ens.ll
        assert np.array_equal(X[:, 0, -1], np.tile(v, (P, 1)))  # endpoint override  src/euler.jl:241-242
        for p in (0, 5, 23):
            Xo, xe = orc.guided_euler(om, og, np.zeros(3), W[p, 0])
            llo = orc.llikelihood(om, og, Xo)
            if exact:
                assert np.array_equal(X[p, 0], Xo) and ll[p] == llo
            else:
                # the last steps divide by H♢ -> 0: compare away from the singular end, and ll relatively
                assert close_x(X[p, 0, :-1], Xo[:-1]) and close_ll(ll[p], llo)
        ens.close()


def test_partialbridge_LMmu_and_time_dependent_aux(B, oracle_ref, oracle_fma):
    """PartialBridge (L, M, μ) drift and a time-dependent auxiliary process (B̃(t), β̃(t) tabulated per grid point;
    the `linearised_startend` choice of partialbridge_fitzhugh.jl:101-104)."""
    N, P = 401, 12
    tt = warped(0.0, 0.5, N)
    v, u0 = -1.0, -0.5
    uv = lambda t: v * (t / 0.5) + u0 * (1 - t / 0.5)
    Bf = lambda t: np.array([[10.0 - 30.0 * uv(t) ** 2, -10.0], [1.5, -1.0]])
    bf = lambda t: np.array([20.0 * uv(t) ** 3, 0.8])
    af = lambda t: np.array([[0.0, 0.0], [0.0, 0.09]])
    Pm = B.FitzhughDiffusion(*FHN_PAR)
    om = O.make_model(O.FHN_HYPO, 2, 1, FHN_PAR)
    Btg = np.stack([Bf(t) for t in tt]); btg = np.stack([bf(t) for t in tt])
    L = np.array([[1.0, 0.0]]); Sg = np.array([[1e-4]])
    for orc, exact in ((oracle_fma, True), (oracle_ref, False)):
        aux = O.staged_aux(tt, Bf, bf, af)
        Lt, Mt, mut = orc.backward_LMmu(tt, aux, L, Sg)
        og = O.GuideHolder(O.GUIDE_LMMU, tt, Lt, mut, Mm=Mt, v=[v], Bt=Btg, betat=btg, aux_const=False, m=1)
        g = B.GuideTables(B.api.K.GUIDE_LMMU, tt, Pm, Lt, mut, Btg, btg, Mm=Mt, v=[v], aux_const=False, m=1)
        ens = B.PathEnsemble(P, 1, N, 2, 1, double_buffer=False)
        ens.set_grid(0, tt); ens.set_start([-0.5, -0.6]); ens.sample_(8, 0)
        W = ens.download(B.W)
        ens.guided_euler_ll_(Pm, [g])
        X = ens.download(B.X); ll = ens.ll
        for p in range(P):
            Xo, _ = orc.guided_euler(om, og, [-0.5, -0.6], W[p, 0])
            llo = orc.llikelihood(om, og, Xo)
            if exact:
                assert np.array_equal(X[p, 0], Xo) and ll[p] == llo
            else:
                assert close_x(X[p, 0], Xo) and close_ll(ll[p], llo)
        ens.close()


# ----------------------------------------------------------------------------------------------- a15: pCN
def test_pcn_replay_bit_exact(B, oracle_fma):
    """Several pCN iterations of a 4-segment FHN chain ensemble, replayed chain by chain with the oracle:
    W°, X°, ll°, log U, the accept decision, the surviving state and the acceptance counter agree exactly."""
    N, P, iters, rho, seed = 121, 96, 6, 0.9, 4
    obs_t, obs_v = (0.5, 1.0, 1.5, 2.0), (-1.0, -0.5, 0.5, 1.1)
    grids = [warped(a, b, N) for a, b in zip((0.0,) + obs_t[:-1], obs_t)]
    S = 4
    Pm = B.FitzhughDiffusion(*FHN_PAR)
    om = O.make_model(O.FHN_HYPO, 2, 1, FHN_PAR)
    tabs = oracle_fhn_chain(oracle_fma, grids, obs_v)
    guides = [B.GuideTables(B.api.K.GUIDE_NUH, grids[s], Pm, tabs[s][1], tabs[s][0], tabs[s][2], tabs[s][3])
              for s in range(S)]
    og = [O.GuideHolder(O.GUIDE_NUH, grids[s], tabs[s][1], tabs[s][0], Bt=tabs[s][2], betat=tabs[s][3])
          for s in range(S)]
    x0 = np.array([-0.5, -0.6])
    ens = B.PathEnsemble(P, S, N, 2, 1, double_buffer=True, chain_offset=500)
    for s in range(S):
        ens.set_grid(s, grids[s])
    ens.set_start(x0)
    ens.sample_(seed, 0xFFFFFFFE)
    ens.guided_euler_ll_(Pm, guides)
    Wc = ens.download(B.W); Xc = ens.download(B.X); ll = ens.ll.copy()
    acc = 0
    for it in range(iters):
        ens.pcn_step_(Pm, guides, rho, seed, it)
        Wp = ens.download(B.W, which=B.PROP); Xp = ens.download(B.X, which=B.PROP)
        llp, logu, flags = ens.ll_prop, ens.logu, ens.accepted
        for p in range(P):
            llo, lu, Wo, Xo, xe = oracle_fma.pcn_propose(om, og, x0, Wc[p], rho, seed, it, 500 + p)
            assert np.array_equal(Wp[p], Wo) and np.array_equal(Xp[p], Xo), (it, p)
            assert llp[p] == llo and logu[p] == lu
            ok = lu <= llo - ll[p]
            assert bool(flags[p]) == ok
            if ok:
                Wc[p], Xc[p], ll[p] = Wo, Xo, llo
                acc += 1
        assert np.array_equal(ens.download(B.W), Wc) and np.array_equal(ens.download(B.X), Xc)
        assert np.array_equal(ens.ll, ll)
        assert ens.acc == acc
    assert 0 < acc < iters * P
    ens.close()


@pytest.mark.parametrize("model", ["fhn", "intdiff"])
def test_pcn_kernels_agree_bit_for_bit(B, model):
    """bb_pcn_step has three kernels for scalar-noise models with X° stored -- one thread per chain, and the two
    warp-specialised ones (noise warps + dynamics warps; one or two chains per dynamics thread) for small ensembles.  Same seeds, same state:
    W°, X°, ll°, log U, flags, surviving state and acceptance counter must be identical, over several iterations, for a
    ragged ensemble size, a skip and a tabulated auxiliary drift."""
    K = B.api.K
    ctx = B.default_context()
    N, P, S = 83, 333, 3
    grids = [warped(0.5 * s, 0.5 * (s + 1), N) for s in range(S)]
    if model == "fhn":
        Pm = B.FitzhughDiffusion(*FHN_PAR)
        obs = (-1.0, -0.5, 0.5)
        ν, Hp = np.zeros(2), np.eye(2) / 1e-3
        ν, Hp = B.gpupdate_νH(ν, Hp, [[1.0, 0.0]], [[1e-4]], [obs[-1]])
        guides = [None] * S
        for i in range(S - 1, -1, -1):
            Bt, bt, at = fhn_aux(obs[i])
            guides[i], ν, Hp, _ = B.partialbridgeνH(grids[i], Pm, B.LinearAux(Bt, bt, at), ν, Hp)
            if i > 0:
                ν, Hp = B.gpupdate_νH(ν, Hp, [[1.0, 0.0]], [[1e-4]], [obs[i - 1]])
        x0, skip = [-0.5, -0.6], 0
    else:  # time-dependent auxiliary process (tabulated B~, beta~) and a skip
        Pm = B.IntegratedDiffusion(0.7)
        Pt = B.LinearAux(lambda t: np.array([[0.0, 1.0], [0.0, -1.0 - 0.1 * t]]), lambda t: np.array([0.0, 0.5]),
                         lambda t: np.array([[0.0, 0.0], [0.0, 0.49]]))
        guides = [B.PartialBridgeνH(g, Pm, Pt, [[1.0, 0.0]], [2.5 - 0.2 * s], 1e-3, [[0.1]]) for s, g in enumerate(grids)]
        x0, skip = [2.0, 1.0], 2
    out = {}
    try:
        for mode in (K.PCN_ONE_THREAD, K.PCN_WARP_SPECIALISED, K.PCN_WARP_SPECIALISED_2):
            ctx.set_pcn_kernel(mode)
            ens = B.PathEnsemble(P, S, N, 2, 1, chain_offset=77)
            for s in range(S):
                ens.set_grid(s, grids[s])
            ens.set_start(x0); ens.sample_(9, 0xFFFFFFFE); ens.guided_euler_ll_(Pm, guides, skip=skip)
            rec = []
            for it in range(4):
                ens.pcn_step_(Pm, guides, 0.9, 9, it, skip=skip)
                rec.append((ens.download(B.W, which=B.PROP), ens.download(B.X, which=B.PROP), ens.ll_prop, ens.logu,
                            ens.accepted, ens.ll, ens.xend_prop))
            rec.append((ens.download(B.W), ens.download(B.X), ens.acc))
            out[mode] = rec
            ens.close()
    finally:
        ctx.set_pcn_kernel(K.PCN_AUTO)
    a = out[K.PCN_ONE_THREAD]
    for other in (K.PCN_WARP_SPECIALISED, K.PCN_WARP_SPECIALISED_2):
        for ra, rb in zip(a, out[other]):
            for xa, xb in zip(ra, rb):
                assert np.array_equal(xa, xb), other
    assert 0 < a[-1][2] < 4 * P


def test_pcn_against_reference_arithmetic(B, oracle_ref):
    """The same iteration against the reference arithmetic (no fma): ll° within 1e-6 relative, decisions
    replayed from the kernel's own ll values are exact, flips against the oracle's ll are counted."""
    N, P, rho, seed = 501, 64, 0.95, 9
    tt = warped(0.0, 0.5, N)
    Pm = B.FitzhughDiffusion(*FHN_PAR)
    om = O.make_model(O.FHN_HYPO, 2, 1, FHN_PAR)
    Bt, bt, at = fhn_aux(-1.0)
    nuT, HpT, C_ = oracle_ref.update_nuHC([[1.0, 0.0]], [[1e-10]], [-1.0], 1e-3)
    nu, H, _, _, _ = oracle_ref.backward_nuH(O.ODE_R3, tt, O.const_aux(Bt, bt, at), nuT, HpT, C_)
    g = B.GuideTables(B.api.K.GUIDE_NUH, tt, Pm, H, nu, Bt, bt)
    og = O.GuideHolder(O.GUIDE_NUH, tt, H, nu, Bt=Bt, betat=bt)
    x0 = np.array([-0.5, -0.6])
    ens = B.PathEnsemble(P, 1, N, 2, 1)
    ens.set_grid(0, tt); ens.set_start(x0); ens.sample_(seed, 0xFFFFFFFE); ens.guided_euler_ll_(Pm, [g])
    flips = 0
    for it in range(4):
        Wc = ens.download(B.W); ll = ens.ll
        ens.pcn_step_(Pm, [g], rho, seed, it)
        llp, logu, flags = ens.ll_prop, ens.logu, ens.accepted
        assert np.array_equal(flags.astype(bool), logu <= llp - ll)  # replay on the kernel's own values
        for p in range(P):
            llo, lu, Wo, Xo, _ = oracle_ref.pcn_propose(om, [og], x0, Wc[p], rho, seed, it, p)
            assert lu == logu[p] and np.isfinite(llo)
            assert close_ll(llp[p], llo), (llp[p], llo)
            flips += int(bool(flags[p]) != (lu <= llo - ll[p]))
    assert flips == 0
    ens.close()


def test_pcn_properties_large(B):
    """Size-independent properties at a production-like ensemble size (4 segments x N = 1001, 20 000 chains):
    rho = 1 reproduces the current state (ll° = ll, every proposal accepted); acc = sum of flags; parity swap
    keeps the accepted proposal as the new current path."""
    import bridge_jl_b200.configs as cfg
    P, n = 20000, 1001
    Pm, guides, x0, rho = cfg.fhn_config4(n)
    S = len(guides)
    ens = B.PathEnsemble(P, S, n, 2, 1)
    for s, g in enumerate(guides):
        ens.set_grid(s, g.tt)
    ens.set_start(x0); ens.sample_(4, 0xFFFFFFFE); ens.guided_euler_ll_(Pm, guides)
    ll0 = ens.ll
    assert np.all(np.isfinite(ll0))
    X = ens.download(B.X, p0=0, np_=50)
    assert np.max(np.abs(X[:, :, -1, 0] - np.array(cfg.FHN_OBS_V))) < 2e-2  # bridges hit the observations
    ens.pcn_step_(Pm, guides, 1.0, 4, 0)
    assert np.array_equal(ens.ll_prop, ll0) and ens.acc == P and np.all(ens.accepted == 1)
    assert np.array_equal(ens.download(B.X, p0=0, np_=50), X)
    ens.reset_acc()
    total = 0
    for it in range(1, 4):
        ll = ens.ll
        ens.pcn_step_(Pm, guides, rho, 4, it)
        flags = ens.accepted.astype(bool)
        assert np.array_equal(flags, ens.logu <= ens.ll_prop - ll)
        assert np.array_equal(ens.ll, np.where(flags, ens.ll_prop, ll))
        total += int(flags.sum())
        assert ens.acc == total
    assert 0.05 * 3 * P < total < 0.999 * 3 * P
    # accepted proposals became the current state: recomputing ll from the stored current X agrees
    ll = ens.ll
    ens.llikelihood_(Pm, guides)
    assert np.array_equal(ens.ll, ll)
    ens.close()


# ----------------------------------------------------------------------------------------------- constructors
def test_backward_constructors_vs_oracle(B, oracle_fma, oracle_ref):
    """updateνH⁺C, partialbridgeodeνH! (R3 and Lyap), gpHinv!/gpV!, partialbridgeode!, gpupdate on the device, in both
    rounding orders of the shared-table constructors: BIT-EXACT against liboracle_ref.so in the default reference
    arithmetic, and against liboracle_fma.so with ARITH_FUSED (the per-chain kernels' order)."""
    K = B.api.K
    ctx = B.default_context()
    tt = warped(0.0, 1.5, 301)
    auxd = dict(B=np.array([[0.0, 1.0], [0.0, -1.0]]), beta=np.array([0.0, 0.5]), a=np.array([[0.0, 0.0], [0.0, 0.49]]))
    Pt = B.LinearAux(auxd["B"], auxd["beta"], auxd["a"])
    Pm = B.IntegratedDiffusion(0.7)
    L, Sg, v = [[1.0, 0.0]], [[0.1]], [2.5]
    Bf = lambda t: np.array([[-1.0 - t, 0.3], [0.1 * t, -0.5]])
    bf = lambda t: np.array([0.2, np.sin(t)])
    af = lambda t: np.array([[0.3 + 0.1 * t, 0.05], [0.05, 0.2]])
    Pt2 = B.LinearAux(Bf, bf, af)
    B3 = -np.array([[1.0, 0.1, 0.0], [-0.2, 1.0, 0.1], [0.0, -0.1, 1.0]])
    P3 = B.LinPro(B3, [0.1, 0.0, -0.1], 0.5 * np.eye(3))
    tt1 = np.linspace(0, 1, 201)
    rng = np.random.default_rng(0)
    A = rng.standard_normal((3, 3)); Hp = A @ A.T + np.eye(3); nu0 = rng.standard_normal(3)
    L2 = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.5]]); S2 = np.diag([0.1, 0.2]); v2 = np.array([0.3, -0.2])
    try:
        for arith, orc, other in ((K.ARITH_REFERENCE, oracle_ref, oracle_fma), (K.ARITH_FUSED, oracle_fma, oracle_ref)):
            ctx.set_arith(arith)
            Po = B.PartialBridgeνH(tt, Pm, Pt, L, v, 1e-5, Sg)
            nuT, HpT, C_ = orc.update_nuHC(L, Sg, v, 1e-5)
            nu, H, _, _, Cc = orc.backward_nuH(O.ODE_R3, tt, O.const_aux(**auxd), nuT, HpT, C_)
            assert np.array_equal(Po.ν, nu) and np.array_equal(Po.H, H)
            assert abs(Po.C - Cc) <= 1e-14 * abs(Cc)  # C passes through the device log
            # the other rounding order agrees to tolerance only (benign case: Σ = 0.1)
            nuT, HpT, C_ = other.update_nuHC(L, Sg, v, 1e-5)
            nu, H, _, _, _ = other.backward_nuH(O.ODE_R3, tt, O.const_aux(**auxd), nuT, HpT, C_)
            assert np.allclose(Po.ν, nu, rtol=1e-9, atol=1e-9) and np.allclose(Po.H, H, rtol=1e-6)
            # Lyapunov variant + time-dependent auxiliary + chaining outputs
            Po2, nul, Hl, C2 = B.partialbridgeνH(tt, Pm, Pt2, [0.1, -0.2], [[2.0, 0.1], [0.1, 1.0]])
            nu, H, nul_o, Hl_o, C_o = orc.backward_nuH(O.ODE_LYAP, tt, O.staged_aux(tt, Bf, bf, af), [0.1, -0.2],
                                                       [[2.0, 0.1], [0.1, 1.0]], 0.0)
            assert np.array_equal(Po2.ν, nu) and np.array_equal(Po2.H, H)
            assert np.array_equal(nul, nul_o) and np.array_equal(Hl, Hl_o) and C2 == C_o
            # GuidedBridge tables, d = 3
            G = B.GuidedBridge(tt1, P3, P3, [0.5, 0.0, -0.5])
            Hd, V = orc.backward_HV(tt1, O.const_aux(B3, -B3 @ np.array([0.1, 0.0, -0.1]), 0.25 * np.eye(3)),
                                    [0.5, 0.0, -0.5])
            assert np.array_equal(G.Hdia, Hd) and np.array_equal(G.V, V)
            # PartialBridge tables
            PB = B.PartialBridge(tt, Pm, Pt, L, v, Sg)
            Lt, Mt, mut = orc.backward_LMmu(tt, O.const_aux(**auxd), L, Sg)
            assert np.array_equal(PB.L, Lt) and np.array_equal(PB.M, Mt) and np.array_equal(PB.μ, mut)
            # gpupdate
            n1, H1 = B.gpupdate_νH(nu0, Hp, L2, S2, v2)
            n2, H2 = orc.gpupdate_nuH(nu0, Hp, L2, S2, v2)
            assert np.array_equal(n1, n2) and np.array_equal(H1, H2)
            Hd2, V2 = B.gpupdate(Hp, nu0, L2, S2, v2)
            assert np.array_equal(Hd2, H2) and np.array_equal(V2, n2)
            n3, H3 = B.gpupdate_νH(np.zeros(2), np.diag([np.inf, np.inf]), np.eye(2), 0.5 * np.eye(2), [1.0, 2.0])
            assert np.allclose(H3, 0.5 * np.eye(2)) and np.allclose(n3, [1.0, 2.0])
    finally:
        ctx.set_arith(K.ARITH_REFERENCE)


def test_config4_end_to_end_vs_reference_arithmetic(B, oracle_ref, oracle_fma):
    """BASELINE config 4 END TO END at the benchmark's own settings (FitzHugh-Nagumo, Σ = 1e-10, ϵ = 1e-3, 4 τ-warped
    segments x N = 1001): tables built ON THE DEVICE by the benchmarked path (configs.fhn_config4) + device forward
    pass and pCN proposal, against tables from the reference-arithmetic oracle's own backward chain + its forward pass.
    The backward recursion is ill-conditioned here: tables in fused order differ by 4e-8 relative and move ll by up to
    2e-6 relative (measured below and on the CPU), which is why the shared-table constructors run in reference
    arithmetic: their tables equal the oracle's bit for bit and ll agrees far inside north_star's 1e-6."""
    import bridge_jl_b200.configs as cfg
    K = B.api.K
    n, P, seed = 1001, 48, 4
    Pm, guides, x0, rho = cfg.fhn_config4(n)
    grids = cfg.fhn_segment_grids(n)
    S = len(guides)
    tabs = oracle_fhn_chain(oracle_ref, grids, cfg.FHN_OBS_V)
    for s in range(S):  # device tables == reference-arithmetic tables, bit for bit
        assert np.array_equal(guides[s].ν, tabs[s][0]) and np.array_equal(guides[s].H, tabs[s][1]), s
    og = [O.GuideHolder(O.GUIDE_NUH, grids[s], tabs[s][1], tabs[s][0], Bt=tabs[s][2], betat=tabs[s][3]) for s in range(S)]
    om = O.make_model(O.FHN_HYPO, 2, 1, cfg.FHN_PAR)
    ens = B.PathEnsemble(P, S, n, 2, 1)
    for s in range(S):
        ens.set_grid(s, grids[s])
    ens.set_start(x0); ens.sample_(seed, 0xFFFFFFFE); ens.guided_euler_ll_(Pm, guides)
    W = ens.download(B.W); X = ens.download(B.X); ll = ens.ll
    worst_ll = worst_x = 0.0
    for p in range(P):
        u, llo = x0, 0.0
        for s in range(S):
            Xo, u = oracle_ref.guided_euler(om, og[s], u, W[p, s])
            llo += oracle_ref.llikelihood(om, og[s], Xo)
            worst_x = max(worst_x, float(np.max(np.abs(X[p, s] - Xo))))
        worst_ll = max(worst_ll, abs(ll[p] - llo) / abs(llo))
    assert worst_ll <= 1e-9 and worst_x <= 1e-9, (worst_ll, worst_x)  # measured: ~1e-13 / ~1e-12
    # what the fused-order tables would have cost (the reason for the reference-arithmetic default)
    ctx = B.default_context()
    try:
        ctx.set_arith(K.ARITH_FUSED)
        _, gf, _, _ = cfg.fhn_config4(n)
        tf = oracle_fhn_chain(oracle_fma, grids, cfg.FHN_OBS_V)
        for s in range(S):
            assert np.array_equal(gf[s].ν, tf[s][0]) and np.array_equal(gf[s].H, tf[s][1])
        ens.guided_euler_ll_(Pm, gf)
        dfused = float(np.max(np.abs(ens.ll - ll) / np.abs(ll)))
        assert 1e-8 < dfused < 1e-4, dfused  # ~2e-6: above 1e-6
    finally:
        ctx.set_arith(K.ARITH_REFERENCE)
    ens.guided_euler_ll_(Pm, guides)
    assert np.array_equal(ens.ll, ll)
    # one pCN iteration of the benchmarked step
    ens.pcn_step_(Pm, guides, rho, seed, 0)
    llp, logu, flags = ens.ll_prop, ens.logu, ens.accepted
    flips = 0
    for p in range(P):
        llo, lu, Wo, Xo, _ = oracle_ref.pcn_propose(om, og, x0, W[p], rho, seed, 0, p)
        assert lu == logu[p] and abs(llp[p] - llo) <= 1e-9 * abs(llo)
        flips += int(bool(flags[p]) != (lu <= llo - ll[p]))
    assert flips == 0
    print(f"config-4 end to end vs reference arithmetic: max rel dll {worst_ll:.2e}, max |dX| {worst_x:.2e}; "
          f"fused-order tables would move ll by {dfused:.2e}")
    ens.close()


def test_chain_backward_in_one_launch(B, oracle_ref):
    """bb_guides_chain_nuH: the whole multi-segment backward pass in one launch, tables written in place on the device.
    Bit-identical to the per-segment constructors (and therefore to the reference-arithmetic oracle), the forward pass on
    its guides gives identical log-likelihoods, and update_ with other observations rewrites the same device tables."""
    import bridge_jl_b200.configs as cfg
    n, P = 257, 64
    Pm, guides, x0, rho = cfg.fhn_config4(n)
    Pm2, chain, _, _ = cfg.fhn_config4_chain(n)
    tabs = oracle_fhn_chain(oracle_ref, cfg.fhn_segment_grids(n), cfg.FHN_OBS_V)
    for s in range(4):
        assert np.array_equal(chain[s].ν[:-1], guides[s].ν[:-1]) and np.array_equal(chain[s].H[:-1], guides[s].H[:-1])
        assert np.array_equal(chain[s].ν[:-1], tabs[s][0][:-1]) and np.array_equal(chain[s].H[:-1], tabs[s][1][:-1])
    ens = B.PathEnsemble(P, 4, n, 2, 1)
    for s in range(4):
        ens.set_grid(s, guides[s].tt)
    ens.set_start(x0); ens.sample_(4, 0xFFFFFFFE)
    ens.guided_euler_ll_(Pm, guides); ll = ens.ll; X = ens.download(B.X)
    ens.guided_euler_ll_(Pm2, chain.segments)
    assert np.array_equal(ens.ll, ll) and np.array_equal(ens.download(B.X), X)
    ens.pcn_step_(Pm2, chain.segments, rho, 4, 0)
    assert 0 < ens.acc < P
    # other observations: the tables are rewritten in place (same device handles) and match a fresh per-segment build
    obs2 = (-0.8, -0.2, 0.4, 0.9)
    handles = [seg._guide.value for seg in chain]
    cfg.fhn_config4_chain(n, obs_v=obs2, chain=chain)
    assert [seg._guide.value for seg in chain] == handles
    _, g2, _, _ = cfg.fhn_config4(n, obs_v=obs2)
    for s in range(4):
        assert np.array_equal(chain[s].ν[:-1], g2[s].ν[:-1]) and np.array_equal(chain[s].H[:-1], g2[s].H[:-1])
    ens.guided_euler_ll_(Pm, g2); ll2 = ens.ll
    ens.guided_euler_ll_(Pm2, chain.segments)
    assert np.array_equal(ens.ll, ll2) and not np.array_equal(ll2, ll)
    ens.close()


def test_lptilde_through_the_abi(B, oracle_ref):
    """lptilde(x, P::PartialBridgeνH) (test/partialbridgenuH.jl:124) and lptilde(P::GuidedBridge, u) (src/guip.jl:206)
    on the device against the oracle, and against the closed form the reference tests compare with."""
    tt = np.arange(1501) / 1000
    Pm = B.IntegratedDiffusion(0.7)
    Pt = B.LinearAux([[0.0, 1.0], [0.0, -1.0]], [0.0, 0.5], [[0.0, 0.0], [0.0, 0.49]])
    Po = B.PartialBridgeνH(tt, Pm, Pt, [[1.0, 0.0]], [2.5], 1e-5, [[0.1]])
    x0 = np.array([2.0, 1.0])
    lp = B.lptilde(x0, Po)
    lo = oracle_ref.lptilde_nuH(Po.ν[0], Po.H[0], Po.C, x0)
    assert abs(lp - lo) <= 1e-13 * abs(lo)
    assert abs(lp - (-0.98368522)) < 1e-6  # SURVEY 8c check value LP2 of the test/partialparam.jl setup
    # GuidedBridge, 1-d LinPro: test/VHK.jl:65  |lptilde(GP, u) - lp(t, u, T, v, Pt)| < 1e-5
    β, μ, σ, T, u, v = 0.8, 0.2, np.sqrt(0.7), 2.0, 0.5, 0.1
    tt1 = np.linspace(0.0, T, 2001)
    P1 = B.LinPro(-β, μ, σ)
    GP = B.GuidedBridge(tt1, P1, P1, [v])
    mean = np.exp(-β * T) * (u - μ) + μ
    var = σ * σ / (2 * β) * (1 - np.exp(-2 * β * T))
    lpc = -0.5 * (v - mean) ** 2 / var - 0.5 * np.log(2 * np.pi * var)
    got = B.lptilde(GP, [u])
    assert abs(got - lpc) < 1e-5
    assert abs(got - oracle_ref.lptilde_HV(tt1, -β, GP.V[0], GP.Hdia[0], [u])) <= 1e-13 * abs(lpc)
    # time-dependent auxiliary: the staged traces
    Bf = lambda t: np.array([[-1.0 - t, 0.3], [0.1 * t, -0.5]])
    Pt2 = B.LinearAux(Bf, lambda t: np.array([0.2, np.sin(t)]), lambda t: np.array([[0.3 + 0.1 * t, 0.05], [0.05, 0.2]]))
    tt2 = warped(0.0, 1.0, 301)
    G2 = B.GuidedBridge(tt2, B.LinPro(-np.eye(2), np.zeros(2), 0.5 * np.eye(2)), Pt2, [0.3, -0.1],
                        hdia=[[0.02, 0.0], [0.0, 0.03]])
    tr = np.array([[np.trace(Bf(tt2[i] + c * (tt2[i + 1] - tt2[i]))) for c in (0.0, 0.5, 0.75)] for i in range(300)])
    want = oracle_ref.lptilde_HV(tt2, tr, G2.V[0], G2.Hdia[0], [0.1, 0.2])
    assert abs(B.lptilde(G2, [0.1, 0.2]) - want) <= 1e-13 * abs(want)


def test_uploaded_path_is_not_refreshed_away(B):
    """sample!(W2) followed by llikelihood(LeftRule(), X, P°) on a d = d' model: both calls share the context's
    one-chain ensemble, and the second one must evaluate the path it was GIVEN, not a path recomputed from the fresh
    noise (X supplied through bb_ens_upload is the chain's current path)."""
    tt = np.linspace(0.0, 1.0, 201)
    P1 = B.LinPro(-0.8, 0.2, 0.7)
    GP = B.GuidedBridge(tt, P1, B.LinPro(-0.5, 0.0, 0.7), [0.1])
    B.seed_(3)
    W = B.sample(tt.copy(), B.Wiener(1))
    X = B.solve(B.Euler(), 0.5, W, GP)
    ll = B.llikelihood(B.LeftRule(), X, GP)
    W2 = B.sample(tt.copy(), B.Wiener(1))  # marks the shared ensemble's X stale
    assert not np.array_equal(W2.yy, W.yy)
    assert B.llikelihood(B.LeftRule(), X, GP) == ll
    X2 = B.solve(B.Euler(), 0.5, W2, GP)
    assert B.llikelihood(B.LeftRule(), X2, GP) != ll


def test_reference_style_sampler_single_chain(B, oracle_ref):
    """The loop of test/partialbridgenuH.jl:155-198 written with the mirrored API on one chain:
    sample!, solve!(Euler(), Xo, x0, Wo, Po) (returns the end point), llikelihood; 1 < acc < iterations."""
    tt = np.arange(301) / 200
    Pm = B.IntegratedDiffusion(0.7)
    Pt = B.LinearAux([[0.0, 1.0], [0.0, -1.0]], [0.0, 0.5], [[0.0, 0.0], [0.0, 0.49]])
    Po = B.PartialBridgeνH(tt, Pm, Pt, [[1.0, 0.0]], [2.5], 1e-5, [[0.1]])
    x0 = np.array([2.0, 1.0])
    B.seed_(1)
    W = B.sample(tt, B.Wiener(1))
    assert W.yy.shape == (301,) and W.yy[0] == 0.0
    X = B.SamplePath(tt, np.zeros((301, 2)))
    xe = B.solve_(B.Euler(), X, x0, W, Po)
    assert np.array_equal(xe, X.yy[-1]) and np.array_equal(X.yy[0], x0)
    ll = B.llikelihood(B.LeftRule(), X, Po)
    W2 = W.copy(); Wo = W.copy(); Xo = X.copy()
    rng = np.random.default_rng(0)
    acc, iters, rho = 0, 60, 0.9
    for it in range(iters):
        B.sample_(W2, B.Wiener(1))
        Wo.yy[...] = rho * W.yy + np.sqrt(1 - rho ** 2) * W2.yy
        B.bridge_(Xo, x0, Wo, Po)
        llo = B.llikelihood(B.LeftRule(), Xo, Po)
        if np.log(rng.random()) <= llo - ll:
            X, Xo = Xo, X
            W, Wo = Wo, W
            ll = llo
            acc += 1
    assert 1 < acc < iters
    # the device llikelihood of the final path equals the oracle's on the same path and tables
    og = O.GuideHolder(O.GUIDE_NUH, tt, Po.H, Po.ν, Bt=[[0.0, 1.0], [0.0, -1.0]], betat=[0.0, 0.5])
    assert close_ll(ll, oracle_ref.llikelihood(O.make_model(O.INTDIFF, 2, 1, [0.7]), og, X.yy))


def test_innovations_roundtrip(B, oracle_fma):
    """innovations!(EulerMaruyama(), W, X, P) inverts solve! (src/euler.jl:357-376)."""
    tt = warped(0, 1, 257)
    Pm = B.FitzHughNagumo(0.1, 0.0, 1.5, 0.8, 0.3, 0.2)
    om = O.make_model(O.FHN_DIAG, 2, 2, [0.1, 0.0, 1.5, 0.8, 0.3, 0.2])
    B.seed_(5)
    W = B.sample(tt, B.Wiener(2))
    X = B.solve(B.Euler(), [-0.5, -0.6], W, Pm)
    W2 = B.innovations_(B.EulerMaruyama(), B.SamplePath(tt, np.zeros((257, 2))), X, Pm)
    assert np.array_equal(W2.yy, oracle_fma.innovations(om, None, tt, X.yy))
    assert np.max(np.abs(W2.yy - W.yy)) < 1e-10
    with pytest.raises(B.BridgeError) as ei:  # hypoelliptic model: sigma is not invertible
        Ph = B.FitzhughDiffusion(*FHN_PAR)
        Wh = B.sample(tt, B.Wiener(1))
        Xh = B.solve(B.Euler(), [-0.5, -0.6], Wh, Ph)
        B.innovations_(B.EulerMaruyama(), Wh.copy(), Xh, Ph)
    assert ei.value.status in (-11, -12)


# ----------------------------------------------------------------------------------------------- error behaviour
def test_reference_error_messages(B):
    tt = np.linspace(0, 1, 11)
    W = B.SamplePath(tt, np.zeros(11))
    Y = B.SamplePath(np.linspace(0, 1, 12), np.zeros(12))
    with pytest.raises(B.BridgeError, match="Y and W differ in length."):
        B.solve_(B.Euler(), Y, 0.1, W, B.OrnsteinUhlenbeck(1.0, 1.0))
    Pm = B.LinPro([[-1.0]], [0.0], [[1.0]])
    G = B.GuidedBridge(tt, Pm, Pm, [0.3])
    with pytest.raises(B.BridgeError, match="Time axis mismatch between bridge P and driving W."):
        Wsame = B.SamplePath.__new__(B.SamplePath); Wsame.tt, Wsame.yy = G.tt, np.zeros(11)
        B.solve(B.Euler(), 0.1, Wsame, G)
    with pytest.raises(B.BridgeError, match="Starting point has wrong length."):
        B.solve(B.Euler(), [0.1, 0.2], W, B.OrnsteinUhlenbeck(1.0, 1.0))
    with pytest.raises(B.BridgeError):
        B.VSamplePath(tt, np.zeros((2, 12)))
    with pytest.raises(B.BridgeError, match="m == length"):
        B.PartialBridgeνH(tt, B.IntegratedDiffusion(0.7), B.LinearAux(np.eye(2), np.zeros(2), np.eye(2)),
                          [[1.0, 0.0]], [1.0, 2.0], 1e-3, [[0.1]])
    ens = B.PathEnsemble(4, 1, 11, 1, 1, double_buffer=False)
    with pytest.raises(B.BridgeError) as ei:  # a model whose dimensions do not match the ensemble
        ens.set_grid(0, tt)
        ens.euler_(B.Lorenz([10.0, 28.0, 8 / 3], 3.0))
    assert ei.value.status == -6
    with pytest.raises(B.BridgeError) as ei:  # pCN needs the proposal buffers
        ens.pcn_step_(Pm, [G], 0.5, 1, 0)
    assert ei.value.status == -7
    ens.close()


# ----------------------------------------------------------------------------------------------- edge cases
@pytest.mark.parametrize("N,P,S", [(2, 1, 1), (3, 31, 2), (16, 33, 1), (17, 257, 3), (33, 5, 16), (48, 300, 1)])
def test_ragged_sizes_bit_exact(B, oracle_fma, N, P, S):
    """Shortest path (one step), grids that end exactly on / one past a 16-point chunk, chain counts around the
    warp and CTA sizes, the maximum number of chained segments: pCN proposal and bookkeeping stay exact."""
    Pm = B.FitzhughDiffusion(*FHN_PAR)
    om = O.make_model(O.FHN_HYPO, 2, 1, FHN_PAR)
    T = 0.05 * S
    edges = np.linspace(0.0, T, S + 1)
    grids = [warped(edges[s], edges[s + 1], N) for s in range(S)]
    obs_v = [-0.5 + 0.02 * s for s in range(S)]
    tabs = oracle_fhn_chain(oracle_fma, grids, obs_v, eps=1e-2, Sig=1e-4)
    guides = [B.GuideTables(B.api.K.GUIDE_NUH, grids[s], Pm, tabs[s][1], tabs[s][0], tabs[s][2], tabs[s][3])
              for s in range(S)]
    og = [O.GuideHolder(O.GUIDE_NUH, grids[s], tabs[s][1], tabs[s][0], Bt=tabs[s][2], betat=tabs[s][3])
          for s in range(S)]
    x0 = np.array([-0.5, -0.6])
    ens = B.PathEnsemble(P, S, N, 2, 1)
    for s in range(S):
        ens.set_grid(s, grids[s])
    ens.set_start(x0); ens.sample_(21, 0xFFFFFFFE); ens.guided_euler_ll_(Pm, guides)
    Wc = ens.download(B.W); ll = ens.ll.copy()
    acc = 0
    for it in range(3):
        ens.pcn_step_(Pm, guides, 0.7, 21, it)
        Wp = ens.download(B.W, which=B.PROP); Xp = ens.download(B.X, which=B.PROP)
        llp, logu, flags = ens.ll_prop, ens.logu, ens.accepted
        for p in sorted(set([0, P // 2, P - 1])):
            llo, lu, Wo, Xo, xe = oracle_fma.pcn_propose(om, og, x0, Wc[p], 0.7, 21, it, p)
            assert np.array_equal(Wp[p], Wo) and np.array_equal(Xp[p], Xo) and llp[p] == llo and logu[p] == lu
        ok = logu <= llp - ll
        assert np.array_equal(flags.astype(bool), ok)
        Wc[ok] = Wp[ok]; ll[ok] = llp[ok]; acc += int(ok.sum())
        assert ens.acc == acc and np.array_equal(ens.download(B.W), Wc)
    ens.close()


def test_skip_and_rho_limits(B, oracle_fma):
    """skip >= N-1 removes every term of the log-likelihood (ll = 0, everything accepted: log U <= 0);
    rho = 0 gives proposals that do not depend on the current W; rho = -1 mirrors it."""
    N, P = 40, 64
    tt = warped(0.0, 0.3, N)
    Pm = B.FitzhughDiffusion(*FHN_PAR)
    tabs = oracle_fhn_chain(oracle_fma, [tt], [-0.6], eps=1e-2, Sig=1e-4)
    g = B.GuideTables(B.api.K.GUIDE_NUH, tt, Pm, tabs[0][1], tabs[0][0], tabs[0][2], tabs[0][3])
    ens = B.PathEnsemble(P, 1, N, 2, 1)
    ens.set_grid(0, tt); ens.set_start([-0.5, -0.6]); ens.sample_(5, 0)
    ens.guided_euler_ll_(Pm, [g], skip=N - 1)
    assert np.all(ens.ll == 0.0)
    ens.pcn_step_(Pm, [g], 0.5, 5, 1, skip=N + 7)
    assert np.all(ens.ll_prop == 0.0) and ens.acc == P
    W1 = ens.download(B.W)
    ens.pcn_step_(Pm, [g], 0.0, 5, 2)
    Wp = ens.download(B.W, which=B.PROP)
    fresh = B.PathEnsemble(P, 1, N, 1, 1, double_buffer=False)
    fresh.set_grid(0, tt); fresh.sample_(5, 2)
    assert np.array_equal(Wp, fresh.download(B.W))          # rho = 0: W° = W2 exactly
    ens2 = B.PathEnsemble(P, 1, N, 2, 1)
    ens2.set_grid(0, tt); ens2.set_start([-0.5, -0.6]); ens2.upload(B.W, W1)
    ens2.pcn_step_(Pm, [g], -1.0, 5, 3)
    assert np.array_equal(ens2.download(B.W, which=B.PROP), -W1)
    with pytest.raises(B.BridgeError):
        ens.pcn_step_(Pm, [g], 1.5, 5, 4)
    ens.close(); ens2.close(); fresh.close()


def test_x_is_last_proposal_and_refresh(B, oracle_fma):
    """X is kept once per chain and holds the last proposal; the current path of a chain that rejected is
    recomputed by bb_ens_refresh_x and equals solve!(Euler(), X, x0, W, P°) on its current W."""
    N, P = 65, 96
    tt = warped(0.0, 0.5, N)
    Pm = B.FitzhughDiffusion(*FHN_PAR)
    om = O.make_model(O.FHN_HYPO, 2, 1, FHN_PAR)
    tabs = oracle_fhn_chain(oracle_fma, [tt], [-1.0])
    g = B.GuideTables(B.api.K.GUIDE_NUH, tt, Pm, tabs[0][1], tabs[0][0], tabs[0][2], tabs[0][3])
    og = O.GuideHolder(O.GUIDE_NUH, tt, tabs[0][1], tabs[0][0], Bt=tabs[0][2], betat=tabs[0][3])
    ens = B.PathEnsemble(P, 1, N, 2, 1)
    ens.set_grid(0, tt); ens.set_start([-0.5, -0.6]); ens.sample_(8, 0); ens.guided_euler_ll_(Pm, [g])
    for it in range(4):
        ens.pcn_step_(Pm, [g], 0.3, 8, it)   # low rho: many rejections
    flags = ens.accepted.astype(bool)
    assert 0 < flags.sum() < P
    import ctypes
    raw = np.empty((P, 1, N, 2))
    st = B.api.lib.bb_ens_download(ens.h, B.X, B.CUR, 0, P, raw.ctypes.data_as(ctypes.c_void_p))
    assert st == -13                                        # BB_ERR_STALE until refreshed
    Xprop = ens.download(B.X, which=B.PROP)
    Wc = ens.download(B.W)
    Xc = ens.download(B.X)                                  # refreshes
    for p in range(P):
        Xo, _ = oracle_fma.guided_euler(om, og, [-0.5, -0.6], Wc[p, 0])
        assert np.array_equal(Xc[p, 0], Xo)
        if flags[p]:
            assert np.array_equal(Xprop[p], Xc[p])          # accepted proposals ARE the current path
    assert np.array_equal(ens.xend, Xc[:, 0, -1])
    ens.close()


def test_sqrt_of_box_muller_radius_is_ieee(B):
    """bb_sqrtf (rsqrt + two fma, no slow path) against IEEE sqrt: the radius sqrt(-2 log u) of every possible
    uniform is reproduced bit for bit by the oracle, which uses sqrtf; checked over a dense sweep of the generator's
    output through sample! with a unit grid (W[1] = normal 1 of quad 0 ...)."""
    N, P = 9, 200000
    ens = B.PathEnsemble(P, 1, N, 1, 1, double_buffer=False)
    ens.set_grid(0, np.arange(N, dtype=float))   # dt = 1: increments are the normals themselves
    ens.sample_(123, 0)
    W = ens.download(B.W)[:, 0, :, 0]
    from oracle import oracle as OO
    orc = OO.load("fma")
    z = np.diff(W, axis=1)
    for p in range(0, P, 997):
        want = np.array([orc.normal(123, 0, p, n) for n in range(1, N)])
        assert np.array_equal(np.cumsum(want), W[p, 1:])
    assert abs(z.mean()) < 5 / np.sqrt(z.size) and abs(z.var() - 1) < 0.01
    assert np.abs(z).max() < 6.8
    ens.close()


def test_config3_full_size_properties(B):
    """BASELINE config 3 at full size (1e5 paths, d = 3, N = 1001): end point = v for every path, finite ll,
    importance weights E[exp(ll) p~/p] within Monte-Carlo error of 1 (test/guip.jl:245-274 at scale)."""
    import bridge_jl_b200.configs as cfg
    from scipy.linalg import expm, solve_continuous_lyapunov
    n, P = 1001, 100000
    Pm, G, u = cfg.linpro_config3(n)
    ens = B.PathEnsemble(P, 1, n, 3, 3, double_buffer=False)
    ens.set_grid(0, G.tt); ens.set_start(u)
    ens.sample_(3, 0)
    ens.guided_euler_ll_(Pm, [G])
    ll = ens.ll
    assert np.all(np.isfinite(ll))
    assert np.array_equal(ens.xend, np.tile(cfg.LIN3_V, (P, 1)))
    a = cfg.LIN3_SIG @ cfg.LIN3_SIG.T

    def logp(Bm):  # Gaussian transition density of dX = B X dt + sigma dW from 0 to v over T = 1
        mean = np.zeros(3)
        lam = solve_continuous_lyapunov(Bm, -a)
        E = expm(Bm)
        cov = lam - E @ lam @ E.T
        dv = cfg.LIN3_V - mean
        return -0.5 * dv @ np.linalg.solve(cov, dv) - 0.5 * np.log(np.linalg.det(2 * np.pi * cov))

    w = np.exp(ll + logp(cfg.LIN3_B2) - logp(cfg.LIN3_B1))
    z = abs(w.mean() - 1) * np.sqrt(P) / w.std()
    assert abs(w.mean() - 1) < 0.02, (w.mean(), z)   # discretisation bias of the left rule at N = 1001 (SURVEY 8c caveat)
    ens.close()


def test_pcn_step_host_buffers_matches_resident(B, oracle_fma):
    """bb_pcn_step_host (W up, W°/X°/ll°/flags down, pipelined over chain slabs) gives exactly what bb_pcn_step +
    downloads give, and leaves the same device state."""
    N, P, S = 49, 700, 2
    grids = [warped(0.0, 0.5, N), warped(0.5, 1.0, N)]
    Pm = B.FitzhughDiffusion(*FHN_PAR)
    tabs = oracle_fhn_chain(oracle_fma, grids, [-1.0, -0.5])
    guides = [B.GuideTables(B.api.K.GUIDE_NUH, grids[s], Pm, tabs[s][1], tabs[s][0], tabs[s][2], tabs[s][3])
              for s in range(S)]
    enss = []
    for _ in range(2):
        ens = B.PathEnsemble(P, S, N, 2, 1)
        for s in range(S):
            ens.set_grid(s, grids[s])
        ens.set_start([-0.5, -0.6]); ens.sample_(6, 0xFFFFFFFE); ens.guided_euler_ll_(Pm, guides)
        enss.append(ens)
    a, b = enss
    for it in range(3):
        Wh = a.download(B.W)
        a.pcn_step_(Pm, guides, 0.8, 6, it)
        Wo = np.empty((P, S, N, 1)); Xo = np.empty((P, S, N, 2)); llo = np.empty(P); acc = np.empty(P, dtype=np.uint8)
        b.pcn_step_host_(Pm, guides, 0.8, 6, it, Wh, Wo, Xo, llo, acc)
        assert np.array_equal(Wo, a.download(B.W, which=B.PROP)) and np.array_equal(Xo, a.download(B.X, which=B.PROP))
        assert np.array_equal(llo, a.ll_prop) and np.array_equal(acc, a.accepted)
        assert np.array_equal(b.download(B.W), a.download(B.W)) and np.array_equal(b.ll, a.ll) and a.acc == b.acc
    a.close(); b.close()


@pytest.mark.parametrize("pinned", [True, False])
def test_pcn_step_host_skips_rejected_rows(B, oracle_fma, pinned):
    """BB_RUN_SKIP_REJECTED: the rows of W°, X° of chains that accept are those of the plain host step, the rows of
    chains that reject are left untouched (pinned, device-mapped arrays: written by a kernel straight into host memory;
    pageable arrays: the flag changes nothing); ll°, flags and the device state are the same either way."""
    import torch
    N, P, S = 49, 900, 2
    grids = [warped(0.0, 0.5, N), warped(0.5, 1.0, N)]
    Pm = B.FitzhughDiffusion(*FHN_PAR)
    tabs = oracle_fhn_chain(oracle_fma, grids, [-1.0, -0.5])
    guides = [B.GuideTables(B.api.K.GUIDE_NUH, grids[s], Pm, tabs[s][1], tabs[s][0], tabs[s][2], tabs[s][3])
              for s in range(S)]
    enss = []
    for _ in range(2):
        ens = B.PathEnsemble(P, S, N, 2, 1)
        for s in range(S):
            ens.set_grid(s, grids[s])
        ens.set_start([-0.5, -0.6]); ens.sample_(6, 0xFFFFFFFE); ens.guided_euler_ll_(Pm, guides)
        enss.append(ens)
    a, b = enss

    def host(shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype, pin_memory=pinned).numpy()

    Wh, Wo, Xo, llo, acc = host((P, S, N, 1)), host((P, S, N, 1)), host((P, S, N, 2)), host(P), host(P, torch.uint8)
    Wr = np.empty((P, S, N, 1)); Xr = np.empty((P, S, N, 2)); llr = np.empty(P); accr = np.empty(P, dtype=np.uint8)
    for it in range(3):
        Wh[...] = a.download(B.W)
        a.pcn_step_host_(Pm, guides, 0.8, 6, it, Wh, Wr, Xr, llr, accr)
        Wo[...] = -7.0; Xo[...] = -7.0
        b.pcn_step_host_(Pm, guides, 0.8, 6, it, Wh, Wo, Xo, llo, acc, skip_rejected=True)
        f = accr.astype(bool)
        assert 0 < f.sum() < P
        assert np.array_equal(llo, llr) and np.array_equal(acc, accr)
        assert np.array_equal(Wo[f], Wr[f]) and np.array_equal(Xo[f], Xr[f])
        if pinned:
            assert np.all(Wo[~f] == -7.0) and np.all(Xo[~f] == -7.0)
        else:
            assert np.array_equal(Wo, Wr) and np.array_equal(Xo, Xr)
        assert np.array_equal(b.download(B.W), a.download(B.W)) and np.array_equal(b.ll, a.ll) and a.acc == b.acc
    # in place: Wo aliases W, the host array follows the chains' current W through the iterations (the loop's swap)
    Wh[...] = b.download(B.W)
    if pinned:
        for it in range(3, 6):
            a.pcn_step_(Pm, guides, 0.8, 6, it)
            b.pcn_step_host_(Pm, guides, 0.8, 6, it, Wh, Wh, Xo, llo, acc, skip_rejected=True)
            assert np.array_equal(Wh, a.download(B.W)) and np.array_equal(acc, a.accepted) and a.acc == b.acc
    else:
        with pytest.raises(B.BridgeError):
            b.pcn_step_host_(Pm, guides, 0.8, 6, 3, Wh, Wh, Xo, llo, acc, skip_rejected=True)
    a.close(); b.close()


@pytest.mark.parametrize("kind", ["nuH", "LMmu"])
def test_nonconstdiff_pair(B, oracle_ref, oracle_fma, kind):
    """a != a~: the extra terms -1/2 tr((a-a~)H)dt + 1/2 r'(a-a~)r dt (src/partialbridge.jl:79-84) in the fused
    log-likelihood, in the second-pass llikelihood and in the pCN accept rule; time-dependent a~(t)."""
    N, P = 161, 40
    tt = np.arange(N) / 1000
    Pm = B.IntegratedDiffusion(0.7)
    om = O.make_model(O.INTDIFF, 2, 1, [0.7])
    Bm, be, a_t = np.array([[0.0, 1.0], [0.0, -1.0]]), np.array([0.0, 0.5]), np.array([[0.0, 0.0], [0.0, 0.49]])
    af = lambda t: np.array([[0.02 + 0.1 * t, 0.0], [0.0, 0.3 + t]])
    Pt = B.LinearAux(lambda t: Bm, lambda t: be, af)
    for orc, exact in ((oracle_fma, True), (oracle_ref, False)):
        aux = O.staged_aux(tt, lambda t: Bm, lambda t: be, af)
        Ad = np.stack([a_t - af(t) for t in tt])
        Btg, btg = np.tile(Bm, (N, 1, 1)), np.tile(be, (N, 1))
        if kind == "nuH":
            nuT, HpT, C_ = orc.update_nuHC([[1.0, 0.0]], [[0.1]], [2.5], 1e-3)
            nu, H, _, _, _ = orc.backward_nuH(O.ODE_R3, tt, aux, nuT, HpT, C_)
            og = O.GuideHolder(O.GUIDE_NUH, tt, H, nu, Bt=Btg, betat=btg, aux_const=False, Adiff=Ad, adiff_const=False)
            g = B.GuideTables(B.api.K.GUIDE_NUH, tt, Pm, H, nu, Btg, btg, aux_const=False, Adiff=Ad, adiff_const=False)
        else:
            Lt, Mt, mut = orc.backward_LMmu(tt, aux, [[1.0, 0.0]], [[0.1]])
            og = O.GuideHolder(O.GUIDE_LMMU, tt, Lt, mut, Mm=Mt, v=[2.5], Bt=Btg, betat=btg, aux_const=False, m=1,
                               Adiff=Ad, adiff_const=False)
            g = B.GuideTables(B.api.K.GUIDE_LMMU, tt, Pm, Lt, mut, Btg, btg, Mm=Mt, v=[2.5], aux_const=False, m=1,
                              Adiff=Ad, adiff_const=False)
        ens = B.PathEnsemble(P, 1, N, 2, 1)
        ens.set_grid(0, tt); ens.set_start([2.0, 1.0]); ens.sample_(13, 0); ens.guided_euler_ll_(Pm, [g])
        W = ens.download(B.W); X = ens.download(B.X); ll = ens.ll.copy()
        og0 = O.GuideHolder(og.kind, tt, og.A, og.b, Mm=og.Mm, v=og.v, Bt=Btg, betat=btg, aux_const=False, m=og.m)
        for p in (0, 17, 39):
            Xo, _ = orc.guided_euler(om, og, [2.0, 1.0], W[p, 0])
            llo = orc.llikelihood(om, og, Xo)
            assert abs(llo - orc.llikelihood(om, og0, Xo)) > 1e-4          # the extra terms are not negligible here
            assert close_x(X[p, 0], Xo) and close_ll(ll[p], llo)           # sin in the drift: tolerance
        ens.llikelihood_(Pm, [g])
        assert np.array_equal(ens.ll, ll)
        ens.pcn_step_(Pm, [g], 0.9, 13, 1)
        assert np.array_equal(ens.accepted.astype(bool), ens.logu <= ens.ll_prop - ll)
        for p in (0, 39):
            llo, lu, Wo, Xo, _ = orc.pcn_propose(om, [og], [2.0, 1.0], W[p], 0.9, 13, 1, p)
            assert close_ll(ens.ll_prop[p], llo) and lu == ens.logu[p]
        ens.close()
    # the mirrored constructors decide by the reference's trait constdiff(Target) && constdiff(Pt)
    # (src/partialbridge.jl:64), not by comparing a and a~
    Ptn = B.LinearAux(lambda t: Bm, lambda t: be, af, constdiff=False)
    Po = B.PartialBridgeνH(tt, Pm, Ptn, [[1.0, 0.0]], [2.5], 1e-3, [[0.1]])
    assert Po.constdiff is False
    Pc = B.PartialBridgeνH(tt, Pm, Pt, [[1.0, 0.0]], [2.5], 1e-3, [[0.1]])
    assert Pc.constdiff is True  # a != a~ but both processes declare constdiff: no extra terms, as in the reference
    ens = B.PathEnsemble(4, 1, N, 2, 1)
    ens.set_grid(0, tt); ens.set_start([2.0, 1.0]); ens.sample_(13, 0)
    ens.guided_euler_ll_(Pm, [Po]); lln = ens.ll.copy()
    ens.guided_euler_ll_(Pm, [Pc]); llc = ens.ll.copy()
    assert np.all(np.abs(lln - llc) > 1e-4)
    ens.close()


def test_config2_full_size(B):
    """BASELINE config 2 at FULL size: 1e6 independent Float64 paths, n = 1001, Wiener + EulerMaruyama on one GPU
    (8 GB W + 8 GB X): fused sample! + solve! with the Wiener process as target gives X = u + W (checked on a strided
    subset of chains and through the end points of all chains), and a second plain solve! of the stored W reproduces X."""
    P, N = 1000000, 1001
    ens = B.PathEnsemble(P, 1, N, 1, 1, double_buffer=False)
    ens.set_grid(0, np.linspace(0, 1, N))
    ens.set_start([0.25])
    ens.sample_euler_(B.Wiener(1), seed=2, stream=0)
    xend = ens.xend[:, 0]
    for p0 in (0, 333333, P - 64):
        W = ens.download(B.W, p0=p0, np_=64)[:, 0, :, 0]
        X = ens.download(B.X, p0=p0, np_=64)[:, 0, :, 0]
        assert np.allclose(X, 0.25 + W, rtol=0, atol=1e-13)
        assert np.array_equal(X[:, -1], xend[p0:p0 + 64])
    assert abs(np.mean(xend) - 0.25) < 5e-3 and abs(np.var(xend) - 1.0) < 5e-3   # W_1 ~ N(0, 1) over 1e6 paths
    Xa = ens.download(B.X, p0=500000, np_=32)
    ens.euler_(B.Wiener(1))
    assert np.array_equal(ens.download(B.X, p0=500000, np_=32), Xa)
    ens.close()


def test_pooled_online_statistics(B, oracle_fma):
    """mcstart / mcnext! / mcstats (src/mclog.jl:22-56, 88-93) pooled over chains and iterations: device moments of the
    CURRENT paths equal numpy's over the downloaded paths."""
    N, P, S = 37, 300, 2
    grids = [warped(0.0, 0.5, N), warped(0.5, 1.0, N)]
    Pm = B.FitzhughDiffusion(*FHN_PAR)
    tabs = oracle_fhn_chain(oracle_fma, grids, [-1.0, -0.5])
    guides = [B.GuideTables(B.api.K.GUIDE_NUH, grids[s], Pm, tabs[s][1], tabs[s][0], tabs[s][2], tabs[s][3])
              for s in range(S)]
    ens = B.PathEnsemble(P, S, N, 2, 1)
    for s in range(S):
        ens.set_grid(s, grids[s])
    ens.set_start([-0.5, -0.6]); ens.sample_(31, 0xFFFFFFFE); ens.guided_euler_ll_(Pm, guides)
    ens.mc_reset_()
    samples = []
    for it in range(3):
        ens.pcn_step_(Pm, guides, 0.5, 31, it)
        ens.mc_update_()                       # refreshes the chains that rejected, then accumulates
        samples.append(ens.download(B.X))
    Xs = np.concatenate(samples, axis=0)       # [3P, S, N, d]
    mean, cov, n = ens.mc_stats()
    assert n == 3 * P
    assert np.allclose(mean, Xs.mean(axis=0), rtol=1e-12, atol=1e-13)
    want = np.einsum("ksni,ksnj->snij", Xs - Xs.mean(axis=0), Xs - Xs.mean(axis=0)) / (n - 1)
    assert np.allclose(cov, want, rtol=1e-9, atol=1e-12)
    # the conditioned end points: variance O(1e-10) (Σ = 1e-10) next to a mean O(1) -- raw moments would cancel there;
    # the device accumulates deviations from a pivot path
    assert np.all(want[:, -1, 0, 0] < 1e-6) and np.all(want[:, -1, 0, 0] > 0)
    assert np.allclose(cov[:, -1, 0, 0], want[:, -1, 0, 0], rtol=1e-6, atol=0.0)
    lo, hi = ens.mc_band()
    assert np.all(lo <= mean) and np.all(mean <= hi)
    assert np.allclose(mean[0, 0], [-0.5, -0.6]) and np.allclose(cov[0, 0], 0.0, atol=1e-12)   # fixed start point
    ens.close()


def test_per_chain_online_statistics_bit_exact(B, oracle_fma):
    """mcstart / mcnext! / mcstats / mcband with the reference's semantics (src/mclog.jl:22-24, 47-56, 75-93): one state
    per chain over the recorded iterations, as `mcstate = [mcnext!(mcstate[i], XX[i].yy) ...]` keeps it
    (partialbridge_fitzhugh.jl:169-189).  The device state equals the oracle's recurrence on the downloaded current paths
    bit for bit (ragged N: the last chunk is padded; ragged P: not a multiple of the CTA size)."""
    from oracle import oracle as O
    N, P, S = 37, 301, 2
    grids = [warped(0.0, 0.5, N), warped(0.5, 1.0, N)]
    Pm = B.FitzhughDiffusion(*FHN_PAR)
    tabs = oracle_fhn_chain(oracle_fma, grids, [-1.0, -0.5])
    guides = [B.GuideTables(B.api.K.GUIDE_NUH, grids[s], Pm, tabs[s][1], tabs[s][0], tabs[s][2], tabs[s][3])
              for s in range(S)]
    ens = B.PathEnsemble(P, S, N, 2, 1)
    for s in range(S):
        ens.set_grid(s, grids[s])
    ens.set_start([-0.5, -0.6]); ens.sample_(33, 0xFFFFFFFE); ens.guided_euler_ll_(Pm, guides)
    ens.chain_mc_reset_()
    mc = None
    for it in range(4):
        ens.pcn_step_(Pm, guides, 0.5, 33, it)
        ens.chain_mc_update_()                 # refreshes the chains that rejected, then mcnext! for every chain
        X = ens.download(B.X)                  # [P, S, N, d]
        mc = O.mcnext(O.mcstart(X) if mc is None else mc, X)
        if it == 0:                            # k = 1: mean = the path itself, m2 = 0, cov = 0/0
            mean, cov, k = ens.chain_mc_stats()
            assert k == 1 and np.array_equal(mean, X) and np.all(np.isnan(cov))
    mean, cov, k = ens.chain_mc_stats()
    assert k == 4
    m_want, cov_want = O.mcstats(mc)
    assert np.array_equal(mean, m_want)
    assert np.array_equal(cov, cov_want)
    assert np.any(cov[:, :, 1:, 0, 0] > 0)     # chains moved between the recorded iterations
    lo, hi = ens.chain_mc_band()
    lo_want, hi_want = O.mcband(mc)
    assert np.array_equal(lo, lo_want) and np.array_equal(hi, hi_want)
    # a sub-range of chains, and the count alone
    m1, c1, _ = ens.chain_mc_stats(17, 5)
    assert np.array_equal(m1, m_want[17:22]) and np.array_equal(c1, cov_want[17:22])
    # a reset starts over
    ens.chain_mc_reset_()
    ens.chain_mc_update_()
    mean, _, k = ens.chain_mc_stats()
    assert k == 1 and np.array_equal(mean, ens.download(B.X))
    # statistics of stale paths are refused through the raw ABI
    ens.pcn_step_(Pm, guides, 0.5, 33, 99)
    assert B.api.lib.bb_ens_chain_mc_update(ens.h) == -13       # BB_ERR_STALE until refreshed
    ens.close()


# ----------------------------------------------------------------------------------------------- rank 4: other SDE schemes
@pytest.mark.parametrize("name,mk,om,exact", MODELS, ids=[m[0] for m in MODELS])
def test_stochastic_heun_vs_oracle(B, oracle_ref, oracle_fma, name, mk, om, exact):
    """solve!(StochasticHeun(), Y, u, W, P)  src/euler.jl:178-198 incl. its quirk: the loop stops at N-2, yy[N-1] gets
    the end point and yy[N] keeps the caller's value."""
    Pm = mk(B)
    d, dp = om.d, om.dprime
    P, N = 70, 83
    tt = warped(0.0, 0.7, N)
    rng = np.random.default_rng(2)
    u = rng.standard_normal((P, d)) * 0.3
    ens = B.PathEnsemble(P, 1, N, d, dp, double_buffer=False)
    ens.set_grid(0, tt)
    ens.set_start(u)
    ens.sample_(seed=5, stream=1)
    W = ens.download(B.W)
    X0 = rng.standard_normal((P, 1, N, d))
    ens.upload(B.X, X0)
    ens.solve_scheme_(Pm, B.StochasticHeun.scheme)
    X = ens.download(B.X)
    assert np.array_equal(X[:, 0, -1], X0[:, 0, -1])  # last point untouched
    for p in (0, 33, 69):
        Xf = oracle_fma.heun(om, tt, u[p], W[p, 0], X0[p, 0])
        Xr = oracle_ref.heun(om, tt, u[p], W[p, 0], X0[p, 0])
        if exact:
            assert np.array_equal(X[p, 0], Xf), (name, p)
        assert close_x(X[p, 0], Xr), (name, p)
    # StratonovichEuler (and StochasticRungeKutta for scalar processes) coincide with Euler-Maruyama for constant σ
    ens.solve_scheme_(Pm, B.StratonovichEuler.scheme)
    Xs = ens.download(B.X)
    ens.euler_(Pm)
    assert np.array_equal(Xs, ens.download(B.X))
    if d == 1:
        ens.solve_scheme_(Pm, B.StochasticRungeKutta.scheme)
        assert np.array_equal(Xs, ens.download(B.X))
    else:
        with pytest.raises(B.BridgeError) as ei:
            ens.solve_scheme_(Pm, B.StochasticRungeKutta.scheme)
        assert ei.value.status == -11
    with pytest.raises(B.BridgeError) as ei:
        ens.solve_scheme_(Pm, B.Mdb.scheme)
    assert ei.value.status == -11
    ens.close()


def test_inplace_solver_on_vsamplepath(B, oracle_fma):
    """solve!(EulerMaruyama!(), Y::VSamplePath, u, W, P) (src/sde!.jl:21-53) against solve!(EulerMaruyama(), ...) on a
    SamplePath and the oracle: test/euler.jl:49-68 (Lorenz, d = 3; the reference asserts agreement to eps(): here the
    two containers go through the same kernel, so it is exact), the d x N container, its DimensionMismatch and the
    "Starting point has wrong length." check of src/sde!.jl:30."""
    K = B.api.K
    N = 501
    tt = np.linspace(0.0, 5.0, N)
    P = B.Lorenz([10.0, 28.0, 8.0 / 3.0], [3.0, 3.0, 3.0])
    om = O.make_model(O.LORENZ, 3, 3, [10.0, 28.0, 8.0 / 3.0, 3.0, 3.0, 3.0])
    u = np.array([1.508870, -1.531271, 25.46091])
    B.seed_(5)
    W = B.sample(tt.copy(), B.Wiener(3))
    X = B.solve(B.EulerMaruyama(), u, W, P)
    Wv = B.VSamplePath(tt.copy(), W.yy.T.copy())      # d x N, as the reference's VSamplePath
    Yv = B.VSamplePath(tt.copy(), np.zeros((3, N)))
    out = B.solve_(B.EulerMaruyama_(), Yv, u, Wv, P)
    assert out is Yv and np.array_equal(Yv.yy, X.yy) and np.array_equal(Yv.tt, W.tt)
    assert np.array_equal(X.yy, oracle_fma.euler(om, tt, u, W.yy))
    with pytest.raises(B.BridgeError) as ei:                      # u of the wrong length
        B.solve_(B.EulerMaruyama_(), Yv, u[:2], Wv, P)
    assert ei.value.status == K.ERR_STARTPOINT and "Starting point has wrong length." in str(ei.value)
    with pytest.raises(B.BridgeError) as ei:                      # DimensionMismatch of the container  src/types.jl:127
        B.VSamplePath(tt, np.zeros((3, N - 1)))
    assert ei.value.status == K.ERR_DIM
    with pytest.raises(B.BridgeError) as ei:                      # Y and W differ in length
        B.solve_(B.EulerMaruyama_(), B.VSamplePath(tt[:-1].copy(), np.zeros((3, N - 1))), u, Wv, P)
    assert ei.value.status == K.ERR_LENGTH


def test_linearappr_as_auxiliary_process(B, oracle_ref):
    """A target with the auxiliary process linearised along a trajectory (linearappr, src/linpro.jl:196; the idea of
    test/smoothing.jl:73-83, here for the FitzHugh-Nagumo model with diagonal noise): the guided proposal built from it
    runs on the device -- tables from the R3 backward solver with the tabulated coefficients (the reference's own
    constructor for this type does not run), tabulated auxiliary drift and the non-constdiff terms in the
    log-likelihood -- and agrees with the oracle fed the same tables.  (For d = d' = 3 the step records of this
    combination do not fit the kernel's shared-memory staging: BB_ERR_UNSUPPORTED, checked below.)"""
    K = B.api.K
    P = B.FitzHughNagumo(0.1, 0.0, 1.5, 0.8, 0.3, 0.25)
    om = O.make_model(O.FHN_DIAG, 2, 2, [0.1, 0.0, 1.5, 0.8, 0.3, 0.25])
    N = 201
    tt = warped(0.0, 0.5, N)
    u = np.array([-0.5, -0.6])
    B.seed_(9)
    Y = B.solve(B.Euler(), u, B.sample(tt.copy(), B.Wiener(2)), P)           # a trajectory to linearise along
    Pt = B.linearappr(Y, P)
    v = np.array([Y.yy[-1, 0] + 0.1])
    Po = B.PartialBridgeνH(tt, P, Pt, [[1.0, 0.0]], v, 1e-3, [[1e-4]])
    assert Po.constdiff is False
    ens = B.PathEnsemble(16, 1, N, 2, 2, double_buffer=False)
    ens.set_grid(0, tt); ens.set_start(u); ens.sample_(10, 0)
    W = ens.download(B.W)
    ens.guided_euler_ll_(P, [Po])
    X = ens.download(B.X); ll = ens.ll
    assert np.max(np.abs(X[:, 0, -1, 0] - v[0])) < 0.05                      # the bridges hit the observation
    Btg = np.stack([Pt.B(t) for t in tt]); btg = np.stack([Pt.β(t) for t in tt])
    Ad = np.stack([np.diag([0.09, 0.0625]) - Pt.a(t) for t in tt])
    og = O.GuideHolder(O.GUIDE_NUH, tt, Po.H, Po.ν, Bt=Btg, betat=btg, aux_const=False, Adiff=Ad, adiff_const=False)
    for p in (0, 7, 15):
        Xo, _ = oracle_ref.guided_euler(om, og, u, W[p, 0])
        assert close_x(X[p, 0], Xo) and close_ll(ll[p], oracle_ref.llikelihood(om, og, Xo))
    ens.close()
    # d = d' = 3 with tabulated auxiliary drift AND non-constdiff terms: refused cleanly
    P3 = B.Lorenz([10.0, 28.0, 8.0 / 3.0], [3.0, 3.0, 3.0])
    t3 = np.linspace(0.0, 0.1, 65)
    Y3 = B.solve(B.Euler(), [1.5, -1.5, 25.0], B.sample(t3.copy(), B.Wiener(3)), P3)
    Po3 = B.PartialBridgeνH(t3, P3, B.linearappr(Y3, P3), np.eye(3), Y3.yy[-1], 1e-4, 1e-2 * np.eye(3))
    e3 = B.PathEnsemble(8, 1, 65, 3, 3, double_buffer=False)
    e3.set_grid(0, t3); e3.set_start([1.5, -1.5, 25.0]); e3.sample_(1, 0)
    with pytest.raises(B.BridgeError) as ei:
        e3.guided_euler_ll_(P3, [Po3])
    assert ei.value.status == K.ERR_UNSUPPORTED
    e3.close()


def test_mdb_on_guided_proposals(B, oracle_fma, oracle_ref):
    """solve!(Mdb(), Y, u, W, P°)  src/euler.jl:308-327 for guided proposals (they carry P.tt and the indexed drift the
    scheme needs): device against the oracle, bit-exact in the kernels' rounding order and to tolerance against the
    reference arithmetic; through the mirrored solve! it returns Y (not the end point), its last step is noise free."""
    K = B.api.K
    N, P = 201, 40
    tt = warped(0.0, 0.5, N)
    Pm = B.FitzhughDiffusion(*FHN_PAR)
    om = O.make_model(O.FHN_HYPO, 2, 1, FHN_PAR)
    Bt, bt, at = fhn_aux(-1.0)
    Po = B.PartialBridgeνH(tt, Pm, B.LinearAux(Bt, bt, at), [[1.0, 0.0]], [-1.0], 1e-3, [[1e-4]])
    og = O.GuideHolder(O.GUIDE_NUH, tt, Po.H, Po.ν, Bt=Bt, betat=bt)
    ens = B.PathEnsemble(P, 1, N, 2, 1, double_buffer=False)
    ens.set_grid(0, tt); ens.set_start([-0.5, -0.6]); ens.sample_(12, 0)
    W = ens.download(B.W)
    ens.guided_mdb_(Pm, [Po])
    X = ens.download(B.X); xe = ens.xend
    for p in (0, 17, 39):
        Xo, xo = oracle_fma.guided_mdb(om, og, [-0.5, -0.6], W[p, 0])
        assert np.array_equal(X[p, 0], Xo) and np.array_equal(xe[p], xo)
        Xr, _ = oracle_ref.guided_mdb(om, og, [-0.5, -0.6], W[p, 0])
        assert close_x(X[p, 0], Xr)
    ens.guided_euler_ll_(Pm, [Po])
    assert not np.array_equal(ens.download(B.X), X)          # it is not the plain guided Euler path
    ens.close()
    # GuidedBridge (H♢, V) with the end-point rule, dense sigma, through the mirrored call
    tt3 = np.linspace(0, 1, 101)
    B1 = -np.array([[1.0, 0.1, 0.0], [-0.2, 1.0, 0.1], [0.0, -0.1, 1.0]])
    sig = 0.5 * np.eye(3) + 0.05 * np.array([[0, 1, 0], [0, 0, 1], [1, 0, 0]])
    v = np.array([0.5, 0.0, -0.5])
    P3 = B.LinPro(B1, np.zeros(3), sig)
    G3 = B.GuidedBridge(tt3, P3, B.LinPro(-np.eye(3), np.zeros(3), sig), v)
    og3 = O.GuideHolder(O.GUIDE_HV, tt3, G3.Hdia, G3.V, Bt=-np.eye(3), betat=np.zeros(3))
    B.seed_(2)
    W3 = B.sample(tt3.copy(), B.Wiener(3))
    Y = B.SamplePath(tt3.copy(), np.zeros((101, 3)))
    out = B.solve_(B.Mdb(), Y, np.zeros(3), W3, G3)
    assert out is Y and np.array_equal(Y.yy[-1], v)
    Xo, _ = oracle_fma.guided_mdb(O.linpro_model(B1, np.zeros(3), sig), og3, np.zeros(3), W3.yy)
    assert np.array_equal(Y.yy, Xo)
    with pytest.raises(B.BridgeError) as ei:                 # a plain target has no time axis of its own
        B.solve_(B.Mdb(), Y, np.zeros(3), W3, P3)
    assert ei.value.status == K.ERR_UNSUPPORTED


def test_schemes_through_solve(B, oracle_fma):
    """solve(StochasticHeun(), u, W, P) / solve(StratonovichEuler(), ...) on SamplePaths (P = 1 plumbing)."""
    tt = np.arange(0, 101) * 0.01
    W = B.sample(tt, B.Wiener())
    Pm = B.OrnsteinUhlenbeck(2.0, 0.7)
    om = O.make_model(O.OU, 1, 1, [2.0, 0.7])
    X = B.solve(B.StochasticHeun(), 0.1, W, Pm)
    assert np.array_equal(X.yy, oracle_fma.heun(om, tt, [0.1], W.yy)[:, 0]) and X.yy[-1] == 0.0
    Xs = B.solve(B.StratonovichEuler(), 0.1, W, Pm)
    assert np.array_equal(Xs.yy, B.solve(B.Euler(), 0.1, W, Pm).yy)
    Xk = B.solve(B.StochasticRungeKutta(), 0.1, W, Pm)
    assert np.array_equal(Xk.yy, Xs.yy)
