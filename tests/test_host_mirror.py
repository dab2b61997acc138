"""Host-side mirror of the reference interface (bridge.jl_b200/api.py, configs.py): containers, coefficient protocol,
auxiliary-process staging and workload definitions.  CPU only: nothing here computes on a device (the package has no CPU
compute path); these are the parts that run on the host in front of the C ABI."""
import numpy as np
import pytest

from oracle import oracle as O


@pytest.fixture(scope="module")
def B():
    import bridge_jl_b200 as B
    return B


def test_samplepath_containers(B):
    tt = np.linspace(0.0, 1.0, 11)
    X = B.SamplePath(tt, np.zeros((11, 3)))
    assert len(X) == 11 and X.dim == 3 and B.SamplePath(tt, np.zeros(11)).dim == 1
    Y = X.copy()
    Y.yy[0, 0] = 1.0
    assert X.yy[0, 0] == 0.0                                   # copy is deep (src/types.jl:84)
    sp = B.samplepath(tt, np.zeros(2))                          # samplepath(tt, v) aliases tt (src/types.jl:78-81)
    assert sp.yy.shape == (11, 2) and sp.tt is not None and np.shares_memory(sp.tt, tt)
    with pytest.raises(B.BridgeError) as ei:
        B.SamplePath(tt, np.zeros(10))
    assert ei.value.status == -4 and "length(tt) != size(yy, 2)" in str(ei.value)   # src/types.jl:127
    V = B.VSamplePath(tt, np.arange(22.0).reshape(2, 11))       # d x N matrix (src/types.jl:123-130)
    assert V.dim == 2 and V.yy[3, 1] == 14.0
    with pytest.raises(B.BridgeError):
        B.VSamplePath(tt, np.zeros((11, 2)))


def test_coefficient_protocol_of_the_registry_models(B):
    x = np.array([0.3, -0.2])
    P = B.FitzhughDiffusion(0.1, 0.0, 1.5, 0.8, 0.3)
    assert np.allclose(P.b(0.0, x), [(x[0] - x[1] - x[0] ** 3) / 0.1, 1.5 * x[0] - x[1] + 0.8])
    assert np.array_equal(P.a(0.0, x), [[0.0, 0.0], [0.0, 0.09]])   # fallback a = σσ'  (src/types.jl:32)
    m = P.cmodel()
    assert (m.id, m.d, m.dprime) == (O.FHN_HYPO, 2, 1) and list(m.par)[:5] == [0.1, 0.0, 1.5, 0.8, 0.3]
    L = B.LinPro([[-1.0, 0.1], [-0.2, -1.0]], [0.1, -0.2], [[0.4, 0.1], [0.2, 0.8]])
    assert np.allclose(L.b(0.0, x), L.B(0.0) @ (x - [0.1, -0.2])) and np.allclose(L.β(0.0), -L.B(0.0) @ [0.1, -0.2])
    Bo = B.BolusDiffusion(100.0, 8.0, 1.25, 1.5, 0.5, 0.2)
    assert np.isclose(Bo.b(2.0, x)[0], 100.0 * 1.0 - (1.25 + 8.0) * x[0] + 1.5 * x[1])   # dose(2) = 1
    Lm = B.Landmarks(0.5, 2.0, 0.5)
    assert Lm.σ(0.0).shape == (16, 8) and Lm.cmodel().id == O.LANDMARKS
    om = O.make_model(O.LANDMARKS, 16, 8, [0.5, 2.0, 0.5])
    xs = np.random.default_rng(1).standard_normal(16)
    b = np.zeros(16)
    O.load("ref").lib.bbo_model_b(O.C.byref(om), O.C.c_double(0.0), O._p(xs), O._p(b))
    assert np.allclose(Lm.b(0.0, xs), b, rtol=1e-13, atol=1e-15)


def test_auxiliary_process_staging_matches_the_oracle(B):
    """_AuxC evaluates B~, β~, a~ at the Ralston stage times t, t + h/2, t + 3h/4 of every backward step
    (src/ode.jl:44-49,92-95) exactly as oracle.staged_aux does."""
    from bridge_jl_b200.api import _AuxC
    tt = np.linspace(0.0, 1.0, 9) ** 2
    Bf = lambda t: np.array([[-1.0 - t, 0.2], [0.1, -2.0]])
    bf = lambda t: np.array([np.sin(t), 0.5 * t])
    af = lambda t: np.array([[0.3 + t, 0.0], [0.0, 0.4]])
    A = _AuxC(B.LinearAux(Bf, bf, af), tt)
    Ao = O.staged_aux(tt, Bf, bf, af)
    assert np.array_equal(A.B, Ao.B) and np.array_equal(A.beta, Ao.beta) and np.array_equal(A.a, Ao.a)
    assert np.array_equal(A.al, Ao.a_left) and A.c.is_const == 0
    C = _AuxC(B.LinearAux(Bf(0.3), bf(0.3), af(0.3)), tt)
    assert C.c.is_const == 1 and np.array_equal(C.B, Bf(0.3))


def test_workload_definitions(B):
    import bridge_jl_b200.configs as cfg
    g = cfg.tau_grid(0.5, 1.0, 1001)                              # τ(T)(x) = x (2 - x/T), partialbridge_fitzhugh.jl:13-14
    assert g[0] == 0.5 and abs(g[-1] - 1.0) < 1e-15 and np.all(np.diff(g) > 0) and np.diff(g)[0] > np.diff(g)[-1]
    Bt, bt, at = cfg.fhn_matching_aux(1.1)
    Bo, bo = O.fhn_aux(O.AUX_FHN_MATCHING, cfg.FHN_PAR, 1.1)
    assert np.array_equal(Bt, Bo) and np.array_equal(bt, bo) and np.array_equal(at, O.fhn_a(O.FHN_HYPO, cfg.FHN_PAR))
    grids = cfg.fhn_segment_grids(101)
    assert len(grids) == 4 and all(abs(grids[k][-1] - grids[k + 1][0]) < 1e-15 for k in range(3))
    qT = np.array([[-0.6, -1.4], [1.5, -0.9], [0.8, 1.4], [-1.2, 0.7]])
    Pt = B.LandmarksTilde(0.5, 2.0, 0.5, qT)
    Bm = Pt.B(0.0)
    k01 = np.exp(-np.dot(qT[0] - qT[1], qT[0] - qT[1]) / 1.0) / np.pi
    assert np.isclose(Bm[0, 6], 0.5 * k01) and np.isclose(Bm[2, 6], -0.25 * k01) and Bm[0, 0] == 0.0
    assert np.array_equal(Pt.a(0.0), np.diag([0, 0, 4.0, 4.0] * 4)) and np.array_equal(Pt.β(0.0), np.zeros(16))


def test_solver_tags_and_sharding(B):
    from bridge_jl_b200.sharding import shard_chains
    assert B.Euler is B.EulerMaruyama and B.StochasticHeun.scheme == 2 and B.StratonovichEuler.scheme == 1
    assert shard_chains(250000, 7, 8) == (218750, 31250)
    # no device in this environment: every compute entry point refuses (the mirror never falls back to NumPy)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(B.BridgeError) as ei:
            B.sample(np.linspace(0, 1, 5), B.Wiener())
        assert ei.value.status == -10


def test_linearappr_container_and_bderiv():
    """LinearAppr / linearappr / bderiv (src/linpro.jl:181-196, src/Models.jl:49-53): the tabulated linearisation of a
    target along a trajectory -- accessors at (index, time) pairs as the reference defines them, interpolation between
    grid points for the backward solvers, and the Jacobians against finite differences."""
    import bridge_jl_b200 as B
    P = B.Lorenz([10.0, 28.0, 8.0 / 3.0], [3.0, 3.0, 3.0])
    tt = np.linspace(0.0, 0.2, 21)
    yy = np.stack([np.array([1.5, -1.5, 25.0]) + 0.1 * k * np.array([1.0, -0.5, 0.2]) for k in range(21)])
    Y = B.SamplePath(tt, yy)
    Pt = B.linearappr(Y, P)
    assert Pt.constdiff is False and Pt.is_const is False and Pt.d == 3
    for P_, x in ((P, yy[3]), (B.FitzhughDiffusion(0.1, 0.0, 1.5, 0.8, 0.3), np.array([0.3, -0.2])),
                  (B.FitzHughNagumo(0.1, 0.0, 1.5, 0.8, 0.3, 0.3), np.array([0.3, -0.2]))):
        J = B.bderiv(0.0, x, P_)
        for k in range(len(x)):
            e = np.zeros(len(x)); e[k] = 1e-6
            fd = (np.asarray(P_.b(0.0, x + e)) - np.asarray(P_.b(0.0, x - e))) / 2e-6
            assert np.allclose(J[:, k], fd, rtol=1e-6, atol=1e-6)
    i = 7
    assert np.array_equal(Pt.B((i, tt[i])), B.bderiv(tt[i], yy[i], P))
    assert np.allclose(Pt.β((i, tt[i])), P.b(tt[i], yy[i]) - B.bderiv(tt[i], yy[i], P) @ yy[i])
    assert np.allclose(Pt.a((i, tt[i])), 9.0 * np.eye(3))
    assert np.allclose(Pt.b((i, tt[i]), yy[i]), P.b(tt[i], yy[i]))            # the linearisation is exact on the trajectory
    tm = 0.5 * (tt[i] + tt[i + 1])
    assert np.allclose(Pt.B(tm), 0.5 * (Pt.Bs[i] + Pt.Bs[i + 1])) and np.array_equal(Pt.B(-1.0), Pt.Bs[0])


def test_readers_of_the_online_statistics(B):
    """mcbandmean / mcmarginalstats (src/mclog.jl:63-73, 100-111) on the arrays chain_mc_stats returns, against the
    oracle's restatement of the state they are derived from."""
    rng = np.random.default_rng(3)
    S, N, d, k = 3, 5, 2, 7
    mc = O.mcstart(np.zeros((S, N, d)))
    for _ in range(k):
        mc = O.mcnext(mc, rng.normal(size=(S, N, d)))
    mean, cov = O.mcstats(mc)
    lo, hi = B.api.mcbandmean(mean, cov, k)
    ste = np.sqrt(np.einsum("snii->sni", mc[1]) * (1.0 / (k - 1))) * np.sqrt(1.0 / k)     # src/mclog.jl:67
    assert np.allclose(lo, mean - O.MCBAND_Q * ste, rtol=1e-14) and np.allclose(hi, mean + O.MCBAND_Q * ste, rtol=1e-14)
    blo, bhi = O.mcband(mc)
    assert np.all(blo <= lo) and np.all(hi <= bhi)            # the band of the mean lies inside the band of the chain
    Xmean, Xstd = B.api.mcmarginalstats(mean, cov)
    assert Xmean.shape == (S * (N - 1) + 1, d) and Xstd.shape == Xmean.shape
    assert np.array_equal(Xmean[:N - 1], mean[0, :-1]) and np.array_equal(Xmean[N - 1], mean[1, 0])   # junction: right segment
    assert np.array_equal(Xmean[-N:], mean[-1]) and np.allclose(Xstd[-1], np.sqrt(np.diag(cov[-1, -1])))
    assert B.api.MCBAND_Q == O.MCBAND_Q
