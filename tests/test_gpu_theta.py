"""Per-chain parameters: device-built guiding tables, forward pass, pCN and the parameter-update
MH step of partialbridge_bolus3.jl:248-365, against the oracle composition in oracle/oracle.py.  GPU only.

  * guiding tables, ν(0), H⁺(0), C vs liboracle_fma: BIT-EXACT; logpdfnormal / Gamma prior (device log): 1e-13 rel;
  * paths and log-likelihoods vs liboracle_fma: BIT-EXACT; vs liboracle_ref: the tolerances of test_gpu_parity.py;
  * accept decisions: replayed exactly from the kernel's own numbers, and against the oracle's diffll;
  * with identical θ in every chain the per-chain kernels reproduce the shared-table kernels bit for bit.
"""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

XTOL = 1e-10
LLREL, LLABS = 1e-6, 1e-9
RW = np.array([0.0, 0.0, 0.02, 0.03, 0.01, 0, 0, 0])  # γ, β, σ are updated
PRIORS = {2: ("gamma", 1.0, 100.0), 3: ("gamma", 1.0, 100.0), 4: ("gamma", 2.0, 50.0)}


@pytest.fixture(scope="module")
def B():
    import bridge_jl_b200 as B
    B.default_context()
    return B


def setup(B, P, n, S, aux_kind=O.AUX_FHN_MATCHING, priors=None, diag=False, seed=11, spread=True, offset=0):
    import bridge_jl_b200.configs as cfg
    obs_t, obs_v = cfg.FHN_OBS_T[:S], cfg.FHN_OBS_V[:S]
    grids = cfg.fhn_segment_grids(n, obs_t)
    if diag:
        Pm = B.FitzHughNagumo(*cfg.FHN_PAR[:4], 0.3, 0.25)
        dp = 2
    else:
        Pm = B.FitzhughDiffusion(*cfg.FHN_PAR)
        dp = 1
    ens = B.PathEnsemble(P, S, n, 2, dp, chain_offset=offset)
    for s, g in enumerate(grids):
        ens.set_grid(s, g)
    ens.set_start(cfg.FHN_X0)
    ens.theta_attach_(Pm, cfg.FHN_L, cfg.FHN_SIGMA, cfg.FHN_EPS, obs_v, aux_kind=aux_kind, priors=priors)
    th = ens.theta()
    if spread:
        rng = np.random.default_rng(5)
        th[:, 2] += 0.2 * rng.standard_normal(P)   # γ
        th[:, 3] += 0.1 * rng.standard_normal(P)   # β
        th[:, 4] *= np.exp(0.2 * rng.standard_normal(P))  # σ
        ens.set_theta(th)
    ens.sample_(seed, 0xFFFFFFF0)
    return ens, Pm, grids, obs_v, th, dp


def oracle_left(o, Pm, th, grids, obs_v, aux_kind, priors):
    import bridge_jl_b200.configs as cfg
    npar = len(Pm.par())
    return O.theta_backward(o, Pm.model_id, th[:npar], grids, cfg.FHN_X0, cfg.FHN_L, cfg.FHN_SIGMA, cfg.FHN_EPS,
                            obs_v, aux_kind, priors)


@pytest.mark.parametrize("aux_kind,diag,n,S", [(O.AUX_FHN_MATCHING, False, 101, 3), (O.AUX_FHN_LINEARISED_END, False, 50, 2),
                                               (O.AUX_FHN_MATCHING, True, 33, 4)])
def test_theta_tables_bit_exact(B, oracle_fma, aux_kind, diag, n, S):
    P = 70
    ens, Pm, grids, obs_v, th, dp = setup(B, P, n, S, aux_kind, PRIORS, diag)
    ens.theta_guides_()
    left = ens.theta_left()
    for p in (0, 1, 31, 32, 69):
        guides, lo = oracle_left(oracle_fma, Pm, th[p], grids, obs_v, aux_kind, PRIORS)
        ν, H = ens.theta_tables(p)
        for s in range(S):
            assert np.array_equal(ν[s], guides[s].b), (p, s)
            assert np.array_equal(H[s], guides[s].A), (p, s)
        assert np.array_equal(left[p, 0:2], lo["nu"]) and np.array_equal(left[p, 2:6].reshape(2, 2), lo["Hp"])
        assert left[p, 6] == lo["C"]
        assert abs(left[p, 7] - lo["lpn"]) <= 1e-13 * abs(lo["lpn"])
        assert left[p, 8] == lo["trsum"]
        assert abs(left[p, 9] - lo["lpri"]) <= 1e-13 * abs(lo["lpri"]) + 1e-15
    ens.close()


@pytest.mark.parametrize("diag", [False, True])
def test_theta_forward_and_pcn_vs_oracle(B, oracle_fma, oracle_ref, diag):
    import bridge_jl_b200.configs as cfg
    P, n, S, seed = 96, 65, 3, 21
    ens, Pm, grids, obs_v, th, dp = setup(B, P, n, S, diag=diag, offset=500)
    ens.theta_guided_euler_ll_()
    Wc = ens.download(B.W); Xc = ens.download(B.X); ll = ens.ll
    npar = len(Pm.par())
    mid = Pm.model_id
    for p in (0, 5, 33, 95):
        # the tables are ill-conditioned in rounding (Σ = 1e-10: H of the fma and the reference arithmetic agree to
        # ~1e-6 only, as in test_gpu_parity.py::test_backward_constructors_vs_oracle), so the forward pass in
        # reference arithmetic is run on the SAME tables
        guides, _ = oracle_left(oracle_fma, Pm, th[p], grids, obs_v, O.AUX_FHN_MATCHING, None)
        for o, exact in ((oracle_fma, True), (oracle_ref, False)):
            Xo, llo, _ = O.theta_forward(o, mid, dp, th[p, :npar], guides, cfg.FHN_X0, Wc[p])
            if exact:
                assert np.array_equal(Xc[p], Xo) and ll[p] == llo
            else:
                assert np.max(np.abs(Xc[p] - Xo)) <= XTOL * (1 + np.max(np.abs(Xo)))
                assert abs(ll[p] - llo) <= LLREL * abs(llo) + LLABS
    # pCN with the chains' own tables
    rho = 0.9
    acc0 = ens.acc
    for it in range(2):
        llc = ens.ll
        ens.theta_pcn_step_(rho, seed, it)
        Wp = ens.download(B.W, which=B.PROP); Xp = ens.download(B.X, which=B.PROP)
        llp, logu, flags = ens.ll_prop, ens.logu, ens.accepted
        assert np.array_equal(flags.astype(bool), logu <= llp - llc)
        for p in (0, 5, 33, 95):
            guides, _ = oracle_left(oracle_fma, Pm, th[p], grids, obs_v, O.AUX_FHN_MATCHING, None)
            mdl = O.make_model(mid, 2, dp, th[p, :npar])
            llo, lu, Wo, Xo, _ = oracle_fma.pcn_propose(mdl, guides, cfg.FHN_X0, Wc[p], rho, seed, it, 500 + p)
            assert np.array_equal(Wp[p], Wo) and np.array_equal(Xp[p], Xo), (it, p)
            assert llp[p] == llo and logu[p] == lu
        Wc = ens.download(B.W)
        assert np.array_equal(ens.ll, np.where(flags.astype(bool), llp, llc))
    assert ens.acc - acc0 > 0
    # X of the chains that rejected is recomputed from their current W with their own tables
    Xcur = ens.download(B.X)
    for p in (0, 5, 33, 95):
        guides, _ = oracle_left(oracle_fma, Pm, th[p], grids, obs_v, O.AUX_FHN_MATCHING, None)
        Xo, llo, _ = O.theta_forward(oracle_fma, mid, dp, th[p, :npar], guides, cfg.FHN_X0, Wc[p])
        assert np.array_equal(Xcur[p], Xo) and ens.ll[p] == llo
    ens.close()


def test_theta_param_step_vs_oracle(B, oracle_fma, oracle_ref):
    import bridge_jl_b200.configs as cfg
    P, n, S, seed = 128, 81, 4, 99
    ens, Pm, grids, obs_v, th, dp = setup(B, P, n, S, priors=PRIORS, offset=7000)
    ens.theta_guided_euler_ll_()
    Wc = ens.download(B.W)
    npar = len(Pm.par()); mid = Pm.model_id
    nacc = 0
    for it in range(3):
        thc = ens.theta(); llc = ens.ll; leftc = ens.theta_left(B.CUR)
        ens.theta_param_step_(RW, seed, 100 + it)
        tho = ens.theta(B.PROP); lefto = ens.theta_left(B.PROP)
        llp, logu, flags = ens.ll_prop, ens.logu, ens.accepted.astype(bool)
        Xp = ens.download(B.X, which=B.PROP)
        # replay of the accept test from the kernel's own numbers (exact)
        diff = (lefto[:, 7] - leftc[:, 7])
        diff = diff + (llp - llc)
        diff = diff + (((lefto[:, 8] - leftc[:, 8]) + lefto[:, 9]) - leftc[:, 9])
        assert np.array_equal(flags, logu <= diff)
        assert np.array_equal(ens.theta(), np.where(flags[:, None], tho, thc))
        assert np.array_equal(ens.ll, np.where(flags, llp, llc))
        assert np.array_equal(ens.theta_left(B.CUR), np.where(flags[:, None], lefto, leftc))
        assert np.array_equal(ens.download(B.W), Wc)  # innovations are held fixed (bolus3.jl:306)
        for p in (0, 17, 64, 127):
            tp = O.theta_propose(oracle_fma, thc[p], RW, seed, 100 + it, 7000 + p)
            gc, lc = oracle_left(oracle_fma, Pm, thc[p], grids, obs_v, O.AUX_FHN_MATCHING, PRIORS)
            go, lo = oracle_left(oracle_fma, Pm, tp, grids, obs_v, O.AUX_FHN_MATCHING, PRIORS)
            for o, exact in ((oracle_fma, True), (oracle_ref, False)):
                assert np.array_equal(tp, O.theta_propose(o, thc[p], RW, seed, 100 + it, 7000 + p))
                Xo, llo, _ = O.theta_forward(o, mid, dp, tp[:npar], go, cfg.FHN_X0, Wc[p])
                _, llcur, _ = O.theta_forward(o, mid, dp, thc[p, :npar], gc, cfg.FHN_X0, Wc[p])
                d_or = O.theta_diffll(lc, lo, llcur, llo)
                lu = o.logu_q(seed, 100 + it, 7000 + p, O.Q_THETA_LOGU)
                if exact:
                    assert np.array_equal(tho[p], tp) and np.array_equal(Xp[p], Xo) and llp[p] == llo
                    assert llc[p] == llcur and logu[p] == lu
                    assert abs(diff[p] - d_or) <= 1e-12 * (1 + abs(d_or))
                else:
                    assert np.max(np.abs(Xp[p] - Xo)) <= XTOL * (1 + np.max(np.abs(Xo)))
                    assert abs(llp[p] - llo) <= LLREL * abs(llo) + LLABS
                if abs(lu - d_or) > 1e-5 * (1 + abs(llo) + abs(llcur)):  # decisions away from the boundary agree
                    assert flags[p] == (lu <= d_or)
        nacc += int(flags.sum())
    assert ens.acc_theta == nacc and 0 < nacc < 3 * P
    # the current path of every chain (rejected proposals recomputed with the tables of the CURRENT θ)
    Xcur = ens.download(B.X); thf = ens.theta()
    for p in (0, 17, 64, 127):
        g, _ = oracle_left(oracle_fma, Pm, thf[p], grids, obs_v, O.AUX_FHN_MATCHING, PRIORS)
        Xo, llo, _ = O.theta_forward(oracle_fma, mid, dp, thf[p, :npar], g, cfg.FHN_X0, Wc[p])
        assert np.array_equal(Xcur[p], Xo) and ens.ll[p] == llo
    # a pCN step after parameter steps uses the tables of the current θ
    llc = ens.ll
    ens.theta_pcn_step_(0.95, seed, 7)
    Wp = ens.download(B.W, which=B.PROP)
    for p in (0, 17, 64, 127):
        g, _ = oracle_left(oracle_fma, Pm, thf[p], grids, obs_v, O.AUX_FHN_MATCHING, PRIORS)
        mdl = O.make_model(mid, 2, dp, thf[p, :npar])
        llo, lu, Wo, Xo, _ = oracle_fma.pcn_propose(mdl, g, cfg.FHN_X0, Wc[p], 0.95, seed, 7, 7000 + p)
        assert np.array_equal(Wp[p], Wo) and ens.ll_prop[p] == llo
    ens.close()


def test_theta_equal_parameters_reproduce_shared_tables(B):
    """All chains at the model's θ: the per-chain kernels must give what the shared-table kernels give."""
    import bridge_jl_b200.configs as cfg
    P, n, seed = 300, 129, 3
    ens, Pm, grids, obs_v, th, dp = setup(B, P, n, 4, spread=False)
    # the per-chain kernels compute in fused order (fp64-pipe bound); the shared-table constructors match them bit for
    # bit in that order (their default is reference arithmetic)
    ctx = B.default_context()
    ctx.set_arith(B.api.K.ARITH_FUSED)
    try:
        Pm2, guides, x0, rho = cfg.fhn_config4(n)
    finally:
        ctx.set_arith(B.api.K.ARITH_REFERENCE)
    ref = B.PathEnsemble(P, 4, n, 2, 1)
    for s, g in enumerate(guides):
        ref.set_grid(s, g.tt)
    ref.set_start(x0)
    ref.sample_(11, 0xFFFFFFF0)
    assert np.array_equal(ref.download(B.W), ens.download(B.W))
    ref.guided_euler_ll_(Pm2, guides)
    ens.theta_guided_euler_ll_()
    ν, H = ens.theta_tables(123)
    for s, g in enumerate(guides):
        assert np.array_equal(ν[s], g.ν) and np.array_equal(H[s], g.H)
    assert np.array_equal(ref.ll, ens.ll) and np.array_equal(ref.download(B.X), ens.download(B.X))
    for it in range(3):
        ref.pcn_step_(Pm2, guides, rho, seed, it)
        ens.theta_pcn_step_(rho, seed, it)
        assert np.array_equal(ref.accepted, ens.accepted) and np.array_equal(ref.ll_prop, ens.ll_prop)
        assert np.array_equal(ref.download(B.W, which=B.PROP), ens.download(B.W, which=B.PROP))
        assert np.array_equal(ref.download(B.X, which=B.PROP), ens.download(B.X, which=B.PROP))
    assert ref.acc == ens.acc
    assert np.array_equal(ref.download(B.X), ens.download(B.X))
    ref.close(); ens.close()


def test_theta_rejects_impossible_parameters(B):
    """σ° <= 0 under a Gamma prior has logπ = -Inf: never accepted; NaN log-likelihoods never accepted."""
    P = 64
    ens, Pm, grids, obs_v, th, dp = setup(B, P, 41, 2, priors=PRIORS)
    ens.theta_guided_euler_ll_()
    th0 = ens.theta()
    ens.theta_param_step_(np.array([0, 0, 0, 0, 5.0, 0, 0, 0]), 1, 0)  # huge steps in σ: about half go negative
    tho = ens.theta(B.PROP); flags = ens.accepted.astype(bool)
    assert (tho[:, 4] <= 0).any()
    assert not flags[tho[:, 4] <= 0].any()
    assert np.array_equal(ens.theta()[~flags], th0[~flags])
    ens.close()


def test_theta_errors(B):
    import bridge_jl_b200.configs as cfg
    ens = B.PathEnsemble(8, 1, 9, 3, 3)
    with pytest.raises(B.BridgeError) as ei:
        ens.theta_attach_(B.Lorenz((10.0, 28.0, 8 / 3), (1.0, 1.0, 1.0)), np.eye(3), np.eye(3), 1e-3, [[0, 0, 0]])
    assert ei.value.status == -11  # BB_ERR_UNSUPPORTED
    with pytest.raises(B.BridgeError):
        ens.theta_guides_()  # nothing attached
    ens.close()
    ens = B.PathEnsemble(8, 1, 9, 2, 1)
    ens.theta_attach_(B.FitzhughDiffusion(*cfg.FHN_PAR), cfg.FHN_L, cfg.FHN_SIGMA, cfg.FHN_EPS, [0.5])
    with pytest.raises(B.BridgeError):
        ens.theta_guides_()  # no grid set
    with pytest.raises(B.BridgeError):
        ens.theta_param_step_(np.array([1, 1, 1, 1, 1.0, 0, 0, 0]), 1, 0)  # more than 4 updated parameters
    ens.close()


@pytest.mark.parametrize("P,n,S,diag", [(37, 18, 1, False), (1, 33, 2, False), (259, 21, 3, True)])
def test_theta_ragged_sizes_skip_and_no_x(B, oracle_fma, P, n, S, diag):
    """Odd numbers of chains (a lane pair with one chain, a single chain), N not a multiple of 16, llikelihood with
    skip > 0, proposals whose X° is not stored (recomputed on demand), the linearised auxiliary process."""
    import bridge_jl_b200.configs as cfg
    aux = O.AUX_FHN_LINEARISED_END
    ens, Pm, grids, obs_v, th, dp = setup(B, P, n, S, aux_kind=aux, priors=PRIORS, diag=diag, offset=11)
    skip, seed = 2, 8
    npar = len(Pm.par()); mid = Pm.model_id
    ens.theta_guided_euler_ll_(skip=skip, store_x=False)
    Wc = ens.download(B.W); ll0 = ens.ll
    chains = sorted({0, P // 2, P - 1})
    for p in chains:
        g, _ = oracle_left(oracle_fma, Pm, th[p], grids, obs_v, aux, PRIORS)
        _, llo, _ = O.theta_forward(oracle_fma, mid, dp, th[p, :npar], g, cfg.FHN_X0, Wc[p], skip=skip)
        assert ll0[p] == llo
    ens.theta_param_step_(RW, seed, 50, skip=skip, store_x=False)
    tho = ens.theta(B.PROP); flags = ens.accepted.astype(bool); llp = ens.ll_prop
    for p in chains:
        tp = O.theta_propose(oracle_fma, th[p], RW, seed, 50, 11 + p)
        assert np.array_equal(tho[p], tp)
        g, _ = oracle_left(oracle_fma, Pm, tp, grids, obs_v, aux, PRIORS)
        _, llo, _ = O.theta_forward(oracle_fma, mid, dp, tp[:npar], g, cfg.FHN_X0, Wc[p], skip=skip)
        assert llp[p] == llo
    ens.theta_pcn_step_(0.8, seed, 51, skip=skip, store_x=False)
    thf = ens.theta(); Wf = ens.download(B.W); Xf = ens.download(B.X)  # X is recomputed for every chain that needs it
    for p in chains:
        g, _ = oracle_left(oracle_fma, Pm, thf[p], grids, obs_v, aux, PRIORS)
        Xo, llo, _ = O.theta_forward(oracle_fma, mid, dp, thf[p, :npar], g, cfg.FHN_X0, Wf[p], skip=skip)
        assert np.array_equal(Xf[p], Xo) and ens.ll[p] == llo
    ens.close()


# ----------------------------------------------------------------------------------------------- the model of bolus3.jl itself
BOLUS_PAR = (70.0 / 0.6, 8.0, 15.0 / (20 * 0.6), 1.5 * 15.0 / 15.0, 0.5, 0.2)  # α, β(init 8.0), λ, μ, σ1(init .5), σ2  :74-77,152-154
BOLUS_L = np.array([[0.5, 0.5]])                                                 # :31
BOLUS_PRIORS = {1: ("gamma", 1.0, 100.0), 4: ("gamma", 1.0, 100.0)}              # logπ, :237
BOLUS_RW = np.array([0.0, 0.02, 0.0, 0.0, 0.02, 0.0, 0, 0])                      # propose(.02, P): β and σ1, :239-242,272


def test_theta_bolus_model_parameter_update_vs_oracle(B, oracle_fma, oracle_ref):
    """The parameter-update branch of partialbridge_bolus3.jl with ITS model: time-dependent drift α dose(t), auxiliary
    process DiffusionAux (β~(t) evaluated at the Ralston stage times on the device, a~ = diag(σ1², σ2²) != a), L = [.5 .5],
    Σ = 1e-4, ϵ = 1e-3, Gamma(1,100) priors on β and σ1, random-walk proposals of sd 0.02 on both."""
    P, n, S, seed = 96, 61, 3, 12
    obs_t = (0.8, 1.7, 2.5); obs_v = (4.0, 9.0, 12.0)
    tcut = (0.0,) + obs_t
    grids = []
    for k in range(S):
        s = np.linspace(0.0, tcut[k + 1] - tcut[k], n)
        grids.append(tcut[k] + s * (2 - s / (tcut[k + 1] - tcut[k])))  # τ(t, T0, Tend), :157
    x0 = np.array([0.5, 0.2])
    Pm = B.BolusDiffusion(*BOLUS_PAR)
    ens = B.PathEnsemble(P, S, n, 2, 2, chain_offset=40)
    for s_, g in enumerate(grids):
        ens.set_grid(s_, g)
    ens.set_start(x0)
    ens.theta_attach_(Pm, BOLUS_L, 1e-4 * np.eye(1), 1e-3, obs_v, aux_kind=O.AUX_BOLUS, priors=BOLUS_PRIORS)
    th = ens.theta()
    rng = np.random.default_rng(2)
    th[:, 1] += 0.5 * rng.standard_normal(P); th[:, 4] *= np.exp(0.1 * rng.standard_normal(P))
    ens.set_theta(th)
    ens.sample_(seed, 0xFFFFFFF0)
    ens.theta_guided_euler_ll_()
    Wc = ens.download(B.W); Xc = ens.download(B.X); llc = ens.ll; leftc = ens.theta_left()

    def orc_left(o, par):
        return O.theta_backward(o, O.BOLUS, par[:6], grids, x0, BOLUS_L, 1e-4 * np.eye(1), 1e-3, obs_v, O.AUX_BOLUS,
                                BOLUS_PRIORS)

    chains = (0, 31, 32, 95)
    for p in chains:
        g, lo = orc_left(oracle_fma, th[p])
        ν, H = ens.theta_tables(p)
        for s_ in range(S):
            assert np.array_equal(ν[s_], g[s_].b) and np.array_equal(H[s_], g[s_].A), (p, s_)
        assert np.array_equal(leftc[p, 0:2], lo["nu"]) and np.array_equal(leftc[p, 2:6].reshape(2, 2), lo["Hp"])
        assert leftc[p, 6] == lo["C"] and leftc[p, 8] == lo["trsum"]
        assert abs(leftc[p, 7] - lo["lpn"]) <= 1e-13 * abs(lo["lpn"]) and abs(leftc[p, 9] - lo["lpri"]) <= 1e-13
        for o, exact in ((oracle_fma, True), (oracle_ref, False)):
            Xo, llo, _ = O.theta_forward(o, O.BOLUS, 2, th[p, :6], g, x0, Wc[p])
            if exact:
                assert np.array_equal(Xc[p], Xo) and llc[p] == llo
            else:
                assert np.max(np.abs(Xc[p] - Xo)) <= XTOL * (1 + np.max(np.abs(Xo)))
                assert abs(llc[p] - llo) <= LLREL * abs(llo) + LLABS
    nacc = 0
    for it in range(3):
        thc = ens.theta(); llc = ens.ll; leftc = ens.theta_left(B.CUR)
        ens.theta_param_step_(BOLUS_RW, seed, 200 + it)
        tho = ens.theta(B.PROP); lefto = ens.theta_left(B.PROP)
        llp, logu, flags = ens.ll_prop, ens.logu, ens.accepted.astype(bool)
        diff = (lefto[:, 7] - leftc[:, 7]) + (llp - llc)
        diff = diff + (((lefto[:, 8] - leftc[:, 8]) + lefto[:, 9]) - leftc[:, 9])
        assert np.array_equal(flags, logu <= diff)
        Xp = ens.download(B.X, which=B.PROP)
        for p in chains:
            tp = O.theta_propose(oracle_fma, thc[p], BOLUS_RW, seed, 200 + it, 40 + p)
            assert np.array_equal(tho[p], tp)
            g, lo = orc_left(oracle_fma, tp)
            Xo, llo, _ = O.theta_forward(oracle_fma, O.BOLUS, 2, tp[:6], g, x0, Wc[p])
            assert np.array_equal(Xp[p], Xo) and llp[p] == llo
            assert logu[p] == oracle_fma.logu_q(seed, 200 + it, 40 + p, O.Q_THETA_LOGU)
        nacc += int(flags.sum())
    assert ens.acc_theta == nacc and 0 < nacc < 3 * P
    # pCN with the tables of the chains' current parameters (ρ = 0 in the script, :28: an independence sampler)
    thf = ens.theta(); llc = ens.ll
    ens.theta_pcn_step_(0.0, seed, 300)
    Wp = ens.download(B.W, which=B.PROP)
    for p in chains:
        g, _ = orc_left(oracle_fma, thf[p])
        mdl = O.make_model(O.BOLUS, 2, 2, thf[p, :6])
        llo, lu, Wo, Xo, _ = oracle_fma.pcn_propose(mdl, g, x0, Wc[p], 0.0, seed, 300, 40 + p)
        assert np.array_equal(Wp[p], Wo) and ens.ll_prop[p] == llo and ens.logu[p] == lu
    # this model is not on the shared-table path (its drift depends on t)
    with pytest.raises(B.BridgeError) as ei:
        ens.euler_(Pm)
    assert ei.value.status == -11
    ens.close()


def test_theta_joint_start_point_update(B, oracle_fma):
    """bolus3.jl:311-318: in a parameter step the starting point is proposed jointly with probability 1/2,
    x0° = x0 + 0.1 (u, -u); logpdfnormal(x0° - ν°(0), H⁺°(0)) - logpdfnormal(x0 - ν(0), H⁺(0)) enters diffll (:319)."""
    P, n, S, seed = 64, 41, 2, 31
    obs_t = (0.8, 1.7); obs_v = (4.0, 9.0)
    tcut = (0.0,) + obs_t
    grids = []
    for k in range(S):
        s = np.linspace(0.0, tcut[k + 1] - tcut[k], n)
        grids.append(tcut[k] + s * (2 - s / (tcut[k + 1] - tcut[k])))
    rng = np.random.default_rng(4)
    x0 = np.array([0.5, 0.2]) + 0.05 * rng.standard_normal((P, 2))
    Pm = B.BolusDiffusion(*BOLUS_PAR)
    ens = B.PathEnsemble(P, S, n, 2, 2, chain_offset=900)
    for s_, g in enumerate(grids):
        ens.set_grid(s_, g)
    ens.set_start(x0)
    ens.theta_attach_(Pm, BOLUS_L, 1e-2 * np.eye(1), 0.1, obs_v, aux_kind=O.AUX_BOLUS, priors=BOLUS_PRIORS,
                      start_sd=0.1, start_dir=[1.0, -1.0])
    ens.sample_(seed, 0xFFFFFFF0)
    ens.theta_guided_euler_ll_()
    Wc = ens.download(B.W)
    moved = 0
    for it in range(2):
        thc = ens.theta(); llc = ens.ll; leftc = ens.theta_left(B.CUR); x0c = ens.theta_start(B.CUR)
        ens.theta_param_step_(BOLUS_RW, seed, 70 + it)
        tho = ens.theta(B.PROP); lefto = ens.theta_left(B.PROP); x0o = ens.theta_start(B.PROP)
        llp, logu, flags = ens.ll_prop, ens.logu, ens.accepted.astype(bool)
        diff = (lefto[:, 7] - leftc[:, 7]) + (llp - llc)
        diff = diff + (((lefto[:, 8] - leftc[:, 8]) + lefto[:, 9]) - leftc[:, 9])
        assert np.array_equal(flags, logu <= diff)
        assert np.array_equal(ens.theta_start(B.CUR), np.where(flags[:, None], x0o, x0c))
        Xp = ens.download(B.X, which=B.PROP)
        assert np.array_equal(Xp[:, 0, 0], x0o)  # the proposal path starts at x0°
        for p in (0, 13, 63):
            xs = O.theta_propose_start(oracle_fma, x0c[p], 0.1, [1.0, -1.0], seed, 70 + it, 900 + p)
            assert np.array_equal(x0o[p], xs)
            tp = O.theta_propose(oracle_fma, thc[p], BOLUS_RW, seed, 70 + it, 900 + p)
            g, lo = O.theta_backward(oracle_fma, O.BOLUS, tp[:6], grids, xs, BOLUS_L, 1e-2 * np.eye(1), 0.1, obs_v,
                                     O.AUX_BOLUS, BOLUS_PRIORS)
            Xo, llo, _ = O.theta_forward(oracle_fma, O.BOLUS, 2, tp[:6], g, xs, Wc[p])
            assert np.array_equal(Xp[p], Xo) and llp[p] == llo
            assert abs(lefto[p, 7] - lo["lpn"]) <= 1e-13 * abs(lo["lpn"])
        moved += int(np.sum(np.any(x0o != x0c, axis=1)))
    assert 0.25 * 2 * P < moved < 0.75 * 2 * P  # about half of the proposals move the starting point
    # current paths start at the chains' current starting points
    Xc = ens.download(B.X)
    assert np.array_equal(Xc[:, 0, 0], ens.theta_start(B.CUR))
    # a broadcast starting point cannot be moved per chain
    e2 = B.PathEnsemble(4, S, n, 2, 2)
    for s_, g in enumerate(grids):
        e2.set_grid(s_, g)
    e2.set_start([0.5, 0.2])
    e2.theta_attach_(Pm, BOLUS_L, 1e-2 * np.eye(1), 0.1, obs_v, aux_kind=O.AUX_BOLUS, start_sd=0.1, start_dir=[1.0, -1.0])
    e2.sample_(1, 0); e2.theta_guided_euler_ll_()
    with pytest.raises(B.BridgeError) as ei:
        e2.theta_param_step_(BOLUS_RW, 1, 0)
    assert ei.value.status == -3
    e2.close(); ens.close()


def test_theta_mcmc_loop_recovers_plausible_parameters(B):
    """The outer loop of bolus3.jl:248-365 (parameter updates and pCN updates alternating at random) runs for many
    chains at once; a statistical smoke check: chains move, both kinds of proposals are accepted at sane rates, the
    log-likelihoods stay finite, θ stays in the support of its priors."""
    P, n, S = 512, 41, 3
    obs_t = (0.8, 1.7, 2.5); obs_v = (4.0, 9.0, 12.0)
    tcut = (0.0,) + obs_t
    grids = []
    for k in range(S):
        s = np.linspace(0.0, tcut[k + 1] - tcut[k], n)
        grids.append(tcut[k] + s * (2 - s / (tcut[k + 1] - tcut[k])))
    Pm = B.BolusDiffusion(*BOLUS_PAR)
    ens = B.PathEnsemble(P, S, n, 2, 2)
    for s_, g in enumerate(grids):
        ens.set_grid(s_, g)
    ens.set_start(np.tile([0.5, 0.2], (P, 1)))
    ens.theta_attach_(Pm, BOLUS_L, 1e-2 * np.eye(1), 0.1, obs_v, aux_kind=O.AUX_BOLUS, priors=BOLUS_PRIORS,
                      start_sd=0.1, start_dir=[1.0, -1.0])
    ens.sample_(5, 0xFFFFFFF0)
    ens.theta_guided_euler_ll_()
    th0 = ens.theta()
    seen = []
    acc, acct = B.theta_mcmc_(ens, 0.7, BOLUS_RW, 40, 5, callback=lambda it, e: seen.append(it))
    assert seen == list(range(40))
    th = ens.theta()
    assert np.all(np.isfinite(ens.ll)) and np.all(th[:, 1] > 0) and np.all(th[:, 4] > 0)
    assert np.array_equal(th[:, [0, 2, 3, 5]], th0[:, [0, 2, 3, 5]])  # only β and σ1 are updated
    assert np.std(th[:, 1]) > 0.01 and np.std(th[:, 4]) > 0.005       # chains have moved apart
    assert 0 < acc and 0 < acct
    X = ens.download(B.X)
    assert np.all(np.isfinite(X))
    ens.close()
