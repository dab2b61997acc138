"""Blocked segment updates of the multi-segment sampler (project_partialbridge/partialbridge_bolus3.jl:258-355, the
`updateparams == false` branch) -- bb_theta_block_step against the oracle composition oracle.theta_block_step.  GPU only.

  * W°, X°, per-segment ll_temp / ll°, log U vs liboracle_fma: BIT-EXACT; the Gaussian start terms (device log): 1e-13 rel;
  * diffll and the accept decisions: replayed exactly from the kernel's own numbers, in the script's summation order;
  * state after the step: the block's segments of W, X (and x0) of the accepting chains are the proposal, everything
    else is untouched;
  * the bolus configuration also against liboracle_ref (reference arithmetic, its own tables): X° within 1e-6 (1 + |X|),
    ll within 1e-6 relative.
"""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

# reference arithmetic builds its OWN tables here (no explicit fma), the device's per-chain constructor is in fused order:
# the bolus tables differ by rounding amplified through the backward recursion, paths by 1.4e-7 absolute at |X| ~ 13
# (measured); the contract is 1e-6 relative on ll
XTOL = 1e-6
LLREL, LLABS = 1e-6, 1e-9


@pytest.fixture(scope="module")
def B():
    import bridge_jl_b200 as B
    B.default_context()
    return B


def replay(blk, s_lo, s_hi):
    """diffll from the kernel's own per-segment numbers, in the script's order (start term, then i in ind descending)"""
    diff = blk[:, 1] - blk[:, 0]
    for s in range(s_hi - 1, s_lo - 1, -1):
        diff = diff + (blk[:, 5 + 2 * s + 1] - blk[:, 5 + 2 * s])
    return diff


def check_block(B, ens, o, oref, mid, dp, npar, grids, L, Sigma, eps, obs_v, aux_kind, s_lo, s_hi, rho, hzero, seed, it,
                offset, start_sd, start_dir, chains):
    S = len(grids)
    Wc = ens.download(B.W); Xc = ens.download(B.X); x0c = ens.theta_start(B.CUR); th = ens.theta()
    acc0 = ens.acc
    ens.theta_block_step_(s_lo, s_hi, rho, seed, it, hzero=hzero)
    blk = ens.theta_block(); flags = ens.accepted.astype(bool); logu = ens.logu
    Wp = ens.download(B.W, which=B.PROP)
    Wn = ens.download(B.W); Xn = ens.download(B.X); x0n = ens.theta_start(B.CUR)
    # accept decisions, exactly, from the kernel's own numbers
    diff = replay(blk, s_lo, s_hi)
    assert np.array_equal(diff, blk[:, 4])
    assert np.array_equal(flags, logu <= diff)
    assert ens.acc - acc0 == int(flags.sum())
    # state: only the block's segments of the accepting chains change
    out = [s for s in range(S) if not (s_lo <= s < s_hi)]
    assert np.array_equal(Wn[:, out], Wc[:, out]) and np.array_equal(Xn[:, out], Xc[:, out])
    assert np.array_equal(Wn[~flags], Wc[~flags]) and np.array_equal(Xn[~flags], Xc[~flags])
    assert np.array_equal(Wn[flags][:, s_lo:s_hi], Wp[flags][:, s_lo:s_hi])
    assert np.array_equal(x0n[~flags], x0c[~flags])
    if s_lo > 0:
        assert np.array_equal(x0n, x0c)
        # the accepted block starts where the current path is
        assert np.array_equal(Xn[flags][:, s_lo, 0], Xc[flags][:, s_lo - 1, -1])
    for s in range(s_lo + 1, s_hi):  # segments of a block are chained by their end points
        assert np.array_equal(Xn[flags][:, s, 0], Xn[flags][:, s - 1, -1])
    for p in chains:
        r = O.theta_block_step(o, mid, dp, th[p, :npar], grids, L, Sigma, eps, obs_v, aux_kind, s_lo, s_hi, hzero,
                               x0c[p], Xc[p], Wc[p], rho, seed, it, offset + p, start_sd, start_dir)
        for s in range(s_lo, s_hi):
            assert np.array_equal(Wp[p, s], r["Wo"][s]), (p, s)
            assert blk[p, 5 + 2 * s] == r["llt"][s] and blk[p, 5 + 2 * s + 1] == r["llo"][s], (p, s)
            if flags[p]:
                assert np.array_equal(Xn[p, s], r["Xo"][s]), (p, s)
        assert logu[p] == r["logu"]
        assert abs(blk[p, 0] - r["lpn"]) <= 1e-13 * abs(r["lpn"]) and abs(blk[p, 1] - r["lpno"]) <= 1e-13 * abs(r["lpno"])
        assert abs(blk[p, 4] - r["diff"]) <= 1e-12 * (1 + abs(r["diff"]))
        if flags[p] and s_lo == 0:
            assert np.array_equal(x0n[p], r["x0o"])
        if oref is not None:
            rr = O.theta_block_step(oref, mid, dp, th[p, :npar], grids, L, Sigma, eps, obs_v, aux_kind, s_lo, s_hi, hzero,
                                    x0c[p], Xc[p], Wc[p], rho, seed, it, offset + p, start_sd, start_dir)
            for s in range(s_lo, s_hi):
                assert np.max(np.abs(r["Xo"][s] - rr["Xo"][s])) <= XTOL * (1 + np.max(np.abs(rr["Xo"][s])))
                for a, b in ((blk[p, 5 + 2 * s], rr["llt"][s]), (blk[p, 5 + 2 * s + 1], rr["llo"][s])):
                    assert abs(a - b) <= LLREL * abs(b) + LLABS
            if abs(rr["logu"] - rr["diff"]) > 1e-5 * (1 + np.sum(np.abs(rr["llo"])) + np.sum(np.abs(rr["llt"]))):
                assert flags[p] == (rr["logu"] <= rr["diff"])
    return flags


def test_block_updates_fhn_vs_oracle(B, oracle_fma):
    """FitzHugh-Nagumo (hypoelliptic), per-chain θ and starting points, 4 segments: middle, left, right, whole-chain and
    single-segment blocks."""
    import bridge_jl_b200.configs as cfg
    P, n, S, seed, offset = 96, 49, 4, 17, 3000
    obs_t, obs_v = cfg.FHN_OBS_T[:S], cfg.FHN_OBS_V[:S]
    grids = cfg.fhn_segment_grids(n, obs_t)
    Pm = B.FitzhughDiffusion(*cfg.FHN_PAR)
    ens = B.PathEnsemble(P, S, n, 2, 1, chain_offset=offset)
    for s, g in enumerate(grids):
        ens.set_grid(s, g)
    rng = np.random.default_rng(8)
    ens.set_start(np.asarray(cfg.FHN_X0) + 0.02 * rng.standard_normal((P, 2)))
    # a milder observation scheme than the benchmark's (Σ = 1e-10), so that block proposals are accepted at a useful rate
    Sigma, eps = 1e-2 * np.eye(1), 1e-1
    ens.theta_attach_(Pm, cfg.FHN_L, Sigma, eps, obs_v, start_sd=0.1, start_dir=[1.0, -1.0])
    th = ens.theta()
    th[:, 2] += 0.2 * rng.standard_normal(P); th[:, 4] *= np.exp(0.2 * rng.standard_normal(P))
    ens.set_theta(th)
    ens.sample_(seed, 0xFFFFFFF0)
    ens.theta_guided_euler_ll_()
    npar = len(Pm.par())
    total = np.zeros(P, dtype=int)
    blocks = [(1, 3), (0, 2), (2, 4), (0, 4), (3, 4), (0, 1), (1, 2)]
    for it, (lo, hi) in enumerate(blocks):
        total += check_block(B, ens, oracle_fma, None, Pm.model_id, 1, npar, grids, cfg.FHN_L, Sigma, eps, obs_v,
                             O.AUX_FHN_MATCHING, lo, hi, 0.9, 0.1, seed, 40 + it, offset, 0.1, [1.0, -1.0],
                             (0, 31, 32, 95))
    assert 0 < total.sum() < P * len(blocks)
    ens.close()


def bolus_setup(B, P, n, S, offset, start_sd=0.1):
    from tests.test_gpu_theta import BOLUS_PAR, BOLUS_L
    obs_t = (0.8, 1.7, 2.5, 3.1)[:S]; obs_v = (4.0, 9.0, 12.0, 13.0)[:S]
    tcut = (0.0,) + obs_t
    grids = []
    for k in range(S):
        s = np.linspace(0.0, tcut[k + 1] - tcut[k], n)
        grids.append(tcut[k] + s * (2 - s / (tcut[k + 1] - tcut[k])))
    Pm = B.BolusDiffusion(*BOLUS_PAR)
    ens = B.PathEnsemble(P, S, n, 2, 2, chain_offset=offset)
    for s_, g in enumerate(grids):
        ens.set_grid(s_, g)
    rng = np.random.default_rng(3)
    ens.set_start(np.array([0.5, 0.2]) + 0.05 * rng.standard_normal((P, 2)))
    ens.theta_attach_(Pm, BOLUS_L, 1e-2 * np.eye(1), 0.1, obs_v, aux_kind=O.AUX_BOLUS, start_sd=start_sd,
                      start_dir=[1.0, -1.0])
    return ens, Pm, grids, obs_v, BOLUS_L


def test_block_updates_bolus_vs_oracle(B, oracle_fma, oracle_ref):
    """The script's own model (time-dependent drift, DiffusionAux, d' = 2), also against reference arithmetic."""
    P, n, S, seed, offset = 70, 37, 3, 5, 100
    ens, Pm, grids, obs_v, L = bolus_setup(B, P, n, S, offset)
    ens.sample_(seed, 0xFFFFFFF0)
    ens.theta_guided_euler_ll_()
    total = 0
    for it, (lo, hi) in enumerate([(0, 1), (1, 3), (0, 3), (2, 3), (1, 2)]):
        total += int(check_block(B, ens, oracle_fma, oracle_ref, O.BOLUS, 2, 6, grids, L, 1e-2 * np.eye(1), 0.1, obs_v,
                                 O.AUX_BOLUS, lo, hi, 0.8, 0.1, seed, 7 + it, offset, 0.1, [1.0, -1.0],
                                 (0, 33, 69)).sum())
    assert total > 0
    ens.close()


def test_blocked_sweeps_run_the_script_loop(B):
    """`while !finished` of bolus3.jl:258-362: blocks drawn by the host cover 1 .. obsnum exactly once per sweep;
    interleaved with whole-path pCN and parameter steps the sampler keeps finite paths and sane acceptance rates."""
    from tests.test_gpu_theta import BOLUS_RW
    P, n, S = 256, 33, 4
    ens, Pm, grids, obs_v, L = bolus_setup(B, P, n, S, 0)
    ens.sample_(9, 0xFFFFFFF0)
    ens.theta_guided_euler_ll_()
    rng = np.random.default_rng(1)
    nblocks, acc0, it = 0, ens.acc, 0
    for sweep in range(12):
        blocks = ens.theta_blocked_sweep_(rng, 0.7, 9, it)
        assert blocks[0][0] == 0 and blocks[-1][1] == S
        assert all(blocks[k][1] == blocks[k + 1][0] for k in range(len(blocks) - 1))
        nblocks += len(blocks); it += len(blocks)
        if sweep % 4 == 3:  # back to whole-path steps: the library re-establishes the running ll first
            ens.theta_param_step_(BOLUS_RW, 9, it); it += 1
            llc = ens.ll
            ens.theta_pcn_step_(0.7, 9, it); it += 1
            assert np.array_equal(ens.accepted.astype(bool), ens.logu <= ens.ll_prop - llc)
    rate = (ens.acc - acc0) / (nblocks * P + 3 * P)
    assert 0.05 < rate < 0.98
    X = ens.download(B.X)
    assert np.all(np.isfinite(X)) and np.array_equal(X[:, 0, 0], ens.theta_start(B.CUR))
    # the outer loop with block sweeps in place of whole-path pCN updates
    acc1, acct1 = ens.acc, ens.acc_theta
    acc, acct = B.theta_mcmc_(ens, 0.7, BOLUS_RW, 16, 9, first_iter=1000, blocked=True)
    assert acc > acc1 and acct > acct1 and np.all(np.isfinite(ens.download(B.X)))
    ens.close()


def test_block_update_size_independent_properties(B):
    """At a size the oracle would not finish (2e4 chains x 4 segments x N = 513): with ρ = 1 and no start move the proposal
    IS the current noise re-run under the block's guide, so ll° = ll_temp segment by segment, diffll = 0 exactly, every
    chain accepts, W is unchanged bit for bit and X of the block becomes that re-run path (continuous inside the block and
    at its left end); the acceptance counter is the sum of the flags; blocks that end the chain reproduce, for a chain whose
    θ is the common one, the whole-chain tables."""
    import bridge_jl_b200.configs as cfg
    P, n, S = 20000, 513, 4
    grids = cfg.fhn_segment_grids(n)
    Pm = B.FitzhughDiffusion(*cfg.FHN_PAR)
    ens = B.PathEnsemble(P, S, n, 2, 1)
    for s, g in enumerate(grids):
        ens.set_grid(s, g)
    ens.set_start(cfg.FHN_X0)
    ens.theta_attach_(Pm, cfg.FHN_L, 1e-2 * np.eye(1), 1e-1, cfg.FHN_OBS_V)   # start_sd = 0: no start move
    ens.sample_(3, 0xFFFFFFF0)
    ens.theta_guided_euler_ll_()
    W0 = ens.download(B.W, p0=0, np_=512); X0 = ens.download(B.X, p0=0, np_=512)
    acc0 = ens.acc
    for it, (lo, hi) in enumerate([(1, 3), (0, 2), (0, 4)]):
        ens.theta_block_step_(lo, hi, 1.0, 3, it)
        blk = ens.theta_block()
        for s in range(lo, hi):
            assert np.array_equal(blk[:, 5 + 2 * s], blk[:, 5 + 2 * s + 1])
        assert np.all(blk[:, 4] == 0.0) and np.all(ens.accepted == 1)
        assert ens.acc - acc0 == P * (it + 1)
        Wn = ens.download(B.W, p0=0, np_=512); Xn = ens.download(B.X, p0=0, np_=512)
        assert np.array_equal(Wn, W0)
        out = [s for s in range(S) if not (lo <= s < hi)]
        assert np.array_equal(Xn[:, out], X0[:, out])
        for s in range(max(lo, 1), hi):
            assert np.array_equal(Xn[:, s, 0], Xn[:, s - 1, -1])
        X0 = Xn
    # the whole-chain block used the whole-chain tables: X equals a fresh solve! with the same noise
    ens.theta_guided_euler_ll_()
    assert np.array_equal(ens.download(B.X, p0=0, np_=512), X0)
    # a genuine proposal at this size: decisions replay exactly, only accepting chains' block rows change
    ens.theta_block_step_(1, 4, 0.9, 3, 10)
    blk = ens.theta_block(); flags = ens.accepted.astype(bool)
    assert np.array_equal(flags, ens.logu <= replay(blk, 1, 4)) and 0 < flags.sum() < P
    Xn = ens.download(B.X, p0=0, np_=512)
    assert np.array_equal(Xn[~flags[:512]], X0[~flags[:512]]) and np.array_equal(Xn[:, 0], X0[:, 0])
    ens.close()


def test_block_step_errors(B):
    ens, Pm, grids, obs_v, L = bolus_setup(B, 8, 17, 3, 0)
    ens.sample_(1, 0); ens.theta_guided_euler_ll_()
    for lo, hi in ((-1, 2), (2, 2), (1, 4), (2, 1)):
        with pytest.raises(B.BridgeError) as ei:
            ens.theta_block_step_(lo, hi, 0.5, 1, 0)
        assert ei.value.status == -7
    with pytest.raises(B.BridgeError):
        ens.theta_block_step_(0, 1, 0.5, 1, 0, hzero=0.0)
    ens.close()
    # a broadcast starting point cannot be moved per chain (blocks that contain the first segment)
    from tests.test_gpu_theta import BOLUS_PAR, BOLUS_L
    e2 = B.PathEnsemble(4, 3, 17, 2, 2)
    for s_, g in enumerate(grids):
        e2.set_grid(s_, g)
    e2.set_start([0.5, 0.2])
    e2.theta_attach_(B.BolusDiffusion(*BOLUS_PAR), BOLUS_L, 1e-2 * np.eye(1), 0.1, obs_v, aux_kind=O.AUX_BOLUS,
                     start_sd=0.1, start_dir=[1.0, -1.0])
    e2.sample_(1, 0); e2.theta_guided_euler_ll_()
    with pytest.raises(B.BridgeError) as ei:
        e2.theta_block_step_(0, 2, 0.5, 1, 0)
    assert ei.value.status == -3
    e2.theta_block_step_(1, 3, 0.5, 1, 0)  # blocks that leave the starting point alone are fine
    e2.close()
    # paths must be stored
    e3 = B.PathEnsemble(4, 3, 17, 2, 2, store_x=False)
    for s_, g in enumerate(grids):
        e3.set_grid(s_, g)
    e3.set_start([0.5, 0.2])
    e3.theta_attach_(B.BolusDiffusion(*BOLUS_PAR), BOLUS_L, 1e-2 * np.eye(1), 0.1, obs_v, aux_kind=O.AUX_BOLUS)
    e3.sample_(1, 0)
    with pytest.raises(B.BridgeError):
        e3.theta_block_step_(1, 3, 0.5, 1, 0)
    e3.close()
