"""User-defined target processes (CUDA C source compiled at run time by NVRTC into the library's own path kernels):
the counterpart of the reference's extension point, Bridge.b / Bridge.σ methods for an own struct (src/types.jl:23,32-33,
project_partialbridge/partialbridge_fitzhugh.jl:44-46).

CPU: the source compiles for every kernel family without a device; a broken source is refused with NVRTC's log.
GPU: a user model that restates the registry's FitzHugh-Nagumo model gives the SAME BITS as the registry model in every
mode (plain / guided Euler, log-likelihood fused and second pass, pCN, Heun), and the oracle's; a model that is not in
the registry (a double-well drift with a skewed diffusion column) matches a NumPy restatement of the Euler loop."""
import numpy as np
import pytest

FHN_DRIFT = ("double x1 = x[0], x2 = x[1]; double c = x1 * x1; double u = x1 - x2; u = fma(-c, x1, u);"
             "o[0] = (u + par[1]) * (1.0 / par[0]); o[1] = fma(par[2], x1, -x2) + par[3];")
FHN_PAR = [0.1, 0.0, 1.5, 0.8, 0.3]


def warped(t0, t1, n):
    s = np.linspace(0, t1 - t0, n)
    return t0 + s * (2 - s / (t1 - t0))


def test_user_source_compiles_without_a_device():
    import bridge_jl_b200 as B
    for gk, gm, auxm, rng in ((0, 0, 1, 0), (0, 0, 1, 2), (1, 0, 1, 1), (1, 0, 1, 3), (1, 0, 0, 0), (2, 0, 1, 0),
                              (3, 1, 2, 1), (1, 0, 1, 10), (0, 0, 1, 12)):
        size, log = B.check_user_source(2, 1, FHN_DRIFT, [-1, 0], [None, "par[4]"], gk, gm, auxm, rng)
        assert size > 10000, (gk, gm, auxm, rng, log)
    # d = d' = 3 with a diagonal sigma
    size, _ = B.check_user_source(3, 3, "o[0] = par[0]*(x[1]-x[0]); o[1] = fma(x[0], par[1]-x[2], -x[1]); "
                                  "o[2] = fma(x[0], x[1], -(par[2]*x[2]));", [0, 1, 2], ["par[3]", "par[4]", "par[5]"], 0, 0, 1, 0)
    assert size > 10000
    with pytest.raises(B.BridgeError) as ei:
        B.check_user_source(2, 1, "o[0] = undefined_symbol(x[0]); o[1] = 0;", [-1, 0], [None, "par[4]"])
    assert ei.value.status == -15 and "undefined_symbol" in str(ei.value)
    with pytest.raises(B.BridgeError) as ei:
        B.check_user_source(2, 1, FHN_DRIFT, [-1, 3], [None, "par[4]"])  # column outside the driving process
    assert ei.value.status == -7


@pytest.mark.gpu
def test_user_model_reproduces_the_registry_model_bit_for_bit():
    import bridge_jl_b200 as B
    import bridge_jl_b200.configs as cfg
    K = B.api.K
    U = B.UserProcess(2, 1, FHN_DRIFT, [-1, 0], [None, "par[4]"], FHN_PAR)
    R = B.FitzhughDiffusion(*FHN_PAR)
    n, P, S = 97, 200, 3
    obs_t, obs_v = (0.5, 1.0, 1.5), (-1.0, -0.5, 0.5)
    _, guides, x0, rho = cfg.fhn_config4(n, obs_t=obs_t, obs_v=obs_v)
    out = {}
    for name, Pm in (("registry", R), ("user", U)):
        ens = B.PathEnsemble(P, S, n, 2, 1, chain_offset=11)
        for s, g in enumerate(guides):
            ens.set_grid(s, g.tt)
        ens.set_start(x0); ens.sample_(6, 0xFFFFFFFE)
        rec = [ens.download(B.W)]
        ens.euler_(Pm); rec.append(ens.download(B.X))
        ens.guided_euler_ll_(Pm, guides, skip=1); rec += [ens.download(B.X), ens.ll, ens.xend]
        ens.set_ll(np.zeros(P)); ens.llikelihood_(Pm, guides, skip=1); rec.append(ens.ll)   # second pass == fused
        assert np.array_equal(rec[-1], rec[-3])
        for it in range(3):
            ens.pcn_step_(Pm, guides, rho, 6, it, skip=1, store_x=(it != 1))
            rec += [ens.ll_prop, ens.logu, ens.accepted, ens.download(B.W, which=B.PROP)]
        rec += [ens.download(B.X), ens.ll, np.array([ens.acc])]
        ens.close()
        one = B.PathEnsemble(64, 1, n, 2, 1, double_buffer=False)
        one.set_grid(0, guides[0].tt); one.set_start(x0); one.sample_(2, 5)
        one.solve_scheme_(Pm, K.SCHEME_HEUN); rec.append(one.download(B.X)[:, :, :-1])
        one.sample_euler_(Pm, 3, 1); rec += [one.download(B.W), one.download(B.X)]
        one.close()
        out[name] = rec
    assert len(out["user"]) == len(out["registry"])
    for a, b in zip(out["user"], out["registry"]):
        assert np.array_equal(a, b)
    # parameters change without recompiling: another sigma, same kernels
    U.p[4] = 0.45
    R2 = B.FitzhughDiffusion(0.1, 0.0, 1.5, 0.8, 0.45)
    e1 = B.PathEnsemble(32, 1, n, 2, 1, double_buffer=False); e2 = B.PathEnsemble(32, 1, n, 2, 1, double_buffer=False)
    for e, Pm in ((e1, U), (e2, R2)):
        e.set_grid(0, guides[0].tt); e.set_start(x0); e.sample_(1, 0); e.euler_(Pm)
    assert np.array_equal(e1.download(B.X), e2.download(B.X))
    e1.close(); e2.close()
    with pytest.raises(B.BridgeError):
        B.UserProcess(2, 1, "o[0] = nonsense;", [-1, 0], [None, "par[4]"], FHN_PAR)
    U.close()


@pytest.mark.gpu
def test_user_model_outside_the_registry_matches_numpy():
    """dX1 = (X1 - X1^3 - θ X2) dt + σ1 dW,  dX2 = (X1 - X2) dt + σ2 dW: one Wiener process drives BOTH components
    (a column-vector σ no registry model has)."""
    import bridge_jl_b200 as B
    U = B.UserProcess(2, 1, "o[0] = (x[0] - x[0]*x[0]*x[0]) - par[0]*x[1]; o[1] = x[0] - x[1];", [0, 0], ["par[1]", "par[2]"],
                      [0.7, 0.4, 0.25])
    n, P = 201, 50
    tt = warped(0.0, 1.0, n)
    ens = B.PathEnsemble(P, 1, n, 2, 1, double_buffer=False)
    ens.set_grid(0, tt); ens.set_start([0.3, -0.2]); ens.sample_(8, 0)
    W = ens.download(B.W)[:, 0, :, 0]
    ens.euler_(U)
    X = ens.download(B.X)[:, 0]
    for p in (0, 17, 49):
        y = np.array([0.3, -0.2]); ref = np.empty((n, 2))
        for i in range(n - 1):
            ref[i] = y
            dt, dw = tt[i + 1] - tt[i], W[p, i + 1] - W[p, i]
            b = np.array([(y[0] - y[0] * y[0] * y[0]) - 0.7 * y[1], y[0] - y[1]])
            y = (y + b * dt) + np.array([0.4, 0.25]) * dw
        ref[n - 1] = y
        assert np.max(np.abs(X[p] - ref)) <= 1e-13 * (1 + np.max(np.abs(ref)))  # the kernel fuses b*dt + y and sigma*dw + .
    ens.close(); U.close()
