"""ctypes binding of the CPU oracle (oracle/bridge_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Two builds are exposed:
    load("ref")  reference arithmetic (no FMA contraction; what Julia computes)
    load("fma")  identical algorithm in the CUDA kernels' rounding order
    load("fast") -O3 -march=native speed build of the reference arithmetic (CPU baseline timing only)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
BB_NPAR = 32

# model ids (include/bridge_b200.h)
WIENER, OU, LINPRO, FHN_DIAG, FHN_HYPO, INTDIFF, NCLAR3, LORENZ, LANDMARKS, BOLUS = range(10)
GUIDE_NUH, GUIDE_HV, GUIDE_LMMU = 1, 2, 3
ODE_R3, ODE_LYAP = 0, 1


class Model(C.Structure):
    _fields_ = [("id", C.c_int32), ("d", C.c_int32), ("dprime", C.c_int32),
                ("reserved", C.c_int32), ("par", C.c_double * BB_NPAR)]


class Aux(C.Structure):
    _fields_ = [("d", C.c_int32), ("is_const", C.c_int32), ("B", C.c_void_p),
                ("beta", C.c_void_p), ("a", C.c_void_p), ("a_left", C.c_void_p)]


class Guide(C.Structure):
    _fields_ = [("kind", C.c_int32), ("N", C.c_int32), ("d", C.c_int32), ("m", C.c_int32),
                ("tt", C.c_void_p), ("A", C.c_void_p), ("b", C.c_void_p), ("Mm", C.c_void_p),
                ("v", C.c_void_p), ("Bt", C.c_void_p), ("betat", C.c_void_p),
                ("aux_const", C.c_int32), ("Adiff", C.c_void_p), ("adiff_const", C.c_int32)]


def build(force: bool = False) -> None:
    """Compile both oracle builds with the committed Makefile."""
    args = ["make", "-C", _HERE]
    if force:
        args.append("-B")
    subprocess.run(args, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)


def rebuild_fast_native() -> None:
    """liboracle_fast.so is compiled with -march=native: rebuild it on the host that is about to TIME it (the built
    .so travels with the repository snapshot and may come from another CPU model).  Once per process."""
    global _fast_rebuilt
    if _fast_rebuilt:
        return
    subprocess.run(["make", "-C", _HERE, "-B", os.path.join(_HERE, "liboracle_fast.so")], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    _fast_rebuilt = True
    _cache.pop("fast", None)


_fast_rebuilt = False


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_model(mid: int, d: int, dprime: int, par=()) -> Model:
    m = Model()
    m.id, m.d, m.dprime, m.reserved = mid, d, dprime, 0
    par = np.asarray(par, dtype=np.float64).ravel()
    assert par.size <= BB_NPAR
    for i, v in enumerate(par):
        m.par[i] = float(v)
    return m


def linpro_model(B, mu, sigma) -> Model:
    B = np.atleast_2d(_f64(B))
    d = B.shape[0]
    par = np.concatenate([B.ravel(), _f64(mu).ravel(), np.atleast_2d(_f64(sigma)).ravel()])
    return make_model(LINPRO, d, d, par)


class AuxHolder:
    """Keeps the numpy arrays behind a bb_aux alive."""

    def __init__(self, d, B, beta, a, a_left=None, is_const=True):
        self.d = d
        self.B, self.beta, self.a = _f64(B), _f64(beta), _f64(a)
        self.a_left = None if a_left is None else _f64(a_left)
        self.c = Aux(d, 1 if is_const else 0, _p(self.B), _p(self.beta), _p(self.a),
                     _p(self.a_left))


def const_aux(B, beta, a) -> AuxHolder:
    B = np.atleast_2d(_f64(B))
    return AuxHolder(B.shape[0], B, np.atleast_1d(_f64(beta)), np.atleast_2d(_f64(a)))


def staged_aux(tt, Bf, betaf, af) -> AuxHolder:
    """Evaluate callables B(t), beta(t), a(t) at the Ralston stage times of every interval."""
    tt = _f64(tt)
    N = tt.size
    d = np.atleast_2d(Bf(tt[0])).shape[0]
    Bs = np.empty((N - 1, 3, d, d)); bs = np.empty((N - 1, 3, d)); As = np.empty((N - 1, 3, d, d))
    al = np.empty((N - 1, d, d))
    for i in range(N - 1):
        t, h = tt[i + 1], tt[i] - tt[i + 1]
        for k, c in enumerate((0.0, 0.5, 0.75)):
            s = t + c * h
            Bs[i, k] = Bf(s); bs[i, k] = betaf(s); As[i, k] = af(s)
        al[i] = af(tt[i])
    return AuxHolder(d, Bs, bs, As, al, is_const=False)


class GuideHolder:
    def __init__(self, kind, tt, A, b, Mm=None, v=None, Bt=None, betat=None, aux_const=True, m=0, Adiff=None,
                 adiff_const=True):
        self.tt = _f64(tt)
        self.A, self.b = _f64(A), _f64(b)
        self.Mm = None if Mm is None else _f64(Mm)
        self.v = None if v is None else _f64(v)
        self.Bt, self.betat = _f64(Bt), _f64(betat)
        N = self.tt.size
        d = self.Bt.shape[-1]
        self.kind, self.N, self.d, self.m = kind, N, d, m
        self.c = Guide(kind, N, d, m, _p(self.tt), _p(self.A), _p(self.b), _p(self.Mm),
                       _p(self.v), _p(self.Bt), _p(self.betat), 1 if aux_const else 0, None, 1)
        self.Adiff = None if Adiff is None else _f64(Adiff)
        if self.Adiff is not None:
            self.c.Adiff = _p(self.Adiff)
            self.c.adiff_const = 1 if adiff_const else 0


class Oracle:
    def __init__(self, variant: str = "ref"):
        path = os.path.join(_HERE, f"liboracle_{variant}.so")
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        self.variant = variant
        L.bbo_logpdfnormal.restype = C.c_double
        L.bbo_normal.restype = C.c_double
        L.bbo_accept_logu.restype = C.c_double
        L.bbo_logu_q.restype = C.c_double
        L.bbo_exp.restype = C.c_double
        L.bbo_exp.argtypes = [C.c_double]
        L.bbo_llikelihood.restype = C.c_double
        L.bbo_lptilde_nuH.restype = C.c_double
        L.bbo_lptilde_HV.restype = C.c_double
        L.bbo_pcn_propose.restype = C.c_double
        L.bbo_pcn_bench.restype = C.c_longlong
        L.bbo_pcn_bench_fhn_tuned.restype = C.c_longlong
        assert L.bbo_gpu_order() == (1 if variant == "fma" else 0)

    # ---- rng
    def philox(self, ctr, key):
        ctr = np.asarray(ctr, dtype=np.uint32); key = np.asarray(key, dtype=np.uint32)
        out = np.zeros(4, dtype=np.uint32)
        self.lib.bbo_philox4x32_10(_p(ctr), _p(key), _p(out))
        return out

    def normal(self, seed, stream, row, n):
        return self.lib.bbo_normal(C.c_uint64(seed), C.c_uint32(stream), C.c_uint64(row),
                                   C.c_uint64(n))

    def accept_logu(self, seed, stream, chain):
        return self.lib.bbo_accept_logu(C.c_uint64(seed), C.c_uint32(stream), C.c_uint64(chain))

    def logu_q(self, seed, stream, chain, q):
        return self.lib.bbo_logu_q(C.c_uint64(seed), C.c_uint32(stream), C.c_uint64(chain), C.c_uint32(q))

    def logpdfnormal(self, x, Sigma):
        x = np.atleast_1d(_f64(x)); S = np.atleast_2d(_f64(Sigma))
        return self.lib.bbo_logpdfnormal(x.size, _p(x), _p(S))

    # ---- paths
    def wiener_sample(self, tt, dprime, seed, stream, row, y1=None):
        tt = _f64(tt)
        W = np.zeros((tt.size, dprime))
        if y1 is not None:
            W[0] = y1
        self.lib.bbo_wiener_sample(tt.size, dprime, _p(tt), C.c_uint64(seed), C.c_uint32(stream),
                                   C.c_uint64(row), _p(W))
        return W

    def euler(self, model, tt, u, W):
        tt = _f64(tt); W = _f64(W).reshape(tt.size, model.dprime); u = np.atleast_1d(_f64(u))
        X = np.zeros((tt.size, model.d))
        self.lib.bbo_euler(C.byref(model), tt.size, _p(tt), _p(u), _p(W), _p(X))
        return X

    def heun(self, model, tt, u, W, X0=None):
        """solve!(StochasticHeun(), Y, u, W, P): the last point of Y is not written by the reference (X0 supplies it)."""
        tt = _f64(tt); W = _f64(W).reshape(tt.size, model.dprime); u = np.atleast_1d(_f64(u))
        X = np.zeros((tt.size, model.d)) if X0 is None else _f64(X0).reshape(tt.size, model.d).copy()
        self.lib.bbo_heun(C.byref(model), tt.size, _p(tt), _p(u), _p(W), _p(X))
        return X

    def guided_euler(self, model, guide: GuideHolder, u, W, store=True):
        W = _f64(W).reshape(guide.N, model.dprime); u = np.atleast_1d(_f64(u))
        X = np.zeros((guide.N, model.d)) if store else None
        xend = np.zeros(model.d)
        self.lib.bbo_guided_euler(C.byref(model), C.byref(guide.c), _p(u), _p(W), _p(X), _p(xend))
        return X, xend

    def guided_mdb(self, model, guide: GuideHolder, u, W):
        W = _f64(W).reshape(guide.N, model.dprime); u = np.atleast_1d(_f64(u))
        X = np.zeros((guide.N, model.d)); xend = np.zeros(model.d)
        self.lib.bbo_guided_mdb(C.byref(model), C.byref(guide.c), _p(u), _p(W), _p(X), _p(xend))
        return X, xend

    def llikelihood(self, model, guide: GuideHolder, X, skip=0):
        X = _f64(X).reshape(guide.N, model.d)
        return self.lib.bbo_llikelihood(C.byref(model), C.byref(guide.c), _p(X), C.c_int(skip))

    def innovations(self, model, guide, tt, X):
        tt = _f64(tt); X = _f64(X).reshape(tt.size, model.d)
        W = np.zeros((tt.size, model.d))
        rc = self.lib.bbo_innovations(C.byref(model), None if guide is None else C.byref(guide.c),
                                      tt.size, _p(tt), _p(X), _p(W))
        assert rc == 0, rc
        return W

    # ---- constructors
    def update_nuHC(self, L, Sigma, v, eps):
        L = np.atleast_2d(_f64(L)); m, d = L.shape
        Sigma = np.atleast_2d(_f64(Sigma)); v = np.atleast_1d(_f64(v))
        nu = np.zeros(d); Hp = np.zeros((d, d)); Cc = C.c_double(0)
        rc = self.lib.bbo_update_nuHC(d, m, _p(L), _p(Sigma), _p(v), C.c_double(eps), _p(nu),
                                      _p(Hp), C.byref(Cc))
        assert rc == 0, rc
        return nu, Hp, Cc.value

    def update_FHC(self, L, Sigma, v, F, H, eps=0.0, C0=0.0):
        L = np.atleast_2d(_f64(L)); m, d = L.shape
        Sigma = np.atleast_2d(_f64(Sigma)); v = np.atleast_1d(_f64(v))
        F = _f64(F).copy(); H = _f64(H).copy(); Cc = C.c_double(C0)
        rc = self.lib.bbo_update_FHC(d, m, _p(L), _p(Sigma), _p(v), _p(F), _p(H), C.c_double(eps),
                                     C.byref(Cc))
        assert rc == 0, rc
        return F, H, Cc.value

    def gpupdate_nuH(self, nu, Hplus, L, Sigma, v):
        L = np.atleast_2d(_f64(L)); m, d = L.shape
        nu = _f64(nu).copy(); Hp = _f64(Hplus).copy()
        Sigma = np.atleast_2d(_f64(Sigma)); v = np.atleast_1d(_f64(v))
        rc = self.lib.bbo_gpupdate_nuH(d, m, _p(nu), _p(Hp), _p(L), _p(Sigma), _p(v))
        assert rc == 0, rc
        return nu, Hp

    def gpupdate_HV(self, Hdia, V, L, Sigma, v):
        V2, H2 = self.gpupdate_nuH(V, Hdia, L, Sigma, v)
        return H2, V2

    def backward_nuH(self, method, tt, aux: AuxHolder, nu_end, Hplus_end, C0=0.0):
        tt = _f64(tt); N, d = tt.size, aux.d
        nu_end = np.atleast_1d(_f64(nu_end)); Hp = np.atleast_2d(_f64(Hplus_end))
        nu = np.zeros((N, d)); H = np.zeros((N, d, d))
        nul = np.zeros(d); Hl = np.zeros((d, d)); Cc = C.c_double(0)
        rc = self.lib.bbo_backward_nuH(method, N, d, _p(tt), C.byref(aux.c), _p(nu_end), _p(Hp),
                                       C.c_double(C0), _p(nu), _p(H), _p(nul), _p(Hl), C.byref(Cc))
        assert rc == 0, rc
        return nu, H, nul, Hl, Cc.value

    def backward_FH(self, tt, aux, F_end, H_end, C0=0.0):
        tt = _f64(tt); N, d = tt.size, aux.d
        F_end = np.atleast_1d(_f64(F_end)); H_end = np.atleast_2d(_f64(H_end))
        F = np.zeros((N, d)); H = np.zeros((N, d, d)); Cc = C.c_double(0)
        rc = self.lib.bbo_backward_FH(N, d, _p(tt), C.byref(aux.c), _p(F_end), _p(H_end),
                                      C.c_double(C0), _p(F), _p(H), C.byref(Cc))
        assert rc == 0, rc
        return F, H, Cc.value

    def backward_HV(self, tt, aux, v, hdia_end=None):
        tt = _f64(tt); N, d = tt.size, aux.d
        v = np.atleast_1d(_f64(v))
        he = None if hdia_end is None else np.atleast_2d(_f64(hdia_end))
        Hd = np.zeros((N, d, d)); V = np.zeros((N, d))
        rc = self.lib.bbo_backward_HV(N, d, _p(tt), C.byref(aux.c), _p(v), _p(he), _p(Hd), _p(V))
        assert rc == 0, rc
        return Hd, V

    def backward_LMmu(self, tt, aux, L, Sigma):
        tt = _f64(tt); N, d = tt.size, aux.d
        L = np.atleast_2d(_f64(L)); m = L.shape[0]; Sigma = np.atleast_2d(_f64(Sigma))
        Lt = np.zeros((N, m, d)); Mt = np.zeros((N, m, m)); mut = np.zeros((N, m))
        rc = self.lib.bbo_backward_LMmu(N, d, m, _p(tt), C.byref(aux.c), _p(L), _p(Sigma), _p(Lt),
                                        _p(Mt), _p(mut))
        assert rc == 0, rc
        return Lt, Mt, mut

    def lptilde_nuH(self, nu0, H0, Cc, x):
        nu0 = np.atleast_1d(_f64(nu0)); H0 = np.atleast_2d(_f64(H0)); x = np.atleast_1d(_f64(x))
        return self.lib.bbo_lptilde_nuH(nu0.size, _p(nu0), _p(H0), C.c_double(Cc), _p(x))

    def lptilde_HV(self, tt, trB, V0, Hdia0, u):
        """trB: a scalar (constant tr B) or [(N-1), 3] values at the forward Ralston stage times."""
        tt = _f64(tt); V0 = np.atleast_1d(_f64(V0)); H0 = np.atleast_2d(_f64(Hdia0)); u = np.atleast_1d(_f64(u))
        const = 1 if np.ndim(trB) == 0 else 0
        trB = np.atleast_1d(_f64(trB))
        return self.lib.bbo_lptilde_HV(tt.size, V0.size, _p(tt), _p(trB), const, _p(V0), _p(H0), _p(u))

    # ---- pCN
    def pcn_propose(self, model, guides, u, Wc, rho, seed, it, chain, skip=0):
        S = len(guides); N = guides[0].N
        garr = (C.POINTER(Guide) * S)(*[C.pointer(g.c) for g in guides])
        Wc = _f64(Wc).reshape(S, N, model.dprime); u = np.atleast_1d(_f64(u))
        Wo = np.zeros_like(Wc); Xo = np.zeros((S, N, model.d)); xend = np.zeros(model.d)
        logu = C.c_double(0)
        ll = self.lib.bbo_pcn_propose(C.byref(model), garr, S, _p(u), _p(Wc), C.c_double(rho),
                                      C.c_uint64(seed), C.c_uint32(it), C.c_uint64(chain),
                                      C.c_int(skip), _p(Wo), _p(Xo), _p(xend), C.byref(logu))
        return ll, logu.value, Wo, Xo, xend

    def pcn_combine(self, tt, Wc, rho, seed, it, row):
        """W° of one segment: sample!(W2, Wiener()) from noise row `row`, W°.yy = ρ W.yy + sqrt(1-ρ²) W2.yy"""
        tt = _f64(tt); Wc = _f64(Wc).reshape(len(tt), -1); Wo = np.zeros_like(Wc)
        self.lib.bbo_pcn_combine(C.c_int(len(tt)), C.c_int(Wc.shape[1]), _p(tt), _p(Wc), C.c_double(rho),
                                 C.c_uint64(seed), C.c_uint32(it), C.c_uint64(row), _p(Wo))
        return Wo

    def pcn_bench(self, model, guides, P, u, rho, seed, iters, skip=0, nthreads=0):
        S = len(guides)
        garr = (C.POINTER(Guide) * S)(*[C.pointer(g.c) for g in guides])
        u = np.atleast_1d(_f64(u)); secs = C.c_double(0); ll = np.zeros(P)
        acc = self.lib.bbo_pcn_bench(C.byref(model), garr, S, C.c_longlong(P), _p(u),
                                     C.c_double(rho), C.c_uint64(seed), C.c_int(iters),
                                     C.c_int(skip), C.c_int(nthreads), C.byref(secs), _p(ll))
        return acc, secs.value, ll

    def pcn_bench_fhn_tuned(self, model, guides, P, u, rho, seed, iters, nthreads=0):
        """The same iteration as pcn_bench, specialised for the FitzHugh-Nagumo / PartialBridgeνH workload (bench.py)."""
        S = len(guides)
        garr = (C.POINTER(Guide) * S)(*[C.pointer(g.c) for g in guides])
        u = np.atleast_1d(_f64(u)); secs = C.c_double(0); ll = np.zeros(P)
        acc = self.lib.bbo_pcn_bench_fhn_tuned(C.byref(model), garr, S, C.c_longlong(P), _p(u), C.c_double(rho),
                                               C.c_uint64(seed), C.c_int(iters), C.c_int(nthreads), C.byref(secs),
                                               _p(ll))
        assert acc >= 0, "bbo_pcn_bench_fhn_tuned: unsupported workload"
        return acc, secs.value, ll

    def max_threads(self):
        return self.lib.bbo_max_threads()


_cache = {}


def load(variant: str = "ref") -> Oracle:
    if variant not in _cache:
        _cache[variant] = Oracle(variant)
    return _cache[variant]


# ======================================================================= per-chain parameters
# The `updateparams` branch of project_partialbridge/partialbridge_bolus3.jl:248-365 for ONE chain, composed from
# the oracle's restatements of the reference functions it calls (partialbridgeνH / Lyap, gpupdate, solve!,
# llikelihood, logpdfnormal).  Mirrors bb_theta.cu; the summation orders of the script-level quantities
# (trace term, diffll) are the ones written here.
AUX_FHN_MATCHING, AUX_FHN_LINEARISED_END, AUX_BOLUS = 1, 2, 3
Q_THETA_NORMALS, Q_THETA_LOGU = 0xFFFFFFFE, 0xFFFFFFFD


def fhn_aux(kind, par, v):
    """B~, beta~, a~ from (θ, v): partialbridge_fitzhugh.jl:98-108 (x^2, x^3 as products: Julia literal_pow)."""
    eps, s, gam, beta = (float(x) for x in par[:4])
    ie = 1.0 / eps
    if kind == AUX_FHN_MATCHING:
        Bt = np.array([[ie, -ie], [gam, -1.0]])
        bt = np.array([s / eps - (v * v * v) / eps, beta])
    else:
        Bt = np.array([[ie - (3.0 * (v * v)) / eps, -ie], [gam, -1.0]])
        bt = np.array([s / eps + (2.0 * (v * v * v)) / eps, beta])
    return Bt, bt


def fhn_a(model_id, par):
    if model_id == FHN_HYPO:
        return np.array([[0.0, 0.0], [0.0, float(par[4]) * float(par[4])]])
    return np.array([[float(par[4]) * float(par[4]), 0.0], [0.0, float(par[5]) * float(par[5])]])


def dose(t):
    """dose(t) = 2*(t/2)/(1+(t/2)^2)   partialbridge_bolus3.jl:73"""
    u = t / 2
    return (2 * u) / (1 + u * u)


def bolus_aux(par):
    """DiffusionAux of partialbridge_bolus3.jl:54-71: B~ (constant), beta~(t), a~ = diag(σ1², σ2²)."""
    alpha, beta, lam, mu, s1, s2 = (float(x) for x in par[:6])
    Bt = np.array([[-lam - beta, mu], [lam, -mu]])
    at = np.array([[s1 * s1, 0.0], [0.0, s2 * s2]])
    return Bt, (lambda t: np.array([alpha * dose(t), 0.0])), at


def theta_backward(o: Oracle, model_id, par, grids, x0, L, Sigma, eps, obs_v, aux_kind, priors=None):
    """-> (guides [S], left dict): bolus3.jl:162-165, 276-291 for one θ."""
    S = len(grids)
    L = np.atleast_2d(_f64(L)); d = L.shape[1]
    nu = np.zeros(d); Hp = np.eye(d) * (1.0 / eps)
    nu, Hp = o.gpupdate_nuH(nu, Hp, L, Sigma, np.atleast_1d(obs_v[S - 1]))
    guides = [None] * S
    Cc = 0.0; trsum = 0.0
    for s in range(S - 1, -1, -1):
        if aux_kind == AUX_BOLUS:
            Bt, bf, at = bolus_aux(par)
            aux = staged_aux(grids[s], lambda t: Bt, bf, lambda t: at)
            nus, Hs, nu, Hp, Cc = o.backward_nuH(ODE_LYAP, grids[s], aux, nu, Hp, Cc)
            n = len(grids[s])
            guides[s] = GuideHolder(GUIDE_NUH, grids[s], Hs, nus, Bt=np.broadcast_to(Bt, (n, d, d)).copy(),
                                    betat=np.stack([bf(t) for t in grids[s]]), aux_const=False)
        else:
            at = fhn_a(model_id, par)
            Bt, bt = fhn_aux(aux_kind, par, float(np.atleast_1d(obs_v[s])[0]))
            aux = const_aux(Bt, bt, at)
            nus, Hs, nu, Hp, Cc = o.backward_nuH(ODE_LYAP, grids[s], aux, nu, Hp, Cc)
            guides[s] = GuideHolder(GUIDE_NUH, grids[s], Hs, nus, Bt=Bt, betat=bt)
        tr = Bt[0, 0]
        for i in range(1, d):
            tr += Bt[i, i]
        trsum += (grids[s][-1] - grids[s][0]) * tr
        if s > 0:
            nu, Hp = o.gpupdate_nuH(nu, Hp, L, Sigma, np.atleast_1d(obs_v[s - 1]))
    lpn = o.logpdfnormal(np.asarray(x0, dtype=np.float64) - nu, Hp)
    lpri = 0.0
    for k in sorted((priors or {})):
        _, a, b = priors[k]
        x = float(par[k])
        import math
        t = -math.lgamma(a) - a * math.log(b)
        if a - 1.0 != 0.0:
            t += (a - 1.0) * math.log(x) if x > 0 else 0.0
        t -= x / b
        lpri += t if x > 0.0 else -math.inf
    return guides, dict(nu=nu, Hp=Hp, C=Cc, lpn=lpn, trsum=trsum, lpri=lpri)


def theta_forward(o: Oracle, model_id, dprime, par, guides, x0, W, skip=0):
    """solve!(Euler(), X, x0, W, Q) over all segments + Σ llikelihood: -> (X [S,N,d], ll, xend)"""
    S = len(guides)
    mdl = make_model(model_id, 2, dprime, par)
    X = []; ll = 0.0; u = np.asarray(x0, dtype=np.float64)
    for s in range(S):
        Xs, u = o.guided_euler(mdl, guides[s], u, W[s])
        ll += o.llikelihood(mdl, guides[s], Xs, skip)
        X.append(Xs)
    return np.stack(X), ll, u


def theta_propose(o: Oracle, par, rw_sd, seed, it, chain):
    """propose(σ, P): θ° = θ + σ .* randn (bolus3.jl:239-242); the n-th updated parameter takes normal n of quad
    0xFFFFFFFE of row = chain."""
    par = np.array(par, dtype=np.float64)
    n = 0
    for k in range(len(rw_sd)):
        if rw_sd[k] != 0.0:
            z = o.normal(seed, it, chain, 4 * Q_THETA_NORMALS + n)
            par[k] = par[k] + rw_sd[k] * z
            n += 1
    return par


Q_THETA_COIN = 0xFFFFFFFC


def theta_propose_start(o: Oracle, x0, sd, direction, seed, it, chain):
    """Joint start-point proposal of a parameter step (bolus3.jl:311-318): with probability 1/2 (bit 0 of word 0 of the
    Philox counter (0xFFFFFFFC, it, chain)) x0° = x0 + (sd u) dir, u = normal 3 of the proposal quad."""
    x0 = np.array(x0, dtype=np.float64)
    coin = int(o.philox([Q_THETA_COIN, it, chain & 0xFFFFFFFF, chain >> 32], [seed & 0xFFFFFFFF, seed >> 32])[0]) & 1
    if coin:
        u = o.normal(seed, it, chain, 4 * Q_THETA_NORMALS + 3)
        x0 = x0 + (sd * u) * np.asarray(direction, dtype=np.float64)
    return x0


def theta_diffll(left_c, left_o, ll_c, ll_o):
    """diffll of bolus3.jl:319,331-336 in the kernel's summation order."""
    diff = left_o["lpn"] - left_c["lpn"]
    diff += ll_o - ll_c
    diff += ((left_o["trsum"] - left_c["trsum"]) + left_o["lpri"]) - left_c["lpri"]
    return diff


# ----------------------------------------------------------------------- blocked segment updates
# The path-update branch (`updateparams == false`) of partialbridge_bolus3.jl:258-355 for ONE chain and ONE block of
# segments s_lo .. s_hi-1 (0-based; the script's ind = (kup-1):-1:klow with klow = s_lo+1, kup = s_hi+1).
def theta_block_backward(o: Oracle, model_id, par, grids, L, Sigma, eps, obs_v, aux_kind, s_lo, s_hi, x_right, hzero):
    """Guides of the block's segments (others None) and (ν, H⁺) at the block's left end.  Right end (:272-275):
    the right-most initialisation (ν = 0, H⁺ = I/ϵ, update with the last observation, :162-165) when the block ends
    the chain, else ν = the chain's own path at the block's right end and H⁺ = Hzero⁺ = hzero I.  An observation update
    follows every segment except the block's left-most one (:288-295)."""
    S = len(grids)
    L = np.atleast_2d(_f64(L)); d = L.shape[1]
    if s_hi == S:
        nu = np.zeros(d); Hp = np.eye(d) * (1.0 / eps)
        nu, Hp = o.gpupdate_nuH(nu, Hp, L, Sigma, np.atleast_1d(obs_v[S - 1]))
    else:
        nu = np.array(x_right, dtype=np.float64); Hp = np.eye(d) * float(hzero)
    guides = [None] * S
    Cc = 0.0
    for s in range(s_hi - 1, s_lo - 1, -1):
        if aux_kind == AUX_BOLUS:
            Bt, bf, at = bolus_aux(par)
            aux = staged_aux(grids[s], lambda t: Bt, bf, lambda t: at)
            nus, Hs, nu, Hp, Cc = o.backward_nuH(ODE_LYAP, grids[s], aux, nu, Hp, Cc)
            n = len(grids[s])
            guides[s] = GuideHolder(GUIDE_NUH, grids[s], Hs, nus, Bt=np.broadcast_to(Bt, (n, d, d)).copy(),
                                    betat=np.stack([bf(t) for t in grids[s]]), aux_const=False)
        else:
            at = fhn_a(model_id, par)
            Bt, bt = fhn_aux(aux_kind, par, float(np.atleast_1d(obs_v[s])[0]))
            nus, Hs, nu, Hp, Cc = o.backward_nuH(ODE_LYAP, grids[s], const_aux(Bt, bt, at), nu, Hp, Cc)
            guides[s] = GuideHolder(GUIDE_NUH, grids[s], Hs, nus, Bt=Bt, betat=bt)
        if s > s_lo:
            nu, Hp = o.gpupdate_nuH(nu, Hp, L, Sigma, np.atleast_1d(obs_v[s - 1]))
    return guides, nu, Hp


def theta_block_step(o: Oracle, model_id, dprime, par, grids, L, Sigma, eps, obs_v, aux_kind, s_lo, s_hi, hzero,
                     x0, Xcur, Wcur, rho, seed, it, chain, start_sd=0.0, start_dir=None, skip=0):
    """One blocked path update of one chain -> dict(Wo, Xo [block segments], x0o, lpn, lpno, llt [S], llo [S], diff,
    logu).  Xcur [S,N,d] / Wcur [S,N,d'] are the chain's current path and driving noise.

    bolus3.jl:300-345: W°[i] = ρW[i] + √(1-ρ²)W2 on the block; the CURRENT noise is re-simulated under the block's
    guide (XXtemp) next to the proposal (XXᵒ); the block starts at the current path's value at its left end -- for
    s_lo = 0 at x0, moved to x0° = x0 + (start_sd u) start_dir with the Gaussian start term (:311-320); diffll sums
    the start term and then, in the script's order (i in ind, descending), ll°[i] - ll_temp[i]; log U as for pCN."""
    S = len(grids)
    x_right = Xcur[s_hi - 1][-1] if s_hi < S else None
    guides, nuL, HpL = theta_block_backward(o, model_id, par, grids, L, Sigma, eps, obs_v, aux_kind, s_lo, s_hi,
                                            x_right, hzero)
    mdl = make_model(model_id, 2, dprime, par)
    lpn = lpno = 0.0
    if s_lo == 0:
        xs = np.asarray(x0, dtype=np.float64); xso = xs
        lpn = o.logpdfnormal(xs - nuL, HpL)
        if start_sd != 0.0:
            u = o.normal(seed, it, chain, 4 * Q_THETA_NORMALS + 3)
            xso = xs + (start_sd * u) * np.asarray(start_dir, dtype=np.float64)
            lpno = o.logpdfnormal(xso - nuL, HpL)
        else:
            lpno = lpn
    else:
        xs = xso = np.array(Xcur[s_lo - 1][-1], dtype=np.float64)
    x0o = xso
    llt = np.zeros(S); llo = np.zeros(S); Wo = {}; Xo = {}
    ut, uo = xs, xso
    for s in range(s_lo, s_hi):
        Wo[s] = o.pcn_combine(grids[s], Wcur[s], rho, seed, it, chain * S + s)
        Xt, ut = o.guided_euler(mdl, guides[s], ut, Wcur[s])
        Xo[s], uo = o.guided_euler(mdl, guides[s], uo, Wo[s])
        llt[s] = o.llikelihood(mdl, guides[s], Xt, skip)
        llo[s] = o.llikelihood(mdl, guides[s], Xo[s], skip)
    diff = lpno - lpn
    for s in range(s_hi - 1, s_lo - 1, -1):
        diff += llo[s] - llt[s]
    return dict(Wo=Wo, Xo=Xo, x0o=x0o, lpn=lpn, lpno=lpno, llt=llt, llo=llo, diff=diff,
                logu=o.accept_logu(seed, it, chain), nuL=nuL, HpL=HpL, guides=guides)


# ----------------------------------------------------------------------------------------------- online statistics
# numpy restatement of src/mclog.jl (element-wise IEEE operations in the reference's order; no fused operations).
# State mc = (m, m2, n): m has the shape of the path array [..., d], m2 = sum of outer products of deviations [..., d, d].
MCBAND_Q = 1.9599639845400538  # sqrt(2.) * erfinv(0.95)  (src/mclog.jl:77,87), evaluated in double precision


def mcstart(yy):
    """mcstart(yy)  src/mclog.jl:22: zeros of the shape of yy (entries ignored), zeros of outer products, n = 0."""
    yy = np.asarray(yy, dtype=np.float64)
    return np.zeros(yy.shape), np.zeros(yy.shape + (yy.shape[-1],)), 0


def mcnext(mc, x):
    """mcnext!(mc, x::Vector{<:AbstractArray})  src/mclog.jl:47-56:
    delta = x[i] - m[i];  m[i] += delta/(n+1);  m2[i] += outer(delta, x[i] - m[i]);  n + 1."""
    m, m2, n = mc
    x = np.asarray(x, dtype=np.float64)
    delta = x - m
    m = m + delta / float(n + 1)
    m2 = m2 + delta[..., :, None] * (x - m)[..., None, :]
    return m, m2, n + 1


def mcstats(mc):
    """mcstats(mc) = (m, m2/(k - 1))  src/mclog.jl:88-93."""
    m, m2, k = mc
    with np.errstate(divide="ignore", invalid="ignore"):
        return m, m2 / float(k - 1)


def mcband(mc):
    """mcband(mc)  src/mclog.jl:75-85: std = sqrt.(diag(v) * (1/(k - 1))); (m - Q std, m + Q std)."""
    m, m2, k = mc
    std = np.sqrt(np.einsum("...ii->...i", m2) * (1.0 / float(k - 1)))
    return m - MCBAND_Q * std, m + MCBAND_Q * std
