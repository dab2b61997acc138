"""Host-side mirror of the Bridge.jl interface for the accelerated path.

Names, argument order and error behaviour follow the reference (mschauer/Bridge.jl @ b09488fe):

    ContinuousTimeProcess{T}, b, σ, a            src/types.jl:23,32-33, src/Bridge.jl:105-106
    SamplePath, VSamplePath, samplepath           src/types.jl:71-81,123-130
    sample, sample!                               src/wiener.jl:11-58
    solve, solve!(::EulerMaruyama / ::Euler, …)   src/euler.jl:117-118,135-152,246-268
    bridge!                                       src/deprecated.jl:16-17, project/partialbridge.jl:63
    llikelihood(::LeftRule, X, P°; skip)          src/partialbridgenuH.jl:171, guip.jl:429, partialbridge.jl:67
    innovations!                                  src/euler.jl:357-376
    GuidedBridge, PartialBridge, PartialBridgeνH, partialbridgeνH, gpupdate
                                                  src/guip.jl:165-243, partialbridge.jl:33-51, partialbridgenuH.jl:122-155

Julia's `f!` is spelled `f_` here.  Every numeric result comes from libbridge_b200.so (CUDA, sm_100a)
through the C ABI in _cabi.py; this module only converts between the reference's containers and the
ABI's buffers.  `PathEnsemble` is the device-resident container a many-chain sampler uses instead of
one SamplePath pair per chain.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _cabi as K
from ._cabi import BridgeError, check, f64, lib, ptr

# ----------------------------------------------------------------------------------------------- context


class Context:
    """One CUDA device + stream (bb_ctx)."""

    def __init__(self, device: Optional[int] = None):
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = C.c_void_p()
        check(lib.bb_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device
        self._small = {}

    def synchronize(self):
        check(lib.bb_ctx_synchronize(self.h))

    def set_stream(self, cuda_stream: int):
        check(lib.bb_ctx_set_stream(self.h, C.c_void_p(cuda_stream)))

    def set_timing(self, on: bool):
        check(lib.bb_ctx_set_timing(self.h, 1 if on else 0))

    def set_arith(self, arith: int):
        """Rounding order of the shared-table constructors: K.ARITH_REFERENCE (default, bit-identical to the CPU
        restatement in reference arithmetic) or K.ARITH_FUSED (the per-chain kernels' explicit fused multiply-adds)."""
        check(lib.bb_ctx_set_arith(self.h, arith))

    def set_pcn_kernel(self, mode: int):
        """K.PCN_AUTO (default) / K.PCN_ONE_THREAD / K.PCN_WARP_SPECIALISED / K.PCN_WARP_SPECIALISED_2: which kernel runs
        pcn_step_ (same results bit for bit)."""
        check(lib.bb_ctx_set_pcn_kernel(self.h, mode))

    @property
    def last_kernel_ms(self) -> float:
        return lib.bb_ctx_last_kernel_ms(self.h)

    @property
    def launch_count(self) -> int:
        return lib.bb_ctx_launch_count(self.h)

    def close(self):
        for e in self._small.values():
            e.close()
        self._small = {}
        if self.h:
            lib.bb_ctx_destroy(self.h)
            self.h = None


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


# ----------------------------------------------------------------------------------------------- solver tags
class SDESolver:
    pass


class EulerMaruyama(SDESolver):  # src/euler.jl:17-21
    pass


Euler = EulerMaruyama  # src/euler.jl:23


class EulerMaruyama_(SDESolver):
    """EulerMaruyama!  (src/euler.jl:62, src/sde!.jl:21-53): the in-place variant for VSamplePath (d x N matrices).  The
    recurrence is the same -- y + b dt + σ dw, evaluated (y + b dt) + σ dw -- so it runs the same kernel; what differs is
    the container and the checks: size(Y.yy) == (length(u), N), else "Starting point has wrong length."."""
    inplace = True


class StratonovichEuler(SDESolver):  # src/euler.jl:26-31
    scheme = K.SCHEME_STRATONOVICH


class StochasticHeun(SDESolver):  # src/euler.jl:38-44
    scheme = K.SCHEME_HEUN


class StochasticRungeKutta(SDESolver):  # src/euler.jl:55-61
    scheme = K.SCHEME_SRK


class Mdb(SDESolver):  # src/euler.jl:46-52 (needs a classic proposal process: not on the accelerated path)
    scheme = K.SCHEME_MDB


class LeftRule:  # src/ode.jl:8
    pass


class R3:  # src/ode.jl:26
    pass


class Lyap:  # src/partialbridgenuH.jl:84
    pass


# ----------------------------------------------------------------------------------------------- processes
class ContinuousTimeProcess:
    """Base of the target processes the device registry knows (src/types.jl:23).  Subclasses give
    the registry id and parameter block, plus host evaluations of b, σ, a for inspection."""

    model_id = -1
    d = 1
    dprime = 1

    def par(self) -> Sequence[float]:
        raise NotImplementedError

    def cmodel(self) -> K.Model:
        m = K.Model()
        m.id, m.d, m.dprime, m.reserved = self.model_id, self.d, self.dprime, 0
        for i, v in enumerate(np.asarray(self.par(), dtype=np.float64).ravel()):
            m.par[i] = float(v)
        return m

    # host-side coefficient protocol (Bridge.b/σ/a); not used by the solvers
    def b(self, t, x):
        raise NotImplementedError

    def σ(self, t, x):
        raise NotImplementedError

    def a(self, t, x=None):
        s = np.atleast_2d(self.σ(t, x))
        return s @ s.T  # fallback a = σσ'  src/types.jl:32

    sigma = σ
    constdiff = True


class Wiener(ContinuousTimeProcess):  # src/wiener.jl
    model_id = K.WIENER

    def __init__(self, d: int = 1):
        self.d = self.dprime = d

    def par(self):
        return []

    def b(self, t, x):
        return np.zeros(self.d)

    def σ(self, t, x=None):
        return np.eye(self.d)


class OrnsteinUhlenbeck(ContinuousTimeProcess):  # docs/src/manual.md:44-46
    model_id = K.OU

    def __init__(self, β: float, σ: float):
        self.β, self.σ_ = float(β), float(σ)

    def par(self):
        return [self.β, self.σ_]

    def b(self, t, x):
        return -self.β * np.asarray(x)

    def σ(self, t, x=None):
        return np.array([[self.σ_]])


class LinPro(ContinuousTimeProcess):  # src/linpro.jl:65-87
    model_id = K.LINPRO
    is_const = True

    def __init__(self, B, μ, σ):
        self.Bm = np.atleast_2d(f64(B))
        self.d = self.dprime = self.Bm.shape[0]
        self.μ = np.atleast_1d(f64(μ))
        self.σm = np.atleast_2d(f64(σ))
        if self.σm.shape == (1, 1) and self.d > 1:
            self.σm = self.σm[0, 0] * np.eye(self.d)

    def par(self):
        return np.concatenate([self.Bm.ravel(), self.μ.ravel(), self.σm.ravel()])

    def b(self, t, x):
        return self.Bm @ (np.asarray(x) - self.μ)

    def σ(self, t, x=None):
        return self.σm

    # auxiliary-process protocol  src/linpro.jl:81-86
    def B(self, t):
        return self.Bm

    def β(self, t):
        return -self.Bm @ self.μ

    def a(self, t, x=None):
        return self.σm @ self.σm.T


class FitzHughNagumo(ContinuousTimeProcess):  # src/Models.jl:9-20 (diagonal noise)
    model_id = K.FHN_DIAG
    d = dprime = 2

    def __init__(self, ϵ, s, γ, β, σ1, σ2):
        self.p = [float(v) for v in (ϵ, s, γ, β, σ1, σ2)]

    def par(self):
        return self.p

    def b(self, t, x):
        ϵ, s, γ, β = self.p[:4]
        return np.array([(x[0] - x[0] ** 3 - x[1] + s) / ϵ, γ * x[0] - x[1] + β])

    def σ(self, t, x=None):
        return np.diag(self.p[4:6])


class FitzhughDiffusion(ContinuousTimeProcess):  # project_partialbridge/partialbridge_fitzhugh.jl:36-46
    model_id = K.FHN_HYPO
    d, dprime = 2, 1

    def __init__(self, ϵ, s, γ, β, σ):
        self.p = [float(v) for v in (ϵ, s, γ, β, σ)]

    def par(self):
        return self.p

    def b(self, t, x):
        ϵ, s, γ, β = self.p[:4]
        return np.array([(x[0] - x[1] - x[0] ** 3 + s) / ϵ, γ * x[0] - x[1] + β])

    def σ(self, t, x=None):
        return np.array([[0.0], [self.p[4]]])


class IntegratedDiffusion(ContinuousTimeProcess):  # test/partialbridge.jl:18-28
    model_id = K.INTDIFF
    d, dprime = 2, 1

    def __init__(self, γ):
        self.γ = float(γ)

    def par(self):
        return [self.γ]

    def b(self, t, x):
        return np.array([x[1], -(x[1] + np.sin(x[1])) + 0.5])

    def σ(self, t, x=None):
        return np.array([[0.0], [self.γ]])


class NclarDiffusion(ContinuousTimeProcess):  # project_partialbridge/partialbridge_nclar.jl:50-60
    model_id = K.NCLAR3
    d, dprime = 3, 1

    def __init__(self, α, ω, σ):
        self.p = [float(α), float(ω), float(σ)]

    def par(self):
        return self.p

    def b(self, t, x):
        return np.array([x[1], x[2], -self.p[0] * np.sin(self.p[1] * x[2])])

    def σ(self, t, x=None):
        return np.array([[0.0], [0.0], [self.p[2]]])


class Lorenz(ContinuousTimeProcess):  # src/Models.jl:38-55
    model_id = K.LORENZ
    d = dprime = 3

    def __init__(self, θ, σ):
        σ = np.atleast_1d(f64(σ))
        if σ.size == 1:
            σ = np.repeat(σ, 3)
        self.p = [float(v) for v in list(θ) + list(σ)]

    def par(self):
        return self.p

    def b(self, t, x):
        θ = self.p
        return np.array([θ[0] * (x[1] - x[0]), x[0] * (θ[1] - x[2]) - x[1], x[0] * x[1] - θ[2] * x[2]])

    def σ(self, t, x=None):
        return np.diag(self.p[3:6])


class BolusDiffusion(ContinuousTimeProcess):  # project_partialbridge/partialbridge_bolus3.jl:38-51 (`Diffusion`)
    """dX1 = (α dose(t) - (λ+β) X1 + μ X2) dt + σ1 dW1,  dX2 = (λ X1 - μ X2) dt + σ1 dW2; σ2 enters the auxiliary process
    only.  Its drift depends on t: it runs on the per-chain-parameter path (PathEnsemble.theta_*)."""
    model_id = K.BOLUS
    d = dprime = 2

    def __init__(self, α, β, λ, μ, σ1, σ2):
        self.p = [float(v) for v in (α, β, λ, μ, σ1, σ2)]

    def par(self):
        return self.p

    @staticmethod
    def dose(t):  # :73
        return 2 * (t / 2) / (1 + (t / 2) ** 2)

    def b(self, t, x):
        α, β, λ, μ = self.p[:4]
        return np.array([α * self.dose(t) - (λ + β) * x[0] + μ * x[1], λ * x[0] - μ * x[1]])

    def σ(self, t, x=None):
        return self.p[4] * np.eye(2)


class Landmarks(ContinuousTimeProcess):  # project_partialbridge/partialbridge_landmarks.jl:47,67-72,86-101
    """n = 4 landmarks in the plane with Gaussian kernel parameter a, noise level σ on the momenta and mean reversion λ.
    State: (q1, p1, ..., q4, p4) flattened (the script's fll(Vector{Point})), d = 16, d' = 8."""
    model_id = K.LANDMARKS
    d, dprime, n = 16, 8, 4

    def __init__(self, a, σ, λ, n: int = 4):
        if n != 4:
            raise ValueError("the device registry holds the n = 4 instance (state SVector{16}, :51)")
        self.a_, self.σ_, self.λ = float(a), float(σ), float(λ)

    def par(self):
        return [self.a_, self.σ_, self.λ]

    def kernel(self, x):  # :47
        return np.exp(-float(np.dot(x, x)) / (2 * self.a_)) / (2 * np.pi * self.a_)

    def b(self, t, x):  # :90-101
        x = np.asarray(x, dtype=np.float64).reshape(self.n, 2, 2)  # [landmark][q|p][coordinate]
        out = np.zeros_like(x)
        for i in range(self.n):
            for j in range(self.n):
                k = self.kernel(x[i, 0] - x[j, 0])
                out[i, 0] += 0.5 * x[j, 1] * k
                out[i, 1] += (-self.λ * 0.5 * x[j, 1] * k
                              + 1 / (2 * self.a_) * np.dot(x[i, 1], x[j, 1]) * (x[i, 0] - x[j, 0]) * k)
        return out.ravel()

    def σ(self, t, x=None):
        S = np.zeros((16, 8))
        for i in range(self.n):
            for k in range(2):
                S[4 * i + 2 + k, 2 * i + k] = self.σ_
        return S


def LandmarksTilde(a, σ, λ, qT):  # partialbridge_landmarks.jl:75-81,126-146
    """Auxiliary process of the landmarks bridge: the drift linearised with the positions frozen at the end
    configuration qT [n, 2]:  B~[q(i),p(j)] = k(qT_i - qT_j)/2 I, B~[p(i),p(j)] = -λ k(qT_i - qT_j)/2 I, β~ = 0, a~ = a."""
    P = Landmarks(a, σ, λ)
    qT = np.asarray(qT, dtype=np.float64).reshape(P.n, 2)
    B = np.zeros((16, 16))
    for i in range(P.n):
        for j in range(P.n):
            k = P.kernel(qT[i] - qT[j])
            for c in range(2):
                B[4 * i + c, 4 * j + 2 + c] += 0.5 * 1.0 * k
                B[4 * i + 2 + c, 4 * j + 2 + c] += -0.5 * 1.0 * P.λ * k
    S = P.σ(0.0)
    return LinearAux(B, np.zeros(16), S @ S.T)


class UserProcess(ContinuousTimeProcess):
    """A target process whose drift and diffusion are given as CUDA C source and compiled at run time (NVRTC) into the
    same path kernels the registry models use -- the counterpart of defining Bridge.b / Bridge.σ for an own struct
    (src/types.jl:23,32-33, src/Bridge.jl:105-106; partialbridge_fitzhugh.jl:44-46).

        drift   statements assigning o[0..d-1] from x[0..d-1] and par[...] (double precision; write fma() where a fused
                multiply-add is wanted: nothing is contracted implicitly)
        col     col[i] = column of the driving Wiener process entering component i, or -1
        sigma   sigma[i] = C expression in par[] for that entry of σ (None where col[i] < 0); constant in x
        par     parameter values (change them freely between calls: no recompilation)
        b, σ    optional Python callables for host-side inspection (Bridge.b, Bridge.σ)"""
    model_id = K.USER

    def __init__(self, d: int, dprime: int, drift: str, col: Sequence[int], sigma: Sequence[Optional[str]], par,
                 b: Optional[Callable] = None, σ: Optional[Callable] = None, ctx: Optional["Context"] = None):
        self.ctx = ctx or default_context()
        self.d, self.dprime = int(d), int(dprime)
        self.p = [float(v) for v in par]
        self._b, self._σ = b, σ
        colarr = (C.c_int32 * d)(*[int(c) for c in col])
        sigarr = (C.c_char_p * d)(*[None if s_ is None else str(s_).encode() for s_ in sigma])
        h = C.c_void_p()
        rc = lib.bb_user_model_create(self.ctx.h, d, dprime, drift.encode(), colarr, sigarr, C.byref(h))
        self.h = h
        self.log = lib.bb_user_model_log(h).decode() if h else ""
        if rc != 0:
            text = lib.bb_strerror(rc).decode() + ("\n" + self.log if self.log else "")
            if h:
                lib.bb_user_model_destroy(h)
                self.h = None
            raise BridgeError(rc, text)
        self.handle = lib.bb_user_model_handle(h)

    def par(self):
        return self.p

    def cmodel(self) -> K.Model:
        m = super().cmodel()
        m.reserved = self.handle
        return m

    def b(self, t, x):
        if self._b is None:
            raise NotImplementedError("no host-side b was given")
        return self._b(t, x, self.p)

    def σ(self, t, x=None):
        if self._σ is None:
            raise NotImplementedError("no host-side σ was given")
        return self._σ(t, x, self.p)

    sigma = σ

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            lib.bb_user_model_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def check_user_source(d, dprime, drift, col, sigma, gk=0, gm=0, auxm=1, rng=0):
    """Compile-only check of a UserProcess definition (works without a device): -> (cubin bytes, NVRTC log); raises
    BridgeError(ERR_USERSRC) with the log if the source does not compile."""
    colarr = (C.c_int32 * d)(*[int(c) for c in col])
    sigarr = (C.c_char_p * d)(*[None if s_ is None else str(s_).encode() for s_ in sigma])
    log = C.create_string_buffer(1 << 16)
    rc = lib.bb_user_source_check(d, dprime, drift.encode(), colarr, sigarr, gk, gm, auxm, rng, log, len(log))
    if rc < 0:
        raise BridgeError(rc, lib.bb_strerror(rc).decode() + "\n" + log.value.decode())
    return rc, log.value.decode()


class LinearAux:
    """An auxiliary process given by B(t), β(t), a(t) (constants or callables), the protocol the
    guided proposals use (Bridge.B, Bridge.β, Bridge.a; partialbridge_fitzhugh.jl:99-116)."""

    def __init__(self, B, β, a, constdiff: bool = True):
        self._B, self._β, self._a = B, β, a
        self.constdiff = constdiff  # Bridge.constdiff(::Aux): the scripts' auxiliary processes all declare `true`
        self.is_const = not (callable(B) or callable(β) or callable(a))
        B0 = np.atleast_2d(f64(B(0.0) if callable(B) else B))
        self.d = B0.shape[0]

    def B(self, t):
        return np.atleast_2d(f64(self._B(t) if callable(self._B) else self._B))

    def β(self, t):
        return np.atleast_1d(f64(self._β(t) if callable(self._β) else self._β))

    def a(self, t, x=None):
        return np.atleast_2d(f64(self._a(t) if callable(self._a) else self._a))


class LinearAppr:
    """LinearAppr(tt, xx, B, b, Σ)  src/linpro.jl:181-194: a linear auxiliary process tabulated on a grid,
        _b((i,s), x) = B[i] (x - xx[i]) + b[i],   B((i,s)) = B[i],   β((i,s)) = b[i] - B[i] xx[i],   a((i,s)) = Σ[i] Σ[i]',
    constdiff = false.  `linearappr(Y, P)` (src/linpro.jl:196) builds it along a trajectory Y from bderiv, b, σ of P.
    As the auxiliary process of GuidedBridge / PartialBridge / PartialBridgeνH its coefficients are needed BETWEEN grid
    points (Ralston stages): they are interpolated linearly.  [The reference's own constructor for this auxiliary type,
    GuidedBridge(tt, P, Pt::LinearAppr, v) -> solvebackwardi!(Heun(), ...), does not run: kerneli uses an undefined `i`
    (src/ode.jl:98-102), so there is no reference table to match; the tables here come from the same R3 scheme as for
    every other auxiliary process.]"""
    is_const = False
    constdiff = False

    def __init__(self, tt, xx, B, b, Σ):
        self.tt = f64(tt)
        N = len(self.tt)
        self.xx = f64(xx).reshape(N, -1)
        self.d = self.xx.shape[1]
        self.Bs = f64(B).reshape(N, self.d, self.d)
        self.bs = f64(b).reshape(N, self.d)
        Σ = f64(Σ)
        self.Σs = Σ.reshape(N, self.d, -1)
        self.βs = self.bs - np.einsum("nij,nj->ni", self.Bs, self.xx)
        self.as_ = np.einsum("nik,njk->nij", self.Σs, self.Σs)

    def _lerp(self, arr, t):
        tt = self.tt
        if t <= tt[0]:
            return arr[0]
        if t >= tt[-1]:
            return arr[-1]
        i = int(np.searchsorted(tt, t, side="right")) - 1
        w = (t - tt[i]) / (tt[i + 1] - tt[i])
        return (1.0 - w) * arr[i] + w * arr[i + 1]

    def B(self, t):
        """B((i,s), P) for an index-time pair (i, s) (0-based i), or the interpolated value at a plain time."""
        return self.Bs[t[0]] if isinstance(t, tuple) else self._lerp(self.Bs, t)

    def β(self, t):
        return self.βs[t[0]] if isinstance(t, tuple) else self._lerp(self.βs, t)

    def a(self, t, x=None):
        return self.as_[t[0]] if isinstance(t, tuple) else self._lerp(self.as_, t)

    def b(self, t, x):  # _b((i,s), x, P)  src/linpro.jl:189
        i = t[0]
        return self.Bs[i] @ (np.asarray(x) - self.xx[i]) + self.bs[i]


def bderiv(t, x, P):
    """Bridge.bderiv(t, x, P): the Jacobian of the drift (src/Models.jl:49-53 Lorenz, src/linpro.jl:82 LinPro; the
    FitzHugh-Nagumo models by differentiation of src/Models.jl:18, partialbridge_fitzhugh.jl:44)."""
    x = np.asarray(x, dtype=np.float64)
    if isinstance(P, LinPro):
        return P.Bm
    if isinstance(P, Lorenz):
        θ = P.p
        return np.array([[-θ[0], θ[0], 0.0], [θ[1] - x[2], -1.0, -x[0]], [x[1], x[0], -θ[2]]])
    if isinstance(P, (FitzHughNagumo, FitzhughDiffusion)):
        ϵ, _, γ = P.p[0], P.p[1], P.p[2]
        return np.array([[(1.0 - 3.0 * x[0] * x[0]) / ϵ, -1.0 / ϵ], [γ, -1.0]])
    if isinstance(P, OrnsteinUhlenbeck):
        return np.array([[-P.β]])
    raise NotImplementedError(f"bderiv is not defined for {type(P).__name__}")


def linearappr(Y: "SamplePath", P) -> LinearAppr:
    """linearappr(Y, P) = LinearAppr(Y.tt, Y.yy, bderiv.(tt, yy), b.(tt, yy), σ.(tt, yy))  src/linpro.jl:196"""
    yy = Y._as2d()
    return LinearAppr(Y.tt, yy, [bderiv(t, x, P) for t, x in zip(Y.tt, yy)], [np.atleast_1d(P.b(t, x)) for t, x in zip(Y.tt, yy)],
                      [np.atleast_2d(P.σ(t, x)) for t, x in zip(Y.tt, yy)])


class _AuxC:
    """bb_aux for a backward solve on grid tt (keeps the arrays alive)."""

    def __init__(self, Pt, tt):
        d = np.atleast_2d(Pt.B(tt[0])).shape[0]
        self.d = d
        if getattr(Pt, "is_const", False):
            self.B, self.beta, self.a = f64(Pt.B(tt[0])), f64(Pt.β(tt[0])), f64(Pt.a(tt[0]))
            self.al = None
            const = 1
        else:
            N = len(tt)
            self.B = np.empty((N - 1, 3, d, d)); self.beta = np.empty((N - 1, 3, d))
            self.a = np.empty((N - 1, 3, d, d)); self.al = np.empty((N - 1, d, d))
            for i in range(N - 1):
                t, h = tt[i + 1], tt[i] - tt[i + 1]
                for k, c in enumerate((0.0, 0.5, 0.75)):  # stage times of src/ode.jl:44-49
                    s = t + c * h
                    self.B[i, k] = Pt.B(s); self.beta[i, k] = Pt.β(s); self.a[i, k] = Pt.a(s)
                self.al[i] = Pt.a(tt[i])
            const = 0
        self.c = K.Aux(d, const, ptr(self.B), ptr(self.beta), ptr(self.a), ptr(self.al))


def _aux_on_grid(Pt, tt):
    """B̃(tt[i]), β̃(tt[i]) for llikelihood (b̃ = B̃x+β̃, src/partialbridgenuH.jl:176)."""
    if getattr(Pt, "is_const", False):
        return f64(Pt.B(tt[0])), f64(Pt.β(tt[0])), 1
    Bt = np.stack([np.atleast_2d(Pt.B(t)) for t in tt])
    bt = np.stack([np.atleast_1d(Pt.β(t)) for t in tt])
    return f64(Bt), f64(bt), 0


# ----------------------------------------------------------------------------------------------- sample paths
class SamplePath:
    """tt::Vector{Float64}, yy::Vector{T}  (src/types.jl:71-76); yy is [N] (T=Float64) or [N, d]."""

    def __init__(self, tt, yy):
        self.tt = f64(tt)
        self.yy = np.array(yy, dtype=np.float64)
        if self.yy.shape[0] != self.tt.shape[0]:
            raise BridgeError(K.ERR_DIM, lib.bb_strerror(K.ERR_DIM).decode())

    def __len__(self):
        return len(self.tt)

    def copy(self):
        return SamplePath(self.tt.copy(), self.yy.copy())

    @property
    def dim(self):
        return 1 if self.yy.ndim == 1 else self.yy.shape[1]

    def _as2d(self):
        return self.yy.reshape(len(self.tt), -1)


class VSamplePath(SamplePath):
    """yy::Matrix of size d x N (src/types.jl:123-130); stored here as [N, d] views of the same data."""

    def __init__(self, tt, yy):
        yy = np.asarray(yy, dtype=np.float64)
        if yy.ndim != 2 or yy.shape[1] != len(tt):
            raise BridgeError(K.ERR_DIM, lib.bb_strerror(K.ERR_DIM).decode())
        super().__init__(tt, yy.T.copy())


def samplepath(tt, v) -> SamplePath:  # src/types.jl:78-81 (aliases tt)
    tt = f64(tt)
    v = np.asarray(v, dtype=np.float64)
    yy = np.zeros((len(tt),) + v.shape)
    yy[...] = v
    sp = SamplePath.__new__(SamplePath)
    sp.tt, sp.yy = tt, yy
    return sp


# ----------------------------------------------------------------------------------------------- ensemble
class PathEnsemble:
    """P chains x S segments x N grid points resident in HBM (bb_ens)."""

    def __init__(self, P: int, S: int, N: int, d: int, dprime: int, double_buffer: bool = True,
                 store_x: bool = True, ctx: Optional[Context] = None, chain_offset: int = 0):
        self.ctx = ctx or default_context()
        self.P, self.S, self.N, self.d, self.dprime = P, S, N, d, dprime
        flags = (K.ENS_DOUBLE_BUFFER if double_buffer else 0) | (0 if store_x else K.ENS_NO_X)
        h = C.c_void_p()
        check(lib.bb_ens_create(self.ctx.h, P, S, N, d, dprime, flags, C.byref(h)))
        self.h = h
        self.has_x = store_x
        self._last = None  # (P, guides) of the last pCN step: what refresh_x needs
        self._theta = False  # per-chain parameters attached (theta_attach_)
        if chain_offset:
            check(lib.bb_ens_set_chain_offset(self.h, chain_offset))

    def close(self):
        if self.h:
            lib.bb_ens_destroy(self.h)
            self.h = None

    # ---- configuration
    def set_grid(self, seg: int, tt):
        tt = f64(tt)
        check(lib.bb_ens_set_grid(self.h, seg, ptr(tt), len(tt)))

    def set_start(self, u):
        u = f64(u)
        if u.ndim <= 1 and u.size == self.d:
            check(lib.bb_ens_set_start(self.h, ptr(u), u.size, 1))
        else:
            check(lib.bb_ens_set_start(self.h, ptr(u), u.size, 0))

    # ---- data movement; host layout [np][S][N][k]
    def upload(self, what, arr, which=K.CUR, p0=0):
        k = self.dprime if what == K.W else self.d
        arr = f64(arr).reshape(-1, self.S, self.N, k)
        check(lib.bb_ens_upload(self.h, what, which, p0, arr.shape[0], ptr(arr)))

    def download(self, what, which=K.CUR, p0=0, np_=None, out=None):
        k = self.dprime if what == K.W else self.d
        n = self.P - p0 if np_ is None else np_
        if what == K.X and which == K.CUR:
            self.refresh_x_()
        if out is None:
            out = np.empty((n, self.S, self.N, k))
        check(lib.bb_ens_download(self.h, what, which, p0, n, ptr(out)))
        return out

    def _f(self, field, width=1):
        out = np.empty((self.P, width) if width > 1 else self.P)
        check(lib.bb_ens_get_f64(self.h, field, 0, self.P, ptr(out)))
        return out

    @property
    def ll(self):
        return self._f(K.F_LL)

    @property
    def ll_prop(self):
        return self._f(K.F_LL_PROP)

    @property
    def logu(self):
        return self._f(K.F_LOGU)

    @property
    def xend(self):
        return self._f(K.F_XEND, self.d).reshape(self.P, self.d)

    @property
    def xend_prop(self):
        return self._f(K.F_XEND_PROP, self.d).reshape(self.P, self.d)

    def set_ll(self, ll):
        ll = f64(ll)
        check(lib.bb_ens_set_ll(self.h, 0, ll.size, ptr(ll)))

    @property
    def accepted(self):
        out = np.empty(self.P, dtype=np.uint8)
        check(lib.bb_ens_get_accepted(self.h, 0, self.P, ptr(out)))
        return out

    @property
    def acc(self) -> int:
        v = C.c_int64(0)
        check(lib.bb_ens_get_acc(self.h, C.byref(v)))
        return v.value

    def reset_acc(self):
        check(lib.bb_ens_reset_acc(self.h))

    @property
    def acc_device_ptr(self) -> int:
        return lib.bb_ens_acc_device_ptr(self.h)

    @property
    def nbytes(self) -> int:
        return lib.bb_ens_bytes(self.h)

    # ---- compute
    @staticmethod
    def _garr(guides):
        hs = [g._guide if hasattr(g, "_guide") else g for g in guides]
        return (C.c_void_p * len(hs))(*[h.value if isinstance(h, C.c_void_p) else h for h in hs])

    def sample_(self, seed: int, stream: int = 0):
        """sample!(W, Wiener()) for every chain and segment."""
        check(lib.bb_wiener_sample(self.h, seed, stream))

    def euler_(self, P: ContinuousTimeProcess):
        m = P.cmodel()
        check(lib.bb_euler(self.h, C.byref(m)))

    def guided_mdb_(self, P: ContinuousTimeProcess, guides):
        """solve!(Mdb(), Y, u, W, P°) on a guided proposal, S = 1 (bb_guided_mdb)."""
        m = P.cmodel()
        self._last = (P, list(guides))
        check(lib.bb_guided_mdb(self.h, C.byref(m), self._garr(guides)))

    def solve_scheme_(self, P: ContinuousTimeProcess, scheme: int):
        """solve! with StratonovichEuler / StochasticHeun / StochasticRungeKutta for a plain target (bb_solve_scheme)."""
        m = P.cmodel()
        check(lib.bb_solve_scheme(self.h, C.byref(m), scheme))

    def sample_euler_(self, P: ContinuousTimeProcess, seed: int, stream: int = 0):
        m = P.cmodel()
        check(lib.bb_sample_euler(self.h, C.byref(m), seed, stream))

    def guided_euler_ll_(self, P, guides, skip: int = 0, store_x: bool = True, ll: bool = True):
        m = P.cmodel()
        flags = (K.RUN_STORE_X if store_x else 0) | (0 if ll else K.RUN_NO_LL)
        self._last = (P, list(guides))
        check(lib.bb_guided_euler_ll(self.h, C.byref(m), self._garr(guides), skip, flags))

    def pcn_step_host_(self, P, guides, ρ: float, seed: int, it: int, W, Wo, Xo=None, llo=None, accepted=None,
                       skip: int = 0, skip_rejected: bool = False):
        """One pCN iteration on HOST arrays (the reference loop's dataflow): W [P,S,N,d'] in; Wo, Xo, llo, accepted
        out.  Pipelined H2D | kernel | D2H over chain slabs; use pinned arrays for full overlap.  skip_rejected: rows of
        Wo / Xo of chains that reject are left as they were (the loop only swaps proposals in on accept); with pinned,
        device-mapped arrays the accepted rows are written straight into them and the rest never crosses the link."""
        m = P.cmodel()
        self._last = (P, list(guides))
        flags = (K.RUN_STORE_X if Xo is not None else 0) | (K.RUN_SKIP_REJECTED if skip_rejected else 0)
        check(lib.bb_pcn_step_host(self.h, C.byref(m), self._garr(guides), ρ, seed, it, skip, flags, ptr(W), ptr(Wo),
                                   ptr(Xo), ptr(llo), ptr(accepted)))

    # ---- pooled online statistics: mcstart / mcnext! / mcstats of src/mclog.jl for the whole ensemble
    def mc_reset_(self):
        check(lib.bb_ens_mc_reset(self.h))

    def mc_update_(self):
        """mcnext!: add the CURRENT path of every chain to the running first and second moments."""
        self.refresh_x_()
        check(lib.bb_ens_mc_update(self.h))

    def mc_stats(self):
        """mcstats: (mean [S,N,d], cov [S,N,d,d], n) over all chains and all mc_update_ calls."""
        mean = np.empty((self.S, self.N, self.d)); cov = np.empty((self.S, self.N, self.d, self.d)); n = C.c_int64(0)
        check(lib.bb_ens_mc_stats(self.h, ptr(mean), ptr(cov), C.byref(n)))
        return mean, cov, n.value

    def mc_band(self):
        """mcband: marginal 95 % band mean -/+ Q std with Q = sqrt(2) erfinv(0.95)  (src/mclog.jl:75-85)."""
        mean, cov, _ = self.mc_stats()
        std = np.sqrt(np.einsum("snii->sni", cov))
        return mean - MCBAND_Q * std, mean + MCBAND_Q * std

    # ---- the same per chain over the recorded iterations: the reference's own mcstart / mcnext! / mcstats / mcband
    # (src/mclog.jl:22-24, 47-56, 75-93) as the scripts keep it, `mcstate = [mcnext!(mcstate[i], XX[i].yy) ...]`
    def chain_mc_reset_(self):
        """mcstart for every chain: m = 0, m2 = 0, k = 0 (allocates (1 + d) x the memory of X on first use)."""
        check(lib.bb_ens_chain_mc_reset(self.h))

    def chain_mc_update_(self):
        """mcnext!(mc_p, X_p.yy) for every chain p with its CURRENT path (Welford, the reference's operation order)."""
        self.refresh_x_()
        check(lib.bb_ens_chain_mc_update(self.h))

    def chain_mc_stats(self, p0: int = 0, np_: Optional[int] = None):
        """mcstats per chain: (mean [np,S,N,d], cov = m2/(k-1) [np,S,N,d,d], k)."""
        n = self.P - p0 if np_ is None else np_
        mean = np.empty((n, self.S, self.N, self.d)); cov = np.empty((n, self.S, self.N, self.d, self.d)); k = C.c_int64(0)
        check(lib.bb_ens_chain_mc_stats(self.h, p0, n, ptr(mean), ptr(cov), C.byref(k)))
        return mean, cov, k.value

    def chain_mc_band(self, p0: int = 0, np_: Optional[int] = None):
        """mcband per chain: (lower, upper) [np,S,N,d] = m -/+ Q sqrt(diag(m2) (1/(k-1)))  (src/mclog.jl:75-85)."""
        n = self.P - p0 if np_ is None else np_
        lo = np.empty((n, self.S, self.N, self.d)); hi = np.empty_like(lo)
        check(lib.bb_ens_chain_mc_band(self.h, p0, n, ptr(lo), ptr(hi)))
        return lo, hi

    def refresh_x_(self):
        """Make X the CURRENT path of every chain again (X holds the last proposal; chains that rejected it
        get their path recomputed from W by the same guided Euler kernel).  No-op if nothing is stale."""
        if self._theta:
            check(lib.bb_theta_refresh_x(self.h))
        elif self._last is not None:
            P, guides = self._last
            m = P.cmodel()
            check(lib.bb_ens_refresh_x(self.h, C.byref(m), self._garr(guides)))

    def llikelihood_(self, P, guides, skip: int = 0):
        self.refresh_x_()
        m = P.cmodel()
        check(lib.bb_llikelihood(self.h, C.byref(m), self._garr(guides), skip))

    def innovations_(self, P, guides=None):
        m = P.cmodel()
        check(lib.bb_innovations(self.h, C.byref(m), None if guides is None else self._garr(guides)))

    def pcn_step_(self, P, guides, ρ: float, seed: int, it: int, skip: int = 0, store_x: bool = True):
        """One pCN / Metropolis-Hastings update of every chain (test/partialbridgenuH.jl:176-191)."""
        m = P.cmodel()
        self._last = (P, list(guides))
        check(lib.bb_pcn_step(self.h, C.byref(m), self._garr(guides), ρ, seed, it, skip,
                              K.RUN_STORE_X if store_x else 0))


    # ---- per-chain parameters: the `updateparams` branch of partialbridge_bolus3.jl:248-365 for P chains
    def theta_attach_(self, P: ContinuousTimeProcess, L, Σ, ϵ: float, obs_v, aux_kind: int = K.AUX_FHN_MATCHING,
                      priors=None, start_sd: float = 0.0, start_dir=None):
        """Give every chain its own copy θ_p of P's parameters (all chains start at P's) and its own guiding tables.
        L (m x d), Σ (m x m), ϵ: observation scheme and H⁺ = I/ϵ right of the last observation (bolus3.jl:162-165);
        obs_v[s]: observation at the right end of segment s; priors: {index: ("gamma", shape, scale)} (logπ, :237)."""
        L = np.atleast_2d(f64(L)); m, d = L.shape
        Σ = np.atleast_2d(f64(Σ))
        sp = K.ThetaSpec()
        sp.m, sp.aux_kind, sp.eps = m, aux_kind, float(ϵ)
        for i in range(m):
            for j in range(d):
                sp.L[i * d + j] = L[i, j]
            for j in range(m):
                sp.Sigma[i * m + j] = Σ[i, j]
        obs_v = np.asarray(obs_v, dtype=np.float64).reshape(self.S, m)
        for s in range(self.S):
            for i in range(m):
                sp.v[s][i] = obs_v[s, i]
        for k, pr in (priors or {}).items():
            if pr[0] != "gamma":
                raise ValueError("priors: ('gamma', shape, scale)")
            sp.prior_kind[k], sp.prior_a[k], sp.prior_b[k] = K.PRIOR_GAMMA, float(pr[1]), float(pr[2])
        if start_sd:  # joint random walk on the starting point in parameter steps (bolus3.jl:311-318)
            sp.start_sd = float(start_sd)
            for k, v in enumerate(f64(start_dir).ravel()):
                sp.start_dir[k] = float(v)
        mdl = P.cmodel()
        check(lib.bb_theta_attach(self.h, C.byref(mdl), C.byref(sp)))
        self._theta = True

    def theta_start(self, which=K.CUR):
        out = np.empty((self.P, self.d))
        check(lib.bb_theta_get_start(self.h, which, 0, self.P, ptr(out)))
        return out

    def set_theta(self, θ, p0: int = 0):
        θ = f64(θ).reshape(-1, K.BB_NTHETA)
        check(lib.bb_theta_set(self.h, p0, θ.shape[0], ptr(θ)))

    def theta(self, which=K.CUR):
        out = np.empty((self.P, K.BB_NTHETA))
        check(lib.bb_theta_get(self.h, which, 0, self.P, ptr(out)))
        return out

    def theta_guides_(self):
        """Backward pass (Lyapunov step + observation updates) for the current θ of every chain."""
        check(lib.bb_theta_guides(self.h))

    def theta_left(self, which=K.CUR):
        """[P, d+d*d+4]: ν(0), H⁺(0), C, logpdfnormal(x0-ν(0), H⁺(0)), trace term, logπ(θ) of the last backward pass."""
        out = np.empty((self.P, self.d + self.d * self.d + 4))
        check(lib.bb_theta_get_left(self.h, which, 0, self.P, ptr(out)))
        return out

    def theta_tables(self, p: int):
        ν = np.empty((self.S, self.N, self.d)); H = np.empty((self.S, self.N, self.d, self.d))
        check(lib.bb_theta_get_tables(self.h, p, ptr(ν), ptr(H)))
        return ν, H

    def theta_guided_euler_ll_(self, skip: int = 0, store_x: bool = True):
        check(lib.bb_theta_guided_euler_ll(self.h, skip, K.RUN_STORE_X if store_x else 0))

    def theta_pcn_step_(self, ρ: float, seed: int, it: int, skip: int = 0, store_x: bool = True):
        check(lib.bb_theta_pcn_step(self.h, ρ, seed, it, skip, K.RUN_STORE_X if store_x else 0))

    def theta_param_step_(self, rw_sd, seed: int, it: int, skip: int = 0, store_x: bool = True):
        """One parameter-update MH iteration with W held fixed; rw_sd[k] = random-walk sd of parameter k (0: fixed)."""
        sd = np.zeros(K.BB_NTHETA); rw = f64(rw_sd).ravel(); sd[:rw.size] = rw
        check(lib.bb_theta_param_step(self.h, ptr(sd), seed, it, skip, K.RUN_STORE_X if store_x else 0))

    def theta_block_step_(self, s_lo: int, s_hi: int, ρ: float, seed: int, it: int, hzero: float = 0.1, skip: int = 0):
        """Blocked path update of segments s_lo .. s_hi-1 (0-based) of every chain: the `updateparams == false` branch of
        project_partialbridge/partialbridge_bolus3.jl:258-355 (ind = (kup-1):-1:klow with klow = s_lo+1, kup = s_hi+1).
        A block that does not end the chain is conditioned on the chain's own path at its right end with H⁺ = hzero·I
        (Hzero⁺, :234); see bb_theta_block_step in include/bridge_b200.h for the step-by-step correspondence."""
        check(lib.bb_theta_block_step(self.h, s_lo, s_hi, ρ, hzero, seed, it, skip))

    def theta_block(self):
        """[P, 5 + 2S] numbers of the last block update: logpdfnormal start terms (current, proposal), Σ ll_temp, Σ ll°,
        diffll, then (ll_temp[s], ll°[s]) per segment."""
        out = np.empty((self.P, 5 + 2 * self.S))
        check(lib.bb_theta_get_block(self.h, 0, self.P, ptr(out)))
        return out

    def theta_blocked_sweep_(self, rng, ρ: float, seed: int, it0: int, hzero: float = 0.1, skip: int = 0):
        """One sweep of block updates over the whole chain, as the `while !finished` loop of bolus3.jl:258-362 with
        updateparams == false runs it: klow = 1; segnum_update = sample(1:obsnum-klow) (drawn from the host generator
        `rng`, one draw for the whole ensemble), kup = klow + segnum_update, update segments klow .. kup-1, klow = kup,
        until klow == obsnum.  Returns the list of (s_lo, s_hi) blocks; noise stream it0 + k drives block k."""
        obsnum = self.S + 1
        klow, blocks = 1, []
        while klow != obsnum:
            kup = klow + int(rng.integers(1, obsnum - klow + 1))
            self.theta_block_step_(klow - 1, kup - 1, ρ, seed, it0 + len(blocks), hzero, skip)
            blocks.append((klow - 1, kup - 1))
            klow = kup
        return blocks

    @property
    def acc_theta(self) -> int:
        v = C.c_int64(0)
        check(lib.bb_theta_get_acc(self.h, C.byref(v)))
        return v.value

    def theta_acc_device_ptr(self) -> int:
        return lib.bb_theta_acc_device_ptr(self.h) or 0


def _small_ens(ctx: Context, S, N, d, dp, double_buffer=False) -> PathEnsemble:
    key = (S, N, d, dp, double_buffer)
    e = ctx._small.get(key)
    if e is None:
        e = PathEnsemble(1, S, N, d, dp, double_buffer=double_buffer, ctx=ctx)
        ctx._small[key] = e
    return e


# ----------------------------------------------------------------------------------------------- guided proposals
class _Proposal:
    kind = 0
    m = 0

    def _make_guide(self, A, b, Mm=None, v=None):
        Bt, bt, const = _aux_on_grid(self.Pt, self.tt)
        h = C.c_void_p()
        A, b = f64(A), f64(b)
        Mm = None if Mm is None else f64(Mm)
        v = None if v is None else f64(v)
        # constdiff(P°) = constdiff(Target) && constdiff(Pt)  (src/partialbridge.jl:64, guip.jl:199, a TRAIT of the two
        # processes, not a comparison of a and a~): only if it is false do the extra log-likelihood terms of
        # src/partialbridge.jl:79-84 apply.  A pair such as Diffusion / DiffusionAux of partialbridge_bolus3.jl:54,70
        # (both constdiff = true, σ1 != σ2) runs WITHOUT them in the reference, and so it does here.
        self.constdiff = bool(getattr(self.Target, "constdiff", True)) and bool(getattr(self.Pt, "constdiff", True))
        if not self.constdiff:
            a_t = np.atleast_2d(f64(self.Target.a(self.tt[0], None)))
            if const:
                Ad = a_t - np.atleast_2d(f64(self.Pt.a(self.tt[0])))
                ad_const = 1
            else:
                Ad = np.stack([a_t - np.atleast_2d(f64(self.Pt.a(t))) for t in self.tt])
                ad_const = 0
            Ad = f64(Ad)
            check(lib.bb_guide_create_ncd(self.ctx.h, self.kind, len(self.tt), self.Target.d, self.m, ptr(self.tt),
                                          ptr(A), ptr(b), ptr(Mm), ptr(v), ptr(Bt), ptr(bt), const, ptr(Ad), ad_const,
                                          C.byref(h)))
        else:
            check(lib.bb_guide_create(self.ctx.h, self.kind, len(self.tt), self.Target.d, self.m, ptr(self.tt), ptr(A),
                                      ptr(b), ptr(Mm), ptr(v), ptr(Bt), ptr(bt), const, C.byref(h)))
        self._guide = h

    def __del__(self):
        g = getattr(self, "_guide", None)
        if g and self.ctx.h:
            lib.bb_guide_destroy(g)
            self._guide = None


class GuideTables(_Proposal):
    """A guided proposal from tables the caller already holds (values on the grid `tt`, layouts of
    bb_guide_create): kind NUH (A=H, b=ν), HV (A=H♢, b=V) or LMMU (A=L, b=μ, Mm=M, v)."""

    def __init__(self, kind, tt, P, A, b, Bt, betat, Mm=None, v=None, aux_const=True, m=0, Adiff=None,
                 adiff_const=True, ctx=None):
        self.ctx = ctx or default_context()
        self.kind, self.m = kind, m
        self.tt, self.Target, self.Pt = np.array(tt, dtype=np.float64), P, None
        h = C.c_void_p()
        A, b, Bt, betat = f64(A), f64(b), f64(Bt), f64(betat)
        Mm = None if Mm is None else f64(Mm)
        v = None if v is None else f64(v)
        if Adiff is None:
            check(lib.bb_guide_create(self.ctx.h, kind, len(self.tt), P.d, m, ptr(self.tt), ptr(A), ptr(b), ptr(Mm),
                                      ptr(v), ptr(Bt), ptr(betat), 1 if aux_const else 0, C.byref(h)))
        else:
            Adiff = f64(Adiff)
            check(lib.bb_guide_create_ncd(self.ctx.h, kind, len(self.tt), P.d, m, ptr(self.tt), ptr(A), ptr(b),
                                          ptr(Mm), ptr(v), ptr(Bt), ptr(betat), 1 if aux_const else 0, ptr(Adiff),
                                          1 if adiff_const else 0, C.byref(h)))
        self._guide = h


def _update_nuHC(ctx, L, Σ, v, ϵ):
    L = np.atleast_2d(f64(L)); m, d = L.shape
    Σ = np.atleast_2d(f64(Σ)); v = np.atleast_1d(f64(v))
    if v.size != m:
        raise BridgeError(K.ERR_ASSERT_M, lib.bb_strerror(K.ERR_ASSERT_M).decode())
    ν = np.zeros(d); Hp = np.zeros((d, d)); Cc = C.c_double(0)
    check(lib.bb_update_nuHC(ctx.h, d, m, ptr(L), ptr(Σ), ptr(v), float(ϵ), ptr(ν), ptr(Hp), C.byref(Cc)))
    return ν, Hp, Cc.value


def _backward_nuH(ctx, method, tt, Pt, νend, Hendp, C0=0.0):
    tt = f64(tt); aux = _AuxC(Pt, tt); N, d = len(tt), aux.d
    νend = np.atleast_1d(f64(νend)); Hendp = np.atleast_2d(f64(Hendp))
    ν = np.zeros((N, d)); H = np.zeros((N, d, d)); νl = np.zeros(d); Hl = np.zeros((d, d)); Cc = C.c_double(0)
    check(lib.bb_backward_nuH(ctx.h, method, N, d, ptr(tt), C.byref(aux.c), ptr(νend), ptr(Hendp), float(C0),
                              ptr(ν), ptr(H), ptr(νl), ptr(Hl), C.byref(Cc)))
    return ν, H, νl, Hl, Cc.value


class PartialBridgeνH(_Proposal):
    """PartialBridgeνH(tt, P, Pt, L, v, ϵ, Σ)  src/partialbridgenuH.jl:134-145; fields Target, Pt, tt, ν, H, C."""
    kind = K.GUIDE_NUH

    def __init__(self, tt, P, Pt, L=None, v=None, ϵ=None, Σ=None, *, _tables=None, ctx=None):
        self.ctx = ctx or default_context()
        self.tt, self.Target, self.Pt = np.array(tt, dtype=np.float64), P, Pt  # tt = collect(tt_)
        if _tables is not None:
            self.ν, self.H, self.C = _tables
        else:
            L = np.atleast_2d(f64(L))
            if Σ is None:
                Σ = np.zeros((L.shape[0], L.shape[0]))
            νT, HpT, C0 = _update_nuHC(self.ctx, L, Σ, v, ϵ)
            self.ν, self.H, _, _, self.C = _backward_nuH(self.ctx, K.ODE_R3, self.tt, Pt, νT, HpT, C0)
        self._make_guide(self.H, self.ν)


PartialBridgenuH = PartialBridgeνH


def partialbridgeνH(tt, P, Pt, νend, Hendp, ctx=None):
    """Bridge.partialbridgeνH(tt, P, Pt, νend, Hend⁺) -> (P°, ν, H⁺, C)  src/partialbridgenuH.jl:148-155
    (Lyapunov backward step; the unbound C of the reference is taken as 0.0)."""
    ctx = ctx or default_context()
    ν, H, νl, Hl, Cc = _backward_nuH(ctx, K.ODE_LYAP, tt, Pt, νend, Hendp, 0.0)
    return PartialBridgeνH(tt, P, Pt, _tables=(ν, H, Cc), ctx=ctx), νl, Hl, Cc


partialbridgenuH = partialbridgeνH


class _ChainSegment(_Proposal):
    """One segment of a PartialBridgeνHChain: a PartialBridgeνH whose tables live on the device only; ν and H are read
    back on first use (values at grid points 0 .. N-2; the terminal values are not kept by the path kernels: NaN)."""
    kind = K.GUIDE_NUH
    constdiff = True

    def __init__(self, ctx, tt, P, Pt, handle):
        self.ctx, self.tt, self.Target, self.Pt, self._guide = ctx, tt, P, Pt, handle
        self._tab = None

    def _fetch(self):
        if self._tab is None:
            N, d = len(self.tt), self.Target.d
            ν = np.empty((N, d)); H = np.empty((N, d, d))
            check(lib.bb_guide_download_nuH(self._guide, ptr(ν), ptr(H)))
            self._tab = (ν, H)
        return self._tab

    @property
    def ν(self):
        return self._fetch()[0]

    @property
    def H(self):
        return self._fetch()[1]


class PartialBridgeνHChain:
    """The backward pass of a chain of segments -- the script loop of partialbridge_bolus3.jl:162-180 -- as ONE device
    launch (bb_guides_chain_nuH): ν = 0, H⁺ = I/ϵ right of the last observation, gpupdate with vs[-1], then for every
    segment from right to left `partialbridgeνH` (Lyapunov step) and the observation update.  The guiding tables are
    written where the path kernels read them; `update_` re-runs the pass in place for new auxiliary processes /
    observations (a parameter update of the sampler) without any table traffic to or from the host.
    `segments` are proposals usable wherever a PartialBridgeνH is (guided_euler_ll_, pcn_step_, ...)."""

    def __init__(self, grids, P, Pts, L, Σ, vs, ϵ, method=Lyap, ctx=None):
        self.ctx = ctx or default_context()
        self.Target = P
        self.tt = f64(np.stack([f64(g) for g in grids]))
        self.S, self.N = self.tt.shape
        self.L = np.atleast_2d(f64(L)); self.m, self.d = self.L.shape
        self.Σ = np.atleast_2d(f64(Σ)); self.ϵ = float(ϵ)
        self.method = K.ODE_LYAP if method in (Lyap, K.ODE_LYAP) or isinstance(method, Lyap) else K.ODE_R3
        self._handles = (C.c_void_p * self.S)()
        self.segments: List[_ChainSegment] = []
        self.update_(Pts, vs)

    def update_(self, Pts, vs):
        vs = f64(np.asarray(vs, dtype=np.float64).reshape(self.S, self.m))
        auxs = [_AuxC(Pt, self.tt[s]) for s, Pt in enumerate(Pts)]
        if any(not a.c.is_const for a in auxs):
            raise BridgeError(K.ERR_UNSUPPORTED, "PartialBridgeνHChain: constant auxiliary processes (one per segment)")
        arr = (K.Aux * self.S)(*[a.c for a in auxs])
        νl = np.zeros(self.d); Hl = np.zeros((self.d, self.d)); Cc = C.c_double(0)
        check(lib.bb_guides_chain_nuH(self.ctx.h, self.method, self.S, self.N, self.d, self.m, ptr(self.tt), arr,
                                      ptr(self.L), ptr(self.Σ), ptr(vs), self.ϵ, self._handles, ptr(νl), ptr(Hl),
                                      C.byref(Cc)))
        if not self.segments:
            self.segments = [_ChainSegment(self.ctx, self.tt[s], self.Target, Pts[s], C.c_void_p(self._handles[s]))
                             for s in range(self.S)]
        else:
            for s, seg in enumerate(self.segments):
                seg.Pt, seg._tab = Pts[s], None
        self.ν_left, self.Hplus_left, self.C = νl, Hl, Cc.value
        return self

    def __iter__(self):
        return iter(self.segments)

    def __len__(self):
        return self.S

    def __getitem__(self, s):
        return self.segments[s]


class GuidedBridge(_Proposal):
    """GuidedBridge(tt, P, Pt, v[, h♢])  src/guip.jl:172-180; fields Target, Pt, tt, H♢, V."""
    kind = K.GUIDE_HV

    def __init__(self, tt, P, Pt, v, hdia=None, *, ctx=None):
        self.ctx = ctx or default_context()
        self.tt, self.Target, self.Pt = np.array(tt, dtype=np.float64), P, Pt  # tt = collect(tt_)
        aux = _AuxC(Pt, self.tt); N, d = len(self.tt), aux.d
        v = np.atleast_1d(f64(v))
        he = None if hdia is None else np.atleast_2d(f64(hdia))
        self.Hdia = np.zeros((N, d, d)); self.V = np.zeros((N, d))
        check(lib.bb_backward_HV(self.ctx.h, N, d, ptr(self.tt), C.byref(aux.c), ptr(v), ptr(he), ptr(self.Hdia),
                                 ptr(self.V)))
        self._make_guide(self.Hdia, self.V)


class PartialBridge(_Proposal):
    """PartialBridge(tt, P, Pt, L, v[, Σ])  src/partialbridge.jl:42-50; fields Target, Pt, tt, v, L, M, μ."""
    kind = K.GUIDE_LMMU

    def __init__(self, tt, P, Pt, L, v, Σ=None, *, ctx=None):
        self.ctx = ctx or default_context()
        self.tt, self.Target, self.Pt = np.array(tt, dtype=np.float64), P, Pt  # tt = collect(tt_)
        L = np.atleast_2d(f64(L)); m, d = L.shape
        self.m = m
        self.v = np.atleast_1d(f64(v))
        if self.v.size != m:
            raise BridgeError(K.ERR_ASSERT_M, lib.bb_strerror(K.ERR_ASSERT_M).decode())
        Σ = np.zeros((m, m)) if Σ is None else np.atleast_2d(f64(Σ))
        aux = _AuxC(Pt, self.tt); N = len(self.tt)
        self.L = np.zeros((N, m, d)); self.M = np.zeros((N, m, m)); self.μ = np.zeros((N, m))
        check(lib.bb_backward_LMmu(self.ctx.h, N, d, m, ptr(self.tt), C.byref(aux.c), ptr(L), ptr(Σ), ptr(self.L),
                                   ptr(self.M), ptr(self.μ)))
        self._make_guide(self.L, self.μ, self.M, self.v)


def gpupdate(*args, ctx=None):
    """Bridge.gpupdate(H♢, V, L, Σ, v) / gpupdate(P°, L, Σ, v) -> (H♢', V')  src/guip.jl:221-243."""
    ctx = ctx or default_context()
    if len(args) == 4:
        Po, L, Σ, v = args
        Hd, V = Po.Hdia[0], Po.V[0]
    else:
        Hd, V, L, Σ, v = args
    L = np.atleast_2d(f64(L)); m, d = L.shape
    Hd = np.atleast_2d(f64(Hd)).copy(); V = np.atleast_1d(f64(V)).copy()
    Σ = np.atleast_2d(f64(Σ)); v = np.atleast_1d(f64(v))
    check(lib.bb_gpupdate_HV(ctx.h, d, m, ptr(Hd), ptr(V), ptr(L), ptr(Σ), ptr(v)))
    return Hd, V


def gpupdate_νH(ν, Hp, L, Σ, v, ctx=None):
    """Observation update of (ν, H⁺) between segments (partialbridge_bolus3.jl:128-137)."""
    ctx = ctx or default_context()
    L = np.atleast_2d(f64(L)); m, d = L.shape
    ν = np.atleast_1d(f64(ν)).copy(); Hp = np.atleast_2d(f64(Hp)).copy()
    Σ = np.atleast_2d(f64(Σ)); v = np.atleast_1d(f64(v))
    check(lib.bb_gpupdate_nuH(ctx.h, d, m, ptr(ν), ptr(Hp), ptr(L), ptr(Σ), ptr(v)))
    return ν, Hp


def lptilde(a, b, ctx=None) -> float:
    """lptilde(P::GuidedBridge, u)   src/guip.jl:206   (proposal first)
    lptilde(x, P::PartialBridgeνH)  src/partialbridgenuH.jl:169 (point first), in the form the reference tests:
    -0.5 (x'H[1]x - 2x'H[1]ν[1]) - C  (test/partialbridgenuH.jl:124)."""
    Po, x = (a, b) if _is_proposal(a) else (b, a)
    ctx = ctx or Po.ctx
    x = np.atleast_1d(f64(x)); d = Po.Target.d
    out = C.c_double(0)
    if isinstance(Po, PartialBridgeνH):
        check(lib.bb_lptilde_nuH(ctx.h, d, ptr(f64(Po.ν[0])), ptr(f64(Po.H[0])), float(Po.C), ptr(x), C.byref(out)))
    elif isinstance(Po, GuidedBridge):
        tt = Po.tt
        if getattr(Po.Pt, "is_const", False):
            tr, const = f64([np.trace(np.atleast_2d(Po.Pt.B(tt[0])))]), 1
        else:  # tr B~ at the forward Ralston stage times of every interval (src/ode.jl:44-49,178-184)
            tr = np.empty((len(tt) - 1, 3)); const = 0
            for i in range(len(tt) - 1):
                h = tt[i + 1] - tt[i]
                for k, c in enumerate((0.0, 0.5, 0.75)):
                    tr[i, k] = np.trace(np.atleast_2d(Po.Pt.B(tt[i] + c * h)))
        check(lib.bb_lptilde_HV(ctx.h, len(tt), d, ptr(tt), ptr(tr), const, ptr(f64(Po.V[0])), ptr(f64(Po.Hdia[0])),
                                ptr(x), C.byref(out)))
    else:
        raise BridgeError(K.ERR_UNSUPPORTED, "lptilde: GuidedBridge or PartialBridgeνH")
    return out.value


# ----------------------------------------------------------------------------------------------- reference calls
class _Rng:
    seed = 0
    stream = 0


def seed_(s: int):
    """Random.seed!(s): seeds the Philox streams used by sample / sample!."""
    _Rng.seed, _Rng.stream = int(s), 0


def sample_(W: SamplePath, P: Wiener, y1=None, ctx=None) -> SamplePath:
    """sample!(W, Wiener{T}(), y1 = W.yy[1])  src/wiener.jl:50-58."""
    ctx = ctx or default_context()
    N, dp = len(W), W.dim
    e = _small_ens(ctx, 1, N, dp, dp)
    w = W._as2d().copy()
    if y1 is not None:
        w[0] = y1
    e.set_grid(0, W.tt)
    e.upload(K.W, w)
    e.sample_(_Rng.seed, _Rng.stream)
    _Rng.stream += 1
    out = e.download(K.W)
    W.yy[...] = out.reshape(W.yy.shape)
    return W


def sample(tt, P: Wiener, y1=None, ctx=None) -> SamplePath:
    """sample(tt, Wiener{T}()[, y1])  src/wiener.jl:11-21."""
    d = P.d
    tt = np.array(tt, dtype=np.float64)  # tt = collect(tt): a fresh vector  (src/wiener.jl:12)
    yy = np.zeros(len(tt)) if d == 1 else np.zeros((len(tt), d))
    return sample_(SamplePath(tt, yy), P, y1, ctx)


def _is_proposal(P):
    return isinstance(P, _Proposal)


def solve_(method: SDESolver, Y: SamplePath, u, W: SamplePath, P, ctx=None):
    """solve!(::EulerMaruyama, Y, u, W, P) -> Y                     src/euler.jl:135-152
    solve!(::Euler, Y, u, W, P°) -> Y.yy[N] (the end point)       src/euler.jl:247-268"""
    ctx = ctx or default_context()
    guided = _is_proposal(P)
    scheme = getattr(method, "scheme", K.SCHEME_EULER)
    if guided and scheme not in (K.SCHEME_EULER, K.SCHEME_STRATONOVICH, K.SCHEME_MDB):
        # guided solve! exists for Euler, StratonovichEuler (src/euler.jl:246-306; the latter equals the former for the
        # registry's constant σ) and Mdb (src/euler.jl:308-327)
        raise BridgeError(K.ERR_UNSUPPORTED, lib.bb_strerror(K.ERR_UNSUPPORTED).decode())
    target = P.Target if guided else P
    if guided and W.tt is P.tt:
        raise BridgeError(K.ERR_TIMEAXIS, lib.bb_strerror(K.ERR_TIMEAXIS).decode())  # src/euler.jl:248
    N = len(W)
    if len(Y) != N or (guided and N != len(P.tt)):
        raise BridgeError(K.ERR_LENGTH, lib.bb_strerror(K.ERR_LENGTH).decode())  # src/euler.jl:137,251
    u = np.atleast_1d(f64(u))
    if getattr(method, "inplace", False):  # solve!(::EulerMaruyama!, Y::VSamplePath, u, W, P)  src/sde!.jl:21-53
        if guided or not isinstance(Y, VSamplePath):
            raise BridgeError(K.ERR_UNSUPPORTED, "EulerMaruyama!: solve!(EulerMaruyama!(), Y::VSamplePath, u, W, P)")
        if Y._as2d().shape != (N, u.size) or u.size != target.d:  # size(Y.yy) != (length(y), N)  src/sde!.jl:30
            raise BridgeError(K.ERR_STARTPOINT, lib.bb_strerror(K.ERR_STARTPOINT).decode())
    if u.size != target.d:
        raise BridgeError(K.ERR_STARTPOINT, lib.bb_strerror(K.ERR_STARTPOINT).decode())
    e = _small_ens(ctx, 1, N, target.d, target.dprime)
    e.set_start(u)
    e.upload(K.W, W._as2d())
    if guided and scheme == K.SCHEME_MDB:
        e.guided_mdb_(target, [P])
        Y.tt[...] = P.tt  # tt[:] = P.tt  src/euler.jl:318
    elif guided:
        e.guided_euler_ll_(target, [P], store_x=True, ll=False)
        Y.tt[...] = P.tt  # tt[:] = P.tt  src/euler.jl:256
    else:
        e.set_grid(0, W.tt)
        if scheme == K.SCHEME_EULER:
            e.euler_(target)
        else:
            if scheme == K.SCHEME_HEUN:  # yy[N] is not written by this scheme (src/euler.jl:188-196)
                e.upload(K.X, Y._as2d())
            e.solve_scheme_(target, scheme)
        Y.tt[...] = W.tt
    X = e.download(K.X)
    Y.yy[...] = X.reshape(Y.yy.shape)
    if guided and scheme != K.SCHEME_MDB:  # the guided Euler method returns the end point, Mdb returns Y (src/euler.jl:326)
        return Y.yy[-1].copy()
    return Y


def solve(method: SDESolver, u, W: SamplePath, P, ctx=None) -> SamplePath:
    """solve(::SDESolver, u, W, P) -> SamplePath  src/euler.jl:117-118, :246."""
    target = P.Target if _is_proposal(P) else P
    d = target.d
    yy = np.zeros(len(W)) if (d == 1 and np.ndim(u) == 0) else np.zeros((len(W), d))
    Y = SamplePath(W.tt.copy(), yy)
    solve_(method, Y, u, W, P, ctx)
    return Y


def bridge_(*args, ctx=None):
    """bridge!(Y, W, P) (src/deprecated.jl:16-17) and the older 4-argument bridge!(X, x0, W, P°)
    (project/partialbridge.jl:63) = solve!(Euler(), X, x0, W, P°)."""
    if len(args) == 4:
        X, x0, W, Po = args
        return solve_(Euler(), X, x0, W, Po, ctx)
    Y, W, Po = args
    return solve_(Euler(), Y, Y.yy[0], W, Po, ctx)


def llikelihood(rule: LeftRule, X: SamplePath, Po, skip: int = 0, ctx=None) -> float:
    """llikelihood(::LeftRule, X, P°; skip = 0) -> Float64."""
    ctx = ctx or default_context()
    target = Po.Target
    N = len(X)
    if N != len(Po.tt):
        raise BridgeError(K.ERR_LENGTH, lib.bb_strerror(K.ERR_LENGTH).decode())
    e = _small_ens(ctx, 1, N, target.d, target.dprime)
    e.upload(K.X, X._as2d())
    e.llikelihood_(target, [Po], skip)
    return float(e.ll[0])


def innovations_(method: SDESolver, W: SamplePath, Y: SamplePath, P, ctx=None) -> SamplePath:
    """innovations!(::EulerMaruyama, W, Y, P) -> W  src/euler.jl:358-376."""
    ctx = ctx or default_context()
    guided = _is_proposal(P)
    target = P.Target if guided else P
    N = len(Y)
    if len(W) != N:
        raise BridgeError(K.ERR_LENGTH, lib.bb_strerror(K.ERR_LENGTH).decode())
    e = _small_ens(ctx, 1, N, target.d, target.dprime)
    e.upload(K.X, Y._as2d())
    if not guided:
        e.set_grid(0, Y.tt)
    e.innovations_(target, [P] if guided else None)
    W.tt[...] = Y.tt
    W.yy[...] = e.download(K.W).reshape(W.yy.shape)
    return W


# ---- host-side readers of the online statistics (src/mclog.jl:58-111); inputs are what mc_stats / chain_mc_stats return
MCBAND_Q = 1.9599639845400538  # sqrt(2.) * erfinv(0.95)  (BB_MCBAND_Q)


def mcbandmean(mean, cov, k: int):
    """mcbandmean(mc)  src/mclog.jl:63-73: band for the chain MEAN, m -/+ Q ste with ste = sqrt(diag(cov)) sqrt(1/k)."""
    mean = np.asarray(mean, dtype=np.float64)
    ste = np.sqrt(np.einsum("...ii->...i", np.asarray(cov, dtype=np.float64))) * np.sqrt(1.0 / k)
    return mean - MCBAND_Q * ste, mean + MCBAND_Q * ste


def mcmarginalstats(mean, cov):
    """mcmarginalstats(mcstates)  src/mclog.jl:100-111 for the segments of one chain (mean [S,N,d], cov [S,N,d,d]):
    (Xmean, Xstd) along the concatenated time axis [S (N-1) + 1, d]; the junction point of two segments is taken
    from the right-hand segment (`pop!` before `append!`)."""
    mean = np.asarray(mean, dtype=np.float64)
    std = np.sqrt(np.einsum("...ii->...i", np.asarray(cov, dtype=np.float64)))
    S = mean.shape[0]
    Xmean = [mean[i, :-1] for i in range(S - 1)] + [mean[S - 1]]
    Xstd = [std[i, :-1] for i in range(S - 1)] + [std[S - 1]]
    return np.concatenate(Xmean, axis=0), np.concatenate(Xstd, axis=0)


def pcn_(ens: PathEnsemble, P, guides, ρ: float, iterations: int, seed: int, first_iter: int = 0,
         skip: int = 0, store_x: bool = True, callback: Optional[Callable] = None) -> int:
    """The sampler loop of test/partialbridgenuH.jl:176-195 for all chains of `ens`: `iterations`
    pCN updates; returns the number of accepted proposals (summed over chains)."""
    for it in range(first_iter, first_iter + iterations):
        ens.pcn_step_(P, guides, ρ, seed, it, skip, store_x)
        if callback is not None:
            callback(it, ens)
    return ens.acc


def theta_mcmc_(ens: PathEnsemble, ρ: float, rw_sd, iterations: int, seed: int, first_iter: int = 0, skip: int = 0,
                store_x: bool = True, param_prob: float = 0.5, callback: Optional[Callable] = None,
                blocked: bool = False, hzero: float = 0.1):
    """The outer loop of project_partialbridge/partialbridge_bolus3.jl:248-365 for all chains of an ensemble with
    per-chain parameters (theta_attach_): every iteration is, with probability `param_prob` (`updateparams = rand(Bool)`,
    :259), a parameter update with the innovations held fixed, otherwise a pCN update of the paths -- all segments
    together, or with blocked = True a sweep of block updates as the script draws them (theta_blocked_sweep_).
    The coin is common to all chains (one launch per iteration) and comes from the host RNG seeded with `seed`; the two
    step kinds use disjoint Philox counters, so `it` can be shared.  Returns (accepted pCN proposals, accepted parameter
    proposals), summed over chains."""
    rng = np.random.default_rng(seed)
    sub = 0  # noise streams of the blocks of path sweeps (blocked = True): disjoint from the iteration numbers
    for it in range(first_iter, first_iter + iterations):
        if rng.random() < param_prob:
            ens.theta_param_step_(rw_sd, seed, it, skip, store_x)
        elif blocked:  # the script's sweep of block updates klow .. kup (:258-275, :358-359)
            sub += len(ens.theta_blocked_sweep_(rng, ρ, seed, 0x40000000 + sub, hzero, skip))
        else:
            ens.theta_pcn_step_(ρ, seed, it, skip, store_x)
        if callback is not None:
            callback(it, ens)
    return ens.acc, ens.acc_theta
