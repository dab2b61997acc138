"""Derives the float32 polynomial coefficients of the Box-Muller transform used by BOTH
oracle/bridge_oracle.c and bridge.jl_b200/csrc/rng.cuh (they are pasted there as hex floats).

Only +, *, fma and IEEE sqrt are used by the transform, so a CPU (fmaf) and a GPU (FFMA)
evaluation are bit-identical.  Run:  python tools/gen_rng_poly.py
"""
import numpy as np

f32 = np.float32


def cheb_fit(fun, a, b, deg, n=4000, powers=None):
    """Least-squares fit on Chebyshev nodes (near-minimax), returns monomial coefficients."""
    k = np.arange(n)
    x = 0.5 * (a + b) + 0.5 * (b - a) * np.cos(np.pi * (k + 0.5) / n)
    A = np.stack([x ** p for p in powers], axis=1)
    y = fun(x)
    c, *_ = np.linalg.lstsq(A, y, rcond=None)
    return c


def hexf(v):
    return float(f32(v)).hex()


def main():
    # ---- log(1+f) = f - f^2/2 + f^3 * P(f),  f in [sqrt(.5)-1, sqrt(2)-1]
    lo, hi = np.sqrt(0.5) - 1, np.sqrt(2) - 1

    def g(f):
        return (np.log1p(f) - f + 0.5 * f * f) / f ** 3

    degP = 7
    cP = cheb_fit(g, lo, hi, degP, powers=range(degP + 1))
    print("log P coefficients (low -> high):")
    print(", ".join(hexf(c) + "f" for c in cP))
    # ---- sin(pi r) = r * S(r^2), cos(pi r) = C(r^2), r in [-1/4, 1/4]
    def gs(s):
        r = np.sqrt(s)
        return np.sin(np.pi * r) / r

    def gc(s):
        return np.cos(np.pi * np.sqrt(s))

    cS = cheb_fit(gs, 1e-12, 1 / 16, 3, powers=range(4))
    cC = cheb_fit(gc, 0.0, 1 / 16, 4, powers=range(5))
    print("sinpi S coefficients:", ", ".join(hexf(c) + "f" for c in cS))
    print("cospi C coefficients:", ", ".join(hexf(c) + "f" for c in cC))

    # ---- accuracy of the float32 evaluation
    def logf_(u):
        u = u.astype(f32)
        ix = u.view(np.uint32).astype(np.int64)
        ix = ix + (0x3F800000 - 0x3F3504F3)
        e = (ix >> 23) - 127
        m = ((ix & 0x007FFFFF) + 0x3F3504F3).astype(np.uint32).view(f32)
        f = (m - f32(1)).astype(f32)
        p = np.full_like(f, f32(cP[-1]))
        for c in cP[-2::-1]:
            p = (p.astype(np.float64) * f + f32(c)).astype(f32)  # fma: one rounding
        f2 = (f * f).astype(f32)
        t = (p.astype(np.float64) * f).astype(f32)
        t = (t.astype(np.float64) * f2 - 0.5 * f2.astype(np.float64)).astype(f32)  # fma(t, f2, -0.5 f2)
        r = (t + f).astype(f32)
        return (e.astype(np.float64) * float(f32(0.6931471805599453)) + r).astype(f32)

    rng = np.random.default_rng(0)
    u = np.concatenate([rng.random(2_000_000), 10.0 ** -rng.uniform(0, 9.9, 500_000)])
    u = u[(u > 0) & (u <= 1)]
    got = logf_(u).astype(np.float64)
    want = np.log(u.astype(f32).astype(np.float64))
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
    big = np.abs(want) > 1e-3
    print("logf max rel err (|log|>1e-3):", rel[big].max(), " max abs err:", np.abs(got - want).max())

    r = rng.uniform(-0.25, 0.25, 2_000_000).astype(f32)
    s = (r * r).astype(f32)

    def horner(cs, s):
        p = np.full_like(s, f32(cs[-1]))
        for c in cs[-2::-1]:
            p = (p.astype(np.float64) * s + f32(c)).astype(f32)
        return p

    sn = (horner(cS, s).astype(np.float64) * r).astype(f32)
    cs = horner(cC, s)
    print("sinpi max abs err:", np.abs(sn - np.sin(np.pi * r.astype(np.float64))).max(),
          " cospi max abs err:", np.abs(cs - np.cos(np.pi * r.astype(np.float64))).max())


if __name__ == "__main__":
    main()
