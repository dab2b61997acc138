#!/usr/bin/env python
"""BASELINE.json configs 2, 3 and the FHN_DIAG variant of config 4 at full size on one GPU (kernel times by the
library's CUDA events; median of 5 after 3 warm-ups)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bridge_jl_b200 as B
import bridge_jl_b200.configs as cfg

ctx = B.default_context()
ctx.set_timing(True)
PEAK = 6549.8


def timeit(name, fn, steps, nbytes):
    ts = []
    for it in range(8):
        fn(100 + it)
        ctx.synchronize()
        ts.append(ctx.last_kernel_ms)
    t = float(np.median(ts[3:]))
    print(f"{name:58s} ms={t:8.3f} path-steps/s={steps / t * 1e3:.3e} alg GB/s={steps * nbytes / t * 1e-6:7.0f} "
          f"({nbytes} B/step) frac={steps * nbytes / t * 1e-6 / PEAK:.3f}", flush=True)


n = 1001
# ---- config 2: Wiener + EulerMaruyama ensemble, 1e6 paths (X = u + W), and OU
P = 1_000_000
ens = B.PathEnsemble(P, 1, n, 1, 1, double_buffer=False)
ens.set_grid(0, np.linspace(0.0, 1.0, n)); ens.set_start([0.0])
steps = P * (n - 1)
timeit("config 2  sample!(W, Wiener)                  (mode: W written)", lambda it: ens.sample_(2, it), steps, 8)
timeit("config 2  solve!(EulerMaruyama, Wiener process) (mode A)", lambda it: ens.euler_(B.Wiener(1)), steps, 16)
timeit("config 2  solve!(EulerMaruyama, OU)             (mode A)", lambda it: ens.euler_(B.OrnsteinUhlenbeck(2.0, 1.0)), steps, 16)
timeit("config 2  fused sample! + solve! (OU)           (mode B)", lambda it: ens.sample_euler_(B.OrnsteinUhlenbeck(2.0, 1.0), 2, it), steps, 16)
ens.close()
# ---- config 3: LinPro d = 3 GuidedBridge, 1e5 paths
P = 100_000
Pm, guide, u = cfg.linpro_config3(n)
ens = B.PathEnsemble(P, 1, n, 3, 3)
ens.set_grid(0, guide.tt); ens.set_start(u); ens.sample_(3, 0)
steps = P * (n - 1)
timeit("config 3  guided Euler + ll, X stored           (mode G)", lambda it: ens.guided_euler_ll_(Pm, [guide]), steps, 48)
timeit("config 3  guided Euler + ll, X not stored       (mode G)", lambda it: ens.guided_euler_ll_(Pm, [guide], store_x=False), steps, 24)
ens.guided_euler_ll_(Pm, [guide])
timeit("config 3  pCN iteration, X° stored              (mode M)", lambda it: ens.pcn_step_(Pm, [guide], 0.9, 3, it), steps, 72)
ens.close()
# ---- config 4, d' = 2 variant: FHN_DIAG (src/Models.jl:18-19), sigma1 = sigma2 = 0.3, 2.5e5 chains x 4 segments
P = 250_000
Pd = B.FitzHughNagumo(*cfg.FHN_PAR[:4], 0.3, 0.3)
grids = cfg.fhn_segment_grids(n)
ν = np.zeros(2); Hp = np.eye(2) / cfg.FHN_EPS
ν, Hp = B.gpupdate_νH(ν, Hp, cfg.FHN_L, cfg.FHN_SIGMA, [cfg.FHN_OBS_V[-1]])
guides = [None] * 4
for i in range(3, -1, -1):
    Bt, bt, _ = cfg.fhn_matching_aux(cfg.FHN_OBS_V[i])
    Pt = B.LinearAux(Bt, bt, np.diag([0.09, 0.09]))
    guides[i], ν, Hp, _ = B.partialbridgeνH(grids[i], Pd, Pt, ν, Hp)
    if i > 0:
        ν, Hp = B.gpupdate_νH(ν, Hp, cfg.FHN_L, cfg.FHN_SIGMA, [cfg.FHN_OBS_V[i - 1]])
ens = B.PathEnsemble(P, 4, n, 2, 2)
for s, g in enumerate(grids):
    ens.set_grid(s, g)
ens.set_start(cfg.FHN_X0); ens.sample_(4, 0xFFFFFFFE); ens.guided_euler_ll_(Pd, guides)
steps = P * 4 * (n - 1)
timeit("config 4' FHN_DIAG (d' = 2) pCN iteration, X° stored (mode M)", lambda it: ens.pcn_step_(Pd, guides, 0.99, 4, it), steps, 48)
print("acc rate", ens.acc / (8 * P))
ens.close()
