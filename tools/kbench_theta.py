#!/usr/bin/env python
"""Kernel micro-benchmarks of the per-chain-parameter path (bb_theta.cu) on the FHN config-4 workload.
usage: kbench_theta.py [chains]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bridge_jl_b200 as B
import bridge_jl_b200.configs as cfg

P = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
DIAG = len(sys.argv) > 2 and sys.argv[2] == "diag"  # the d' = 2 model (src/Models.jl:18-19) instead of the hypoelliptic one
ctx = B.default_context()
ctx.set_timing(True)
n, S = 1001, 4
grids = cfg.fhn_segment_grids(n)
Pm = B.FitzHughNagumo(*cfg.FHN_PAR[:4], 0.3, 0.3) if DIAG else B.FitzhughDiffusion(*cfg.FHN_PAR)
ens = B.PathEnsemble(P, S, n, 2, 2 if DIAG else 1)
for s, g in enumerate(grids):
    ens.set_grid(s, g)
ens.set_start(cfg.FHN_X0)
ens.theta_attach_(Pm, cfg.FHN_L, cfg.FHN_SIGMA, cfg.FHN_EPS, cfg.FHN_OBS_V,
                  priors={2: ("gamma", 1.0, 100.0), 3: ("gamma", 1.0, 100.0), 4: ("gamma", 1.0, 100.0)})
th = ens.theta()
rng = np.random.default_rng(1)
th[:, 2] += 0.1 * rng.standard_normal(P); th[:, 3] += 0.05 * rng.standard_normal(P)
ens.set_theta(th)
ens.sample_(4, 0xFFFFFFFE)
ens.theta_guided_euler_ll_()
print("device bytes: %.1f GB" % (ens.nbytes / 1e9), flush=True)
steps = P * S * (n - 1)
RW = [0, 0, 0.01, 0.01, 0.005]
# algorithmic bytes per path-step: tables 48 (d + d*d doubles), W 8 d', X 16
w = 16 if DIAG else 8
calls = {
    "backward (tables of θ)": (lambda it: ens.theta_guides_(), 48),
    "forward guided Euler+ll": (lambda it: (ens.set_theta(th[:1]), ens.theta_guides_(), ens.theta_guided_euler_ll_())[-1], 64 + w),
    "pCN, own tables": (lambda it: ens.theta_pcn_step_(cfg.FHN_RHO, 4, it), 64 + 2 * w),
    "parameter step (backward + forward)": (lambda it: ens.theta_param_step_(RW, 4, 1000 + it), 112 + w),
}
for name, (fn, nbytes) in calls.items():
    ts = []
    for it in range(7):
        fn(100 + it)
        ctx.synchronize()
        ts.append(ctx.last_kernel_ms)
    t = float(np.median(ts[2:]))
    print(f"{name:38s} P={P} ms={t:8.3f} steps/s={steps / t * 1e3:.3e} alg GB/s={steps * nbytes / t * 1e-6:7.0f} "
          f"({nbytes} B/step) frac={steps * nbytes / t * 1e-6 / 6549.8:.3f}", flush=True)
print("acc_theta", ens.acc_theta, "acc", ens.acc)
# blocked segment updates (bb_theta_block_step): per-chain backward over the block + two forward passes + commit of the
# accepted rows.  Per path-step of the block: tables 48 written + 48 read, W 8 d' read, W° 8 d' and X° 16 written, plus the
# commit (read + write of W° and X° for the accepting chains).  Needs per-chain starting points only for blocks with s_lo = 0.
if os.environ.get("KBENCH_BLOCKS", "1") == "1":
    ens.set_start(np.tile(cfg.FHN_X0, (P, 1)))
    ens.theta_guided_euler_ll_()
    for lo, hi in ((0, 4), (1, 3), (2, 4), (0, 1)):
        ts = []
        for it in range(5):
            ens.theta_block_step_(lo, hi, cfg.FHN_RHO, 4, 2000 + it)
            ctx.synchronize()
            ts.append(ctx.last_kernel_ms)
        t = float(np.median(ts[1:]))
        bsteps = P * (hi - lo) * (n - 1)
        rate = float(np.mean(ens.accepted))
        nb = 96 + 2 * w + 16 + rate * 2 * (w + 16)
        print(f"block update, segments [{lo},{hi})         P={P} ms={t:8.3f} steps/s={bsteps / t * 1e3:.3e} "
              f"alg GB/s={bsteps * nb / t * 1e-6:7.0f} ({nb:.0f} B/step at acceptance {rate:.2f}) frac={bsteps * nb / t * 1e-6 / 6549.8:.3f}",
              flush=True)
ens.close()
