#!/usr/bin/env python
"""Kernel micro-benchmarks (tuning aid, run under gpurun): times individual C-ABI calls of the FHN config-4
workload with the library's own CUDA events.  usage: kbench.py [chains] [modes...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bridge_jl_b200 as B
import bridge_jl_b200.configs as cfg

P = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
modes = sys.argv[2:] or ["pcn", "pcn_nox", "guided", "guided_nox", "sample", "sample_euler"]
ctx = B.default_context()
ctx.set_timing(True)
n = 1001
Pm, guides, x0, rho = cfg.fhn_config4(n)
ens = B.PathEnsemble(P, 4, n, 2, 1)
for s, g in enumerate(guides):
    ens.set_grid(s, g.tt)
ens.set_start(x0)
ens.sample_(4, 0xFFFFFFFE)
ens.guided_euler_ll_(Pm, guides)
steps = P * 4 * (n - 1)
calls = {
    "pcn": (lambda it: ens.pcn_step_(Pm, guides, rho, 4, it), 32),
    "pcn_nox": (lambda it: ens.pcn_step_(Pm, guides, rho, 4, it, store_x=False), 16),
    "guided": (lambda it: ens.guided_euler_ll_(Pm, guides), 24),
    "guided_nox": (lambda it: ens.guided_euler_ll_(Pm, guides, store_x=False), 8),
    "sample": (lambda it: ens.sample_(4, it), 8),
    "sample_euler": (lambda it: ens.sample_euler_(Pm, 4, it), 24),
    "euler": (lambda it: ens.euler_(Pm), 24),
    "llik": (lambda it: ens.llikelihood_(Pm, guides), 16),
    "mc": (lambda it: ens.mc_update_(), 16),                 # pooled moments: X read once
    "chain_mc": (lambda it: ens.chain_mc_update_(), 112),    # per-chain Welford: X read, m and m2 read + written, d = 2
}
for m in modes:
    fn, nbytes = calls[m]
    ts = []
    for it in range(8):
        fn(100 + it)
        ctx.synchronize()
        ts.append(ctx.last_kernel_ms)
    t = float(np.median(ts[3:]))
    print(f"{os.environ.get('BB_LIB','default').split('/')[-1]:24s} {m:13s} P={P} ms={t:8.3f} steps/s={steps / t * 1e3:.3e} "
          f"alg GB/s={steps * nbytes / t * 1e-6:7.0f} ({nbytes} B/step) frac={steps * nbytes / t * 1e-6 / 6546.6:.3f}", flush=True)
ens.close()
