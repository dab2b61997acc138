#!/usr/bin/env python
"""Sustained rate of the config-4 pCN kernel: 60 launches back to back, three repetitions (the regime of bench.py's timed
region: the kernel runs into the board's power cap within ~100 ms), plus the same launches separated by idle pauses."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bridge_jl_b200 as B
import bridge_jl_b200.configs as cfg

P = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
ctx = B.default_context()
n = 1001
Pm, guides, x0, rho = cfg.fhn_config4(n)
ens = B.PathEnsemble(P, 4, n, 2, 1)
for s, g in enumerate(guides):
    ens.set_grid(s, g.tt)
ens.set_start(x0); ens.sample_(4, 0xFFFFFFFE); ens.guided_euler_ll_(Pm, guides)
it = 0
for _ in range(5):
    ens.pcn_step_(Pm, guides, rho, 4, it); it += 1
ctx.synchronize()
res = []
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(60):
        ens.pcn_step_(Pm, guides, rho, 4, it); it += 1
    ctx.synchronize()
    res.append((time.perf_counter() - t0) / 60 * 1e3)
time.sleep(3.0)
ctx.set_timing(True)
alone = []
for _ in range(3):
    ens.pcn_step_(Pm, guides, rho, 4, it); it += 1
    ctx.synchronize(); alone.append(ctx.last_kernel_ms); time.sleep(0.5)
print(f"{os.path.basename(os.environ.get('BB_LIB', 'libbridge_b200.so')):24s} P={P} sustained (60 back to back) "
      + " ".join(f"{x:.3f}" for x in res) + f" ms; alone after a pause {min(alone):.3f} ms", flush=True)
