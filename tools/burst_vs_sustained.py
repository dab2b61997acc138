#!/usr/bin/env python
"""pCN iteration of BASELINE config 4 timed two ways on one GPU: every launch followed by a host synchronisation (as
tools/kbench.py does) and K launches back to back (as bench.py's timed region does)."""
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
for rep in range(3):
    ctx.set_timing(True)
    ts = []
    for _ in range(10):
        ens.pcn_step_(Pm, guides, rho, 4, it); it += 1
        ctx.synchronize(); ts.append(ctx.last_kernel_ms)
    ctx.set_timing(False)
    t0 = time.perf_counter()
    K = 40
    for _ in range(K):
        ens.pcn_step_(Pm, guides, rho, 4, it); it += 1
    ctx.synchronize()
    t1 = time.perf_counter()
    gap = []
    for _ in range(10):
        t2 = time.perf_counter()
        ens.pcn_step_(Pm, guides, rho, 4, it); it += 1
        ctx.synchronize()
        gap.append((time.perf_counter() - t2) * 1e3)
    print(f"P={P}: synchronised after every launch {np.median(ts):.3f} ms (library events), wall {np.median(gap):.3f} ms; "
          f"{K} launches back to back {(t1 - t0) / K * 1e3:.3f} ms each", flush=True)
# where does the host spend its time in a back-to-back loop?  (a launch that returns in microseconds lets the host run ahead)
d = []
for _ in range(12):
    t2 = time.perf_counter()
    ens.pcn_step_(Pm, guides, rho, 4, it); it += 1
    d.append((time.perf_counter() - t2) * 1e3)
ctx.synchronize()
print("host time of consecutive pcn_step_ calls (ms):", " ".join(f"{x:.3f}" for x in d), flush=True)
# does an idle pause bring the burst figure back?  (yes -> a sustained-load effect of the memory system, not of the chains' state)
for pause in (0.0, 2.0, 0.0, 5.0):
    time.sleep(pause)
    ctx.set_timing(True)
    ts = []
    for _ in range(6):
        ens.pcn_step_(Pm, guides, rho, 4, it); it += 1
        ctx.synchronize(); ts.append(ctx.last_kernel_ms)
    ctx.set_timing(False)
    t0 = time.perf_counter()
    for _ in range(40):
        ens.pcn_step_(Pm, guides, rho, 4, it); it += 1
    ctx.synchronize()
    print(f"after a pause of {pause:.0f} s: first launches {ts[0]:.3f} {ts[1]:.3f} {ts[2]:.3f} ... median {np.median(ts):.3f} ms; "
          f"then 40 back to back {(time.perf_counter() - t0) / 40 * 1e3:.3f} ms each", flush=True)
