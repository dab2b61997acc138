#!/usr/bin/env python
"""BASELINE config 5 on one GPU: Landmarks (d = 16, d' = 8) PartialBridgeνH, N = 1001; guided Euler + ll (mode G) and
one pCN iteration (mode M), for the full 1e4 paths and for the 2500-path share of one of 4 GPUs; the CPU oracle timed
beside it.  usage: kbench_landmarks.py [paths ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bridge_jl_b200 as B

A_K, SIG, LAM = 0.5, 2.0, 0.5
Q0 = np.array([[-1.0, -1.0], [1.0, -1.2], [1.1, 0.9], [-0.8, 1.0]])
P0 = np.array([[0.5, 0.1], [-0.2, 0.4], [-0.3, -0.3], [0.2, -0.5]])
QT = np.array([[-0.6, -1.4], [1.5, -0.9], [0.8, 1.4], [-1.2, 0.7]])
x0 = np.concatenate([np.concatenate([Q0[i], P0[i]]) for i in range(4)])
L = np.zeros((8, 16))
for i in range(4):
    for c in range(2):
        L[2 * i + c, 4 * i + c] = 1.0
n, T = 1001, 1.0
s = np.linspace(0.0, T, n); tt = s * (2 - s / T)
ctx = B.default_context(); ctx.set_timing(True)
Pm = B.Landmarks(A_K, SIG, LAM); Pt = B.LandmarksTilde(A_K, SIG, LAM, QT)
t0 = time.time()
Po = B.PartialBridgeνH(tt, Pm, Pt, L, QT.ravel(), 1e-3, 1e-4 * np.eye(8))
print(f"constructor (updateνH⁺C + backward R3, d = 16, N = {n}): {1e3 * (time.time() - t0):.1f} ms incl. transfers", flush=True)
PEAK = 6549.8
for P in [int(a) for a in sys.argv[1:]] or [10000, 2500]:
    ens = B.PathEnsemble(P, 1, n, 16, 8)
    ens.set_grid(0, tt); ens.set_start(x0); ens.sample_(5, 0xFFFFFFF0)
    ens.guided_euler_ll_(Pm, [Po])
    steps = P * (n - 1)
    for name, fn, nb in (("guided Euler + ll, X stored (mode G)", lambda it: ens.guided_euler_ll_(Pm, [Po]), 8 * (8 + 16)),
                         ("pCN iteration, X° stored (mode M)", lambda it: ens.pcn_step_(Pm, [Po], 0.9, 5, it), 16 * 8 + 8 * 16),
                         ("pCN iteration, X° not stored", lambda it: ens.pcn_step_(Pm, [Po], 0.9, 5, it, store_x=False), 16 * 8)):
        ts = []
        for it in range(8):
            fn(100 + it); ctx.synchronize(); ts.append(ctx.last_kernel_ms)
        t = float(np.median(ts[3:]))
        print(f"P={P:6d} {name:40s} ms={t:7.3f} path-steps/s={steps / t * 1e3:.3e} alg GB/s={steps * nb / t * 1e-6:6.0f} "
              f"({nb} B/step) frac={steps * nb / t * 1e-6 / PEAK:.3f}", flush=True)
    print("   acc rate", ens.acc / (16 * P), " miss", float(np.max(np.abs(ens.xend.reshape(P, 4, 2, 2)[:, :, 0] - QT))))
    ens.close()
# CPU oracle (test infrastructure) beside it: same workload, all host threads
from oracle import oracle as O
orc = O.load("fast")
om = O.make_model(O.LANDMARKS, 16, 8, [A_K, SIG, LAM])
og = O.GuideHolder(O.GUIDE_NUH, tt, Po.H, Po.ν, Bt=Pt.B(0.0), betat=Pt.β(0.0))
cores = len(os.sched_getaffinity(0))
pc = 8 * cores
acc, secs, _ = orc.pcn_bench(om, [og], pc, x0, 0.9, 5, 2, nthreads=cores)
print(f"CPU oracle (-O3 -march=native, OpenMP, {cores} threads): {pc * 2 * (n - 1) / secs:.3e} path-steps/s "
      f"({pc} chains x 2 pCN iterations, {secs:.2f} s)")
