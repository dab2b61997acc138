#!/usr/bin/env python
"""CPU baseline of the parameter-update step (per-chain parameters) on the FHN config-4 shape: the oracle's OpenMP
driver bbo_theta_param_bench (oracle/bridge_oracle.c) on all host threads.  Test infrastructure; prints path-steps/s
(a parameter step = backward tables + forward guided Euler + ll for 4 x 1000 steps per chain)."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O

O.build()
orc = O.load("fast")
n, S = 1001, 4
obs_t = (0.0, 0.5, 1.0, 1.5, 2.0); obs_v = np.array([-1.0, -0.5, 0.5, 1.1])
grids = np.stack([obs_t[k] + np.linspace(0, obs_t[k + 1] - obs_t[k], n) * (2 - np.linspace(0, obs_t[k + 1] - obs_t[k], n) / (obs_t[k + 1] - obs_t[k])) for k in range(S)])
par = np.array([0.1, 0.0, 1.5, 0.8, 0.3]); x0 = np.array([-0.5, -0.6]); L = np.array([1.0, 0.0])
rw = np.array([0.0, 0.0, 0.01, 0.01, 0.005])
cores = len(os.sched_getaffinity(0))
P = int(sys.argv[1]) if len(sys.argv) > 1 else 16 * cores
iters = 2
secs = C.c_double(0)
orc.lib.bbo_theta_param_bench.restype = C.c_longlong
acc = orc.lib.bbo_theta_param_bench(O._p(par), S, n, O._p(grids), O._p(x0), O._p(L), C.c_double(1e-10), C.c_double(1e-3),
                                    O._p(obs_v), O._p(rw), C.c_longlong(P), C.c_uint64(4), iters, cores, C.byref(secs))
steps = P * iters * S * (n - 1)
print(f"CPU oracle parameter step (-O3 -march=native, OpenMP, {cores} threads): {steps / secs.value:.3e} path-steps/s "
      f"({P} chains x {iters} iterations, {secs.value:.2f} s, {acc} accepted)")
