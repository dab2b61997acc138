#!/usr/bin/env python
"""Small pCN / guided / Euler run for compute-sanitizer (memcheck, racecheck, synccheck) under gpurun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bridge_jl_b200 as B
import bridge_jl_b200.configs as cfg

n, P = 49, 700   # ragged: N not a multiple of 16, P not a multiple of 256
# SAN_PCN_KERNEL=1: one thread per chain everywhere (small ensembles otherwise pick the warp-specialised kernel, whose
# warp pairs are synchronised by mbarriers only -- racecheck does not model those and reports every hand-over)
ONE_THREAD = os.environ.get("SAN_PCN_KERNEL") == "1"
if ONE_THREAD:
    B.default_context().set_pcn_kernel(1)
Pm, guides, x0, rho = cfg.fhn_config4(n, obs_t=(0.5, 1.0, 1.5), obs_v=(-1.0, -0.5, 0.5))
ens = B.PathEnsemble(P, len(guides), n, 2, 1)
for s, g in enumerate(guides):
    ens.set_grid(s, g.tt)
ens.set_start(x0)
ens.sample_(1, 0)
ens.guided_euler_ll_(Pm, guides)
for it in range(3):
    ens.pcn_step_(Pm, guides, rho, 1, it)
ens.pcn_step_(Pm, guides, rho, 1, 3, store_x=False)
X = ens.download(B.X)
ens.llikelihood_(Pm, guides)
ens.euler_(Pm)
L3 = B.Lorenz([10.0, 28.0, 8 / 3], 3.0)
e3 = B.PathEnsemble(300, 2, 33, 3, 3, double_buffer=False)
for s in range(2):
    e3.set_grid(s, np.linspace(s, s + 1, 33))
e3.set_start([1.0, 0.0, 0.0]); e3.sample_euler_(L3, 3, 0); e3.innovations_(L3)
print("ok", float(np.sum(X)) != 0.0, ens.acc)
# per-chain parameters: backward tables, forward, pCN, parameter step, X refresh (ragged: P = 333, N = 49)
et = B.PathEnsemble(333, 3, n, 2, 1)
for s, g in enumerate(guides):
    et.set_grid(s, g.tt)
et.set_start(x0)
et.theta_attach_(Pm, cfg.FHN_L, cfg.FHN_SIGMA, cfg.FHN_EPS, (-1.0, -0.5, 0.5), priors={4: ("gamma", 2.0, 50.0)})
et.sample_(1, 0)
et.theta_guided_euler_ll_()
et.theta_param_step_([0, 0, 0.02, 0.02, 0.01], 1, 10)
et.theta_pcn_step_(rho, 1, 11)
Xt = et.download(B.X)
print("theta ok", et.acc_theta, et.acc, float(np.sum(Xt)) != 0.0)
# wide path: Landmarks d = 16, d' = 8 (constructors at d = 16, sample, guided Euler + ll, pCN with and without X; P = 70, N = 21)
A_K, SIG, LAM = 0.5, 2.0, 0.5
QT = np.array([[-0.6, -1.4], [1.5, -0.9], [0.8, 1.4], [-1.2, 0.7]])
xl = np.array([-1.0, -1.0, 0.5, 0.1, 1.0, -1.2, -0.2, 0.4, 1.1, 0.9, -0.3, -0.3, -0.8, 1.0, 0.2, -0.5])
Ll = np.zeros((8, 16))
for i in range(4):
    for c in range(2):
        Ll[2 * i + c, 4 * i + c] = 1.0
ttl = np.linspace(0.0, 1.0, 21)
Pl = B.Landmarks(A_K, SIG, LAM); Ptl = B.LandmarksTilde(A_K, SIG, LAM, QT)
Pol = B.PartialBridgeνH(ttl, Pl, Ptl, Ll, QT.ravel(), 1e-3, 1e-4 * np.eye(8))
el = B.PathEnsemble(70, 1, 21, 16, 8)
el.set_grid(0, ttl); el.set_start(xl); el.sample_(1, 0); el.guided_euler_ll_(Pl, [Pol])
el.pcn_step_(Pl, [Pol], 0.9, 1, 0); el.pcn_step_(Pl, [Pol], 0.9, 1, 1, store_x=False)
Xl = el.download(B.X)
print("landmarks ok", el.acc, bool(np.all(np.isfinite(Xl))))
# the bolus model of partialbridge_bolus3.jl on the per-chain path: time-dependent auxiliary drift, joint start-point move
Pb = B.BolusDiffusion(70.0 / 0.6, 8.0, 1.25, 1.5, 0.5, 0.2)
gb = [np.linspace(0.0, 0.8, 27), np.linspace(0.8, 1.7, 27)]
eb = B.PathEnsemble(101, 2, 27, 2, 2)
for s, g in enumerate(gb):
    eb.set_grid(s, g)
eb.set_start(np.tile([0.5, 0.2], (101, 1)))
eb.theta_attach_(Pb, [[0.5, 0.5]], 1e-2 * np.eye(1), 0.1, (4.0, 9.0), aux_kind=3, priors={1: ("gamma", 1.0, 100.0)},
                 start_sd=0.1, start_dir=[1.0, -1.0])
eb.sample_(2, 0); eb.theta_guided_euler_ll_()
B.theta_mcmc_(eb, 0.5, [0, 0.02, 0, 0, 0.02], 4, 3)
print("bolus ok", eb.acc, eb.acc_theta, bool(np.all(np.isfinite(eb.download(B.X)))))
# ---- round 2 kernels
# warp-specialised pCN (noise warps -> dynamics warps through an mbarrier-guarded ring), ragged sizes
ctx = B.default_context()
if not ONE_THREAD:
    ctx.set_pcn_kernel(2)  # BB_PCN_WARP_SPECIALISED
    for it in range(2):
        ens.pcn_step_(Pm, guides, rho, 1, 20 + it)
    ctx.set_pcn_kernel(0)
# the whole backward chain in one launch, lptilde, Mdb on a guided proposal
Pm2, chain, x02, _ = cfg.fhn_config4_chain(n, ctx=ctx)
cfg.fhn_config4_chain(n, ctx=ctx, chain=chain)
e4 = B.PathEnsemble(130, len(chain.segments), n, 2, 1)
for s, g in enumerate(chain.segments):
    e4.set_grid(s, g.tt)
e4.set_start(x02); e4.sample_(5, 0); e4.guided_euler_ll_(Pm2, chain.segments); e4.pcn_step_(Pm2, chain.segments, rho, 5, 1)
print("chain ok", e4.acc, float(B.lptilde(np.array(x02), guides[0])))
# blocked segment updates (per-chain right-end conditioning, two passes, commit of the accepted rows) and the sweep
rngb = np.random.default_rng(0)
eb.theta_block_step_(0, 1, 0.7, 3, 100); eb.theta_block_step_(1, 2, 0.7, 3, 101); eb.theta_block_step_(0, 2, 0.7, 3, 102)
eb.theta_blocked_sweep_(rngb, 0.7, 3, 110)
B.theta_mcmc_(eb, 0.5, [0, 0.02, 0, 0, 0.02], 3, 3, first_iter=50, blocked=True)
print("blocks ok", eb.acc, bool(np.all(np.isfinite(eb.download(B.X)))))
# tensor-core Landmarks kernel is the default wide path (above); host-buffer step with skipped rejected rows (pinned, mapped)
import torch
hW = torch.empty((P, len(guides), n, 1), dtype=torch.float64, pin_memory=True).numpy()
hX = torch.empty((P, len(guides), n, 2), dtype=torch.float64, pin_memory=True).numpy()
hl = torch.empty(P, dtype=torch.float64, pin_memory=True).numpy(); ha = torch.empty(P, dtype=torch.uint8, pin_memory=True).numpy()
hW[...] = ens.download(B.W)
ens.pcn_step_host_(Pm, guides, rho, 1, 30, hW, hW, hX, hl, ha, skip_rejected=True)
print("host ok", int(ha.sum()), bool(np.array_equal(hW, ens.download(B.W))))
# a user-defined model compiled at run time (NVRTC) through the same path kernel
um = B.UserProcess(2, 1, "double u = x[0] - x[1]; u = fma(-(x[0]*x[0]), x[0], u); o[0] = (u + par[1]) * (1.0/par[0]);"
                         "o[1] = fma(par[2], x[0], -x[1]) + par[3];", [-1, 0], [None, "par[4]"], cfg.FHN_PAR)
ens.guided_euler_ll_(um, guides); ens.pcn_step_(um, guides, rho, 1, 40)
print("user ok", ens.acc)
