#!/usr/bin/env python
"""The sampler loop of project_partialbridge/partialbridge_fitzhugh.jl:125-189 for many independent chains, written against
the host mirror (`import bridge_jl_b200`): FitzHugh-Nagumo (hypoelliptic) target, one PartialBridgeνH per observation
segment (backward chain right to left), pCN updates of the innovations, and the script's online statistics
`mcstate = [mcnext!(mcstate[i], XX[i].yy) ...]` kept on the device for every chain.

    python examples/fitzhugh_smoothing.py [chains] [grid points per segment] [iterations]

Needs a B200 (the library has no CPU path).  Prints one line per report interval and a summary the GPU test parses."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

import bridge_jl_b200 as B  # noqa: E402
import bridge_jl_b200.configs as cfg  # noqa: E402


def main(P: int = 1000, N: int = 257, iterations: int = 50, seed: int = 5, record_from: int = 10) -> dict:
    # target, proposals (guiding tables built on the device in reference arithmetic), start point, ρ   (:36-50, :96-120)
    Pm, guides, x0, ρ = cfg.fhn_config4(N)
    S = len(guides)
    ens = B.PathEnsemble(P, S, N, 2, 1)
    for s, g in enumerate(guides):
        ens.set_grid(s, g.tt)
    ens.set_start(x0)
    ens.sample_(seed, 0xFFFFFFFE)                     # sample!(W, Wiener())                              :126
    ens.guided_euler_ll_(Pm, guides)                  # solve!(Euler(), X, x0, W, P°); ll = llikelihood() :127-128
    ens.reset_acc()
    ens.chain_mc_reset_()                             # mcstate = [mcstart(XX[i].yy) ...]                 :169
    for it in range(iterations):
        ens.pcn_step_(Pm, guides, ρ, seed, it)        # W° = ρ W + sqrt(1-ρ²) W2; X°, ll°; accept/reject  :139-163
        if it >= record_from:
            ens.chain_mc_update_()                    # mcstate = [mcnext!(mcstate[i], XX[i].yy) ...]     :174
        if (it + 1) % 10 == 0:
            print(f"iteration {it + 1:5d}  acceptance {ens.acc / ((it + 1) * P):.3f}  mean ll {float(np.mean(ens.ll)):.4f}",
                  flush=True)
    mean, cov, k = ens.chain_mc_stats(0, min(P, 8))   # mcstats of the first chains
    lo, hi = ens.chain_mc_band(0, min(P, 8))          # mcband
    out = dict(acc=ens.acc, k=k, ll_sum=float(np.sum(ens.ll)), mean_mid=mean[0, S // 2, N // 2].tolist(),
               band_mid=[lo[0, S // 2, N // 2].tolist(), hi[0, S // 2, N // 2].tolist()],
               obs_fit=float(np.max(np.abs(mean[:, :, -1, 0] - np.asarray(cfg.FHN_OBS_V)))))
    print("summary acc", out["acc"], "k", k, "ll_sum", repr(out["ll_sum"]), "obs_fit", repr(out["obs_fit"]))
    print("chain 0, mid point of segment", S // 2, ": mean", out["mean_mid"], "95 % band", out["band_mid"])
    ens.close()
    return out


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:4]]
    main(*a)
