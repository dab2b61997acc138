"""The drop-in boundary from plain C: examples/pcn_fhn.c is compiled against include/bridge_b200.h with gcc (so the header,
not the ctypes table, is what is checked) and linked to libbridge_b200.so.  CPU: the program must refuse to compute
(BB_ERR_NODEVICE, no CPU path).  GPU: its pCN run must agree with the same run through the Python mirror."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "bridge.jl_b200", "lib")


def build(tmp_path):
    exe = str(tmp_path / "pcn_fhn")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-std=c11", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "pcn_fhn.c"), "-o", exe, "-L", LIBDIR, "-lbridge_b200",
                    f"-Wl,-rpath,{LIBDIR}", "-lm"], check=True)
    return exe


def test_c_example_builds_and_refuses_without_a_gpu(tmp_path):
    import torch
    exe = build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([exe, "10", "33", "2"], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_c_example_matches_python_mirror(tmp_path):
    import bridge_jl_b200 as B
    import bridge_jl_b200.configs as cfg
    exe = build(tmp_path)
    P, N, iters = 777, 129, 4
    r = subprocess.run([exe, str(P), str(N), str(iters)], capture_output=True, text=True, check=True)
    tok = r.stdout.split()
    acc_c, ll_c, x_c = int(tok[tok.index("acc") + 1]), float(tok[tok.index("ll_sum") + 1]), float(tok[tok.index("xend_sum") + 1])
    Pm, guides, x0, rho = cfg.fhn_config4(N, obs_t=(0.5, 1.0), obs_v=(-1.0, -0.5))
    ens = B.PathEnsemble(P, 2, N, 2, 1)
    for s, g in enumerate(guides):
        ens.set_grid(s, g.tt)
    ens.set_start(x0)
    ens.sample_(44, 0xFFFFFFFE)
    ens.guided_euler_ll_(Pm, guides)
    for it in range(iters):
        ens.pcn_step_(Pm, guides, rho, 44, it)
    assert ens.acc == acc_c and 0 < acc_c < iters * P
    assert abs(float(np.sum(ens.ll)) - ll_c) <= 1e-10 * abs(ll_c)
    assert abs(float(np.sum(ens.xend)) - x_c) <= 1e-10 * abs(x_c)
    ens.close()


@pytest.mark.gpu
def test_python_example_of_the_script_loop():
    """examples/fitzhugh_smoothing.py: the partialbridge_fitzhugh.jl:125-189 loop (pCN + per-chain mcnext!) through the host
    mirror runs, is reproducible (same seed, same chains: bit-identical), and its statistics make sense."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("fitzhugh_smoothing", os.path.join(ROOT, "examples", "fitzhugh_smoothing.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    a = ex.main(300, 65, 20, seed=5, record_from=8)
    b = ex.main(300, 65, 20, seed=5, record_from=8)
    assert a == b
    assert a["k"] == 12 and 0 < a["acc"] < 20 * 300
    lo, hi = a["band_mid"]
    assert all(l <= m <= h for l, m, h in zip(lo, a["mean_mid"], hi)) and hi[0] > lo[0]
    assert a["obs_fit"] < 0.05        # the chains' mean paths pass through the observations (Σ = 1e-10)
    assert np.isfinite(a["ll_sum"])
