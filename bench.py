#!/usr/bin/env python
"""bench.py -- guided-bridge path-steps/s of the pCN sampler (BASELINE.json config 4) on N B200s.

A "step" is ONE pCN / Metropolis-Hastings iteration of every chain on every rank: fresh Wiener noise,
W° = ρW + sqrt(1-ρ²)W2, guided Euler through 4 chained FitzHugh-Nagumo bridge segments (N = 1001 each),
Girsanov log-likelihood, accept/reject -- one fused kernel launch per rank, plus (N > 1) the all-reduce of the
acceptance counter (bb_allreduce_acc: NCCL on its own stream, behind an event).  path-step = one Euler step of one
chain incl. its ll increment.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--chains P]
                    [--scaling weak|strong] [--config 2|3|4|5]

--scaling weak (default): every rank owns --chains chains (default 250 000 = BASELINE config 4 in full on each GPU;
48 GB of path state per GPU, far beyond the 126 MB L2, so no cache flush is needed between steps).
--scaling strong: --chains chains IN TOTAL (250 000 = config 4 as BASELINE.json names it), sharded over the ranks.
--config 2 | 3 | 5: the other BASELINE configurations at full size on one GPU (kernel value, roofline, e2e through
the public API); they are parity-test cases first, these lines exist so that their numbers are driver-runnable.
--impl reference times the CPU restatement of the reference loop (oracle/, OpenMP over chains on all host
cores); the reference itself is Julia and cannot run in this image.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEGMENTS, NGRID = 4, 1001
ALG_BYTES_PER_STEP = 32  # mode M with X° kept: read W 8 d' + write W° 8 d' + write X° 8 d  (d' = 1, d = 2)
METRIC = "guided_bridge_path_steps_per_s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=[2, 3, 4, 5])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--chains", type=int, default=None, help="chains per GPU (weak) or in total (strong)")
    ap.add_argument("--n", type=int, default=NGRID, help="grid points per segment")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        if os.environ.get("BB_BENCH_NO_SAMPLER") == "1":  # A/B of the sampler's own footprint (tools only)
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for ts, ln in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.05:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:  # region shorter than the sampling period: use every sample we have
            for ts, ln in self.lines:
                f = [x.strip() for x in ln.split(",")]
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------- CPU arm
def oracle_workload(n):
    """The same workload for the CPU restatement: tables from the oracle's own backward chain."""
    from oracle import oracle as O
    O.build()
    O.rebuild_fast_native()  # -march=native must mean THIS host (the .so travels with the snapshot)
    orc = O.load("fast")  # -O3 -march=native build of the same restatement
    par = (0.1, 0.0, 1.5, 0.8, 0.3)
    obs_t, obs_v = (0.5, 1.0, 1.5, 2.0), (-1.0, -0.5, 0.5, 1.1)

    def tau(t0, t1):
        s = np.linspace(0.0, t1 - t0, n)
        return t0 + s * (2.0 - s / (t1 - t0))

    grids = [tau(a, b) for a, b in zip((0.0,) + obs_t[:-1], obs_t)]
    L = np.array([[1.0, 0.0]]); Sg = np.array([[1e-10]])
    nu = np.zeros(2); Hp = np.eye(2) / 1e-3
    nu, Hp = orc.gpupdate_nuH(nu, Hp, L, Sg, [obs_v[-1]])
    guides = [None] * 4
    for i in range(3, -1, -1):
        v = obs_v[i]
        Bt = np.array([[10.0, -10.0], [1.5, -1.0]]); bt = np.array([0.0 - v ** 3 / 0.1, 0.8])
        at = np.array([[0.0, 0.0], [0.0, 0.09]])
        nut, Ht, nu, Hp, _ = orc.backward_nuH(O.ODE_LYAP, grids[i], O.const_aux(Bt, bt, at), nu, Hp, 0.0)
        guides[i] = O.GuideHolder(O.GUIDE_NUH, grids[i], Ht, nut, Bt=Bt, betat=bt)
        if i > 0:
            nu, Hp = orc.gpupdate_nuH(nu, Hp, L, Sg, [obs_v[i - 1]])
    model = O.make_model(O.FHN_HYPO, 2, 1, par)
    return orc, model, guides, np.array([-0.5, -0.6])


_CPU_NOTE = ("C restatement of the Julia loop (oracle/bridge_oracle.c), OpenMP over chains; Julia itself is not "
             "installed in this image.  value = the restatement specialised the way Julia compiles this workload "
             "(SVector{2} arithmetic inlined, constants hoisted, all four normals of a Philox call used, batched "
             "normals; same passes and operation order); as_restated = the generic-dimension restatement the parity "
             "tests use")


def cpu_run(n, seconds, iters=None, chains=None, tuned=True):
    """Times the oracle's OpenMP pCN driver on a bounded sample; returns (steps/s, cores, sample text, ...)."""
    orc, model, guides, x0 = oracle_workload(n)
    cores = len(os.sched_getaffinity(0))  # all host threads (torchrun presets OMP_NUM_THREADS=1; it is overridden)
    steps_per_chain_iter = SEGMENTS * (n - 1)
    run = (lambda P, it: orc.pcn_bench_fhn_tuned(model, guides, P, x0, 0.99, 4, it, nthreads=cores)) if tuned else \
          (lambda P, it: orc.pcn_bench(model, guides, P, x0, 0.99, 4, it, nthreads=cores))
    if chains is None:
        pc = max(64, 8 * cores)
        _, secs, _ = run(pc, 2)
        rate = pc * 2 * steps_per_chain_iter / max(secs, 1e-6)
        iters = 4
        chains = int(max(8 * cores, min(200000, rate * seconds / (iters * steps_per_chain_iter))))
        chains = max(cores, chains - chains % cores)
    acc, secs, _ = run(chains, iters)
    value = chains * iters * steps_per_chain_iter / secs
    sample = (f"{chains} chains x {SEGMENTS} segments x {n - 1} steps x {iters} pCN iterations "
              f"({chains * iters * steps_per_chain_iter:.3g} path-steps, {secs:.2f} s)")
    return value, cores, sample, secs, chains, iters


def cpu_baseline_record(n, seconds):
    """Both CPU numbers: the tuned port (the baseline quoted) and the generic restatement."""
    v, cores, sample, _, _, _ = cpu_run(n, seconds, tuned=True)
    vg, _, sample_g, _, _, _ = cpu_run(n, max(2.0, seconds / 3), tuned=False)
    return {"value": v, "unit": "path-steps/s", "cores": cores, "kind": "port", "sample": sample,
            "as_restated": {"value": vg, "sample": sample_g}, "note": _CPU_NOTE}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config != 4:
        print(json.dumps({"impl": "reference", "unavailable": f"the CPU arm is defined for config 4 only (got {args.config})"}))
        return
    per_step_budget = max(0.5, min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup)))
    value, cores, sample, secs, chains, iters = cpu_run(args.n, per_step_budget)
    for _ in range(max(0, args.warmup - 1)):
        cpu_run(args.n, 0, iters, chains)
    tot_steps, tot_secs = 0.0, 0.0
    for _ in range(args.steps):
        v, _, _, s, c, it = cpu_run(args.n, 0, iters, chains)
        tot_steps += c * it * SEGMENTS * (args.n - 1); tot_secs += s
    value = tot_steps / tot_secs
    vg, _, sample_g, _, _, _ = cpu_run(args.n, 3.0, tuned=False)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "path-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_secs / max(1, args.steps), "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[3]: FitzHugh-Nagumo PartialBridgeνH pCN, 4 segments x N=1001, rho=0.99",
                   "per_step_sample": sample},
        "cpu_baseline": {"value": value, "unit": "path-steps/s", "cores": cores, "kind": "port", "sample": sample,
                         "as_restated": {"value": vg, "sample": sample_g}, "note": _CPU_NOTE},
        "e2e": {"value": value, "unit": "path-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def bind_near_gpu(torch, local):
    """Pin this rank's host threads to the CPUs next to its GPU (NVML's ideal affinity), so that the pinned staging
    buffers of the host-buffer leg are allocated on the GPU's NUMA node.  Returns the original mask (the CPU baseline
    leg runs on ALL host cores and restores it first) or None if NVML is not usable."""
    try:
        import pynvml
        orig = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        try:
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:  # noqa: BLE001
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        if not os.sched_getaffinity(0):
            os.sched_setaffinity(0, orig)
        return orig
    except Exception:  # noqa: BLE001
        return None


KERNEL_SOURCES = {  # the files whose content decides the dominant kernel's DRAM traffic, per config
    4: ["bb_chain.cuh", "bb_device.cuh", "bb_inst_fhn_hypo.cu", "Makefile"],
    2: ["bb_chain.cuh", "bb_device.cuh", "bb_inst_ou.cu", "Makefile"],
    3: ["bb_chain.cuh", "bb_device.cuh", "bb_inst_linpro3.cu", "Makefile"],
    5: ["bb_wide_mma.cuh", "bb_wide.cuh", "bb_device.cuh", "bb_inst_landmarks.cu", "Makefile"],
}


def kernel_source_hash(config):
    h = hashlib.sha256()
    for f in KERNEL_SOURCES[config]:
        with open(os.path.join(ROOT, "bridge.jl_b200", "csrc", f), "rb") as fh:
            h.update(f.encode()); h.update(fh.read())
    return h.hexdigest()[:16]


def measured_traffic(config, units_per_launch):
    """dram bytes per launch from the committed ncu capture -- ONLY if that capture was taken from the kernel sources
    of this very tree (hash of the kernel's source files) and scaled by the units of this launch; else None."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except (OSError, ValueError):
        return None, "no profiles/ncu_traffic.json"
    rec = rec.get(str(config))
    if not rec:
        return None, f"no ncu capture for config {config}"
    now = kernel_source_hash(config)
    if rec.get("src_hash") != now:
        return None, f"ncu capture is of other kernel sources ({rec.get('src_hash')} != {now}): traffic not claimed"
    return rec["dram_bytes_per_unit"] * units_per_launch, f"{rec.get('source')}; scaled per path-step"


def build_workload(args, B, cfg, ctx, rank, world):
    """-> dict(ens, step(it), init(), steps_per_iter (this rank), alg_bytes, kernel, workload text, P)"""
    n = args.n
    if args.config == 4:
        total = args.chains or 250000
        if args.scaling == "strong":
            from bridge_jl_b200.sharding import shard_chains
            first, P = shard_chains(total, rank, world)
        else:
            first, P = rank * total, total
        Pm, guides, x0, rho = cfg.fhn_config4(n, ctx=ctx)
        S = len(guides)
        ens = B.PathEnsemble(P, S, n, 2, 1, double_buffer=True, store_x=True, ctx=ctx, chain_offset=first)
        for s, g in enumerate(guides):
            ens.set_grid(s, g.tt)
        ens.set_start(x0)
        ens.sample_(4, 0xFFFFFFFE)
        ens.guided_euler_ll_(Pm, guides)
        return dict(ens=ens, step=lambda it: ens.pcn_step_(Pm, guides, rho, 4, it), P=P, S=S, Pm=Pm, guides=guides,
                    rho=rho, steps=P * S * (n - 1), alg_bytes=ALG_BYTES_PER_STEP,
                    kernel="bb_chain_kernel<MFhnHypo, NUH, aux const, pCN>",
                    text="configs[3]: FitzHugh-Nagumo (hypoelliptic, d=2, d'=1) PartialBridgeνH pCN, 4 segments x "
                         "N=1001 (tau-warped), rho=0.99, X° stored")
    if world > 1 and args.config != 5:
        raise SystemExit("--config 2 / 3 are single-GPU lines")
    if args.config == 2:
        P = args.chains or 1_000_000
        ens = B.PathEnsemble(P, 1, n, 1, 1, double_buffer=False, ctx=ctx)
        ens.set_grid(0, np.linspace(0.0, 1.0, n)); ens.set_start([0.0])
        ou = B.OrnsteinUhlenbeck(2.0, 1.0)
        ens.sample_(2, 0)
        return dict(ens=ens, step=lambda it: ens.euler_(ou), P=P, S=1, steps=P * (n - 1), alg_bytes=16,
                    kernel="bb_chain_kernel<MOU, plain, read W>",
                    text="configs[1]: EulerMaruyama ensemble, 1e6 independent Float64 paths (OU b=-2x, sigma=1), N=1001, "
                         "mode A: W read, X stored")
    if args.config == 3:
        P = args.chains or 100_000
        Pm, guide, u = cfg.linpro_config3(n, ctx=ctx)
        ens = B.PathEnsemble(P, 1, n, 3, 3, ctx=ctx)
        ens.set_grid(0, guide.tt); ens.set_start(u); ens.sample_(3, 0)
        ens.guided_euler_ll_(Pm, [guide])
        return dict(ens=ens, step=lambda it: ens.guided_euler_ll_(Pm, [guide]), P=P, S=1, steps=P * (n - 1), alg_bytes=48,
                    kernel="bb_chain_kernel<MLinPro<3>, HV, aux const, read W>", Pm=Pm, guides=[guide],
                    text="configs[2]: LinPro d=3 GuidedBridge, 1e5 paths, N=1001, guided Euler + llikelihood "
                         "(mode G: W read, X stored)")
    # config 5: Landmarks d = 16, d' = 8, 1e4 paths in total over the ranks (BASELINE: 4 GPUs)
    total = args.chains or 10_000
    from bridge_jl_b200.sharding import shard_chains
    first, P = shard_chains(total, rank, world) if args.scaling == "strong" or world > 1 else (0, total)
    Pm, Po, x0 = cfg.landmarks_config5(n, ctx=ctx)
    ens = B.PathEnsemble(P, 1, n, 16, 8, ctx=ctx, chain_offset=first)
    ens.set_grid(0, Po.tt); ens.set_start(x0); ens.sample_(5, 0xFFFFFFF0)
    ens.guided_euler_ll_(Pm, [Po])
    return dict(ens=ens, step=lambda it: ens.pcn_step_(Pm, [Po], 0.9, 5, it), P=P, S=1, steps=P * (n - 1),
                alg_bytes=16 * 8 + 8 * 16, kernel="bb_wide4_kernel<NUH, pCN>", Pm=Pm, guides=[Po], rho=0.9,
                text="configs[4]: Landmarks d=16, d'=8 PartialBridgeνH pCN, 1e4 paths in total, N=1001, rho=0.9, X° stored")


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return
    # libraries (NCCL's version banner, ...) write to fd 1: keep the real stdout for the ONE JSON line, send the rest
    # to stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    orig_affinity = bind_near_gpu(torch, local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.config == 5 and world > 1:
        args.scaling = "strong"  # BASELINE names 1e4 paths over 4 GPUs

    import bridge_jl_b200 as B
    import bridge_jl_b200.configs as cfg
    from bridge_jl_b200.sharding import Communicator

    ctx = B.Context(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)  # kernels run on torch's current stream: torch events time them
    dev = torch.device("cuda", local)

    n = args.n
    wl = build_workload(args, B, cfg, ctx, rank, world)
    ens, step, P, S = wl["ens"], wl["step"], wl["P"], wl["S"]
    # the path's one collective: all-reduce of the acceptance counter, issued by the library on its own stream
    comm = Communicator(ctx, rank, world) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    it = 0
    for _ in range(max(3, args.warmup)):
        step(it); it += 1
        if comm:
            comm.allreduce_acc_(ens)
    steps_rank = wl["steps"]
    steps_all = allsum(steps_rank)
    sampler = ClockSampler(local)
    sampler.start()
    t_wait = time.time()
    while not sampler.lines and time.time() - t_wait < 5.0:  # nvidia-smi takes a moment to deliver its first sample
        time.sleep(0.05)
    time.sleep(0.2)
    barrier()
    launches0 = ctx.launch_count
    acc0 = ens.acc
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 2)]
    t0 = time.time()
    ev[0].record(stream)
    for k in range(args.steps):
        ev[2 + 2 * k].record(stream)
        step(it)
        ev[3 + 2 * k].record(stream)
        if comm:
            comm.allreduce_acc_(ens)
        it += 1
    ev[1].record(stream)
    if comm:
        acc_global = comm.acc  # waits for the last all-reduce (side stream): inside the timed wall, after the kernels
    barrier()
    t1 = time.time()
    launches = ctx.launch_count - launches0
    total_ms = allmax(ev[0].elapsed_time(ev[1]))
    kern_ms = [ev[2 + 2 * k].elapsed_time(ev[3 + 2 * k]) for k in range(args.steps)]
    clocks = sampler.stop(t0, t1)
    value = steps_all * args.steps / (total_ms * 1e-3)
    acc_rate = (ens.acc - acc0) / max(1, P * args.steps)  # this rank's acceptance rate over the timed steps
    if comm:
        assert acc_global == int(allsum(ens.acc)), "all-reduced acceptance counter != sum of the ranks' counters"

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    k_ms = float(np.mean(kern_ms))
    achieved = steps_rank * wl["alg_bytes"] / (k_ms * 1e-3) / 1e9
    traffic, traffic_note = measured_traffic(args.config, steps_rank)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_note, "kernel": wl["kernel"],
                "kernel_ms": k_ms, "kernel_ms_max_over_ranks": allmax(k_ms), "alg_bytes_per_path_step": wl["alg_bytes"],
                "alg_bytes_per_launch": steps_rank * wl["alg_bytes"],
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650"}

    # ---- end to end through host buffers: the reference loop keeps W, X in host memory
    e2e = None
    if not args.no_e2e and args.config == 4:
        Pm, guides, rho = wl["Pm"], wl["guides"], wl["rho"]
        # host-buffer leg: every rank pins 4 x (W-sized) arrays; all ranks agree first whether that succeeded, so that a
        # rank that cannot allocate never leaves the others waiting in a collective
        bufs, ok, why = None, 1, ""
        try:
            import psutil
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
            need = P * S * n * 8 * 4
            if psutil.virtual_memory().available / local_world < 1.2 * need:
                raise MemoryError(f"host RAM: need {need / 2**30:.0f} GiB per rank")
            bufs = [torch.empty((P, S, n, k), dtype=torch.float64, pin_memory=True).numpy() for k in (1, 1, 2)]
            bufs.append(torch.empty(P, dtype=torch.float64, pin_memory=True).numpy())
            bufs.append(torch.empty(P, dtype=torch.uint8, pin_memory=True).numpy())
        except Exception as ex:  # noqa: BLE001
            ok, why, bufs = 0, f"{type(ex).__name__}: {ex}", None
        okt = torch.tensor([ok], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        e2e = {}
        if int(okt.item()) == 1:
            hostW, hostWo, hostXo, hostLL, hostAcc = bufs
            ens.download(B.W, out=hostW)

            def e2e_leg(skip_rejected):
                # one call per step on host buffers: W up; W°, X°, ll°, accept flags down (pipelined over chain slabs).
                # skip_rejected: Wo is W itself -- the host array is updated in place for the chains that accept, which
                # is the reference loop's swap of W and Wo, so it holds the chains' current W in every iteration
                nonlocal it
                ens.download(B.W, out=hostW)
                Wo = hostW if skip_rejected else hostWo
                ens.pcn_step_host_(Pm, guides, rho, 4, it, hostW, Wo, hostXo, hostLL, hostAcc,
                                   skip_rejected=skip_rejected); it += 1
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                nacc = 0
                for _ in range(args.e2e_steps):
                    ens.pcn_step_host_(Pm, guides, rho, 4, it, hostW, Wo, hostXo, hostLL, hostAcc,
                                       skip_rejected=skip_rejected); it += 1
                    _ = ens.acc
                    nacc += int(hostAcc.sum())
                e1.record(stream)
                barrier()
                return allmax(e0.elapsed_time(e1)), nacc / args.e2e_steps

            row_bytes = int((hostWo.nbytes + hostXo.nbytes) // P)
            ems, acc_per_step = e2e_leg(True)   # first: the chains' state is consistent with the host's W throughout
            ems_all, _ = e2e_leg(False)
            # the reference loop reads W°, X° only to swap them in on accept (test/partialbridgenuH.jl:184-186), so the rows
            # of the chains that reject do not have to come back: headline = accepted rows only, all rows beside it
            e2e = {"value": steps_all * args.e2e_steps / (ems * 1e-3),
                   "unit": "path-steps/s", "h2d_bytes_per_step": int(hostW.nbytes),
                   "d2h_bytes_per_step": int(acc_per_step * row_bytes + P * 9 + 8),
                   "steps": args.e2e_steps,
                   "what": "bb_pcn_step_host(BB_RUN_SKIP_REJECTED) on pinned host buffers: per step W of every chain up; "
                           "ll°, accept flags of every chain and W°, X° of the chains that ACCEPT down, written straight "
                           "into the mapped host arrays (W in place: the loop's swap); H2D | kernel | D2H pipelined over "
                           "chain slabs",
                   "acc_rate": acc_per_step / P,
                   "all_rows": {"value": steps_all * args.e2e_steps / (ems_all * 1e-3), "unit": "path-steps/s",
                                "d2h_bytes_per_step": int(hostWo.nbytes + hostXo.nbytes + P * 9 + 8),
                                "what": "the same with W°, X° of every chain coming back (rejected proposals too)"}}
        else:
            e2e = {"value": None, "unit": "path-steps/s", "h2d_bytes_per_step": P * S * n * 8,
                   "d2h_bytes_per_step": P * S * n * 24 + P * 9 + 8,
                   "unavailable": "host-buffer leg skipped on all ranks: " + (why or "another rank could not pin memory")}
        bufs = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # device-resident ensemble API: per step only the guide tables go up and ll°/flags/acc come back
        barrier()
        e0.record(stream)
        nres = args.e2e_steps * 2
        Pm2, chain, _, _ = cfg.fhn_config4_chain(n, ctx=ctx)
        for _ in range(nres):
            cfg.fhn_config4_chain(n, ctx=ctx, chain=chain)  # the backward pass of all 4 segments: one launch, in place
            ens.pcn_step_(Pm2, chain.segments, rho, 4, it); it += 1
            _ = ens.ll_prop, ens.accepted, ens.acc
        e1.record(stream)
        barrier()
        ems = allmax(e0.elapsed_time(e1))
        e2e["resident"] = {"value": steps_all * nres / (ems * 1e-3),
                           "unit": "path-steps/s", "d2h_bytes_per_step": P * 9 + 8, "steps": nres,
                           "what": "chain state stays in HBM (PathEnsemble); per step: rebuild the 4 guide tables on the "
                                   "device (bb_guides_chain_nuH, one launch, in place), bb_pcn_step, read back ll°, "
                                   "accept flags, acc"}
    elif not args.no_e2e:
        # configs 2 / 3 / 5 through the public API with host arrays: W up (pinned), step, X and ll down
        k_w, k_x = ens.dprime, ens.d
        hostW = torch.empty((P, S, n, k_w), dtype=torch.float64, pin_memory=True).numpy()
        hostX = torch.empty((P, S, n, k_x), dtype=torch.float64, pin_memory=True).numpy()
        ens.download(B.W, out=hostW)
        which = B.PROP if args.config == 5 else B.CUR

        def api_step(itn):
            ens.upload(B.W, hostW)
            step(itn)
            ens.download(B.X, which=which, out=hostX)
            return ens.ll if args.config != 2 else None

        api_step(it); it += 1
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.e2e_steps):
            api_step(it); it += 1
        e1.record(stream)
        barrier()
        ems = allmax(e0.elapsed_time(e1))
        e2e = {"value": steps_all * args.e2e_steps / (ems * 1e-3), "unit": "path-steps/s",
               "h2d_bytes_per_step": int(hostW.nbytes), "d2h_bytes_per_step": int(hostX.nbytes + (P * 8 if args.config != 2 else 0)),
               "steps": args.e2e_steps,
               "what": "public API on pinned host arrays: bb_ens_upload(W), the step, bb_ens_download(X), ll"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and args.config == 4:
        if orig_affinity is not None:
            os.sched_setaffinity(0, orig_affinity)  # the CPU baseline uses every host core
        cpu = cpu_baseline_record(n, args.cpu_seconds)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "path-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["text"],
                       "chains_this_rank": P, "chains_total": int(round(steps_all / (S * (n - 1)))), "segments": S,
                       "grid_points": n, "path_steps_per_step": steps_all,
                       "state_bytes_per_gpu": ens.nbytes,
                       "l2": ("working set >> 126 MB L2: no flush needed" if ens.nbytes > 4 * 126e6 else
                              "working set comparable to the 126 MB L2: every step rewrites it in full (W°, X°), "
                              "no separate flush"),
                       "host_binding": ("rank threads bound to the GPU's NUMA-local CPUs (NVML) for the host-buffer leg"
                                        if orig_affinity is not None else "none"),
                       "parallelism": f"chains sharded over {world} GPU(s); the acceptance counter is all-reduced by "
                                      f"bb_allreduce_acc (NCCL, side stream) after every step" if world > 1 else
                                      "1 GPU"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "acc_rate": acc_rate,
        }
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if comm:
        comm.close()
    ens.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
