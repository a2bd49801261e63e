#!/usr/bin/env python
"""Benchmark of the pyBMC inference hot path on B200 (driver contract: see README / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

Headline workload = BASELINE.json configs[2]: synthetic ensemble of 16 models x 3000 points,
K = 8 SVD components, 65,536 conjugate-Gibbs chains x 10,000 iterations per GPU (chains shard
across ranks with no data-path collective: weak scaling).  A "step" is one full run of those chains.

  value  chain-iterations / s, whole job, problem constants resident in HBM (CUDA events)
  e2e    the same metric through the public API `pybmc_b200.run_gibbs` with host NumPy inputs:
         H2D of (y, X), sufficient statistics, sampler, D2H of moments + the kept samples
  extra  the other metric of BASELINE.json (posterior-pred samples x points / s, configs[3]),
         the simplex sampler (configs[1]), the fp64 sampler and an HBM-bound kernel of the path

One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gibbs_chain_iters_per_sec"
UNIT = "chain-iters/s"
CHAINS_PER_GPU = 65536
ITERATIONS = 10000
KEEP_PER_CHAIN = 10
SEED = 0xB200 + 3

# ------------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------------
def config3_ensemble():
    rng = np.random.default_rng(1003)
    n, m = 3000, 16
    t = np.cumsum(rng.uniform(5, 15, n))
    preds = t[:, None] * (1 + rng.normal(0, 0.003, m))[None, :] + rng.normal(0, 2, m)[None, :] \
        + rng.normal(0, 0.5, (n, m))
    truth = t + rng.normal(0, 0.15, n)
    return preds, truth


def config1_ensemble():
    """Seeded surrogate of the missing selected_data.h5: 629 nuclei x 15 mass models."""
    rng = np.random.default_rng(1001)
    pts = [(nn, z) for z in range(8, 111, 2) for nn in range(z, min(161, int(1.6 * z) + 12), 2)]
    pts = np.array(pts[:: max(1, len(pts) // 629)][:629], dtype=float)
    nn, z = pts[:, 0], pts[:, 1]
    a = nn + z

    def semf(c):
        av, as_, ac, aa, ap = c
        return av * a - as_ * a ** (2 / 3) - ac * z * (z - 1) / a ** (1 / 3) - aa * (nn - z) ** 2 / a + ap / np.sqrt(a)
    base = np.array([15.8, 18.3, 0.714, 23.2, 12.0])
    truth = semf(base) + rng.normal(0, 0.15, len(a))
    preds = np.column_stack([semf(base * (1 + rng.normal(0, 0.003, 5))) + rng.normal(0, 2.0)
                             + rng.normal(0, 0.5, len(a)) for _ in range(15)])
    return preds, truth


def config4_inputs(n_points, n_draws, k=16):
    rng = np.random.default_rng(1004)
    m = 24
    preds = rng.uniform(100, 2000, n_points)[:, None] + rng.normal(0, 3.0, (n_points, m))
    vt = rng.normal(size=(k, m)) * 0.03
    beta_star = rng.normal(0, 1, k)
    theta = np.column_stack([beta_star[None, :] + 0.1 * rng.normal(size=(n_draws, k)),
                             np.abs(rng.normal(0.15, 0.01, n_draws))])
    u = preds @ vt.T
    truth = preds.mean(axis=1) + u @ beta_star + rng.normal(0, 0.15, n_points)
    return preds, vt, theta, truth


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks and throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def flush_l2(torch, buf):
    buf.add_(1)          # 512 MiB read + write > 126 MB L2


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port on NumPy's generators (= the reference algorithm)
# ------------------------------------------------------------------------------------------------
def _cpu_chain(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    y, X, prior, iters, seed = args
    from oracle import bmc_oracle as oc
    np.random.seed(seed)
    t0 = time.perf_counter()
    oc.gibbs_conjugate(y, X, iters, prior)
    return time.perf_counter() - t0


def cpu_problem():
    from oracle import bmc_oracle as oc
    preds, truth = config3_ensemble()
    r = oc.orthogonalize_arrays(preds, truth, 8, full_matrices=False)
    prior = [np.zeros(8), np.diag(r["S_hat"] ** 2), 1.0, 0.02]
    return r["y"], np.ascontiguousarray(r["U_hat"]), prior


def cpu_sampler_rate(iters_per_chain, pool, cores, problem):
    y, X, prior = problem
    t0 = time.perf_counter()
    pool.map(_cpu_chain, [(y, X, prior, iters_per_chain, 1000 + c) for c in range(cores)])
    dt = time.perf_counter() - t0
    return cores * iters_per_chain / dt, dt


def run_reference(args):
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = "1"
    cores = os.cpu_count() or 1
    problem = cpu_problem()
    iters = 1500
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_sampler_rate(200, pool, cores, problem)
        times = []
        for _ in range(args.steps):
            _, dt = cpu_sampler_rate(iters, pool, cores, problem)
            times.append(dt)
    total = float(np.sum(times))
    value = cores * iters * args.steps / total
    sample = f"{cores} chains x {iters} iterations per step on {cores} host processes (oracle port, NumPy RNG)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": "BASELINE configs[2]: synthetic ensemble 16 models x 3000 points, K=8 SVD components, "
                        "conjugate Gibbs, 65536 chains x 10000 iterations per GPU",
            "n_points": 3000, "n_models": 16, "components": 8, "chains_per_gpu": CHAINS_PER_GPU,
            "chains_total": CHAINS_PER_GPU * n_gpus, "iterations": ITERATIONS, "kept_per_chain": KEEP_PER_CHAIN,
            "moments": "full cross moments in fp64", "sharding": f"chains over {n_gpus} GPU(s), no data-path collective; "
            "one all-reduce of the 54 moment sums per step",
            "l2": "512 MiB buffer rewritten between timed steps (working set is K-sized, not L2-resident data)"}


# ------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist
    import pybmc_b200 as pb
    from pybmc_b200 import _lib
    from pybmc_b200.inference_utils import ConjugateSampler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pybmc_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if not os.path.exists(_lib.LIB_PATH):       # snapshot without the built library: compile it (rank 0), never fall back
        if rank == 0:
            from pybmc_b200.build import build_library
            build_library(force=True)
        if world > 1:
            dist.barrier()
    lib = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- set-up (untimed): orthogonalise on the device, build the resident sampler -----------------
    preds, truth = config3_ensemble()
    orth = pb.orthogonalize_arrays(preds, truth, 8)
    y, X = orth["y"], orth["U_hat"]
    prior = [np.zeros(8), np.diag(orth["S_hat"] ** 2), 1.0, 0.02]
    sampler = ConjugateSampler(y, X, prior, device=dev)
    thin = ITERATIONS // KEEP_PER_CHAIN
    chain0 = rank * CHAINS_PER_GPU
    flush = torch.zeros(128 * 2 ** 20, dtype=torch.float32, device=dev)

    def device_step(dtype="float32"):
        samples, cstats, meta = sampler.run(ITERATIONS, CHAINS_PER_GPU, SEED, dtype, thin, 0, True, "full", chain0)
        total = cstats.sum(dim=1)
        if world > 1:
            dist.all_reduce(total)              # the only collective: 54 fp64 moment sums
        return samples, total

    def timed(step_fn, steps, warmup, robust=False):
        """K timed steps (CUDA events, L2 rewritten before each), max over ranks of the total.  The
        headline uses the plain mean of exactly K steps; the extras (robust=True) take the median step so
        that one disturbed step of a 2-3 step sample does not decide the number."""
        for _ in range(warmup):
            step_fn()
        barrier()
        ms = []
        for _ in range(steps):
            flush_l2(torch, flush)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step_fn()
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        barrier()
        if robust:
            return max_over_ranks(float(np.median(ms)))
        return max_over_ranks(float(np.sum(ms))) / steps

    if args.only == "predict":
        out = extras(args, torch, dist, dev, world, rank, sampler, pb, timed, max_over_ranks, barrier, flush, 6551.0,
                     only_predict=True)
        if rank == 0:
            print(json.dumps(out), flush=True)
        return
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_step = timed(device_step, args.steps, args.warmup)
    clock_info = clocks.stop() if rank == 0 else None
    units = CHAINS_PER_GPU * world * ITERATIONS
    value = units / (ms_step * 1e-3)

    # ---- e2e through the public API, host buffers in, host results out ------------------------------
    y_h, X_h = np.ascontiguousarray(y), np.ascontiguousarray(X)

    def e2e_step():
        res = pb.run_gibbs(y_h, X_h, ITERATIONS, prior, n_chains=CHAINS_PER_GPU, seed=SEED, dtype="float32",
                           thin=thin, stats="full", device=dev, chain_offset=chain0)
        return res
    e2e_results = [e2e_step() for _ in range(max(3, args.warmup))]    # warm the pinned-host block cache too
    del e2e_results
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = e2e_step()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / args.steps
    barrier()
    h2d = y_h.nbytes + X_h.nbytes + 8 * (8 + 4 * 8)            # design + OLS vector + problem constants
    d2h = res.samples.nbytes + 8 * (54 + 9 * 9 + 1 + 9 * CHAINS_PER_GPU)
    e2e = {"value": units / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": 1e3 * e2e_s}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "e2e": e2e, "gpu_launches": args.steps * 1 + args.steps * 5,
            "gpu_launches_detail": "value region: 1 kernel per step (gibbs_conjugate_kernel); e2e region: 5 per step "
                                   "(gram_partial, sum_partials, rss_partial, sum_partials, gibbs_conjugate_kernel); "
                                   "reductions of the moment rows and the sample transpose are torch ops",
            "clocks": clock_info}

    # ---- roofline of the dominant kernel --------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    sm_mhz = (clock_info or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    per_gpu_rate = value / world
    alg_bytes = CHAINS_PER_GPU * (KEEP_PER_CHAIN * 9 * 4 + 54 * 8 * 2 * (ITERATIONS // 64))
    kernel_s = ms_step * 1e-3
    inst = load_inst_per_iter()
    nominal_issue = 148 * 4 * sm_mhz * 1e6 / 1e9                 # 148 SMs x 4 schedulers x clock, Gwarp-inst/s
    pipe_peaks = measure_pipe_peaks(torch, dev)
    issue_peak = max(pipe_peaks["issue"]["Gwarp_inst_per_s"], pipe_peaks["ffma"]["Gwarp_inst_per_s"])
    _PEAKS["issue"] = issue_peak
    fma = fma_pipe_view(per_gpu_rate, pipe_peaks) or {}
    line["roofline"] = {
        "kernel": "gibbs_conjugate_kernel<float,8,2>",
        "bound": "fma-pipe", "unit": "Gwarp-inst/s (FFMA slots)",
        "achieved": fma.get("achieved"), "peak": fma.get("peak"), "frac": fma.get("frac"),
        "peak_source": "FFMA rate measured in this run by bmc_probe (independent streams, all SMs); nominal "
                       "148 SMs x 4 schedulers x sampled SM clock = %.0f" % nominal_issue,
        "work": "FMA-pipe instructions of one chain-iteration priced in FFMA slots (IMAD.WIDE and packed fp32 "
                "cost more than one, measured): SURVEY.md section 8d counts the same items -- 2.5 Philox calls, "
                "9 normals, the K-component update, the moment sums",
        "fma_pipe": fma,
        "issue": {"bound": "issue", "unit": "Gwarp-inst/s",
                  "achieved": (inst * per_gpu_rate / 32 / 1e9) if inst else None, "peak": issue_peak,
                  "frac": (inst * per_gpu_rate / 32 / 1e9 / issue_peak) if inst else None,
                  "warp_inst_per_chain_iter": inst,
                  "note": "all executed warp-instructions against the measured issue peak; fell from 0.63 as "
                          "instructions were removed (427 -> 261 per chain-iteration) faster than time"},
        "pipe_peaks_measured": pipe_peaks,
        "traffic": load_dram_bytes(),
        "hbm": {"bound": "hbm", "achieved": alg_bytes / kernel_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes / kernel_s / 1e9 / hbm_peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs"
                if peaks else "fallback 6650 GB/s",
                "note": "algorithmic bytes = kept samples + fp64 moment flushes; the sampler's state is K-sized "
                        "and lives in registers, so HBM is not the bound (SURVEY.md section 8d)"}}

    # ---- the rest of BASELINE.json's metric, measured the same way -------------------------------------
    if args.only == "sampler":
        if rank == 0:
            print(json.dumps(line), flush=True)
        return
    if rank == 0 or world > 1:
        line["extra"] = extras(args, torch, dist, dev, world, rank, sampler, pb, timed, max_over_ranks, barrier,
                               flush, hbm_peak)
    if rank == 0 and world == 1:
        line["cpu_baseline"] = cpu_baseline()
        line["extra"]["predict_cpu_baseline"] = cpu_predict_baseline()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_pipe_peaks(torch, dev):
    """FP32 FMA, MUFU, Philox-mix and dual-issue peaks of this GPU, measured now (bmc_probe)."""
    from pybmc_b200 import _lib
    lib = _lib.load()
    sink = torch.zeros(1, dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    blocks, threads, iters = 148 * 8, 256, 20000
    out = {}
    for kind, name in ((0, "ffma"), (1, "mufu"), (2, "philox_mix"), (3, "issue"), (4, "imad_wide_plus_iadd"),
                       (5, "imad_hi"), (6, "imad"), (7, "lop3"), (8, "ffma2"), (9, "imad_wide"),
                       (11, "imad_wide_plus_2ffma")):
        ops = lib.bmc_probe_ops_per_iteration(kind)
        best = None
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.bmc_probe(kind, iters, blocks, threads, sink.data_ptr(), st))
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        thread_ops = float(blocks) * threads * iters * ops
        out[name] = {"Gwarp_inst_per_s": thread_ops / 32 / (best * 1e-3) / 1e9, "ms": best}
    out["ffma"]["TFLOP_per_s"] = out["ffma"]["Gwarp_inst_per_s"] * 32 * 2 / 1e3
    return out


_PEAKS = {}


def predict_roofline(torch, dev, units_per_s_per_gpu, ms_step):
    """Issue roofline of the fused prediction step: the tensor-core pass is bound by its consumer
    (Philox, Box-Muller, compares, predicated counts), not by the MMAs or HBM."""
    try:
        c = json.load(open(os.path.join(ROOT, "profiles", "kernel_constants.json")))["predict_pass_tc_f32_k16_q5"]
    except (OSError, KeyError, ValueError):
        return None
    if "issue" not in _PEAKS:
        pk = measure_pipe_peaks(torch, dev)
        _PEAKS["issue"] = max(pk["issue"]["Gwarp_inst_per_s"], pk["ffma"]["Gwarp_inst_per_s"])
    per_unit = float(c["thread_inst_per_sample_point"]) / 32.0
    achieved = per_unit * units_per_s_per_gpu / 1e9
    flops = 2.0 * 16 * 3 * units_per_s_per_gpu / 1e12          # three TF32 products per component
    return {"kernel": "predict_pass_tc_kernel<16,5> (+ predict_select_kernel)", "bound": "issue", "unit": "Gwarp-inst/s",
            "achieved": achieved, "peak": _PEAKS["issue"], "frac": achieved / _PEAKS["issue"],
            "thread_inst_per_sample_point": float(c["thread_inst_per_sample_point"]),
            "traffic": float(c["dram_bytes_per_launch"]),
            "tensor": {"kind": "tcgen05.mma kind::tf32, split TF32 (3 products)", "TFLOP_per_s": flops,
                       "note": "K=16: the MMAs keep the tensor pipe ~3 % busy (ncu); the 16 consumer warps are the bound"},
            "note": "whole step (pass + select + window set-up) against the pass kernel's instruction count"}


def fma_pipe_view(rate_per_gpu, pipe_peaks):
    """The sampler's bounding unit is the FMA pipe.  bmc_probe shows that on B200 a 32x32->64 multiply
    (IMAD.WIDE, two per Philox round) holds the whole pipe for ~4 cycles -- no FFMA flows beside it
    (IMAD.WIDE + 2 FFMA = 6.1 cycles) -- and that a packed FFMA2 is two FFMA slots.  Work = the kernel's
    FMA-pipe instructions per warp-iteration (ncu opcode mix, profiles/kernel_constants.json), each class
    priced in FFMA slots by the rates measured in this run; peak = the measured FFMA rate."""
    try:
        mix = json.load(open(os.path.join(ROOT, "profiles", "kernel_constants.json")))[
            "gibbs_conjugate_f32_k8_full"]["fma_pipe_mix"]
    except (OSError, KeyError, ValueError):
        return None
    ffma = pipe_peaks["ffma"]["Gwarp_inst_per_s"]
    price = {"fp32": 1.0, "fp32x2": ffma / pipe_peaks["ffma2"]["Gwarp_inst_per_s"],
             "imad": ffma / pipe_peaks["imad"]["Gwarp_inst_per_s"],
             "imad_wide": ffma / pipe_peaks["imad_wide"]["Gwarp_inst_per_s"]}
    slots = sum(float(mix[k]) * price[k] for k in price)                  # FFMA slots per warp-iteration
    achieved = slots * rate_per_gpu / 32 / 1e9
    return {"achieved": achieved, "peak": ffma, "frac": achieved / ffma, "ffma_slots_per_warp_iter": slots,
            "price_in_ffma_slots": price, "mix_per_warp_iter": mix}


def load_inst_per_iter():
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "kernel_constants.json")))[
            "gibbs_conjugate_f32_k8_full"]["warp_inst_per_chain_iter"])
    except (OSError, KeyError, ValueError):
        return None


def load_dram_bytes():
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "kernel_constants.json")))[
            "gibbs_conjugate_f32_k8_full"]["dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        return None


def extras(args, torch, dist, dev, world, rank, sampler, pb, timed, max_over_ranks, barrier, flush, hbm_peak,
           only_predict=False):
    out = {}
    steps = 3
    timed_main = timed
    timed = lambda fn, k, w: timed_main(fn, k, max(w, 2), robust=True)   # noqa: E731
    if not only_predict:
        extras_samplers(out, timed, sampler, pb, dev, world, rank, steps)
    extras_predict(out, args, torch, dev, world, rank, timed, barrier, flush, hbm_peak, steps, only_predict)
    if not only_predict:
        extras_config5(out, torch, dev, world, rank, timed, pb, steps, hbm_peak)
        if rank == 0:
            extras_config0(out, torch, pb, dev)
    barrier()
    return out


def extras_config0(out, torch, pb, dev):
    """BASELINE configs[0], the reference's own workflow through the drop-in class (surrogate for the absent
    selected_data.h5: 629 nuclei x 15 models, 377 training rows, K = 3): wall-clock of each public call,
    second call of a process (the first also pays one-time CUDA / cuSOLVER initialisation)."""
    import contextlib
    import io
    import pandas as pd
    preds, truth = config1_ensemble()
    models = [f"m{i}" for i in range(preds.shape[1])]
    df = pd.DataFrame(preds, columns=models)
    df["N"] = np.arange(len(truth))
    df["Z"] = np.arange(len(truth)) // 3
    df["truth"] = truth
    train = df.iloc[np.random.default_rng(1).permutation(len(df))[:377]]
    bmc = pb.BayesianModelCombination(models, {"BE": df}, "truth")

    def wall(fn):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize(dev)
        return 1e3 * (time.perf_counter() - t0)
    calls = (("orthogonalize", lambda: bmc.orthogonalize("BE", train, 3)),
             ("train_50000_iterations", lambda: bmc.train({"iterations": 50000})),
             ("predict2", lambda: bmc.predict2("BE")),
             ("evaluate", lambda: bmc.evaluate()))
    ms = {}
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(2):
            for name, fn in calls:
                ms[name] = wall(fn)
    out["config0_workflow_ms"] = dict(ms, config="configs[0] surrogate: 629 x 15, K=3, one fp64 chain of 50,000 iterations, "
                                                 "10,000 draws materialised by predict2, 21 coverage levels; reference "
                                                 "(NumPy, one core): ~7000 / 840 / 2500 ms for train / predict2 / evaluate")


def extras_samplers(out, timed, sampler, pb, dev, world, rank, steps):
    from pybmc_b200.inference_utils import SimplexSampler
    chain0 = rank * CHAINS_PER_GPU
    thin = ITERATIONS // KEEP_PER_CHAIN
    # conjugate sampler in fp64 (the reference's arithmetic type)
    ms = timed(lambda: sampler.run(ITERATIONS // 4, CHAINS_PER_GPU, SEED, "float64", thin, 0, True, "full", chain0),
               steps, 1)
    out["gibbs_f64"] = {"value": CHAINS_PER_GPU * world * (ITERATIONS // 4) / (ms * 1e-3), "unit": UNIT,
                        "ms_per_step": ms, "config": "same problem, fp64 arithmetic, 2500 iterations per chain"}

    # simplex sampler, BASELINE configs[1]: nuclear-mass surrogate, 4096 chains per GPU
    preds, truth = config1_ensemble()
    rng = np.random.default_rng(1)
    idx = rng.permutation(len(truth))[:377]
    o = pb.orthogonalize_arrays(preds[idx], truth[idx], 3)
    simplex = SimplexSampler(o["y"], o["U_hat"], o["Vt_hat"], o["S_hat"], [1.0, 0.02], 0.001, device=dev)
    burn, iters, chains = 10000, 50000, 4096
    ms = timed(lambda: simplex.run(iters, burn, chains, SEED, "float32", iters // 10, True, "full", rank * chains),
               steps, 1)
    out["simplex_f32"] = {"value": chains * world * (burn + iters) / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                          "config": "configs[1] surrogate: 377 x 15, K=3, 4096 chains/GPU x (10000 burn + 50000)"}

    # the literal one-chain-per-warp kernel (parity anchor): redoes the O(nK) residual every iteration
    import torch
    from pybmc_b200 import _lib
    lib = _lib.load()
    n, k = sampler.n, sampler.k
    xt = sampler._Xd.t().contiguous().to(torch.float32)
    yr = sampler._yd.to(torch.float32)
    consts = torch.from_numpy(np.concatenate([sampler.lam.reshape(-1), sampler.lam @ sampler.b0])).to(dev)
    lit_chains, lit_iters = 8192, 200
    buf = torch.empty((lit_iters, k + 1, lit_chains), dtype=torch.float32, device=dev)

    def lit_step():
        _lib.check(lib.bmc_gibbs_literal_run(_lib.F32, xt.data_ptr(), yr.data_ptr(), n, k, consts.data_ptr(),
                                             consts.data_ptr() + 8 * k * k, sampler.nu0, sampler.sigma20,
                                             sampler.sigma2_init, SEED, rank * lit_chains, lit_chains, lit_iters,
                                             buf.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    ms = timed(lit_step, steps, 1)
    rate = lit_chains * world * lit_iters / (ms * 1e-3)
    out["literal_f32"] = {"value": rate, "unit": UNIT, "ms_per_step": ms,
                          "reference_equivalent_tflops": rate * (4 * n * k + 3 * n) / 1e12,
                          "config": "same problem, one chain per warp, X in shared memory via TMA, residual over all "
                                    "3000 rows each iteration; 8192 chains x 200 iterations per GPU"}



def extras_predict(out, args, torch, dev, world, rank, timed, barrier, flush, hbm_peak, steps, only_predict):
    from pybmc_b200.sampling_utils import PredictiveProblem
    from pybmc_b200 import _lib
    lib = _lib.load()
    # fused prediction, BASELINE configs[3]: 1e5 nuclei (sharded over ranks) x 1e5 draws x K=16
    n_total, n_draws = 100_000, 100_000
    per = -(-n_total // world) // 4 * 4 + (4 if (-(-n_total // world)) % 4 else 0)
    lo, hi = min(rank * per, n_total), min((rank + 1) * per, n_total)
    preds, vt, theta, truth = config4_inputs(n_total, n_draws)
    prob = PredictiveProblem(preds[lo:hi], theta, vt, truth=truth[lo:hi], dtype="float32", device=dev, point0=lo)
    nbytes = int(lib.bmc_predict_workspace_bytes(_lib.F32, hi - lo, 5, n_draws))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    q = [2.5, 16.0, 50.0, 84.0, 97.5]
    holder = {}

    def pred_step():
        holder["r"] = prob.run(percentiles=q, seed=SEED, as_numpy=False, workspace=ws)
    ms = timed(pred_step, steps, 1)
    out["predict_f32"] = {"metric": "posterior_pred_samples_x_points_per_sec",
                          "value": n_total * n_draws / (ms * 1e-3), "unit": "samples*points/s", "ms_per_step": ms,
                          "passes": holder["r"].passes,
                          "roofline": predict_roofline(torch, dev, n_total * n_draws / (ms * 1e-3) / world, ms),
                          "config": "configs[3]: 1e5 nuclei x 1e5 draws x K=16, mean/var/5 percentiles/coverage "
                                    "counts, no S x N matrix, nuclei sharded over ranks"}
    del prob, ws

    # the same step end to end through the public call: host arrays in (this rank's predictions, truth, all
    # posterior rows), host results out (mean, variance, percentiles, order counts)
    from pybmc_b200.sampling_utils import predictive_summary
    p_h, t_h = np.ascontiguousarray(preds[lo:hi]), np.ascontiguousarray(truth[lo:hi])

    def pred_e2e():
        return predictive_summary(p_h, theta, vt, truth=t_h, percentiles=q, seed=SEED, dtype="float32",
                                  subsample=False, device=dev, point0=lo)
    for _ in range(2):
        r = pred_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        r = pred_e2e()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / steps
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    out["predict_f32"]["e2e"] = {
        "value": n_total * n_draws / e2e_s, "unit": "samples*points/s", "ms_per_step": 1e3 * e2e_s,
        "h2d_bytes_per_step": int(p_h.nbytes + t_h.nbytes + theta.nbytes + vt.nbytes),
        "d2h_bytes_per_step": int((hi - lo) * 8 * (2 + len(q) + 2))}
    del r

    # an HBM-bound kernel of the path: order counts of a materialised matrix (coverage())
    if rank == 0 and not only_predict:
        s_rows, n_cols = 10000, 65536
        mat = torch.randn((s_rows, n_cols), dtype=torch.float64, device=dev)
        tr = torch.zeros(n_cols, dtype=torch.float64, device=dev)
        c1 = torch.empty(n_cols, dtype=torch.int64, device=dev)
        c2 = torch.empty(n_cols, dtype=torch.int64, device=dev)

        def cov_step():
            _lib.check(lib.bmc_coverage_counts(mat.data_ptr(), s_rows, n_cols, n_cols, tr.data_ptr(), c1.data_ptr(),
                                               c2.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        for _ in range(2):
            cov_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush_l2(torch, flush)
        e0.record(); cov_step(); e1.record(); e1.synchronize()
        gbs = mat.numel() * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9
        out["coverage_counts_hbm"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                      "frac": gbs / hbm_peak, "config": "10000 x 65536 fp64 matrix read once"}
        del mat
    barrier()
    return out


def extras_config5(out, torch, dev, world, rank, timed, pb, steps, hbm_peak):
    """BASELINE configs[4]: 256 models x 1e5 points, K = 64 (rows sharded over ranks for the projection)."""
    from pybmc_b200 import _lib
    lib = _lib.load()
    n_total, m, k = 100_000, 256, 64
    n = n_total // world
    gen = torch.Generator(device=dev).manual_seed(1005 + rank)
    latent = torch.randn((n, k), generator=gen, device=dev, dtype=torch.float64) * torch.logspace(
        0, -3, k, device=dev, dtype=torch.float64)
    mix = torch.randn((k, m), generator=torch.Generator(device=dev).manual_seed(7), device=dev, dtype=torch.float64)
    preds = (torch.rand((n, 1), generator=gen, device=dev, dtype=torch.float64) * 1900 + 100
             + 30 * latent @ mix + 0.05 * torch.randn((n, m), generator=gen, device=dev, dtype=torch.float64))
    truth = preds.mean(dim=1)
    mu = torch.empty(n, dtype=torch.float64, device=dev)
    y = torch.empty(n, dtype=torch.float64, device=dev)
    xc = torch.empty((n, m), dtype=torch.float64, device=dev)
    gram = torch.empty((m, m), dtype=torch.float64, device=dev)
    ws = torch.empty(int(lib.bmc_gram_workspace_bytes(n, m)), dtype=torch.uint8, device=dev)
    vt = torch.randn((k, m), generator=torch.Generator(device=dev).manual_seed(8), device=dev, dtype=torch.float64)
    u = torch.empty((n, k), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream

    def center():
        _lib.check(lib.bmc_center_rows(preds.data_ptr(), n, m, m, truth.data_ptr(), mu.data_ptr(), y.data_ptr(),
                                       xc.data_ptr(), m, st))

    def gram_step():
        _lib.check(lib.bmc_gram(xc.data_ptr(), n, m, m, None, None, gram.data_ptr(), ws.data_ptr(), ws.numel(), st))

    def project():
        _lib.check(lib.bmc_project_rows(xc.data_ptr(), n, m, m, None, vt.data_ptr(), k, u.data_ptr(), k, st))
    res = {}
    for name, fn, bytes_, flops in (("center_rows", center, 8 * n * m * 2, 0),
                                    ("gram", gram_step, 8 * n * m, 2.0 * n * m * m),
                                    ("project_rows", project, 8 * n * (m + k), 2.0 * n * m * k)):
        ms = timed(fn, steps, 1)
        res[name] = {"ms": ms, "GB/s": bytes_ / (ms * 1e-3) / 1e9, "frac_of_hbm": bytes_ / (ms * 1e-3) / 1e9 / hbm_peak,
                     "fp64_TFLOP/s": flops / (ms * 1e-3) / 1e12}
    out["config5_orthogonalize_f64"] = {"config": f"configs[4]: {n} x 256 fp64 rows per GPU, K=64", **res}
    del preds, xc

    # prediction at K = 64: the contraction runs on the tensor cores (tcgen05 kind::tf32, split TF32)
    from pybmc_b200.sampling_utils import PredictiveProblem
    rng = np.random.default_rng(1005)
    n_draws = 10000
    lo, hi = rank * n, (rank + 1) * n
    pr = rng.uniform(100, 2000, n)[:, None] + rng.normal(0, 3.0, (n, 80))
    vt64 = rng.normal(size=(k, 80)) * 0.02
    vt64 -= vt64.mean(axis=1, keepdims=True)          # Vt_hat of row-centred predictions is orthogonal to 1
    theta = np.column_stack([rng.normal(size=k)[None, :] + 0.1 * rng.normal(size=(n_draws, k)),
                             np.abs(rng.normal(0.15, 0.01, n_draws))])
    prob = PredictiveProblem(pr, theta, vt64, truth=pr.mean(axis=1), dtype="float32", device=dev, point0=lo)
    ws = torch.empty(int(lib.bmc_predict_workspace_bytes(_lib.F32, n, 3, n_draws)), dtype=torch.uint8, device=dev)
    ms = timed(lambda: prob.run(percentiles=[2.5, 50.0, 97.5], seed=SEED, as_numpy=False, workspace=ws), steps, 1)
    out["config5_predict_f32"] = {"metric": "posterior_pred_samples_x_points_per_sec",
                                  "value": float(n) * world * n_draws / (ms * 1e-3), "unit": "samples*points/s",
                                  "ms_per_step": ms, "config": f"configs[4]: {n} nuclei per GPU x 10000 draws x K=64, "
                                  "3 percentiles + coverage counts"}


def _cpu_predict(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    preds, theta, vt, truth, seed = args
    from oracle import bmc_oracle as oc
    rndm_m, _ = oc.predictive_draws(preds, theta, vt, np.random.default_rng(seed))
    # the reference re-sorts every column for each of the 21 levels (pybmc/sampling_utils.py:24-28); the
    # port sorts once, so this baseline is faster than the reference itself
    oc.coverage_levels(np.arange(0, 101, 5), rndm_m, truth)
    return rndm_m.shape[0] * rndm_m.shape[1]


def cpu_predict_baseline():
    """rndm_m_random_calculator + coverage (oracle port) at the reference's S = 10^4 on all host cores."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    preds, vt, theta, truth = config4_inputs(629, 12000)
    jobs = [(preds, theta, vt, truth, 50 + c) for c in range(cores)]
    with mp.get_context("fork").Pool(cores) as pool:
        t0 = time.perf_counter()
        units = sum(pool.map(_cpu_predict, jobs))
        dt = time.perf_counter() - t0
    return {"value": units / dt, "unit": "samples*points/s", "cores": cores, "kind": "port",
            "sample": f"{cores} x (629 points x 10000 draws x K=16: predictive matrix, 3 percentiles, 21 coverage "
                      f"levels), one process per core, {dt:.1f} s"}


def cpu_baseline():
    """Reference algorithm (oracle port) on the host cores, bounded sample of the same workload."""
    import multiprocessing as mp
    os.environ["OMP_NUM_THREADS"] = "1"
    cores = os.cpu_count() or 1
    problem = cpu_problem()
    iters = 100000
    with mp.get_context("fork").Pool(cores) as pool:
        cpu_sampler_rate(100, pool, cores, problem)
        rate, dt = cpu_sampler_rate(iters, pool, cores, problem)
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{cores} chains x {iters} iterations of the same 3000 x 8 problem, one process per core, "
                      f"{dt:.1f} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--only", default="all", choices=["all", "sampler", "predict"],
                    help="profiling aid: restrict the run to one kernel family (no JSON contract)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_native(args)


if __name__ == "__main__":
    main()
