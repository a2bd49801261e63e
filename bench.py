#!/usr/bin/env python
"""Benchmark of the pyBMC inference hot path on B200 (driver contract: see README / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
                    [--metric gibbs|predict] [--dtype f32|f64]

BASELINE.json quotes two metrics; both are first-class here, in both arms:

  gibbs    chain-iterations / s on BASELINE configs[2]: synthetic ensemble of 16 models x 3000 points, K = 8 SVD
           components, 65,536 conjugate-Gibbs chains x 10,000 iterations PER GPU (chains shard over ranks through
           pybmc_b200.parallel.sharded_gibbs, no data-path collective: weak scaling).  One step = one full run.
           Measured in fp32 (the default headline, admitted by BASELINE.json's 1e-5 / 3-MCSE tolerances) AND in
           fp64 (the reference's own arithmetic type) at the same full size: the fp64 numbers are the top-level
           object "f64" of the default line, or the whole line with --dtype f64.
  predict  posterior-pred samples x points / s on BASELINE configs[3]: 1e5 nuclei x 1e5 posterior draws x K = 16,
           mean / variance / 5 percentiles / coverage counts, the S x N matrix never stored; nuclei shard over
           ranks.  Top-level object "predict" of the default line, or the whole line with --metric predict.

Per metric:  value  device-timed (CUDA events), inputs resident in HBM;
             e2e    the same metric through the public API with HOST arrays in and host results out;
             roofline  live rate against the measured pipe peaks + the hardware counters of the committed ncu
                       capture of the same kernel (profiles/kernel_constants.json, regenerated from the .ncu-rep);
             cpu_baseline  the reference algorithm (oracle port) on the host cores (N = 1, rank 0).
`--impl reference` times the reference's CPU path for the same metric(s) on all host cores.

One JSON line on stdout (rank 0).
"""
import argparse
import gc
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
PRED_METRIC = "posterior_pred_samples_x_points_per_sec"
PRED_UNIT = "samples*points/s"
CHAINS_PER_GPU = 65536
ITERATIONS = 10000
KEEP_PER_CHAIN = 10
HIST_EVERY = 64           # marginal histograms of (b, sigma): every 64th state (the kernels' flush points)
SEED = 0xB200 + 3
PRED_POINTS, PRED_DRAWS, PRED_K = 100_000, 100_000, 16
PRED_Q = [2.5, 16.0, 50.0, 84.0, 97.5]


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


def config5_rows(lo, hi, m=256, k=64):
    """Rows [lo, hi) of BASELINE configs[4]'s table (1e5 points x 256 models, 64 latent factors with a graded
    spectrum 1 .. 1e-2 + noise), generated per row so that every rank can build exactly its own block."""
    mix = np.random.default_rng(1005).normal(size=(k, m))
    spec = np.logspace(0, -2, k)
    out = np.empty((hi - lo, m))
    truth = np.empty(hi - lo)
    step = 8192
    for a in range(lo, hi, step):
        b = min(hi, a + step)
        rng = np.random.default_rng([1005, a])
        base = rng.uniform(100, 2000, b - a)
        latent = rng.normal(size=(b - a, k)) * spec
        out[a - lo:b - lo] = base[:, None] + 30.0 * latent @ mix + rng.normal(0, 0.05, (b - a, m))
        truth[a - lo:b - lo] = base + 30.0 * latent @ mix[:, 0] * 0.5 + rng.normal(0, 0.15, b - a)
    return out, truth


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
            # nvidia-smi attaches to every GPU of the node while it starts (hundreds of ms, and it holds driver locks
            # that delay launches on ALL of them: with the start-up inside the five timed steps the 8-GPU headline
            # step read 9.2-9.9 ms instead of 8.5).  The recipe says "start before": wait for its first line, then
            # only the 100 ms polls of this one GPU fall into the timed region.
            t_end = time.time() + 5.0
            while not self.lines and time.time() < t_end and self.proc.poll() is None:
                time.sleep(0.02)
        except OSError:
            self.proc = None
        return self

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            try:
                pw.append(float(f[3]))
            except ValueError:
                pass
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


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


def _cpu_predict(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    preds, theta, vt, truth, seed = args
    from oracle import bmc_oracle as oc
    rndm_m, _ = oc.predictive_draws(preds, theta, vt, np.random.default_rng(seed))       # sampling_utils.py:40-84
    # the reference re-sorts every column for each of the 21 levels (pybmc/sampling_utils.py:24-28); the
    # port sorts once, so this baseline is faster than the reference itself
    oc.coverage_levels(np.arange(0, 101, 5), rndm_m, truth)                              # sampling_utils.py:4-37
    return rndm_m.shape[0] * rndm_m.shape[1]


CPU_PRED_POINTS, CPU_PRED_DRAWS = 629, 10000      # the reference's own sizes: S = 10^4 is hard-coded upstream (:57)


def cpu_predict_rate(pool, cores, inputs, seed0=50):
    preds, vt, theta, truth = inputs
    t0 = time.perf_counter()
    units = sum(pool.map(_cpu_predict, [(preds, theta, vt, truth, seed0 + c) for c in range(cores)]))
    dt = time.perf_counter() - t0
    return units / dt, dt


def gibbs_sample_text(cores, iters):
    return (f"{cores} chains x {iters} iterations of the same 3000 x 8 problem per step, one process per host core "
            f"(oracle port of pybmc/inference_utils.py:4-56 on NumPy's generators; {cores * iters / (CHAINS_PER_GPU * ITERATIONS):.1e} "
            "of one GPU step's chain-iterations: a RATE extrapolation, and conservative -- the port runs ~1.8x the "
            "iterations/s/core of the live reference (no multivariate_normal SVD check, BASELINE.md))")


def predict_sample_text(cores, dt):
    return (f"{cores} x ({CPU_PRED_POINTS} points x {CPU_PRED_DRAWS} draws x K={PRED_K}: predictive matrix, 3 percentiles, 21 "
            f"coverage levels -- oracle port of pybmc/sampling_utils.py:40-84 + :4-37), one process per core, {dt:.1f} s; "
            "rate extrapolation (the reference cannot hold a 1e5 x 1e5 matrix: 80 GB), conservative (the port sorts "
            "each column once, the reference 21 times)")


def pred_config(n_gpus):
    return {"workload": "BASELINE configs[3]: synthetic posterior prediction, 1e5 nuclei x 1e5 posterior draws x K=16, fused "
                        "mean / variance / 5 percentiles / coverage counts, no materialised S x N matrix",
            "n_points": PRED_POINTS, "n_draws": PRED_DRAWS, "components": PRED_K, "percentiles": PRED_Q,
            "sharding": f"nuclei over {n_gpus} GPU(s); e2e: each rank uploads 1/N of the posterior rows, one all-gather "
                        "over NVLink completes them everywhere, one gather of the packed per-nucleus outputs to rank 0",
            "l2": "512 MiB buffer rewritten between timed steps"}


def workload_config(n_gpus, dtype="f32"):
    return {"workload": "BASELINE configs[2]: synthetic ensemble 16 models x 3000 points, K=8 SVD components, "
                        "conjugate Gibbs, 65536 chains x 10000 iterations per GPU",
            "n_points": 3000, "n_models": 16, "components": 8, "chains_per_gpu": CHAINS_PER_GPU,
            "chains_total": CHAINS_PER_GPU * n_gpus, "iterations": ITERATIONS, "kept_per_chain": KEEP_PER_CHAIN,
            "arithmetic": dtype,
            "moments": "all first and cross second moments per chain, summed in the kernel's arithmetic type over 64 "
                       "iterations, then added to fp64 rows in HBM",
            "histograms": f"marginal histograms of (b, sigma), 512 bins, every {HIST_EVERY}th state",
            "sharding": f"chains over {n_gpus} GPU(s) via pybmc_b200.parallel.sharded_gibbs, no data-path collective; "
                        "per step ONE all-reduce: the 54 moment sums + count and the 9 x 512 histogram counts "
                        "(as fp64, exact)",
            "l2": "512 MiB buffer rewritten between timed steps (working set is K-sized, not L2-resident data)"}


def run_reference(args):
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = "1"
    cores = os.cpu_count() or 1
    problem = cpu_problem()
    pred_inputs = config4_inputs(CPU_PRED_POINTS, CPU_PRED_DRAWS + 2000)
    iters = 1500
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_sampler_rate(200, pool, cores, problem)
        times = []
        for _ in range(args.steps):
            _, dt = cpu_sampler_rate(iters, pool, cores, problem)
            times.append(dt)
        p_times = []
        for i in range(max(1, min(args.steps, 3))):
            _, dt = cpu_predict_rate(pool, cores, pred_inputs, 50 + 100 * i)
            p_times.append(dt)
    total = float(np.sum(times))
    value = cores * iters * args.steps / total
    p_total = float(np.sum(p_times))
    p_units = cores * CPU_PRED_POINTS * CPU_PRED_DRAWS * len(p_times)
    p_value = p_units / p_total
    gibbs = {"metric": METRIC, "value": value, "unit": UNIT, "ms_per_step": 1e3 * total / args.steps,
             "higher_is_better": True, "dtype": "f64", "config": workload_config(args.gpus, "f64"),
             "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                              "sample": gibbs_sample_text(cores, iters)},
             "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    predict = {"metric": PRED_METRIC, "value": p_value, "unit": PRED_UNIT, "ms_per_step": 1e3 * p_total / len(p_times),
               "higher_is_better": True, "dtype": "f64", "config": pred_config(args.gpus),
               "cpu_baseline": {"value": p_value, "unit": PRED_UNIT, "cores": cores, "kind": "port",
                                "sample": predict_sample_text(cores, p_total / len(p_times))},
               "e2e": {"value": p_value, "unit": PRED_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    main_, other = (predict, gibbs) if args.metric == "predict" else (gibbs, predict)
    line = {"impl": "reference", **main_, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "scaling": "weak", "vs_baseline": None, "data": "synthetic", "gpu_launches": 0}
    if args.metric == "predict":
        line["gibbs"] = other
    else:
        line["predict"] = other
        line["f64"] = {k: gibbs[k] for k in ("metric", "value", "unit", "ms_per_step", "dtype", "e2e", "cpu_baseline")}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
class Bench:
    """Shared plumbing of the native arm: ranks, barriers, the timing loop."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: pybmc_b200 has no CPU path")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        from pybmc_b200 import _lib
        if not (os.path.exists(_lib.LIB_PATH) and os.path.exists(_lib.PROBE_LIB_PATH)):
            if self.rank == 0:      # snapshot without the built libraries: compile them, never fall back
                from pybmc_b200.build import build_library, build_probe_library
                build_library(force=not os.path.exists(_lib.LIB_PATH))
                build_probe_library()
            if self.world > 1:
                dist.barrier()
        self.lib = _lib.load()
        self.flush = torch.zeros(128 * 2 ** 20, dtype=torch.float32, device=self.dev)
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            self.peaks = {}
        try:
            self.constants = json.load(open(os.path.join(ROOT, "profiles", "kernel_constants.json")))
        except (OSError, ValueError):
            self.constants = {}
        self._pipe_peaks = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, step_fn, steps, warmup, robust=False):
        """K timed steps (CUDA events on the launching stream, L2 rewritten before each), max over ranks of the
        total.  The headline metrics use the plain mean of exactly K steps; the extras (robust=True) take the
        median step so that one disturbed step of a 3-step sample does not decide the number."""
        torch = self.torch
        for _ in range(warmup):
            step_fn()
        self.barrier()
        ms = []
        gc.collect()
        gc.disable()                    # no collector pause inside a timed step (the steps allocate a few tensors each)
        try:
            for _ in range(steps):
                self.flush.add_(1)          # 512 MiB read + write > 126 MB L2
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                step_fn()
                e1.record()
                e1.synchronize()
                ms.append(e0.elapsed_time(e1))
        finally:
            gc.enable()
        self.barrier()
        self.last_local_ms = float(np.median(ms)) if robust else float(np.sum(ms)) / steps
        if robust:
            return self.max_over_ranks(float(np.median(ms)))
        return self.max_over_ranks(float(np.sum(ms))) / steps

    def per_rank(self, info):
        """One small dict per rank, gathered on every rank (the spread between GPUs of one node: a step ends with
        collectives, so every rank's step time contains the slowest GPU's kernel)."""
        if self.world == 1:
            return None
        out = [None] * self.world
        self.dist.all_gather_object(out, info)
        return out

    def wall(self, step_fn, steps, warmup):
        """End-to-end steps by the host clock (host arrays in, host results out), max over ranks.  The warm-up
        results are kept alive together, so that PyTorch's caching host allocator ends up holding as many
        page-locked result buffers as the timed loop can have in flight (a fresh cudaHostAlloc costs ~1 ms per MB)."""
        keep = [step_fn() for _ in range(max(warmup, 3))]
        del keep
        self.barrier()
        gc.collect()
        gc.disable()
        try:
            t0 = time.perf_counter()
            for _ in range(steps):
                out = step_fn()
            self.torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        finally:
            gc.enable()
        s = self.max_over_ranks(dt) / steps
        self.barrier()
        return s, out

    def pipe_peaks(self):
        if self._pipe_peaks is None:
            self._pipe_peaks = measure_pipe_peaks(self.torch, self.dev)
        return self._pipe_peaks


def measure_pipe_peaks(torch, dev):
    """FP32 FMA, FP64-free integer and MUFU issue peaks of this GPU, measured now (libbmc_probe.so)."""
    from pybmc_b200 import _lib
    lib = _lib.load_probes()
    sink = torch.zeros(1, dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    blocks, threads, iters = sms * 8, 256, 20000
    out = {}
    for kind, name in ((0, "ffma"), (1, "mufu"), (2, "philox_mix"), (3, "issue"), (6, "imad"), (7, "lop3"),
                       (8, "ffma2"), (9, "imad_wide"), (11, "imad_wide_plus_2ffma"), (16, "dfma"),
                       (17, "dfma_three_registers")):
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
    out["dfma"]["TFLOP_per_s"] = out["dfma"]["Gwarp_inst_per_s"] * 32 * 2 / 1e3
    out["sms"] = sms
    return out


COUNTER_KEYS = (("fma_pipe_cycles_active_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                ("fma_pipe_cycles_active_pct_busiest_scheduler", "smsp__pipe_fma_cycles_active.max.pct_of_peak_sustained_active"),
                ("fmaheavy_cycles_active_pct", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active"),
                ("fmalite_cycles_active_pct", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active"),
                ("alu_pipe_cycles_active_pct", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
                ("fp64_pipe_cycles_active_pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                ("xu_pipe_inst_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                ("tensor_pipe_cycles_active_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                ("issue_active_pct_busiest_scheduler", "smsp__issue_active.max.pct_of_peak_sustained_active"),
                ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
                ("registers_per_thread", "launch__registers_per_thread"),
                ("ncu_duration_ms", "gpu__time_duration.sum"))


def counters_view(entry):
    """Hardware counters of the committed ncu capture of a kernel (profiles/kernel_constants.json)."""
    if not entry:
        return None
    c = entry.get("counters", {})
    out = {name: (c[key] * 1e3 if name == "ncu_duration_ms" else c[key]) for name, key in COUNTER_KEYS if key in c}
    out["stall_share_pct"] = entry.get("stall_share_pct")
    out["source"] = f"{entry.get('report')} ({entry.get('report_mtime')}): ncu --set full + pipe-cycle counters, cold clocks " \
                    "under the profiler: compare SHARES with the live numbers, not absolutes"
    return out


def sampler_roofline(b, key, kernel_name, rate_per_gpu, ms_step, kept_bytes_per_chain, real_bytes):
    """The sampler has no HBM or tensor traffic to speak of (state in registers, no GEMM): its bound is the FMA
    pipe (Philox 32x32->64 multiplies + fp32/fp64 arithmetic).  Live view: FMA-pipe instructions of one
    warp-iteration (opcode mix of the committed ncu capture) priced in FFMA issue slots by the probe rates of THIS
    run, against the FFMA rate of this run.  Counter view: what the hardware counted in the capture
    (sm__pipe_fma_cycles_active and friends) -- the corroboration the live model has to live with."""
    entry = b.constants.get(key) or {}
    pk = b.pipe_peaks()
    ffma = pk["ffma"]["Gwarp_inst_per_s"]
    out = {"kernel": kernel_name, "bound": "fma-pipe", "unit": "Gwarp-inst/s (FFMA issue slots)",
           "achieved": None, "peak": ffma, "frac": None,
           "peak_source": "FFMA rate measured in this run by libbmc_probe (independent streams, all SMs)"}
    mix = entry.get("fma_pipe_mix")
    if mix:
        price = {"fp32": 1.0, "fp32x2": ffma / pk["ffma2"]["Gwarp_inst_per_s"],
                 "imad": ffma / pk["imad"]["Gwarp_inst_per_s"], "imad_wide": ffma / pk["imad_wide"]["Gwarp_inst_per_s"]}
        slots = sum(float(mix[k]) * price[k] for k in price)
        fp64_slots = float(entry.get("fp64_per_unit") or 0.0)
        out.update({"achieved": slots * rate_per_gpu / 32 / 1e9, "frac": slots * rate_per_gpu / 32 / 1e9 / ffma,
                    "ffma_slots_per_warp_iteration": slots, "price_in_ffma_slots": price, "mix_per_warp_iteration": mix,
                    "fp64_instructions_per_warp_iteration": fp64_slots,
                    "mufu_per_warp_iteration": entry.get("mufu_per_unit"),
                    "warp_inst_per_warp_iteration": entry.get("warp_inst_per_warp_unit")})
        if entry.get("mufu_per_unit"):
            xu = float(entry["mufu_per_unit"]) * rate_per_gpu / 32 / 1e9
            out["xu"] = {"achieved": xu, "peak": pk["mufu"]["Gwarp_inst_per_s"], "frac": xu / pk["mufu"]["Gwarp_inst_per_s"],
                         "unit": "Gwarp-inst/s (MUFU)", "note": "second pipe above half load: 8 cycles per MUFU and scheduler"}
        inst = entry.get("warp_inst_per_warp_unit")
        if inst:
            issue_peak = max(pk["issue"]["Gwarp_inst_per_s"], ffma)
            out["issue"] = {"achieved": inst * rate_per_gpu / 32 / 1e9, "peak": issue_peak,
                            "frac": inst * rate_per_gpu / 32 / 1e9 / issue_peak, "unit": "Gwarp-inst/s"}
    if real_bytes == 8 and entry.get("fp64_per_unit"):
        # the fp64 kernels: the same construction on the fp64 pipe (DFMA / DMUL / DADD per warp-iteration of the
        # capture x the live rate, against the DFMA rate measured in this run); the FMA-pipe view (Philox) beside it
        dp, dpeak = float(entry["fp64_per_unit"]), pk["dfma"]["Gwarp_inst_per_s"]
        out["fma_pipe"] = {k: out[k] for k in ("achieved", "peak", "frac", "unit")}
        out.update({"bound": "fp64-pipe", "unit": "Gwarp-inst/s (fp64 instructions)", "achieved": dp * rate_per_gpu / 32 / 1e9,
                    "peak": dpeak, "frac": dp * rate_per_gpu / 32 / 1e9 / dpeak,
                    "peak_source": "DFMA rate measured in this run by libbmc_probe (independent streams, all SMs)"})
        # The quoted peak is DFMA with its multiplier and addend shared by every instruction (one per 2.05
        # scheduler-cycles).  With three DISTINCT register operands -- what the sampler's arithmetic mostly is -- B200
        # issues one DFMA per 3.03 cycles (register-operand bandwidth; profiles/probe_dfma.py): the realistic ceiling.
        d3 = pk.get("dfma_three_registers", {}).get("Gwarp_inst_per_s")
        if d3:
            out["fp64_three_register_operands"] = {
                "achieved": out["achieved"], "peak": d3, "frac": out["achieved"] / d3, "unit": out["unit"],
                "note": "same instruction count against the DFMA rate with three distinct register operands "
                        "(measured in this run): the ceiling real fp64 code has on this part"}
    out["counters"] = counters_view(entry)
    ckey = "fp64_pipe_cycles_active_pct" if out["bound"] == "fp64-pipe" else "fma_pipe_cycles_active_pct"
    if out["counters"] and ckey in out["counters"]:
        out["frac_by_counter"] = out["counters"][ckey] / 100.0
    out["traffic"] = entry.get("dram_bytes_per_launch")
    # what HAS to cross HBM: the kept draws and one copy of the moment rows (the 156 flushes of a launch are
    # reductions into rows that stay in L2: 28 MB)
    alg = CHAINS_PER_GPU * (KEEP_PER_CHAIN * kept_bytes_per_chain * real_bytes + 54 * 8)
    out["hbm"] = {"algorithmic_bytes_per_launch": alg, "achieved": alg / (ms_step * 1e-3) / 1e9,
                  "peak": float(b.peaks.get("hbm_gbs", 6650.0)), "unit": "GB/s",
                  "frac": alg / (ms_step * 1e-3) / 1e9 / float(b.peaks.get("hbm_gbs", 6650.0)),
                  "note": "kept samples + one copy of the fp64 moment rows; not the bound (SURVEY.md section 8d)"}
    out["pipe_peaks_measured"] = pk
    return out


class SamplerWorkload:
    """BASELINE configs[2] through pybmc_b200.parallel.sharded_gibbs."""

    def __init__(self, b):
        import pybmc_b200 as pb
        from pybmc_b200.inference_utils import ConjugateSampler
        self.b = b
        preds, truth = config3_ensemble()
        orth = pb.orthogonalize_arrays(preds, truth, 8, device=b.dev)
        self.y, self.X = np.ascontiguousarray(orth["y"]), np.ascontiguousarray(orth["U_hat"])
        self.prior = [np.zeros(8), np.diag(orth["S_hat"] ** 2), 1.0, 0.02]
        self.sampler = ConjugateSampler(self.y, self.X, self.prior, device=b.dev)
        self.thin = ITERATIONS // KEEP_PER_CHAIN

    def step_resident(self, dtype, hist=HIST_EVERY, iterations=ITERATIONS):
        """One step with the problem resident in HBM: this rank's chains, the all-reduces, device results."""
        from pybmc_b200 import parallel as par
        return par.sharded_gibbs(None, None, iterations, self.prior, CHAINS_PER_GPU * self.b.world, seed=SEED,
                                 dtype=dtype, thin=iterations // KEEP_PER_CHAIN, keep_samples=True, hist_every=hist,
                                 as_numpy=False, sampler=self.sampler, device=self.b.dev)

    def step_e2e(self, dtype):
        """The public call from host arrays to host results (set-up, sufficient statistics, sampler, all-reduces,
        D2H of moments, histograms and kept samples)."""
        from pybmc_b200 import parallel as par
        return par.sharded_gibbs(self.y, self.X, ITERATIONS, self.prior, CHAINS_PER_GPU * self.b.world, seed=SEED,
                                 dtype=dtype, thin=self.thin, keep_samples=True, hist_every=HIST_EVERY,
                                 device=self.b.dev)

    def measure(self, dtype, steps, warmup, with_clocks=False):
        b = self.b
        torch_dtype = {"f32": "float32", "f64": "float64"}[dtype]
        clocks = ClockSampler(b.local).start() if (with_clocks and b.rank == 0) else None
        ms_step = b.timed(lambda: self.step_resident(torch_dtype), steps, warmup)
        clock_info = clocks.stop() if clocks else None
        ranks = None
        if with_clocks and b.world > 1:
            # the kernel alone on every GPU (no collective inside the events): which GPU sets the pace, and at what clock
            ks = ClockSampler(b.local).start()
            k_ms = b.timed(lambda: self.sampler.run(ITERATIONS, CHAINS_PER_GPU, SEED, torch_dtype, self.thin, 0, True,
                                                    "full", b.rank * CHAINS_PER_GPU, None, HIST_EVERY), steps, 1)
            kc = ks.stop()
            ranks = b.per_rank({"rank": b.rank, "step_ms": round(float(np.float64(b.last_local_ms)), 3),
                                "kernel_only_ms": None, "sm_mhz": kc.get("sm_mhz"), "power_w_max": kc.get("power_w_max"),
                                "reasons": kc.get("reasons")})
            for r in ranks:
                r["kernel_only_ms"] = r.pop("step_ms")
            del k_ms
        units = CHAINS_PER_GPU * b.world * ITERATIONS
        value = units / (ms_step * 1e-3)
        e2e_s, out = b.wall(lambda: self.step_e2e(torch_dtype), steps, max(3, warmup))
        mean, cov, local = out
        real_bytes = 4 if dtype == "f32" else 8
        h2d = self.y.nbytes + self.X.nbytes + 8 * (8 + 6 * 8 + 2)          # design + OLS vector + problem constants
        d2h = local.samples.nbytes + 8 * (55 + 9 * 9 + 1) + 8 * 9 * 512
        kernel = f"gibbs_conjugate_kernel<{'float' if dtype == 'f32' else 'double'}, 8, 2, true>"
        block = {"metric": METRIC, "value": value, "unit": UNIT, "ms_per_step": ms_step, "dtype": dtype,
                 "higher_is_better": True, "config": workload_config(b.world, dtype),
                 "e2e": {"value": units / e2e_s, "unit": UNIT, "ms_per_step": 1e3 * e2e_s,
                         "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                         "api": "pybmc_b200.parallel.sharded_gibbs(y, X, ...) from host NumPy arrays"},
                 "gpu_launches": steps * 1 + steps * 5,
                 "gpu_launches_detail": "value region: 1 kernel per step (gibbs_conjugate_kernel); e2e region: 5 per "
                                        "step (gram_partial, sum_partials, rss_partial, sum_partials, "
                                        "gibbs_conjugate_kernel); the reduction of the moment rows over chains and the "
                                        "sample transpose are torch ops",
                 "posterior_check": {"sigma_mean": float(mean[-1]), "b0_mean": float(mean[0]),
                                     "sigma_q50_from_histogram": float(local.quantiles([50.0])[0, -1])}}
        key = "gibbs_conjugate_f32_k8_full" if dtype == "f32" else "gibbs_conjugate_f64_k8_full"
        block["roofline"] = sampler_roofline(b, key, kernel, value / b.world, ms_step, 9, real_bytes)
        if clock_info is not None:
            block["clocks"] = clock_info
        if ranks is not None:
            block["per_rank"] = ranks
        return block


class PredictWorkload:
    """BASELINE configs[3]: fused prediction, nuclei sharded over ranks."""

    def __init__(self, b):
        from pybmc_b200 import _lib
        from pybmc_b200 import parallel as par
        from pybmc_b200.sampling_utils import PredictiveProblem
        self.b = b
        self.lo, self.hi = par.point_range(PRED_POINTS, b.rank, b.world)
        self.preds, self.vt, self.theta, self.truth = config4_inputs(PRED_POINTS, PRED_DRAWS, PRED_K)
        self.prob = PredictiveProblem(self.preds[self.lo:self.hi], self.theta, self.vt, truth=self.truth[self.lo:self.hi],
                                      dtype="float32", device=b.dev, point0=self.lo)
        nbytes = int(b.lib.bmc_predict_workspace_bytes(_lib.F32, self.hi - self.lo, len(PRED_Q), PRED_DRAWS))
        self.ws = b.torch.empty(nbytes, dtype=b.torch.uint8, device=b.dev)
        self.p_h = np.ascontiguousarray(self.preds[self.lo:self.hi])
        self.t_h = np.ascontiguousarray(self.truth[self.lo:self.hi])

    def step_resident(self):
        self.last = self.prob.run(percentiles=PRED_Q, seed=SEED, as_numpy=False, workspace=self.ws)

    def step_e2e(self):
        """Host arrays in (this rank's block of predictions and truth; the posterior rows -- 1/N uploaded per rank,
        all-gathered), host results out: rank 0 ends with the full-length outputs (one reader on the host link instead of N)."""
        from pybmc_b200 import parallel as par
        return par.sharded_predictive_summary(self.p_h, self.theta, self.vt, truth=self.t_h, percentiles=PRED_Q,
                                              seed=SEED, dtype="float32", device=self.b.dev,
                                              n_points_total=PRED_POINTS, gather="root")

    def measure(self, steps, warmup):
        b = self.b
        ms = b.timed(self.step_resident, steps, warmup)
        units = PRED_POINTS * PRED_DRAWS
        value = units / (ms * 1e-3)
        e2e_s, res = b.wall(self.step_e2e, steps, max(2, warmup))
        assert b.rank != 0 or res.mean.shape[0] == PRED_POINTS
        h2d = self.p_h.nbytes + self.t_h.nbytes + self.vt.nbytes + -(-self.theta.shape[0] // b.world) * self.theta.shape[1] * 8
        block = {"metric": PRED_METRIC, "value": value, "unit": PRED_UNIT, "ms_per_step": ms, "dtype": "f32",
                 "higher_is_better": True, "passes": self.last.passes, "config": pred_config(b.world),
                 "e2e": {"value": units / e2e_s, "unit": PRED_UNIT, "ms_per_step": 1e3 * e2e_s,
                         "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(PRED_POINTS * 8 * (4 + len(PRED_Q))),
                         "collectives": "all_gather(theta: each rank uploads 1/N of the 13.6 MB fp64 rows) + gather(9 x nuclei fp64) to rank 0, which alone reads the full-length results back" if b.world > 1
                         else "none (1 GPU)",
                         "api": "pybmc_b200.parallel.sharded_predictive_summary from host NumPy arrays"},
                 "roofline": self.roofline(value / b.world)}
        return block

    def roofline(self, units_per_s_per_gpu):
        """The fused pass is bound by its consumer loop (Philox, Box-Muller, compares, predicated counts): issue
        slots.  Live: thread-instructions per sample x point of the committed capture x the live rate against the
        issue peak measured in this run; counters of the capture beside it."""
        b = self.b
        entry = b.constants.get("predict_pass_tc_f32_k16_q5") or {}
        c = entry.get("counters", {})
        pk = b.pipe_peaks()
        issue_peak = max(pk["issue"]["Gwarp_inst_per_s"], pk["ffma"]["Gwarp_inst_per_s"])
        out = {"kernel": "predict_pass_tc_kernel<16,5> (+ predict_select_kernel)", "bound": "issue", "unit": "Gwarp-inst/s",
               "achieved": None, "peak": issue_peak, "frac": None,
               "note": "whole step (pass + select + window set-up) against the pass kernel's instruction count"}
        per_unit = entry.get("thread_inst_per_unit")
        if per_unit is None and entry.get("warp_inst_per_warp_unit"):
            per_unit = entry["warp_inst_per_warp_unit"]          # all lanes active: thread-inst per unit = warp-inst per 32
        if per_unit:
            ach = per_unit / 32.0 * units_per_s_per_gpu / 1e9
            out.update({"achieved": ach, "frac": ach / issue_peak, "thread_inst_per_sample_point": per_unit})
        out["tensor"] = {"kind": "tcgen05.mma kind::tf32, split TF32 (3 products)",
                         "TFLOP_per_s": 2.0 * PRED_K * 3 * units_per_s_per_gpu / 1e12,
                         "note": "K=16: the MMAs keep the tensor pipe a few % busy (ncu); the 16 consumer warps are the bound"}
        out["counters"] = counters_view(entry)
        if out["counters"] and "issue_active_pct" in out["counters"]:
            out["frac_by_counter"] = out["counters"]["issue_active_pct"] / 100.0
        out["traffic"] = entry.get("dram_bytes_per_launch")
        out["algorithmic_bytes_per_step"] = 8 * (PRED_DRAWS * (PRED_K + 1) + PRED_POINTS * (PRED_K + 2)) + 40 * PRED_POINTS
        sel = b.constants.get("predict_select_f32")
        if sel:
            out["select_kernel_counters"] = counters_view(sel)
        return out


def run_native(args):
    b = Bench(args)
    torch, world, rank = b.torch, b.world, b.rank
    line = {}
    want_gibbs = args.metric == "gibbs" or args.only == "all"
    want_predict = args.metric == "predict" or args.only == "all"
    if args.only == "sampler":
        want_predict = False
    if args.only == "predict":
        want_gibbs = False

    sw = SamplerWorkload(b) if want_gibbs else None
    gibbs = {}
    if sw is not None:
        order = ["f32", "f64"] if args.dtype == "f32" else ["f64", "f32"]
        if args.only == "sampler":
            order = order[:1]
        for i, dt in enumerate(order):
            gibbs[dt] = sw.measure(dt, args.steps, args.warmup, with_clocks=(i == 0 and args.metric == "gibbs"))
    predict = None
    if want_predict:
        pw = PredictWorkload(b)
        clocks = ClockSampler(b.local).start() if (args.metric == "predict" and rank == 0) else None
        predict = pw.measure(args.steps if args.metric == "predict" else min(args.steps, 3), max(args.warmup, 2)
                             if args.metric == "predict" else 2)
        if clocks:
            predict["clocks"] = clocks.stop()
        del pw
        torch.cuda.empty_cache()

    common = {"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "scaling": "weak", "vs_baseline": None,
              "data": "synthetic"}
    if args.metric == "predict":
        line = {**predict, **common}
        if gibbs:
            line["gibbs"] = gibbs.get("f32")
            line["f64"] = gibbs.get("f64")
        line["scaling"] = "strong"           # 1e5 nuclei in total, split over the ranks
    else:
        head = gibbs[args.dtype]
        line = {**head, **common}
        other = "f64" if args.dtype == "f32" else "f32"
        if other in gibbs:
            line[other] = gibbs[other]
        if predict is not None:
            line["predict"] = predict

    if args.only == "all" and not args.no_extras:
        line["extra"] = extras(b, sw)
    if rank == 0 and world == 1 and not args.no_cpu:
        cg, cp = cpu_baselines(want_gibbs, want_predict)
        if cg:
            (line if args.metric == "gibbs" else line.get("gibbs", {}))["cpu_baseline"] = cg
            if "f64" in line and isinstance(line["f64"], dict):
                line["f64"]["cpu_baseline"] = cg
            if "f32" in line and isinstance(line["f32"], dict):
                line["f32"]["cpu_baseline"] = cg
        if cp:
            (line if args.metric == "predict" else line.get("predict", {}))["cpu_baseline"] = cp
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        b.dist.destroy_process_group()


def cpu_baselines(want_gibbs, want_predict):
    """Reference algorithm (oracle port) on the host cores, bounded samples of the same workloads."""
    import multiprocessing as mp
    os.environ["OMP_NUM_THREADS"] = "1"
    cores = os.cpu_count() or 1
    cg = cp = None
    with mp.get_context("fork").Pool(cores) as pool:
        if want_gibbs:
            problem = cpu_problem()
            iters = 60000
            cpu_sampler_rate(100, pool, cores, problem)
            rate, dt = cpu_sampler_rate(iters, pool, cores, problem)
            cg = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "seconds": dt,
                  "sample": gibbs_sample_text(cores, iters)}
        if want_predict:
            inputs = config4_inputs(CPU_PRED_POINTS, CPU_PRED_DRAWS + 2000)
            rate, dt = cpu_predict_rate(pool, cores, inputs)
            cp = {"value": rate, "unit": PRED_UNIT, "cores": cores, "kind": "port", "seconds": dt,
                  "sample": predict_sample_text(cores, dt)}
    return cg, cp


# ------------------------------------------------------------------------------------------------
# extras: the other configurations of BASELINE.json and the pieces around the two metrics
# ------------------------------------------------------------------------------------------------
def extras(b, sw):
    out = {}
    steps = 3
    timed = lambda fn, k=steps, w=2: b.timed(fn, k, w, robust=True)   # noqa: E731
    if sw is not None:
        extras_samplers(out, b, sw, timed)
    extras_coverage(out, b)
    extras_config4(out, b)
    if b.rank == 0:
        extras_config0(out, b)
    b.barrier()
    return out


def extras_samplers(out, b, sw, timed):
    import pybmc_b200 as pb
    from pybmc_b200 import _lib
    from pybmc_b200.inference_utils import SimplexSampler
    torch, dev, world, rank = b.torch, b.dev, b.world, b.rank
    # what the marginal histograms cost: the same step without them
    for dt, name in (("float32", "f32"), ("float64", "f64")):
        ms = timed(lambda: sw.step_resident(dt, hist=0))
        out[f"gibbs_{name}_without_histograms"] = {"value": CHAINS_PER_GPU * world * ITERATIONS / (ms * 1e-3), "unit": UNIT,
                                                   "ms_per_step": ms}
    # simplex sampler, BASELINE configs[1]: nuclear-mass surrogate, 4096 chains per GPU
    preds, truth = config1_ensemble()
    idx = np.random.default_rng(1).permutation(len(truth))[:377]
    o = pb.orthogonalize_arrays(preds[idx], truth[idx], 3, device=dev)
    simplex = SimplexSampler(o["y"], o["U_hat"], o["Vt_hat"], o["S_hat"], [1.0, 0.02], 0.001, device=dev)
    burn, iters, chains = 10000, 50000, 4096
    from pybmc_b200 import parallel as par

    def simplex_step(dt):      # this rank's chains + the all-reduce of moment sums, count and accepted proposals
        return par.sharded_gibbs_simplex(None, None, None, None, iters, [1.0, 0.02], chains * world, burn=burn,
                                         stepsize=0.001, seed=SEED, dtype=dt, thin=iters // 10, keep_samples=True,
                                         as_numpy=False, sampler=simplex, device=dev)
    ms = timed(lambda: simplex_step("float32"))
    acc_rate = simplex_step("float32")[2]
    out["simplex_f32"] = {"value": chains * world * (burn + iters) / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                          "acceptance_rate": acc_rate, "api": "pybmc_b200.parallel.sharded_gibbs_simplex (device-resident problem)",
                          "config": "configs[1] surrogate: 377 x 15, K=3, 4096 chains/GPU x (10000 burn + 50000)",
                          "counters": counters_view(b.constants.get("gibbs_simplex_group16_f32_k4"))}
    ms = timed(lambda: simplex_step("float64"))
    out["simplex_f64"] = {"value": chains * world * (burn + iters) / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                          "config": "the same in fp64"}
    # the literal one-chain-per-warp kernel (parity anchor): redoes the O(nK) residual every iteration
    s = sw.sampler
    n, k = s.n, s.k
    xt = s._Xd.t().contiguous().to(torch.float32)
    yr = s._yd.to(torch.float32)
    consts = torch.from_numpy(np.concatenate([s.lam.reshape(-1), s.lam @ s.b0])).to(dev)
    lit_chains, lit_iters = 8192, 200
    buf = torch.empty((lit_iters, k + 1, lit_chains), dtype=torch.float32, device=dev)

    def lit_step():
        _lib.check(b.lib.bmc_gibbs_literal_run(_lib.F32, xt.data_ptr(), yr.data_ptr(), n, k, consts.data_ptr(),
                                               consts.data_ptr() + 8 * k * k, s.nu0, s.sigma20, s.sigma2_init, SEED,
                                               rank * lit_chains, lit_chains, lit_iters, buf.data_ptr(),
                                               torch.cuda.current_stream(dev).cuda_stream))
    ms = timed(lit_step)
    rate = lit_chains * world * lit_iters / (ms * 1e-3)
    out["literal_f32"] = {"value": rate, "unit": UNIT, "ms_per_step": ms,
                          "reference_equivalent_tflops": rate * (4 * n * k + 3 * n) / 1e12,
                          "config": "same problem, one chain per warp, X in shared memory via TMA, residual over all "
                                    "3000 rows each iteration; 8192 chains x 200 iterations per GPU"}


def extras_coverage(out, b):
    """An HBM-bound kernel of the path: order counts of a materialised matrix (coverage())."""
    from pybmc_b200 import _lib
    torch, dev = b.torch, b.dev
    if b.rank != 0:
        return
    s_rows, n_cols = 10000, 65536
    mat = torch.randn((s_rows, n_cols), dtype=torch.float64, device=dev)
    tr = torch.zeros(n_cols, dtype=torch.float64, device=dev)
    c1 = torch.empty(n_cols, dtype=torch.int64, device=dev)
    c2 = torch.empty(n_cols, dtype=torch.int64, device=dev)

    def cov_step():
        _lib.check(b.lib.bmc_coverage_counts(mat.data_ptr(), s_rows, n_cols, n_cols, tr.data_ptr(), c1.data_ptr(),
                                             c2.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    for _ in range(2):
        cov_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.flush.add_(1)
    e0.record(); cov_step(); e1.record(); e1.synchronize()
    hbm_peak = float(b.peaks.get("hbm_gbs", 6650.0))
    gbs = mat.numel() * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    out["coverage_counts_hbm"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                  "frac": gbs / hbm_peak, "config": "10000 x 65536 fp64 matrix read once",
                                  "peak_source": "MEASURED_PEAKS.json hbm_gbs" if b.peaks else "fallback 6650 GB/s"}
    del mat


def extras_config4(out, b):
    """BASELINE configs[4] END TO END through pybmc_b200.parallel: 256 models x 1e5 points (rows sharded over the
    ranks), K = 64: sharded_orthogonalize (all-reduce of the 256 x 256 Gram matrix) -> sharded_gibbs(rows_sharded)
    at K = 64 (all-reduce of X'X, X'y, y'y, n, RSS_min; then of the moment sums) -> sharded_predictive_summary on
    the same 1e5 points with 1e4 posterior draws (broadcast + all-gather).  Host arrays in, host results out."""
    from pybmc_b200 import parallel as par
    torch, world, rank, dev = b.torch, b.world, b.rank, b.dev
    n_total, m, k = 100_000, 256, 64
    lo, hi = par.point_range(n_total, rank, world)
    preds, truth = config5_rows(lo, hi, m, k)
    chains_total, iters, keep = 16384 * world, 1000, 10
    n_draws = 10000

    def wall(fn):
        b.barrier()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        return b.max_over_ranks(time.perf_counter() - t0) * 1e3, r

    res, reps = {}, []
    for rep in range(4):            # first pass: warm allocator, NCCL channels and cuSOLVER handles; then three timed
        ms_o, orth = wall(lambda: par.sharded_orthogonalize(preds, truth, k, device=dev))
        prior = [np.zeros(k), np.diag(orth["S_hat"] ** 2), 1.0, 0.02]
        ms_g, (mean, cov, local) = wall(lambda: par.sharded_gibbs(
            orth["y"], orth["U_hat"], iters, prior, chains_total, seed=SEED + 5, dtype="float32", thin=iters // keep,
            keep_samples=True, device=dev, rows_sharded=True))
        # 1e4 posterior rows for the prediction: rank r contributes rows [r S/N, (r+1) S/N) of ITS kept draws; the
        # all-gather inside sharded_predictive_summary makes the table the same on every rank
        theta = np.ascontiguousarray(local.samples[:n_draws]).astype(np.float64)
        ms_p, pred = wall(lambda: par.sharded_predictive_summary(
            preds, theta, orth["Vt_hat"], truth=truth, percentiles=[2.5, 50.0, 97.5], seed=SEED + 6, dtype="float32",
            device=dev, n_points_total=n_total))
        if rep:
            reps.append((ms_o, ms_g, ms_p))
    # median of the three timed passes per stage (one wall-clock sample per stage was at the mercy of a single
    # page-locked allocation or a stray retry pass: 12 ms vs 73 ms for the same prediction on two GPUs)
    ms_o, ms_g, ms_p = (float(np.median([r[i] for r in reps])) for i in range(3))
    res = {"orthogonalize_ms": ms_o, "gibbs_ms": ms_g, "predict_ms": ms_p, "total_ms": ms_o + ms_g + ms_p,
           "passes_ms": [[round(v, 2) for v in r] for r in reps]}
    from pybmc_b200.sampling_utils import coverage_from_counts
    cover = coverage_from_counts([68, 95], n_draws, pred.c_lt, pred.c_le, device=dev) if rank == 0 else None
    out["config4_end_to_end"] = {
        "config": f"configs[4]: 1e5 x 256 fp64 table, rows over {world} GPU(s) ({hi - lo} here), K=64; sampler "
                  f"{chains_total} chains x {iters} iterations fp32 (marginal moments); prediction 1e5 nuclei x {n_draws} "
                  "draws x K=64, 3 percentiles + coverage counts",
        **res, "method": orth["method"],
        "gibbs_chain_iters_per_sec": chains_total * iters / (res["gibbs_ms"] * 1e-3),
        "predict_samples_x_points_per_sec": float(n_total) * n_draws / (res["predict_ms"] * 1e-3),
        "collectives": "all_reduce(Gram 256x256 fp64 = 512 KB); all_reduce([X|y]'[X|y] 65x65, n, RSS_min); "
                       "all_reduce(moment sums); all_gather(theta 5.2 MB, 1/N uploaded per rank); all_gather(7 x nuclei fp64)" if world > 1
        else "none (1 GPU)",
        "check": {"sigma_posterior_mean": float(mean[-1]), "S_hat_first_last": [float(orth["S_hat"][0]), float(orth["S_hat"][-1])],
                  "coverage_68_95_pct": cover}}
    # device-only stage times of the same problem (resident inputs), for the share of the collectives / copies
    from pybmc_b200.inference_utils import ConjugateSampler
    from pybmc_b200 import _lib
    smp = ConjugateSampler(orth["y"], orth["U_hat"], prior, device=dev, reduce=par.sum_over_ranks())
    per = chains_total // world
    ms = b.timed(lambda: smp.run(iters, per, SEED + 5, "float32", iters // keep, 0, True, "diag", rank * per), 3, 2, robust=True)
    out["config4_end_to_end"]["gibbs_kernel_only"] = {
        "ms": ms, "chain_iters_per_sec": chains_total * iters / (ms * 1e-3),
        "kernel": "gibbs_conjugate_kernel<float, 64, 1>", "counters": counters_view(b.constants.get("gibbs_conjugate_f32_k64_diag"))}
    ms64 = b.timed(lambda: smp.run(iters // 4, per, SEED + 5, "float64", iters // keep, 0, False, "diag", rank * per), 3, 1,
                   robust=True)
    out["config4_end_to_end"]["gibbs_kernel_only_f64"] = {"ms": ms64, "chain_iters_per_sec": chains_total * (iters // 4) / (ms64 * 1e-3)}
    del smp, preds


def extras_config0(out, b):
    """BASELINE configs[0], the reference's own workflow through the drop-in class (surrogate for the absent
    selected_data.h5: 629 nuclei x 15 models, 377 training rows, K = 3): wall-clock of each public call,
    second call of a process (the first also pays one-time CUDA / cuSOLVER initialisation)."""
    import contextlib
    import io
    import pandas as pd
    import pybmc_b200 as pb
    torch, dev = b.torch, b.dev
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--metric", default="gibbs", choices=["gibbs", "predict"],
                    help="which of BASELINE.json's two metrics is the top-level line (the other is a sibling object)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"],
                    help="arithmetic type of the top-level gibbs line (the other type is a sibling object)")
    ap.add_argument("--only", default="all", choices=["all", "sampler", "predict"],
                    help="profiling aid: restrict the run to one kernel family")
    ap.add_argument("--no-extras", action="store_true", help="skip the `extra` block")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    args = ap.parse_args()
    if args.only == "predict":
        args.metric = "predict"
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_native(args)


if __name__ == "__main__":
    main()
