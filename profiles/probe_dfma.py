"""fp64 pipe probes: DFMA with shared operands (the peak bench.py quotes), with three distinct register operands, and
with a LOP3 beside each; at 8, 3 and 2 warps per scheduler.  usage: python profiles/probe_dfma.py"""
import sys
import torch
sys.path.insert(0, "/root/repo")
from pybmc_b200 import _lib
lib = _lib.load_probes()
dev = torch.device("cuda", 0)
sink = torch.zeros(1, dtype=torch.float32, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
sms = torch.cuda.get_device_properties(dev).multi_processor_count
for kind, name in ((16, "DFMA a = fma(a, m, c), m and c shared"), (17, "DFMA, three distinct registers"), (18, "DFMA + LOP3"),
                   (0, "FFMA a = fma(a, m, c), m and c shared"), (19, "FFMA, three distinct registers"),
                   (8, "FFMA2, m and c shared"), (20, "FFMA2, three distinct register pairs")):
    for blocks_per_sm, threads in ((8, 256), (3, 128), (2, 128), (1, 128)):
        ops = lib.bmc_probe_ops_per_iteration(kind)
        iters = 20000
        best = None
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.bmc_probe(kind, iters, sms * blocks_per_sm, threads, sink.data_ptr(), st))
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        warps_per_sched = blocks_per_sm * threads / 32 / 4
        rate = sms * blocks_per_sm * threads * iters * ops / 32 / (best * 1e-3) / 1e9
        cyc = sms * 4 * 1.965e9 / (rate * 1e9)
        print(f"{name:45s} {warps_per_sched:4.1f} warps/scheduler: {rate:7.1f} Gwarp-inst/s = one per {cyc:.2f} scheduler-cycles")
