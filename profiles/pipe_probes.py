"""Instruction-class throughput on this GPU (bmc_probe): single streams and the mixes that tell which classes
share a pipe.  Prints warp-instructions per clock and scheduler (148 SMs x 4 schedulers, SM clock from nvidia-smi
max; all streams independent, 8 chains per thread, 8 x 256-thread blocks per SM)."""
import subprocess
import sys
import torch
sys.path.insert(0, "/root/repo")
from pybmc_b200 import _lib

lib = _lib.load_probes()
names = {0: "FFMA", 1: "MUFU (ex2, lg2)", 2: "IMAD.WIDE + LOP3", 3: "FFMA + LOP3", 4: "IMAD.WIDE + IADD", 5: "IMAD.HI",
         6: "IMAD", 7: "LOP3", 8: "FFMA2", 9: "IMAD.WIDE", 10: "IMAD.WIDE + FFMA", 11: "IMAD.WIDE + 2 FFMA",
         12: "IMAD.WIDE + FFMA2", 13: "MUFU + 4 FFMA", 14: "MUFU + 2 IMAD.WIDE", 15: "FFMA2 + FFMA"}
try:
    mhz = float(subprocess.run(["nvidia-smi", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits", "-i", "0"],
                               capture_output=True, text=True).stdout.split()[0])
except Exception:
    mhz = 1965.0
sink = torch.zeros(1, dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
sms = torch.cuda.get_device_properties(0).multi_processor_count
blocks, threads, iters = sms * 8, 256, 20000
print(f"SM clock {mhz:.0f} MHz")
for kind in range(16):
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
    rate = blocks * threads * iters * ops / 32 / (best * 1e-3)
    per_clk = rate / (sms * 4 * mhz * 1e6)
    print(f"{names[kind]:22s} {rate / 1e9:8.1f} Gwarp-inst/s  {per_clk:5.3f} inst/clk/scheduler  "
          f"{ops / 8 / per_clk:5.2f} cycles per group")
