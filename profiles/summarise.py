"""Summarise ncu reports brought back in gpurun_out/ (run here, no GPU needed).

    python profiles/summarise.py gpurun_out/prof_gibbs.ncu-rep [more.ncu-rep ...]
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]


EXTRA = ["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
         "smsp__pipe_fma_cycles_active.max.pct_of_peak_sustained_active",
         "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
         "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active",
         "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
         "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
         "smsp__issue_active.max.pct_of_peak_sustained_active"]
KEYS = KEYS + EXTRA


def main(paths):
    for path in paths:
        if path.endswith(".csv"):        # the export profiles/ncu_capture.sh made on the GPU box
            raw = open(path).read()
        else:
            raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, units))
            print(f"## {d.get('Kernel Name')}  ({path})")
            for k in KEYS:
                if k in d and d[k] not in ("", "n/a"):
                    print(f"  {k} = {d[k]} {u.get(k, '')}")
            stalls = [(k.split("stalled_")[1], float(v.replace(",", ""))) for k, v in d.items()
                      if "pcsamp_warps_issue_stalled" in k and v not in ("", "n/a") and not k.endswith("not_issued")]
            tot = sum(v for _, v in stalls) or 1.0
            print("  stalls: " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in sorted(stalls, key=lambda kv: -kv[1])[:7]))


if __name__ == "__main__":
    main(sys.argv[1:])
