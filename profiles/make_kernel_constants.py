"""Regenerate profiles/kernel_constants.json FROM the .ncu-rep files of a capture (run here, no GPU needed).

    python profiles/make_kernel_constants.py [--dir gpurun_out] [--tag r2b,r2c]

bench.py divides by nothing in this file; it QUOTES it: the pipe counters that corroborate (or correct) the
live roofline of each kernel, the per-unit instruction counts, the opcode mix and the DRAM traffic, each with
the report it came from.  Because the file is rewritten from the reports (never edited by hand), it cannot go
stale silently: every entry carries the report's name, size and modification time, and the git revision of
the tree at regeneration time.  profiles/ncu_capture.sh is the gpurun command that produces the reports.

Each capture is described in CAPTURES: (key, report stem, kernel-name regex, work units per launch).
"""
import argparse
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# key -> (report stem, kernel regex, units per launch, unit name)
CAPTURES = {
    "gibbs_conjugate_f32_k8_full": ("gibbs_f32", r"gibbs_conjugate_kernel<float, ?8, ?2", 65536 * 10000, "chain-iteration"),
    "gibbs_conjugate_f64_k8_full": ("gibbs_f64", r"gibbs_conjugate_kernel<double, ?8, ?2", 65536 * 10000, "chain-iteration"),
    "gibbs_conjugate_f32_k64_diag": ("gibbs_k64", r"gibbs_conjugate_kernel<float, ?64, ?1", 16384 * 1000, "chain-iteration"),
    "gibbs_simplex_group16_f32_k4": ("simplex_f32", r"gibbs_simplex_group16_kernel<float", 4096 * 6000, "chain-iteration"),
    # one launch = one chunk of 25,088 nuclei (1e5 in four equal chunks, rounded up to 256) x all 1e5 draws; the
    # launch with the largest grid is a full chunk
    "predict_pass_tc_f32_k16_q5": ("predict_f32", r"predict_pass_tc_kernel<16, ?5", 25088 * 100000, "sample x point"),
    "predict_select_f32": ("predict_f32", r"predict_select_kernel<float", None, "nucleus x percentile"),
}

COUNTERS = [
    "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.max.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__pipe_fma_cycles_active.max.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_active.avg", "smsp__cycles_active.avg",
]
# what ncu_capture.sh adds to --set full (not part of any section on gb100)
EXTRA_METRICS = [m for m in COUNTERS if "_cycles_active." in m and "tensor" not in m and "sm__cycles" not in m
                 and "smsp__cycles" not in m] + [
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.max.pct_of_peak_sustained_active", "smsp__thread_inst_executed.sum"]

UNIT_SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
              "second": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}


def _num(v):
    try:
        return float(v.replace(",", ""))
    except (ValueError, AttributeError):
        return None


def _page(path, page):
    """CSV text of a report page: from the .ncu-rep, or from the export ncu_capture.sh made on the GPU box
    (<stem>_raw.csv / <stem>_source.csv.gz) when the report itself was too large to bring back."""
    if os.path.exists(path):
        return subprocess.run(["ncu", "-i", path, "--page", page, "--csv"], capture_output=True, text=True).stdout
    stem = path[:-len(".ncu-rep")]
    if page == "raw" and os.path.exists(stem + "_raw.csv"):
        return open(stem + "_raw.csv").read()
    if page == "source" and os.path.exists(stem + "_source.csv.gz"):
        import gzip
        return gzip.open(stem + "_source.csv.gz", "rt").read()
    return ""


def raw_rows(path):
    out = _page(path, "raw")
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    return [({h: v for h, v in zip(hdr, vals)}, {h: u for h, u in zip(hdr, units)}) for vals in rows[2:]]


def opcode_mix(path, kernel_re):
    """Executed warp-instructions by opcode from the source page (last launch whose kernel matches)."""
    out = _page(path, "source")
    rows = list(csv.reader(io.StringIO(out)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": re.sub(r"\(int\)|\(bool\)|bmc::", "", r[1] if len(r) > 1 else ""), "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None and "Source" in r and "Instructions Executed" in r:
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] is not None:
            cur["rows"].append(r)
    blocks = [b for b in blocks if b["hdr"] and re.search(kernel_re, b["name"])]
    if not blocks:
        return None
    b = blocks[-1]
    src, ex = b["hdr"].index("Source"), b["hdr"].index("Instructions Executed")
    tot = collections.Counter()
    for r in b["rows"]:
        if len(r) <= max(src, ex):
            continue
        e = _num(r[ex])
        parts = r[src].split()
        if e is None or not parts:
            continue
        op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
        tot[op.rstrip(";")] += e
    return tot


def summarise(key, stem, kernel_re, units, unit_name, directory, tag):
    path = os.path.join(directory, f"{tag}_{stem}.ncu-rep")
    evidence = path if os.path.exists(path) else path[:-len(".ncu-rep")] + "_raw.csv"
    if not os.path.exists(evidence):
        return None
    rows = [(d, u) for d, u in raw_rows(path) if re.search(kernel_re, d.get("Kernel Name", ""))]
    if not rows:
        return None
    # the last captured launch (warm); for chunked kernels the one with the largest grid (a full chunk)
    d, u = max(reversed(rows), key=lambda du: _num(du[0].get("launch__grid_size", "0")) or 0.0)
    counters = {}
    for name in COUNTERS:
        v = _num(d.get(name, ""))
        if v is None:
            continue
        counters[name] = v * UNIT_SCALE.get(u.get(name, ""), 1.0) if name.endswith(".sum") and u.get(name) in UNIT_SCALE else v
    stalls = {k.split("stalled_")[1]: _num(v) for k, v in d.items()
              if "pcsamp_warps_issue_stalled" in k and _num(v) is not None and not k.endswith("not_issued")}
    tot = sum(stalls.values()) or 1.0
    entry = {
        "kernel": d.get("Kernel Name"), "report": os.path.relpath(evidence, ROOT),
        "report_bytes": os.path.getsize(evidence),
        "report_mtime": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime(os.path.getmtime(evidence))),
        "launches_in_report": len(rows), "counters": counters,
        "stall_share_pct": {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]},
        "dram_bytes_per_launch": (counters.get("dram__bytes_read.sum", 0.0) + counters.get("dram__bytes_write.sum", 0.0)),
        "unit": unit_name,
    }
    inst = counters.get("smsp__inst_executed.sum")
    if units and inst:
        entry["units_per_launch"] = units
        entry["warp_inst_per_warp_unit"] = inst * 32.0 / units      # warp-instructions per 32 units (one warp's worth)
        tinst = counters.get("smsp__thread_inst_executed.sum")
        if tinst:
            entry["thread_inst_per_unit"] = tinst / units
    mix = opcode_mix(path, kernel_re)
    if mix:
        total = sum(mix.values())
        norm = (units / 32.0) if units else total / 100.0
        groups = collections.Counter()
        for op, n in mix.items():
            base = op.split(".")[0]
            if op.startswith("IMAD.WIDE") or op.startswith("IMAD.HI"):
                groups["imad_wide"] += n
            elif base == "IMAD":
                groups["imad"] += n
            elif base in ("FFMA2", "FMUL2", "FADD2"):
                groups["fp32x2"] += n
            elif base in ("FFMA", "FMUL", "FADD"):
                groups["fp32"] += n
            elif base == "MUFU":
                groups["mufu"] += n
            elif base in ("DFMA", "DMUL", "DADD", "DSETP"):
                groups["fp64"] += n
        entry["opcode_mix_note"] = ("warp-instructions per warp and unit" if units else "percent of executed warp-instructions")
        entry["opcode_mix_top"] = {op: round(n / norm, 2) for op, n in mix.most_common(24)}
        entry["fma_pipe_mix"] = {k: round(groups[k] / norm, 2) for k in ("imad_wide", "imad", "fp32x2", "fp32")}
        entry["mufu_per_unit"] = round(groups["mufu"] / norm, 2)
        entry["fp64_per_unit"] = round(groups["fp64"] / norm, 2)
    return entry


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", default=os.path.join(ROOT, "gpurun_out"))
    ap.add_argument("--tag", default="r2")
    ap.add_argument("--print-metrics", action="store_true", help="print the --metrics list for ncu_capture.sh")
    args = ap.parse_args()
    if args.print_metrics:
        print(",".join(EXTRA_METRICS))
        return
    rev = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    out = {"_source": f"profiles/make_kernel_constants.py --tag {args.tag} over {os.path.relpath(args.dir, ROOT)}/*.ncu-rep "
                      f"(ncu --set full + pipe-cycle counters, --clock-control none, B200); regenerated "
                      f"{time.strftime('%Y-%m-%dT%H:%M:%SZ', time.gmtime())} at git {rev}"}
    tags = args.tag.split(",")                   # several captures: the LAST tag that has a report for a stem wins
    for key, (stem, kre, units, unit_name) in CAPTURES.items():
        e = None
        for tag in reversed(tags):
            e = summarise(key, stem, kre, units, unit_name, args.dir, tag)
            if e is not None:
                break
        if e is None:
            print(f"[skip] {key}: no {args.tag}_{stem}.ncu-rep with a kernel matching /{kre}/", file=sys.stderr)
            continue
        out[key] = e
    path = os.path.join(ROOT, "profiles", "kernel_constants.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=False)
        f.write("\n")
    print("wrote", path, "with", [k for k in out if not k.startswith("_")])


if __name__ == "__main__":
    main()
