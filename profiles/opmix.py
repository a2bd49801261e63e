"""Per-opcode executed-instruction mix from an `ncu --page source --csv` dump.

    ncu -i prof.ncu-rep --page source --csv > src.csv
    python profiles/opmix.py src.csv <warp_iterations>

Prints warp-level instructions per warp-iteration by opcode and the stall-sample share of the
hottest instructions (needs -lineinfo builds for the source mapping, not for this table)."""
import collections
import csv
import sys


def main(path, denom):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    src, ex, smp = ci["Source"], ci["Instructions Executed"], ci["# Samples"]
    tot, samples = collections.Counter(), collections.Counter()
    n = ns = 0.0
    for r in rows[2:]:
        if len(r) <= ex:
            continue
        try:
            e, s = float(r[ex]), float(r[smp])
        except ValueError:
            continue
        parts = r[src].split()
        if not parts:
            continue
        op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
        op = op.split(".")[0]
        tot[op] += e
        samples[op] += s
        n += e
        ns += s
    print(f"{'opcode':12s} {'inst/warp-iter':>14s} {'stall samples %':>16s}")
    for k, v in tot.most_common(28):
        print(f"{k:12s} {v / denom:14.1f} {100 * samples[k] / max(ns, 1):16.1f}")
    print(f"{'total':12s} {n / denom:14.1f}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]))
