"""Per-opcode view of an ncu source page (the *_source.csv.gz files ncu_capture.sh writes): executed
warp-instructions and stall samples by opcode, and the instructions with the most samples.
usage: python profiles/source_stalls.py gpurun_out/<tag>_<stem>_source.csv.gz [units per launch]"""
import collections
import csv
import gzip
import re
import sys


def main(path, units=None):
    rows = list(csv.reader(gzip.open(path, "rt")))
    # several kernels may be concatenated: take the one with the most executed instructions
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and len(r) > 5:
            cur["rows"].append(r)
    for blk in blocks:
        h = {n: i for i, n in enumerate(blk["hdr"])}
        stall_cols = [n for n in blk["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
        ex = collections.Counter()
        smp = collections.Counter()
        why = collections.defaultdict(collections.Counter)
        total_ex = total_s = 0
        top = []
        for r in blk["rows"]:
            txt = re.sub(r"^@!?U?P\d+\s+", "", r[h["Source"]].strip())
            op = txt.split()[0] if txt else "?"
            op = ".".join(op.split(".")[:2]) if op.startswith(("IMAD", "MUFU")) else op.split(".")[0]
            n = int(r[h["Instructions Executed"]] or 0)
            s = int(r[h["# Samples"]] or 0)
            ex[op] += n
            smp[op] += s
            total_ex += n
            total_s += s
            for c in stall_cols:
                v = int(r[h[c]] or 0)
                if v:
                    why[op][c[6:]] += v
            top.append((s, n, r[h["Source"]].strip(), {c[6:]: int(r[h[c]] or 0) for c in stall_cols if int(r[h[c]] or 0)}))
        print("##", blk["name"])
        per = f" = {total_ex * 32 / float(units):.1f} thread-inst per unit" if units else ""
        print(f"   warp-instructions executed {total_ex}{per}; samples {total_s}")
        print("   opcode: executed share | sample share | samples per 1000 executed | main reasons")
        for op, n in ex.most_common(22):
            rs = ", ".join(f"{k} {v * 100 // max(smp[op], 1)}%" for k, v in why[op].most_common(3))
            print(f"   {op:14s} {100.0 * n / total_ex:5.1f}% | {100.0 * smp[op] / max(total_s, 1):5.1f}% | "
                  f"{1000.0 * smp[op] / max(n, 1):7.2f} | {rs}")
        print("   hottest instructions:")
        for s, n, txt, w in sorted(top, key=lambda t: -t[0])[:14]:
            print(f"   {s:6d} samples, {n:9d} executed  {txt[:60]:60s} {dict(sorted(w.items(), key=lambda kv: -kv[1])[:3])}")


if __name__ == "__main__":
    main(*sys.argv[1:])
