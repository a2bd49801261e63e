"""Opcode histogram of the innermost hot loop (largest backward branch span below a limit) of one kernel in
libbmc_b200.so.  usage: python profiles/sass_loop.py <mangled-name substring> [lib]"""
import collections
import re
import subprocess
import sys


def main(pattern, lib="pybmc_b200/csrc/libbmc_b200.so"):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    for b in out.split("Function : ")[1:]:
        name = b.split("\n", 1)[0]
        if pattern not in name:
            continue
        ins = []
        for l in b.splitlines():
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        loops = []
        for a, t in ins:
            m = re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                loops.append((int(m.group(1), 16), a))
        print(name, "instructions", len(ins), "loops", [(hex(x), hex(y), (y - x) // 16 + 1) for x, y in loops])
        for lo, hi in loops:
            n = (hi - lo) // 16 + 1
            if n < 200:
                continue
            c = collections.Counter()
            for a, t in ins:
                if lo <= a <= hi:
                    c[re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]] += 1
            print(f"  loop {hex(lo)}-{hex(hi)} ({n}):", ", ".join(f"{k} {v}" for k, v in c.most_common(24)))


if __name__ == "__main__":
    main(*sys.argv[1:])
