"""Static SASS instruction count of one kernel in libbmc_b200.so (quick proxy for the dynamic
per-iteration count between GPU runs).  usage: python profiles/sass_count.py <substring of mangled name>"""
import collections
import re
import subprocess
import sys

LIB = "pybmc_b200/csrc/libbmc_b200.so"


def main(pattern):
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    blocks = out.split("Function : ")
    for b in blocks[1:]:
        name = b.split("\n", 1)[0]
        if pattern not in name:
            continue
        ops = collections.Counter()
        n = 0
        for line in b.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                ops[m.group(2).split(".")[0]] += 1
                n += 1
        print(name, "total", n)
        print("  " + ", ".join(f"{k} {v}" for k, v in ops.most_common(16)))


if __name__ == "__main__":
    main(sys.argv[1])
