#!/usr/bin/env python
"""tests/sass_summary.py [lib.so] — per-kernel resource usage (cuobjdump -res-usage) and instruction mix of the shipped SASS
(cuobjdump -sass) as markdown: what the compiled sm_100a code actually consists of.  Runs without a GPU."""
import collections
import re
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "bwtc_b200", "libbwtc_cuda.so")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    res = []
    for d in out:
        d = re.sub(r"\(.*", "", d).replace("void ", "").replace("bwtc_b200::", "").replace("unsigned long long", "u64").replace("unsigned int", "u32")
        d = d.replace("(bool)1", "true").replace("(bool)0", "false").replace("(int)", "")
        res.append(d)
    return res


def main():
    ru = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    rows = []
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", ru):
        rows.append(m.groups())
    names = demangle([r[0] for r in rows])
    print("## Resource usage (`cuobjdump -res-usage`)\n")
    print("| kernel | registers | stack B | static smem B | local B |\n|---|---:|---:|---:|---:|")
    for n, r in sorted(zip(names, rows)):
        print(f"| `{n}` | {r[1]} | {r[2]} | {r[3]} | {r[4]} |")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    groups = [("LDG", ("LDG",)), ("STG", ("STG",)), ("LDS", ("LDS", "LDSM")), ("STS", ("STS",)), ("ATOMS/ATOMG/RED", ("ATOMS", "ATOMG", "RED", "ATOM")),
              ("VOTE", ("VOTE", "VOTEU")), ("SHFL", ("SHFL",)), ("BAR", ("BAR",)), ("R2P/P2R", ("R2P", "P2R")), ("LOP3", ("LOP3",)),
              ("POPC/FLO", ("POPC", "FLO", "BREV")), ("NANOSLEEP", ("NANOSLEEP",)), ("UTMA*/UBLKCP", ("UTMALDG", "UTMASTG", "UBLKCP"))]
    dn = demangle(list(kernels.keys()))
    print("\n## Instruction mix of the shipped SASS (`cuobjdump -sass`, static counts)\n")
    print("| kernel | total | " + " | ".join(g for g, _ in groups) + " |\n|---|---:|" + "---:|" * len(groups))
    for n, (k, c) in sorted(zip(dn, kernels.items())):
        tot = sum(c.values())
        print(f"| `{n}` | {tot} | " + " | ".join(str(sum(c[o] for o in ops)) for _, ops in groups) + " |")


if __name__ == "__main__":
    main()
