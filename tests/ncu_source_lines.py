#!/usr/bin/env python
"""tests/ncu_source_lines.py REP KERNEL [min_pct] — PC-sample share per CUDA source line of one kernel of an
ncu report captured with --set full --import-source on (compile with -lineinfo)."""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[hi]
    cs, ci = hdr.index("# Samples"), hdr.index("Instructions Executed")
    lines = {}
    cur = None
    for r in rows[hi + 1:]:
        if len(r) <= cs:
            continue
        if r[0] not in ("", "-"):
            cur = (r[0], r[1])
            if r[2] == "-":  # the line's own summary row duplicates its instructions
                continue
        if cur is None:
            continue
        try:
            s, ie = int(r[cs] or 0), int(r[ci] or 0)
        except ValueError:
            continue
        a = lines.setdefault(cur, [0, 0])
        a[0] += s
        a[1] += ie
    tot = sum(a[0] for a in lines.values()) or 1
    print(f"total samples {tot}")
    for (ln, src), a in sorted(lines.items(), key=lambda kv: int(kv[0][0])):
        if 100.0 * a[0] / tot >= min_pct:
            print(f"{100.0 * a[0] / tot:5.1f}%  inst={a[1]:>10}  L{ln:>5}  {src.strip()[:110]}")


if __name__ == "__main__":
    main()
