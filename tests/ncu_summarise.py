#!/usr/bin/env python
"""tests/ncu_summarise.py CSV [name ...] — per-kernel totals of an `ncu --metrics gpu__time_duration.sum[,dram__bytes_*]
--csv` launch list, one table per transformed block (a block starts at k_prep_block).  Prints markdown."""
import collections
import csv
import re
import sys


def main():
    path, names = sys.argv[1], sys.argv[2:]
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    launch, order = {}, []
    for row in csv.DictReader(lines):
        i = int(row["ID"])
        if i not in launch:
            launch[i] = {"name": row["Kernel Name"], "grid": row["Grid Size"]}
            order.append(i)
        v = float(row["Metric Value"].replace(",", ""))
        u, m = row["Metric Unit"], row["Metric Name"]
        if m == "gpu__time_duration.sum":
            launch[i]["us"] = {"ns": v / 1e3, "us": v, "ms": v * 1e3}[u]
        else:
            launch[i][m] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    blocks = []
    for i in order:
        if launch[i]["name"].startswith("k_prep_block"):
            blocks.append([])
        if blocks:
            blocks[-1].append(launch[i])
    for bi, b in enumerate(blocks):
        agg = collections.OrderedDict()
        for l in b:
            k = re.sub(r"\(.*", "", l["name"]).replace("void ", "").replace("unsigned long long", "u64").replace("unsigned int", "u32")
            a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
            a[0] += 1
            a[1] += l["us"]
            a[2] += l.get("dram__bytes_read.sum", 0)
            a[3] += l.get("dram__bytes_write.sum", 0)
        tot = sum(a[1] for a in agg.values())
        tb = sum(a[2] + a[3] for a in agg.values())
        nm = names[bi] if bi < len(names) else f"block {bi}"
        print(f"\n### {nm}: {len(b)} launches, {tot:.0f} us summed, DRAM read+write {tb / 1e6:.0f} MB\n")
        print("| kernel | launches | us | share | DRAM read MB | DRAM write MB |\n|---|---:|---:|---:|---:|---:|")
        for k, a in agg.items():
            print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f}% | {a[2] / 1e6:.1f} | {a[3] / 1e6:.1f} |")


if __name__ == "__main__":
    main()
