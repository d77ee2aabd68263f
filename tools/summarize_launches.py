"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`) into a share-per-kernel markdown table.

    python tools/summarize_launches.py gpurun_out/launches_v4.csv [first_n_rows]
"""
import collections
import csv
import re
import sys


def load(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit") in ("us", "usecond"):
            ns *= 1e3
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("ard::", "")
        rows.append((name, r["Grid Size"], ns))
    return rows


def table(rows):
    agg = collections.OrderedDict()
    for name, grid, ns in rows:
        k = (name, grid)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    out = ["| share | launches | avg us | kernel | grid |", "|---|---|---|---|---|"]
    for (name, grid), (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {100 * ns / tot:.1f}% | {n} | {ns / n / 1e3:.1f} | `{name}` | {grid} |")
    return "\n".join(out), tot


if __name__ == "__main__":
    rows = load(sys.argv[1])
    if len(sys.argv) > 2:
        rows = rows[:int(sys.argv[2])]
    md, tot = table(rows)
    print(f"{len(rows)} launches, {tot / 1e6:.3f} ms of kernel time\n")
    print(md)
