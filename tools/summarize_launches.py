#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of the step."""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    out = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        out.append((re.sub(r"\(.*", "", row["Kernel Name"])[:100], v, row["Grid Size"]))
    return out


def main():
    rows = load(sys.argv[1])
    tot = collections.defaultdict(lambda: [0, 0.0])
    for n, v, _ in rows:
        tot[n][0] += 1
        tot[n][1] += v
    T = sum(v for _, v in tot.values())
    ours = sum(v for k, (_, v) in tot.items() if "gtos::" in k)
    print(f"launches {len(rows)}, total {T:.1f} us (cold-cache, serialised); gtos_b200 kernels {100 * ours / T:.1f}% of device time")
    print("| share | time us | launches | kernel |\n|---|---|---|---|")
    for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
        print(f"| {100 * v / T:.1f}% | {v:.1f} | {n} | `{k}` |")


if __name__ == "__main__":
    main()
