#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares.

usage: python tools/summarize_launches.py gpurun_out/launches.csv [--skip N] > profiles/<name>.md
The per-launch times are cold-cache and serialised (ncu replays each kernel alone), so only the
SHARE of each kernel is meaningful, not the absolute step time.
"""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*\)$", "", name)
    name = name.replace("nvae::", "").replace("(anonymous namespace)::", "")
    return name[:90]


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((short(r["Kernel Name"]), ns, r["Grid Size"], r["Block Size"]))
    rows = rows[skip:]
    tot = sum(r[1] for r in rows)
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for n, ns, *_ in rows:
        a = agg[n]
        a[0] += 1
        a[1] += ns
        a[2] = max(a[2], ns)
    print(f"# launch list summary: {path}")
    print(f"\n{len(rows)} launches, {tot / 1e6:.3f} ms summed device time (serialised, cold cache)\n")
    print("| kernel | launches | total ms | share | avg us | max us |")
    print("|---|---:|---:|---:|---:|---:|")
    for n, (c, t, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {t / 1e6:.3f} | {100 * t / tot:.1f}% | {t / c / 1e3:.1f} | {mx / 1e3:.1f} |")
    print("\n## 15 longest launches\n")
    print("| kernel | grid | block | us |")
    print("|---|---|---|---:|")
    for n, ns, g, b in sorted(rows, key=lambda r: -r[1])[:15]:
        print(f"| `{n}` | {g} | {b} | {ns / 1e3:.1f} |")


if __name__ == "__main__":
    main()
