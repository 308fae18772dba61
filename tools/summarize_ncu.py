#!/usr/bin/env python
"""Summarise an `ncu --set full` report (one row per profiled launch) into the counters the roofline discussion uses.
usage: python tools/summarize_ncu.py gpurun_out/r02_kernels.ncu-rep > profiles/r02_ncu_kernels.md   (runs without a GPU)"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAKS = {}
try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except OSError:
    pass
HBM = PEAKS.get("hbm_gbs", 6650.0)
WANT = {
    "gpu__time_duration.sum": "us",
    "dram__bytes_read.sum": "rd",
    "dram__bytes_write.sum": "wr",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram%",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2%",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm%",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor%",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ%",
    "launch__registers_per_thread": "regs",
    "launch__shared_mem_per_block_dynamic": "dsmem",
    "launch__shared_mem_per_block_static": "ssmem",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__cycles_active.avg": "cyc",
    "sm__cycles_elapsed.avg.per_second": "clk",
}
STALL = re.compile(r"smsp__average_warps?_issue_stalled_(\w+?)_per_issue_active\.ratio|"
                   r"smsp__average_warp_latency_issue_stalled_(\w+?)\.ratio")


def to_num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*\)$", "", name)
    return name.replace("nvae::", "").replace("(anonymous namespace)::", "")[:60]


def main():
    rep = sys.argv[1]
    if rep.endswith(".csv"):  # `ncu -i x.ncu-rep --page raw --csv > x.csv` exported on the GPU box
        txt = open(rep).read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    print(f"# ncu --set full summary: {os.path.basename(rep)}\n")
    print(f"One row per profiled launch (`--clock-control none`; times are cold-cache single launches).  `DRAM MB` = "
          f"dram__bytes_read.sum + dram__bytes_write.sum; `GB/s` = that / duration; `of HBM` against the measured copy peak "
          f"{HBM:.0f} GB/s (MEASURED_PEAKS.json).  Stall columns: the three largest `issue_stalled_*` reasons per issued "
          f"instruction.\n")
    print("| kernel | grid x block | regs | smem KB | us | DRAM MB | GB/s | of HBM | dram% | L2% | SM% | tensor% | occupancy% | top stalls |")
    print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
    for r in data:
        if len(r) < len(head) or r[col["Kernel Name"]].lstrip("void ").startswith("at::"):
            continue  # (torch's own fill / randn kernels of the driver script)
        g = {}
        for m, k in WANT.items():
            if m in col:
                v = to_num(r[col[m]])
                u = units[col[m]]
                if v is not None and m == "gpu__time_duration.sum":
                    v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
                if v is not None and k in ("rd", "wr"):
                    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                if v is not None and k in ("dsmem", "ssmem"):
                    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6}.get(u.split("/")[0], 1)
                g[k] = v
        stalls = []
        for h, i in col.items():
            mm = STALL.match(h)
            if mm and "not_issued" not in h:
                v = to_num(r[i])
                if v is not None:
                    stalls.append((v, (mm.group(1) or mm.group(2))))
        stalls.sort(reverse=True)
        seen, top = set(), []
        for v, n in stalls:
            if n not in seen:
                seen.add(n)
                top.append(f"{n} {v:.2f}")
            if len(top) == 3:
                break
        mb = ((g.get("rd") or 0) + (g.get("wr") or 0)) / 1e6
        us = g.get("us") or 0
        gbs = mb * 1e6 / (us * 1e-6) / 1e9 if us else 0
        f = lambda k, p=1: (f"{g[k]:.{p}f}" if g.get(k) is not None else "-")
        smem = ((g.get("dsmem") or 0) + (g.get("ssmem") or 0)) / 1000
        print(f"| `{short(r[col['Kernel Name']])}` | {f('grid', 0)} x {f('block', 0)} | {f('regs', 0)} | {smem:.0f} | {us:.1f} | "
              f"{mb:.1f} | {gbs:.0f} | {100 * gbs / HBM:.0f}% | {f('dram%')} | {f('l2%')} | {f('sm%')} | {f('tensor%')} | {f('occ%')} | "
              f"{', '.join(top)} |")


if __name__ == "__main__":
    main()
