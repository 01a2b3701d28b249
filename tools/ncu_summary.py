#!/usr/bin/env python
"""Condense an `ncu --set full` report and/or a launch list into the small text files kept
under profiles/ (the .ncu-rep itself stays in gpurun_out/, which is scratch).

    python tools/ncu_summary.py --rep gpurun_out/X.ncu-rep|X_raw.csv --launches gpurun_out/X_launches.csv \
        --out profiles/r01_X

writes <out>_kernels.csv (one row per profiled launch: duration, registers, occupancy, DRAM
bytes / %, SM %, tensor %, fp64 %, fma %, L2 hit rate, top stall reasons) and
<out>_launches.csv (kernel, grid, block, ns for every launch of the launch-list pass, plus
per-kernel totals and shares).
"""
from __future__ import annotations

import argparse
import collections
import csv
import io
import subprocess

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_static", "smem_static"),
    ("launch__shared_mem_per_block_dynamic", "smem_dyn"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_cycles_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
]


def raw_rows(rep):
    if rep.endswith(".csv"):  # already exported on the GPU box (`ncu -i X.ncu-rep --page raw --csv`)
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def kernels_table(rep, path):
    hdr, units, data = raw_rows(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel"] + [f"{n}" for _, n in METRICS] + ["top_stalls(warps per issue)"])
        for d in data:
            name = d[idx["Kernel Name"]].split("(")[0]
            row = [d[idx["ID"]], name]
            for m, _ in METRICS:
                row.append(f"{d[idx[m]]} {units[idx[m]]}".strip() if m in idx else "")
            st = []
            for c in stall_cols:
                try:
                    st.append((float(d[idx[c]]), c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
            st.sort(reverse=True)
            row.append("; ".join(f"{n}={v:.2f}" for v, n in st[:4]))
            w.writerow(row)


def launches_table(src, path):
    lines = [l for l in open(src) if not l.startswith("==")]
    if lines and not lines[0].startswith('"ID"'):  # filtered list: put the header line first
        hdr = [l for l in lines if l.startswith('"ID"')]
        lines = hdr[:1] + [l for l in lines if not l.startswith('"ID"')]
    rows = list(csv.DictReader(lines))
    tot = collections.OrderedDict()
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "ns"])
        for r in rows:
            name = r["Kernel Name"].split("(")[0]
            ns = float(r["Metric Value"].replace(",", ""))
            if r["Metric Unit"] in ("us", "usecond"):
                ns *= 1e3
            elif r["Metric Unit"] in ("ms", "msecond"):
                ns *= 1e6
            w.writerow([r["ID"], name, r["Grid Size"], r["Block Size"], f"{ns:.0f}"])
            t = tot.setdefault(name, [0, 0.0])
            t[0] += 1
            t[1] += ns
        allns = sum(v[1] for v in tot.values()) or 1.0
        w.writerow([])
        w.writerow(["kernel", "launches", "total_ns", "share"])
        for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, f"{ns:.0f}", f"{ns / allns:.4f}"])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rep")
    ap.add_argument("--launches")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    if a.rep:
        kernels_table(a.rep, a.out + "_kernels.csv")
    if a.launches:
        launches_table(a.launches, a.out + "_launches.csv")
