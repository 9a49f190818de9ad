#!/usr/bin/env python
"""Turn ncu output brought back from a gpurun call into the small summaries kept under profiles/.

  python tools/ncu_summary.py launches gpurun_out/<dir>/launches.csv  > profiles/<name>_launches.md
  python tools/ncu_summary.py kernel   gpurun_out/<dir>/prof.ncu-rep  > profiles/<name>_kernel.json
  python tools/ncu_summary.py stalls   gpurun_out/<dir>/prof.ncu-rep  [kernel-id]   (source page: hottest lines)

`launches` reads the CSV of `ncu --metrics gpu__time_duration.sum --csv --log-file ...` (one row per
launch) and prints per-kernel launch counts, total / mean device time and the share of the captured
window.  `kernel` reads a `--set full` report through `ncu -i ... --page raw --csv` and keeps the
metrics DESIGN.md argues from.  Runs here (no GPU needed)."""
from __future__ import annotations

import collections
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
    "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.per_cycle_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
]


def launches(path: str) -> None:
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"source: {path}\n")
    print("| kernel | launches | total ms | mean us | share of captured device time |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {v[0]} | {v[1] / 1e3:.3f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f} % |")
    print(f"| all | {sum(v[0] for v in agg.values())} | {tot / 1e3:.3f} | | 100 % |")
    print("\nPer-launch times are ncu's (serialised, cold cache, clocks not controlled): compare shares, not absolutes.")


def _raw(rep: str):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def kernel(rep: str) -> None:
    hdr, units, rows = _raw(rep)
    out = []
    for r in rows:
        d = collections.OrderedDict()
        d["kernel"] = r[hdr.index("Kernel Name")].split("(")[0]
        for h, u, v in zip(hdr, units, r):
            if h in KEEP and v != "":
                d[h] = f"{v} {u}".strip()
        out.append(d)
    print(json.dumps({"source": rep, "launches": out}, indent=1))


def stalls(rep: str, which: str | None) -> None:
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    print(out[:200000])


if __name__ == "__main__":
    if len(sys.argv) < 3:
        sys.exit(__doc__)
    {"launches": lambda: launches(sys.argv[2]), "kernel": lambda: kernel(sys.argv[2]),
     "stalls": lambda: stalls(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)}[sys.argv[1]]()
