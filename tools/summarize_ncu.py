#!/usr/bin/env python
"""Turn ncu output brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv  profiles/r01_launches.md  ["title"]
    python tools/summarize_ncu.py kernel   gpurun_out/prof.ncu-rep  profiles/r01_kernel.md    ["title"]

`launches`: per-kernel count / mean duration / share of the step from a `--metrics gpu__time_duration.sum` launch list
(cold-cache, serialised: compare SHARES, not absolutes).  `kernel`: the roofline-relevant counters of every launch in
an `ncu --set full` report (needs the `ncu` binary to read the .ncu-rep; no GPU needed).
"""
from __future__ import annotations

import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__warps_active.avg.per_cycle_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def launches(src: str, dst: str, title: str) -> None:
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "")
        agg.setdefault(name, []).append((float(r[vi].replace(",", "")), r[gi], r[bi]))
    tot = sum(v[0] for vs in agg.values() for v in vs)
    out = ["# %s" % title, "", "Source: `%s` (ncu `--metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache and" % src,
           "serialised, so only the SHARES are comparable with the CUDA-event stage times in the bench line).", "",
           "| kernel | launches | grid | block | mean us | share of all launches |", "|---|---|---|---|---|---|"]
    for name, vs in agg.items():
        s = sum(v[0] for v in vs)
        out.append("| `%s` | %d | %s | %s | %.1f | %.1f %% |" % (name, len(vs), vs[-1][1], vs[-1][2], s / len(vs) / 1e3, 100 * s / tot))
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


def kernel(src: str, dst: str, title: str) -> None:
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = ["# %s" % title, "", "Source: `%s` (`ncu --set full --clock-control none --import-source on`), one block per profiled launch." % src, ""]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        out.append("## `%s`  grid %s block %s" % (d.get("Kernel Name", "?")[:110], d.get("Grid Size"), d.get("Block Size")))
        out.append("")
        out.append("| metric | value | unit |")
        out.append("|---|---|---|")
        for k in KEYS:
            if k in d and d[k] != "":
                out.append("| %s | %s | %s |" % (k, d[k], units[hdr.index(k)]))
        out.append("")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    {"launches": launches, "kernel": kernel}[mode](src, dst, title)
