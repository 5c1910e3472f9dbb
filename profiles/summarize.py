#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the small text summaries kept under
profiles/.  Usage: python profiles/summarize.py <launches.csv> <prof.ncu-rep> <out.md> [title]"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
           "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = next(i for i, r in enumerate(rows) if r[0] == "ID")
    H, data = rows[h], rows[h + 1:]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    total = sum(sum(v) for v in agg.values())
    out = ["| kernel | launches | total us | mean us | share |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k[:90]}` | {len(v)} | {sum(v) / 1e3:.1f} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / total:.1%} |")
    return out


def full(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    H, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        out.append(f"\n**{r[H.index('Kernel Name')]}**\n")
        out.append("| metric | value | unit |")
        out.append("|---|---|---|")
        for m in METRICS:
            if m in H:
                out.append(f"| {m} | {r[H.index(m)]} | {units[H.index(m)]} |")
    return out


if __name__ == "__main__":
    lcsv, rep, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else "ncu summary"
    lines = [f"# {title}", "", "## launch list (`--metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)", ""]
    lines += launches(lcsv)
    lines += ["", "## `--set full` capture", ""]
    lines += full(rep)
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst)
