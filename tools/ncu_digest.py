"""Digest an .ncu-rep: headline metrics, stall breakdown and opcode mix per kernel (debug/optimisation aid)."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
stalls = [h for h in H if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print("==", r[H.index("Kernel Name")][:100])
    for w in want:
        if w in H:
            print(f"  {w:75s} {r[H.index(w)]} {rows[1][H.index(w)]}")
    st = sorted(((float(r[H.index(s)]), s) for s in stalls), reverse=True)[:7]
    print("  stalls/issue:", ", ".join(f"{s.split('stalled_')[1].split('_per_issue')[0]}={v:.2f}" for v, s in st))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
sec = None
for r in csv.reader(src.splitlines()):
    if r and r[0] == "Kernel Name":
        if sec:
            pass
        sec = {"name": r[1], "H": None, "agg": collections.Counter(), "tot": 0}
        secs = globals().setdefault("SECS", [])
        secs.append(sec)
    elif sec is not None and sec["H"] is None:
        sec["H"] = r
    elif sec is not None:
        Hs = sec["H"]
        try:
            n = int(r[Hs.index("Instructions Executed")])
        except Exception:
            continue
        toks = r[Hs.index("Source")].split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        sec["agg"][op.split(".")[0]] += n
        sec["tot"] += n
for sec in globals().get("SECS", []):
    print("== opcode mix:", sec["name"][:80], "total warp instr", sec["tot"])
    print("  ", ", ".join(f"{op}={n / max(sec['tot'], 1):.1%}" for op, n in sec["agg"].most_common(14)))
