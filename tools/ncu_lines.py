#!/usr/bin/env python
"""Per-source-line digest of an ncu report: aligns the report's SASS listing (``--page source --csv``) with
``nvdisasm -g`` of the object file (built with -lineinfo) and sums executed warp instructions and stall samples per
CUDA source line.  usage: ncu_lines.py report.ncu-rep object.o kernel-substring [top]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# instructions of the wanted function with their source line
lines, cur, infn = [], None, False
for l in sass:
    if l.startswith(".text."):
        infn = kname in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((cur, m.group(2).strip()))
ncu_cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
if os.environ.get("NCU_K"):  # report holds several kernels: NCU_K=regex keeps one (and NCU_C=n takes n launches, default 1)
    ncu_cmd += ["-k", "regex:" + os.environ["NCU_K"], "-c", os.environ.get("NCU_C", "1")]
raw = subprocess.run(ncu_cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
st = rows.index(H) + 1
si, ie, ws = H.index("Source"), H.index("Instructions Executed"), H.index("Warp Stall Sampling (All Samples)")
prof = []
for r in rows[st:]:
    try:
        prof.append((r[si].strip(), int(r[ie]), int(r[ws])))
    except Exception:
        pass
print(f"sass instr: nvdisasm {len(lines)} / ncu {len(prof)}")
n = min(len(lines), len(prof))
agg = collections.defaultdict(lambda: [0, 0])
for (loc, _), (_, ni, nw) in zip(lines[:n], prof[:n]):
    agg[loc][0] += ni
    agg[loc][1] += nw
ti = sum(v[0] for v in agg.values())
tw = sum(v[1] for v in agg.values())
src_cache = {}
def src(loc):
    if loc is None:
        return "?"
    f, ln = loc
    for d in ("ct_vae_b200/csrc", "."):
        p = os.path.join(d, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][ln - 1].strip()[:100]
    return ""
print(f"total warp instr {ti}, stall samples {tw}")
for loc, (ni, nw) in sorted(agg.items(), key=lambda kv: -kv[1][int(os.environ.get("SORT_INSTR", "0")) ^ 1])[:top]:
    print(f"{100*nw/max(tw,1):5.1f}% stall {100*ni/max(ti,1):5.1f}% instr  {loc}: {src(loc)}")

# optional region breakdown: NCU_REGIONS="name:lo-hi,name:lo-hi" over lines of the kernel's own .cu file
regs = os.environ.get("NCU_REGIONS")
if regs:
    main_file = os.environ.get("NCU_FILE", os.path.basename(obj).replace(".o", ".cu"))
    print("-- regions of", main_file)
    spec = [(r.split(":")[0], *map(int, r.split(":")[1].split("-"))) for r in regs.split(",")]
    tot = collections.OrderedDict((nm, [0, 0]) for nm, _, _ in spec)
    tot["(headers / other)"] = [0, 0]
    for loc, (ni, nw) in agg.items():
        hit = "(headers / other)"
        if loc is not None and loc[0] == main_file:
            for nm, lo, hi in spec:
                if lo <= loc[1] <= hi:
                    hit = nm
                    break
        tot[hit][0] += ni
        tot[hit][1] += nw
    for nm, (ni, nw) in tot.items():
        print(f"  {nm:28s} {100*ni/max(ti,1):5.1f}% instr ({ni/1e6:7.2f} M)  {100*nw/max(tw,1):5.1f}% stall samples")
