#!/usr/bin/env python
"""Per-source-line executed warp instructions of one kernel: joins the SASS page of an .ncu-rep with `nvdisasm -g` of the
library the capture ran (same instruction order).  usage: tools/line_hot.py rep lib.so kernel-regex mangled-prefix"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import Counter
rep, lib, pat, mangled = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
blk = rows[starts[0]:starts[1]]
hdr = blk[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in blk[2:] if r and r[0].startswith("0x")]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout.splitlines()
i0 = next(i for i, l in enumerate(sass) if l.startswith(".text." + mangled))
cur = None
lines = []
for l in sass[i0 + 1:]:
    if l.startswith(".text.") or l.startswith("\t.section") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
print("sass in report", len(data), "in disassembly", len(lines))
n = min(len(data), len(lines))
per = Counter()
samp = Counter()
for r, ln in zip(data[:n], lines[:n]):
    per[ln] += int(float(r[idx["Instructions Executed"]] or 0))
    samp[ln] += int(float(r[idx["# Samples"]] or 0))
tot = sum(per.values())
ts = sum(samp.values())
src = {}
for (f, ln), v in sorted(per.items(), key=lambda t: -t[1])[:60]:
    if f not in src:
        p = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][ln - 1].strip()[:90] if 0 < ln <= len(src[f]) else ""
    print(f"{v / tot * 100:5.1f}% inst {samp[(f, ln)] / max(ts, 1) * 100:5.1f}% smp  {f}:{ln}  {text}")
