#!/usr/bin/env python
"""Hot spots of one kernel from an .ncu-rep (source page, SASS level): per-opcode shares and the instructions with the
most executed warps / stall samples.  usage: tools/sass_hot.py rep kernel-regex [instance]"""
import csv, io, subprocess, sys
from collections import Counter
rep, pat = sys.argv[1], sys.argv[2]
inst_no = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
blk = rows[starts[inst_no]:starts[inst_no + 1]]
print(blk[0][1][:100])
hdr = blk[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in blk[2:] if r and r[0].startswith("0x")]
I = lambda r, k: int(float(r[idx[k]] or 0))
tot_i = sum(I(r, "Instructions Executed") for r in data)
tot_s = sum(I(r, "# Samples") for r in data)
print("warp instructions", tot_i, "samples", tot_s, "sass lines", len(data))
c, cs = Counter(), Counter()
for r in data:
    op = r[idx["Source"]].split()
    o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
    c[o] += I(r, "Instructions Executed")
    cs[o] += I(r, "# Samples")
for o, v in c.most_common(22):
    print(f"  {o:10s} inst {v / tot_i * 100:5.1f}%  samples {cs[o] / max(tot_s, 1) * 100:5.1f}%")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("top by samples:")
for n, r in sorted(((I(r, "# Samples"), r) for r in data), key=lambda t: -t[0])[:40]:
    st = sorted(((I(r, s), s[6:]) for s in stall_cols), reverse=True)[:2]
    i = data.index(r)
    print(f"  #{i:5d} {n / max(tot_s, 1) * 100:5.2f}% inst {I(r, 'Instructions Executed'):9d} {r[idx['Source']].strip()[:70]:70s} {st}")
