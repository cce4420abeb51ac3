#!/usr/bin/env python
"""profiles/traffic.json from `ncu --set full` captures of one bench step: DRAM bytes (read + write) of the hot kernels,
stamped with the hash of the kernel sources (bench.py reports `roofline.traffic` only when the stamp matches the sources
the loaded library was built from).   usage: tools/traffic_from_ncu.py key=rep.ncu-rep [key=rep ...]
key = "<workload>-<dist>", e.g. cfg4-zipf"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_sha16

HOT = ("k_row_touch", "k_row_materialise", "k_ffm_tile", "k_ffm_staged_rows", "k_ffm_combine",
       "k_lrfm_sample", "k_lrfm_rows", "k_lrfm_combine")
out = {}
for arg in sys.argv[1:]:
    key, rep = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    detail = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        hot = next((h for h in HOT if h in name), None)
        if not hot or hot in detail:  # the first instance of each kernel = one step
            continue
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v, u = float(r[idx[m]]), units[idx[m]].lower()
            tot += v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        detail[hot] = tot
    out[key] = sum(detail.values())
    out[key + "-detail"] = detail
out["_kernel_source_sha16"] = kernel_source_sha16()
out["_source"] = ("ncu --set full capture of one bench step (tools/profile.sh), dram__bytes_read.sum + dram__bytes_write.sum "
                  "summed over the hot kernels of a step -- the same kernel group as roofline.achieved")
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
