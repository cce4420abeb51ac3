#!/bin/bash
# run ON the GPU box: one short bench line per BASELINE.json configuration that fits one GPU (profiles/<tag>_*.json)
#   tools/bench_all.sh <tag>
tag=${1:-rX}
mkdir -p gpurun_out
for w in "cfg4" "cfg4 --dist uniform" "cfg3" "cfg3 --dist uniform" "cfg2-fm" "cfg2-lr"; do
  f=gpurun_out/${tag}_$(echo $w | tr " -" "__" | sed 's/___dist_/_/').json
  python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > $f 2> gpurun_out/${tag}_err.log
  python - "$w" "$f" <<'PY'
import json,sys
w,f=sys.argv[1:3]
try:
    d=json.load(open(f)); r=d["roofline"]
    print(w, "| ms", round(d["ms_per_step"],3), "| Msamples/s", round(d["value"]/1e6,2), "| e2e", round(d["e2e"]["value"]/1e6,2), "| frac", round(r["frac"],4),
          "| U/nnz", round(r["U_over_nnz"],3), {k:round(v,3) for k,v in r["phase_ms_per_step"].items()})
except Exception as e:
    print(w,"ERR",e)
PY
done
