#!/bin/bash
# run ON the GPU box: A/B matrix of env knobs over short bench runs.  usage: tools/ab.sh <tag> "<ENV1>" "<ENV2>" ...
tag=$1; shift
i=0
for e in "$@"; do
  extra="--no-secondary"; [ $i -eq 0 ] && extra=""
  env $e python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e $extra > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  python - "$tag" "$i" "$e" <<'PY'
import json,sys
tag,i,e=sys.argv[1:4]
try:
    d=json.load(open(f"gpurun_out/{tag}_{i}.json")); r=d["roofline"]; u=d.get("roofline_uniform_ids") or {}
    print(i, e or "-", "ms", round(d["ms_per_step"],3), "frac", round(r["frac"],4), {k:round(v,3) for k,v in r["phase_ms_per_step"].items() if k in("sample","rows","combine","materialise")}, "uniform", u.get("frac") and round(u["frac"],4), u.get("kernel_ms_per_step") and round(u["kernel_ms_per_step"],3))
except Exception as ex:
    print(i, e, "ERR", ex, open(f"gpurun_out/{tag}_{i}.err").read()[-400:])
PY
  i=$((i+1))
done
