#!/bin/bash
# run ON the GPU box (through gpurun): compute-sanitizer over the final kernels.
#   tools/sanitize.sh <tag>
# memcheck + racecheck of __graft_entry__.smoke() (sequential kernel, tile path, generic path, LR/FM kernels, predict, AUC)
# and memcheck of one sharded run on logical shards (tests/test_sharded.py: peer barriers, pipelined index phase,
# owner-side kernels).  Summaries go to gpurun_out/<tag>_sanitizer_*.log (copy to profiles/).
tag=${1:-rX}
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
run() {  # name, tool, command...
  local name=$1 tool=$2; shift 2
  timeout 900 $S --tool $tool --print-limit 20 "$@" > gpurun_out/${tag}_sanitizer_${name}.full 2>&1
  echo "rc=$?" >> gpurun_out/${tag}_sanitizer_${name}.full
  # keep the verdict lines and any error records
  grep -E "=========|rc=|passed|failed|error" gpurun_out/${tag}_sanitizer_${name}.full | head -80 > gpurun_out/${tag}_sanitizer_${name}.log
  rm -f gpurun_out/${tag}_sanitizer_${name}.full
  tail -3 gpurun_out/${tag}_sanitizer_${name}.log
}
run smoke_memcheck memcheck python -c "import __graft_entry__ as g; g.smoke()"
run smoke_racecheck racecheck python -c "import __graft_entry__ as g; g.smoke()"
run sharded_memcheck memcheck python -m pytest tests/test_sharded.py -x -q -m gpu -k "ragged or empty_share"
