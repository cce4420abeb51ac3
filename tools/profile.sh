#!/bin/bash
# run ON the GPU box (through gpurun): launch list + one full ncu capture of the hot kernels of one bench step.
#   tools/profile.sh <tag> [zipf|uniform]
# outputs gpurun_out/<tag>_launches_<dist>.csv and gpurun_out/<tag>_<dist>.ncu-rep ; numbers printed by a run
# under ncu are never bench values.
tag=${1:-rX}; dist=${2:-zipf}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --dist $dist"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_${dist}.csv $B > gpurun_out/${tag}_launches_${dist}.log 2>&1
# the last bench step: skip the launches of warm-up (hot kernels only, by name)
ncu --set full --clock-control none --import-source on \
    -k regex:'k_ffm_tile|k_ffm_staged_rows|k_ffm_combine|k_row_touch|k_row_materialise' --launch-skip 15 -c 5 \
    -f -o gpurun_out/${tag}_${dist} $B > gpurun_out/${tag}_${dist}.log 2>&1
ls -la gpurun_out/ | tail -5
