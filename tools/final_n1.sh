#!/bin/bash
# run ON the GPU box: the evidence set of a round on one GPU.   tools/final_n1.sh <tag>
tag=${1:-rX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
bash tools/bench_all.sh ${tag}
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench_n1.err
bash tools/profile.sh ${tag} zipf > /dev/null 2>&1
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary"
ncu --set full --clock-control none --import-source on -k regex:"k_ffm_tile|k_ffm_staged_rows|k_ffm_combine|k_row_touch|k_row_materialise|k_chunk_desc" --launch-skip 18 -c 6 -f -o gpurun_out/${tag}_uniform $B --dist uniform > gpurun_out/${tag}_uniform.log 2>&1
for w in cfg2-lr cfg2-fm; do
  ncu --set full --clock-control none --import-source on -k regex:"k_lrfm_sample|k_lrfm_rows|k_lrfm_combine" --launch-skip 9 -c 3 -f -o gpurun_out/${tag}_${w} $B --workload $w > gpurun_out/${tag}_${w}.log 2>&1
done
bash tools/sanitize.sh ${tag}
ls gpurun_out | grep ${tag} | wc -l
