#!/bin/bash
# run ON the GPU box: the evidence set of a round on one GPU (small files only: gpurun_out is capped at 64 MiB per
# call; the ncu reports come from tools/profile.sh in a call of their own).   tools/final_n1.sh <tag>
tag=${1:-rX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
FTRL_B200_PRECISE=1 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_sharded.py -x -q -m gpu > gpurun_out/${tag}_pytest_precise.log 2>&1; tail -1 gpurun_out/${tag}_pytest_precise.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
bash tools/bench_all.sh ${tag}
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench_n1.err
bash tools/sanitize.sh ${tag}
ls gpurun_out | grep ${tag} | wc -l
