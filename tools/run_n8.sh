#!/bin/bash
# run ON an 8-GPU box (gpurun --gpus 8): real multi-GPU evidence for the feature-sharded path.
#   tools/run_n8.sh <tag>
#  1. tests/test_multi_gpu.py at world 4 and 8 (torchrun sharded == single GPU; main --n_gpus 4)
#  2. BASELINE.json configs[4]: 100 M features, k 8 (374 GB of w/z/n over 8 GPUs): size-independent properties of one
#     sharded step (tools/mgpu_check.py --big) and the bench line
#  3. configs[3] (cfg4) bench lines at 8 and 4 GPUs
tag=${1:-rX}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -k "8 or 4" > gpurun_out/${tag}_mgpu8_pytest.log 2>&1; tail -3 gpurun_out/${tag}_mgpu8_pytest.log
timeout 600 $TR --nproc-per-node 8 --master-port 29701 tools/mgpu_check.py --big 100000000 > gpurun_out/${tag}_cfg5_properties.log 2>&1; grep -E "MGPU_BIG|rank 0" gpurun_out/${tag}_cfg5_properties.log
timeout 600 $TR --nproc-per-node 8 --master-port 29702 bench.py --gpus 8 --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${tag}_bench_cfg5_n8.json 2> gpurun_out/${tag}_bench_cfg5_n8.err
timeout 600 $TR --nproc-per-node 8 --master-port 29703 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${tag}_bench_cfg4_n8.json 2> gpurun_out/${tag}_bench_cfg4_n8.err
timeout 600 $TR --nproc-per-node 4 --master-port 29704 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${tag}_bench_cfg4_n4.json 2> gpurun_out/${tag}_bench_cfg4_n4.err
for f in cfg5_n8 cfg4_n8 cfg4_n4; do python - gpurun_out/${tag}_bench_$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "N", d["n_gpus"], "ms", round(d["ms_per_step"],3), "Msamples/s", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2),
          {k:round(v,3) for k,v in r["phase_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
