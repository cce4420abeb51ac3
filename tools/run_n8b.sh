#!/bin/bash
# run ON an 8-GPU box: final multi-GPU bench lines + an A/B of the pipelined index phase at 8 GPUs.  tools/run_n8b.sh <tag>
tag=${1:-rX}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
A="--steps 10 --warmup 3 --no-cpu-baseline --no-secondary"
timeout 300 $TR --nproc-per-node 8 --master-port 29703 bench.py --gpus 8 $A > gpurun_out/${tag}_bench_cfg4_n8.json 2> /dev/null
FTRL_B200_PIPELINE=0 timeout 300 $TR --nproc-per-node 8 --master-port 29704 bench.py --gpus 8 $A --no-e2e > gpurun_out/${tag}_bench_cfg4_n8_unpiped.json 2> /dev/null
timeout 300 $TR --nproc-per-node 8 --master-port 29705 bench.py --gpus 8 --workload cfg5 $A > gpurun_out/${tag}_bench_cfg5_n8.json 2> /dev/null
timeout 300 $TR --nproc-per-node 4 --master-port 29706 bench.py --gpus 4 $A > gpurun_out/${tag}_bench_cfg4_n4.json 2> /dev/null
for f in cfg4_n8 cfg4_n8_unpiped cfg5_n8 cfg4_n4; do python - gpurun_out/${tag}_bench_$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "N", d["n_gpus"], "ms", round(d["ms_per_step"],3), "Msamples/s", round(d["value"]/1e6,2), "e2e", d["e2e"] and round(d["e2e"]["value"]/1e6,2),
          {k:round(v,3) for k,v in r["phase_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
