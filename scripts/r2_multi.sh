#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_parallel.py -q -x > gpurun_out/r2_parallel_n$N.log 2>&1; tail -5 gpurun_out/r2_parallel_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; tail -3 gpurun_out/r2_bench_n$N.err
python - <<P
import json
d=json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().split("\n")[-1])
print("value", d["value"], "n_gpus", d["n_gpus"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"] if d.get("e2e") else None)
print(json.dumps(d.get("sharded"), indent=1)[:3000])
P
