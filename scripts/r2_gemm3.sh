#!/bin/bash
mkdir -p gpurun_out
timeout 180 python scripts/gemm_cta_times.py > gpurun_out/r2_gemm_cg2.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_gemm_cg2.txt
grep -v "kb[ 0-9]*:" gpurun_out/r2_gemm_cg2.txt
timeout 900 python -m pytest tests/test_gpu_gemm_tiles.py tests/test_gpu_ot_step.py tests/test_gpu_matchers.py tests/test_gpu_cov.py -x -q > gpurun_out/r2_pytest_sub.log 2>&1; tail -5 gpurun_out/r2_pytest_sub.log
for cg in 2; do OPTEX_CTA_GROUP=$cg timeout 300 python bench.py --steps 50 --warmup 5 --no-all-modes --no-layers --no-synthesis --no-cpu-baseline --no-tf32-peak > gpurun_out/r2_bench_cg$cg.json 2>gpurun_out/r2_bench_cg$cg.err; python - <<P
import json
d=json.load(open("gpurun_out/r2_bench_cg$cg.json"))
print("cg$cg", round(d["value"]), "it/s", round(d["ms_per_step"]*1e3,1), "us/step", {k:round(v["ms"]*1e3,1) for k,v in d["kernels"].items()}, "e2e", round(d["e2e"]["value"]))
P
done
