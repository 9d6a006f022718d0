#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ot_step.py tests/test_gpu_baseline_shapes.py tests/test_gpu_gemm_tiles.py tests/test_gpu_cov.py -x -q > gpurun_out/r2_pytest_sub.log 2>&1; tail -5 gpurun_out/r2_pytest_sub.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-all-modes --no-layers --no-synthesis --no-cpu-baseline --no-tf32-peak > gpurun_out/r2_bench_dual.json 2>gpurun_out/r2_bench_dual.err; tail -2 gpurun_out/r2_bench_dual.err; python - <<P
import json
d=json.load(open("gpurun_out/r2_bench_dual.json"))
print(round(d["value"]), "it/s", round(d["ms_per_step"]*1e3,1), "us/step", {k:(round(v["ms"]*1e3,1),v["launches"]) for k,v in d["kernels"].items()}, "launches", d["gpu_launches"], "e2e", round(d["e2e"]["value"]), d.get("warning"))
P
