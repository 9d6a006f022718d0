#!/bin/bash
# round-2 first GPU call: race-fix check on the 64-wide TMEM-A path, gpu tests, bench at the driver's step count
mkdir -p gpurun_out
for bn in 64 128 256; do OPTEX_FORCE_BN=$bn timeout 300 python scripts/debug_gemm_multi.py; done > gpurun_out/r2_gemm_multi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench20.json 2> gpurun_out/r2_bench20.err
timeout 600 python bench.py --steps 200 --warmup 5 --no-all-modes --no-layers --no-synthesis --no-cpu-baseline > gpurun_out/r2_bench200.json 2> gpurun_out/r2_bench200.err
tail -3 gpurun_out/r2_pytest.log; cut -c1-400 gpurun_out/r2_bench20.json; tail -5 gpurun_out/r2_bench20.err
