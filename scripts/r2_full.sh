#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -15 gpurun_out/r2_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench20.json 2> gpurun_out/r2_bench20.err
cut -c1-300 gpurun_out/r2_bench20.json; tail -3 gpurun_out/r2_bench20.err
