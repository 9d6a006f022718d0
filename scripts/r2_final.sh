#!/bin/bash
# round-2 final validation: smoke, full gpu suite, driver-style bench, ncu launch list + full capture, configs on real inputs
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log; tail -3 gpurun_out/r2_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench20.json 2> gpurun_out/r2_bench20.err; cut -c1-200 gpurun_out/r2_bench20.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_ref.json
timeout 600 python scripts/run_configs.py > gpurun_out/r2_configs.log 2>&1; tail -4 gpurun_out/r2_configs.log | cut -c1-160
bash scripts/capture_profiles_r02.sh > /dev/null 2>&1
ls gpurun_out | grep r02_
