#!/bin/bash
# compute-sanitizer evidence: initcheck on the synthesis loop, racecheck / synccheck on the GEMM + matcher step
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
timeout 900 compute-sanitizer --tool initcheck --print-limit 30 python -m pytest "tests/test_gpu_texture.py::test_end_to_end_sanity" -q -x > gpurun_out/r2_initcheck.txt 2>&1
grep -E "ERROR SUMMARY|Uninitialized|at .*\+0x|passed|failed" gpurun_out/r2_initcheck.txt | head -30
unset PYTORCH_NO_CUDA_MEMORY_CACHING
cat > /tmp/step.py <<'P'
import sys; sys.path.insert(0, ".")
import torch, optimaltextures_b200 as ob
g = torch.Generator().manual_seed(0)
p = torch.relu(torch.randn(1, 64, 64, 512, generator=g)).cuda(); s = torch.relu(torch.randn(1, 64, 64, 512, generator=g)).cuda()
big = torch.relu(torch.randn(1, 128, 128, 512, generator=g)).cuda()
r = ob.random_rotation(512, "cuda", seed=1, counter=0)
for mode in ("cdf", "sort", "pca"):
    ob.optimal_transport(p, s, mode, rotation=r)
    ob.optimal_transport(big, big, mode, rotation=r)
torch.cuda.synchronize(); print("done")
P
for tool in synccheck racecheck; do timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/step.py > gpurun_out/r2_$tool.txt 2>&1; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Barrier error|done" gpurun_out/r2_$tool.txt | head -12; done
