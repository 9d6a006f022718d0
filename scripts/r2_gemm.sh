#!/bin/bash
mkdir -p gpurun_out
OPTEX_CTA_GROUP=1 timeout 180 python scripts/gemm_cta_times.py > gpurun_out/r2_gemm_cg1.txt 2>&1
timeout 180 python scripts/gemm_cta_times.py > gpurun_out/r2_gemm_cg2.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_gemm_cg2.txt
tail -12 gpurun_out/r2_gemm_cg1.txt; tail -14 gpurun_out/r2_gemm_cg2.txt
