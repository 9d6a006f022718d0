# round 2: launch list + full capture of the cdf step (GEMMs, matcher) at the headline shape; source-level CSV of the matcher
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_cdf.csv python scripts/prof_step.py cdf auto 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rotate_gemm|cdf_|split_fill" -s 5 -c 5 -o gpurun_out/r02_full_cdf -f python scripts/prof_step.py cdf auto 2 > gpurun_out/full_cdf.log 2>&1
ncu -i gpurun_out/r02_full_cdf.ncu-rep --page raw --csv > gpurun_out/r02_raw_cdf.csv
ncu -i gpurun_out/r02_full_cdf.ncu-rep --page source --csv -k regex:cdf_channel > gpurun_out/r02_source_cdf_channel.csv 2>/dev/null
ls -la gpurun_out | tail -8
