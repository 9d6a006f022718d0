# run on the GPU box: ncu launch lists of every mode + one full capture of the cdf step and of the sort kernels
set -x
mkdir -p gpurun_out
for m in cdf sort pca chol; do
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_$m.csv python scripts/prof_step.py $m auto 3 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:"rotate_gemm|cdf_|split_fill|rot_" -o gpurun_out/r01_full_cdf -f python scripts/prof_step.py cdf auto 2 > gpurun_out/full_cdf.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sort_" -o gpurun_out/r01_full_sort -f python scripts/prof_step.py sort auto 1 > gpurun_out/full_sort.log 2>&1
ncu -i gpurun_out/r01_full_cdf.ncu-rep --page raw --csv > gpurun_out/raw_cdf.csv
ncu -i gpurun_out/r01_full_sort.ncu-rep --page raw --csv > gpurun_out/raw_sort.csv
ls -la gpurun_out | tail -12
