# round 2, second session: full captures of the new kernels - the blocked-order Jacobi solver (pca.cu), the cooperative
# covariance chain (cov_chain.cu) and the narrow covariance path (cov_small.cu) - plus the launch list of a pca loop
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches_pca_loop.csv python scripts/time_loop_layers.py pca 4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pca_jacobi_blk" -s 1 -c 2 -o gpurun_out/r02b_full_jacobi -f python scripts/pca_stamps.py 512 4096 > gpurun_out/full_jacobi.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ns_coop|small_" -s 30 -c 12 -o gpurun_out/r02b_full_cov -f python scripts/time_loop_layers.py pca 4 > gpurun_out/full_cov.log 2>&1
ncu -i gpurun_out/r02b_full_jacobi.ncu-rep --page raw --csv > gpurun_out/r02b_raw_jacobi.csv
ncu -i gpurun_out/r02b_full_cov.ncu-rep --page raw --csv > gpurun_out/r02b_raw_cov.csv
ls -la gpurun_out | tail -8
