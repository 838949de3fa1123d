set -x
cd $GRAFT_REPO_ROOT
rm -f gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests/test_gpu_2d.py tests/test_gpu_stats_1d.py -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2k_tests.log
timeout 600 python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2k_launches_ncu.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2k_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'k_shear_hist|k_hist2d_records|k_bucket_records|k_conv2d|k_shear_minmax_tiled|k_bw2d|k_xform_rows|k_xform_cols|k_qhist|k_hist1d_tma|k_bin8c' -c 22 -o /tmp/r2k_full python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2k_ncu_full.log 2>&1
ncu -i /tmp/r2k_full.ncu-rep --page raw --csv > gpurun_out/r2k_full_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8; du -sh gpurun_out
