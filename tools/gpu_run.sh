# Evidence run for profiles/ (one B200): GPU tests, the default bench line, the ncu launch list and a full capture of
# the data-path kernels exported as CSV (the .ncu-rep stays on the box: gpurun copies back at most 64 MiB).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_run.sh'
set -x
cd $GRAFT_REPO_ROOT
TAG=${TAG:-run}
rm -f gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/${TAG}_tests.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_ncu.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'k_shear_hist|k_hist2d_records|k_bucket_records|k_conv2d|k_shear_minmax_tiled|k_bw2d|k_xform_rows|k_xform_cols|k_qhist|k_hist1d_tma|k_bin8c' -c 22 -o /tmp/${TAG}_full python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
du -sh gpurun_out
