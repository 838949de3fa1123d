# Evidence run for profiles/ (one B200): the default bench line, the ncu launch list of the same command and one full
# capture of every data-path kernel, exported as CSV (the .ncu-rep stays on the box: gpurun copies back at most 64 MiB).
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'TAG=r4 bash tools/gpu_profile.sh'
cd $GRAFT_REPO_ROOT
TAG=${TAG:-run}
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches_ncu.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_shear_hist_w|k_hist2d_records|k_bucket_records|k_conv2d|k_shear_minmax_tma|k_bw2d|k_xform_rows|k_xform_cols|k_qhist|k_hist1d_tma|k_bin8c|k_contours2d|k_stats_fused|k_kde1d' -c 26 -o /tmp/${TAG}_full python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
python tools/peaks.py > gpurun_out/${TAG}_peaks.json 2>/dev/null
du -sh gpurun_out; tail -c 400 gpurun_out/${TAG}_bench.err; head -c 300 gpurun_out/${TAG}_bench.json
