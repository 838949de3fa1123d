# full ncu capture of the data-path kernels of one step (the statistics sweep separately, from tools/stats_probe.py)
cd $GRAFT_REPO_ROOT
TAG=${TAG:-run}
timeout 900 ncu --set full --clock-control none -k regex:'k_shear_hist_w|k_hist2d_records|k_bucket_records|k_conv2d|k_shear_minmax_tma|k_bw2d|k_xform_rows|k_xform_cols|k_qhist|k_hist1d_tma|k_bin8c|k_contours2d|k_kde1d' -c 30 -o /tmp/${TAG}_full python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:'k_stats_fused' -s 40 -c 1 -o /tmp/${TAG}_stats python tools/stats_probe.py 10000000 64 > gpurun_out/${TAG}_ncu_stats.log 2>&1
ncu -i /tmp/${TAG}_stats.ncu-rep --page raw --csv > gpurun_out/${TAG}_stats_raw.csv 2>/dev/null
python tools/stats_probe.py 10000000 128 4 > gpurun_out/${TAG}_stats_c4.json 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log | cut -c1-200; cat gpurun_out/${TAG}_stats_c4.json
