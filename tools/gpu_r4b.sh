cd $GRAFT_REPO_ROOT
python tools/stats_probe.py 4000000 64 > gpurun_out/r4b_stats64.json 2>&1
python tools/stats_probe.py 4000000 128 4 > gpurun_out/r4b_stats128.json 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_stats_fused -s 2 -c 1 -o /tmp/r4b_stats python tools/stats_probe.py 4000000 64 > gpurun_out/r4b_ncu.log 2>&1
ncu -i /tmp/r4b_stats.ncu-rep --page details > gpurun_out/r4b_stats_details.txt 2>/dev/null
ncu -i /tmp/r4b_stats.ncu-rep --page source --csv > gpurun_out/r4b_stats_source.csv 2>/dev/null
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 > gpurun_out/r4b_tests.log
cat gpurun_out/r4b_stats64.json gpurun_out/r4b_stats128.json; tail -3 gpurun_out/r4b_tests.log
