cd $GRAFT_REPO_ROOT
python tools/stats_probe.py 4000000 64 > gpurun_out/r4c_stats64.json 2>&1
python tools/stats_probe.py 4000000 128 4 > gpurun_out/r4c_stats128.json 2>&1
python tools/stats_probe.py 3000000 32 > gpurun_out/r4c_stats32.json 2>&1
cat gpurun_out/r4c_stats*.json
