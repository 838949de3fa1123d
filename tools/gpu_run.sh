cd $GRAFT_REPO_ROOT
timeout 30 python -m pytest tests/test_gpu_stats_1d.py -x -q -m gpu -k "marge_limits" 2>&1 | tail -5 > gpurun_out/r2n_tests.log
