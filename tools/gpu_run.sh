set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2g_tests.log
timeout 600 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
ls -la gpurun_out
