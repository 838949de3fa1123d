set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2d_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
ls -la gpurun_out
