set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_2d.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2i_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
timeout 600 python bench.py --n 1000000 --p 256 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2i_bench_p256.json 2> gpurun_out/r2i_bench_p256.err
tail -3 gpurun_out/r2i_bench_p256.err
ls -la gpurun_out
