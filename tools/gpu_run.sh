set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_2d.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2c_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_hist2d_records|k_bucket_records' -c 4 -o gpurun_out/r2c_sorted python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2c_ncu.log 2>&1
ls -la gpurun_out
