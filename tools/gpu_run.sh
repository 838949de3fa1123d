set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_2d.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2a_tests.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2a_bench_sorted.json 2> gpurun_out/r2a_bench_sorted.err
GDK_SORTED=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2a_bench_hot.json 2> gpurun_out/r2a_bench_hot.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_hist2d_sorted|k_bucket_scatter|k_bin8c|k_bin8_rowmajor' -c 4 -o gpurun_out/r2a_sorted python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2a_ncu.log 2>&1
ls -la gpurun_out
