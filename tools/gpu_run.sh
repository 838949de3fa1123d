set -x
cd $GRAFT_REPO_ROOT
for t in 256 512 768; do
GDK_BW2D_THREADS=$t timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r2j_bench_bw$t.json 2> gpurun_out/r2j_bench_bw$t.err
done
GDK_BW2D_THREADS=768 timeout 600 python -m pytest tests/test_gpu_2d.py -x -q -m gpu -k "density_2d or contour" 2>&1 | tail -4 > gpurun_out/r2j_tests768.log
ls gpurun_out | head -50
