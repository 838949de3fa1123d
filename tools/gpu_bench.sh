# bench line only: gpurun --timeout 900 -- 'TAG=x BENCH_ARGS="--no-cpu" bash tools/gpu_bench.sh'
cd $GRAFT_REPO_ROOT
TAG=${TAG:-b}
timeout 800 python bench.py ${BENCH_ARGS} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 1500 gpurun_out/${TAG}_bench.err; head -c 600 gpurun_out/${TAG}_bench.json
