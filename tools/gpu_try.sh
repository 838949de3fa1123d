# quick GPU check: tests + default bench line.  gpurun --timeout 1500 -- 'TAG=r4a bash tools/gpu_try.sh'
cd $GRAFT_REPO_ROOT
TAG=${TAG:-try}
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 > gpurun_out/${TAG}_tests.log
timeout 900 python bench.py ${BENCH_ARGS} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -5 gpurun_out/${TAG}_tests.log; tail -c 600 gpurun_out/${TAG}_bench.err; head -c 1500 gpurun_out/${TAG}_bench.json
