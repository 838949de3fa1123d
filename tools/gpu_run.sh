set -x
cd $GRAFT_REPO_ROOT
timeout 600 python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
tail -2 gpurun_out/r2l_bench.err
