# Multi-GPU bench line (two B200): /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- "bash tools/gpu_run2.sh"
set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2h_bench_2gpu.json 2> gpurun_out/r2h_bench_2gpu.err
tail -5 gpurun_out/r2h_bench_2gpu.err
