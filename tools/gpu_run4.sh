# Multi-GPU bench line (four B200): /usr/local/graft/bin/gpurun --gpus 4 --timeout 400 -- "bash tools/gpu_run4.sh"
set -x
cd $GRAFT_REPO_ROOT
free -g | head -2 > gpurun_out/r2m_mem.txt; nproc >> gpurun_out/r2m_mem.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2m_bench_4gpu.json 2> gpurun_out/r2m_bench_4gpu.err
tail -3 gpurun_out/r2m_bench_4gpu.err
