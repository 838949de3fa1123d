# multi-GPU bench line: gpurun --gpus N --timeout 900 -- 'NP=N TAG=r4d bash tools/gpu_multi.sh'
cd $GRAFT_REPO_ROOT
NP=${NP:-2}
TAG=${TAG:-multi}
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NP --steps 3 --warmup 3 ${BENCH_ARGS} > gpurun_out/${TAG}_bench_${NP}gpu.json 2> gpurun_out/${TAG}_bench_${NP}gpu.err
tail -c 3000 gpurun_out/${TAG}_bench_${NP}gpu.err; head -c 3000 gpurun_out/${TAG}_bench_${NP}gpu.json
