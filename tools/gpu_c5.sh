# C5 (BASELINE.json configs[4]: P=256 triangle, 32640 2D densities) on NP GPUs:
#   gpurun --gpus 8 --timeout 1200 -- 'NP=8 TAG=r4z bash tools/gpu_c5.sh'
cd $GRAFT_REPO_ROOT
NP=${NP:-8}
TAG=${TAG:-c5}
free -g | head -2 > gpurun_out/${TAG}_mem.txt; nproc >> gpurun_out/${TAG}_mem.txt; df -h /dev/shm | tail -1 >> gpurun_out/${TAG}_mem.txt
AVAIL=$(free -g | awk '/Mem:/{print $7}')
if [ "$AVAIL" -lt $((NP * 30 + 60)) ]; then echo "not enough host memory ($AVAIL GB) for $NP ranks of C5" | tee gpurun_out/${TAG}_bench_${NP}gpu.err; exit 0; fi
OMP_NUM_THREADS=8 timeout 1000 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NP --workload c5 --steps 2 --warmup 1 --no-cpu ${BENCH_ARGS} > gpurun_out/${TAG}_bench_${NP}gpu.json 2> gpurun_out/${TAG}_bench_${NP}gpu.err
cat gpurun_out/${TAG}_mem.txt; tail -c 1500 gpurun_out/${TAG}_bench_${NP}gpu.err; head -c 1200 gpurun_out/${TAG}_bench_${NP}gpu.json
