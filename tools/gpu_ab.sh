# A/B of an environment knob: gpurun -- 'TAG=x VAR=GDK_BW2D_THREADS VALS="256 512" bash tools/gpu_ab.sh'
cd $GRAFT_REPO_ROOT
for v in $VALS; do
  env $VAR=$v timeout 600 python bench.py --no-cpu --no-e2e ${BENCH_ARGS} > gpurun_out/${TAG}_${v}.json 2> gpurun_out/${TAG}_${v}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_${v}.json').read().strip().splitlines()[-1])
print('$VAR=$v', round(d['ms_per_step'],2), d['phases_ms'])
PY
done
