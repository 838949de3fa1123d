# one ncu --set full capture of kernel $K (regex) from a short bench run; summaries into gpurun_out/${TAG}_*
cd $GRAFT_REPO_ROOT
TAG=${TAG:-ncu}
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"$K" -s ${SKIP:-0} -c ${COUNT:-1} -o /tmp/${TAG} python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}.ncu-rep --page details > gpurun_out/${TAG}_details.txt 2>/dev/null
ncu -i /tmp/${TAG}.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
ncu -i /tmp/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
tail -3 gpurun_out/${TAG}_ncu.log | cut -c1-300
