#!/bin/bash
# Round-end evidence, short form (a few GPU-minutes): bench lines of every workload, the CUPTI timeline of graph-replayed
# steps, the ncu launch list of one eager step and one `ncu --set full` capture of the optimiser kernel.  tools/final_profile.sh
# is the long form (full captures of every kernel family).  Outputs in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in ${BENCH_WORKLOADS:-cond_grid sample vae cond_grid1024 cond256}; do
  timeout 170 python bench.py --workload $w --steps 30 --warmup 5 > gpurun_out/bench_final_$w.log 2>&1
  grep "^{" gpurun_out/bench_final_$w.log > gpurun_out/bench_final_$w.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_final_$w.json')); print('$w', round(d['value'],1), d['unit'], round(d['ms_per_step'],4), 'ms/step  e2e', round(d['e2e']['value'],1))"
done
timeout 120 python tools/timeline.py > gpurun_out/timeline_final.log 2>&1
grep -E "step span|concurrency" gpurun_out/timeline_final.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_final.csv \
    python tools/ncu_step.py > gpurun_out/ncu_launches_final.log 2>&1
timeout 200 ncu --set full --clock-control none --profile-from-start off -k regex:adam_multi_kernel -c 1 -f \
    -o gpurun_out/ncu_full_adam_multi python tools/ncu_step.py > gpurun_out/ncu_full_adam_multi.log 2>&1
ncu -i gpurun_out/ncu_full_adam_multi.ncu-rep --page raw --csv > gpurun_out/ncu_full_adam_multi_raw.csv 2>/dev/null
rm -f gpurun_out/ncu_full_adam_multi.ncu-rep
ls -la gpurun_out/launches_final.csv gpurun_out/ncu_full_adam_multi_raw.csv
