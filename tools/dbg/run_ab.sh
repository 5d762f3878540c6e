#!/bin/bash
for i in 1 2; do
timeout 100 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-profile 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))"
done
timeout 120 python tools/timeline.py > gpurun_out/timeline_final.log 2>&1
grep -E "step span" gpurun_out/timeline_final.log
grep -A8 "timeline of the step" gpurun_out/timeline_final.log | cut -c1-90
timeout 60 python -m pytest tests/test_model_gpu.py -x -q 2>&1 | tail -1
