#!/bin/bash
# scratch: full GPU suite + default bench line
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_full.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/gpu_tests_full.log
timeout 200 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_cur.log 2>&1
grep "^{" gpurun_out/bench_cur.log > gpurun_out/bench_cur.json
python - <<'P'
import json
d = json.load(open("gpurun_out/bench_cur.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
print(d["roofline_hbm"]["kernels"]["adam_multi"])
P
