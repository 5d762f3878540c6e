#!/bin/bash
# scratch: adam_multi tests + timing
timeout 300 python -m pytest tests/test_kernels_r02_gpu.py -x -q -k "adam_multi_direct or fused_tail_equals" -s > gpurun_out/adam_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/adam_tests.log
grep -E "adam|fused tail|passed|failed|rc=|Error|error" gpurun_out/adam_tests.log | tail -20
timeout 120 python tools/adam_bench.py 2>&1 | tail -4
SVRS_ADAM_BULK=0 timeout 120 python tools/adam_bench.py 2>&1 | tail -1
