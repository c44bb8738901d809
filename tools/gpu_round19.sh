#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests" ; timeout 1200 python -m pytest tests -m gpu -q -x -s > gpurun_out/test_all.log 2>&1 ; echo "rc=$?" ; grep -E "passed|failed|Error|error|enc_p" gpurun_out/test_all.log | tail -6
b() { timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], [x.get('ms_per_step') for x in d['roofline_secondary']])"; }
echo "=== bench default (fused prefix tails)"; b
echo "=== bench SVB_PREFIX_FUSED=0"; SVB_PREFIX_FUSED=0 b
