#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== fused resblock tests (variants 0,1,2)"; timeout 300 python -m pytest tests -m gpu -q -x -s -k "fused_resblock" > gpurun_out/test_skew.log 2>&1; echo "rc=$?"; grep -E "passed|failed|Error|error" gpurun_out/test_skew.log | tail -3; grep "v2" gpurun_out/test_skew.log | head -9
b() { timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], [x.get('ms_per_step') for x in d['roofline_secondary']])"; }
echo "=== bench default"; b
echo "=== bench SVB_RB_VARIANT=2"; SVB_RB_VARIANT=2 b
