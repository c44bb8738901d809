#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests" ; timeout 1200 python -m pytest tests -m gpu -q -x -s > gpurun_out/test_all.log 2>&1 ; echo "rc=$?" ; grep -E "passed|failed|Error|error" gpurun_out/test_all.log | tail -6
b() { timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], [x.get('ms_per_step') for x in d['roofline_secondary']])"; }
echo "=== bench default"; b
echo "=== bench default again"; b
echo "=== traces"; (cd so-vits-svc_b200/csrc/build; for c in "32 11" "32 3" "16 7"; do timeout 60 ./bench_rb $c | head -9; done; for c in "128 11 3" "128 3 1" "128 7 5 1" "64 7 3"; do timeout 60 ./bench_pairtrace $c | head -3; done) > gpurun_out/traces_r17.log 2>&1; grep -E "ms per launch|avg cycles" gpurun_out/traces_r17.log | cut -c1-200
