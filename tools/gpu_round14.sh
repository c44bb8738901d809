#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests" ; timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/test_all.log 2>&1 ; echo "rc=$?" ; grep -E "passed|failed|Error|error" gpurun_out/test_all.log | tail -6
b() { timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], [x.get('ms_per_step') for x in d['roofline_secondary']])"; }
echo "=== bench default"; b
echo "=== bench no dephase"; SVB_PAIR_DEPHASE=0 SVB_RB_SKEW=0 b
echo "=== bench no red"; SVB_PAIR_RED=0 b
echo "=== bench no vec4"; SVB_PAIR_VEC4=0 b
echo "=== bench no vec4 no dephase"; SVB_PAIR_VEC4=0 SVB_PAIR_DEPHASE=0 SVB_RB_SKEW=0 b
echo "=== pair trace"; (cd so-vits-svc_b200/csrc/build; for c in "128 11 3" "128 3 1" "128 7 5 1" "64 7 3"; do timeout 60 ./bench_pairtrace $c | head -4; done; for c in "32 11" "32 3" "16 7"; do timeout 60 ./bench_rb $c | head -2; done ) > gpurun_out/bench_pairtrace3.log 2>&1; cut -c1-330 gpurun_out/bench_pairtrace3.log
