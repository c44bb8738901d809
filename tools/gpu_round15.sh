#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests" ; timeout 1200 python -m pytest tests -m gpu -q -x -s > gpurun_out/test_all.log 2>&1 ; echo "rc=$?" ; grep -E "passed|failed|Error|error" gpurun_out/test_all.log | tail -6; grep "schedule" gpurun_out/test_all.log
b() { timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], [x.get('ms_per_step') for x in d['roofline_secondary']])"; }
echo "=== bench default (tma)"; b
echo "=== bench tma=0 vec4 red"; SVB_TC_TMA=0 b
echo "=== bench tma=0 vec4 nored"; SVB_TC_TMA=0 SVB_PAIR_RED=0 b
echo "=== bench tma=0 scalar loader"; SVB_TC_TMA=0 SVB_PAIR_VEC4=0 b
echo "=== bench tma, no dephase"; SVB_PAIR_DEPHASE=0 SVB_RB_SKEW=0 b
