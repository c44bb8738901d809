#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
b() { timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], [x.get('ms_per_step') for x in d['roofline_secondary']])"; }
echo "=== default"; b
echo "=== fuse_maxc=64"; SVB_FUSE_MAXC=64 b
echo "=== tc variant 0 (1 CTA/SM big tiles)"; SVB_TC_VARIANT=0 b
echo "=== rb variant 0"; SVB_RB_VARIANT=0 b
echo "=== pair dephase 20000"; SVB_PAIR_DEPHASE=20000 b
echo "=== pair dephase 70000"; SVB_PAIR_DEPHASE=70000 b
echo "=== rb skew 25000"; SVB_RB_SKEW=25000 b
echo "=== rb skew 70000"; SVB_RB_SKEW=70000 b
