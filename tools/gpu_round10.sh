#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests (TMA on)" ; timeout 1200 python -m pytest tests -m gpu -q -s -x > gpurun_out/test_all.log 2>&1 ; echo "rc=$?" ; grep -E "passed|failed|Error|error" gpurun_out/test_all.log | tail -5; grep -E "tc stage|full|b2_t24 tc|b1_t33 tc" gpurun_out/test_all.log | grep parity
for M in 1 0; do
echo "=== bench TMA=$M" ; SVB_TC_TMA=$M timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tma$M.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_tma$M.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline_secondary'])"
done
