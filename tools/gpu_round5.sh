#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests" ; timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/test_all.log 2>&1 ; echo "rc=$?" ; grep -E "passed|failed|Error|error" gpurun_out/test_all.log | tail -10; grep -E "full|fixture|b2_t24|b1_t33" gpurun_out/test_all.log | grep parity
echo "=== pair sweep" ; timeout 900 python tools/bench_pair.py --variants 0,1 --check 0 > gpurun_out/bench_pair.log 2>&1 ; echo "rc=$?" ; grep -E "FUSED|sum of|k=11 d=5|k= 3 d=1" gpurun_out/bench_pair.log
for M in 32 64; do
echo "=== bench tc fuse maxc=$M" ; SVB_FUSE_MAXC=$M timeout 600 python bench.py --precision tc --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc_m$M.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_tc_m$M.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline'], d['roofline_secondary'])"
done
