#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests" ; timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/test_all.log 2>&1 ; echo "rc=$?" ; grep -E "parity|passed|failed|Error" gpurun_out/test_all.log | tail -20
echo "=== pair sweep" ; timeout 900 python tools/bench_pair.py --variants 0,1 > gpurun_out/bench_pair.log 2>&1 ; echo "rc=$?" ; tail -50 gpurun_out/bench_pair.log
echo "=== bench tc v0" ; SVB_TC_VARIANT=0 timeout 600 python bench.py --precision tc --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc_v0.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_tc_v0.log | cut -c1-600
echo "=== bench tc v1" ; SVB_TC_VARIANT=1 timeout 600 python bench.py --precision tc --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc_v1.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_tc_v1.log | cut -c1-600
