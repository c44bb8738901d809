#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests" ; timeout 1200 python -m pytest tests -m gpu -q -s -x > gpurun_out/test_all.log 2>&1 ; echo "rc=$?" ; grep -E "parity|passed|failed|Error|error" gpurun_out/test_all.log | tail -40
echo "=== bench tc" ; timeout 600 python bench.py --precision tc --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_tc.log | cut -c1-300; tail -1 gpurun_out/bench_tc.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline'], d['roofline_secondary'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_tc.csv \
    python bench.py --precision tc --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_tc.log 2>&1
echo "launch list rc=$?"
