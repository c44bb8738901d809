#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests" ; timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/test_all.log 2>&1 ; echo "rc=$?" ; grep -E "passed|failed|Error|error" gpurun_out/test_all.log | tail -6
echo "=== bench" ; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_tc.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline_secondary'])"
echo "=== pair trace"; (cd so-vits-svc_b200/csrc/build; for c in "128 11 3" "128 3 1" "128 7 5 1" "64 7 3" "256 7 3"; do timeout 60 ./bench_pairtrace $c | head -4; done) > gpurun_out/bench_pairtrace2.log 2>&1; cut -c1-330 gpurun_out/bench_pairtrace2.log
