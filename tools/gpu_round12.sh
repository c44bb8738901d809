#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests" ; timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/test_all.log 2>&1 ; echo "rc=$?" ; grep -E "passed|failed|Error|error" gpurun_out/test_all.log | tail -6
echo "=== bench" ; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_tc.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline_secondary'])"
echo "=== rb trace"; (cd so-vits-svc_b200/csrc/build; for sk in 0 2500; do echo "== skew $sk"; for c in "32 11" "32 3" "16 7"; do SVB_RB_SKEW=$sk timeout 60 ./bench_rb $c | head -9; done; done) > gpurun_out/bench_rb_tight.log 2>&1; grep -E "ms per launch|skew" gpurun_out/bench_rb_tight.log
