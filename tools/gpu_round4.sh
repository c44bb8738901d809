#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests" ; timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/test_all.log 2>&1 ; echo "rc=$?" ; grep -E "parity|passed|failed|Error|error" gpurun_out/test_all.log | tail -60
echo "=== pair sweep" ; timeout 900 python tools/bench_pair.py --variants 0,1 --check 0 > gpurun_out/bench_pair.log 2>&1 ; echo "rc=$?" ; grep -E "FUSED|sum of|k=11 d=5|k= 3 d=1" gpurun_out/bench_pair.log
for F in 1 0; do
echo "=== bench tc fuse=$F" ; SVB_FUSE_RESBLOCK=$F timeout 600 python bench.py --precision tc --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc_f$F.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_tc_f$F.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline'], d['roofline_secondary'])"
done
