#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests" ; timeout 1200 python -m pytest tests -m gpu -q -x -s > gpurun_out/test_final.log 2>&1 ; echo "rc=$?" ; grep -E "passed|failed|Error|error" gpurun_out/test_final.log | tail -4
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "=== bench"; timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_default.log | cut -c1-400
