#!/bin/bash
# Round-1 evidence run: GPU tests, smoke, the default bench line (with the CPU baseline leg), the launch list of the bench
# command and full ncu captures of the dominant kernels.  Outputs under gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== tests"; timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/test_final.log 2>&1; echo "rc=$?"; grep -E "passed|failed" gpurun_out/test_final.log | tail -2
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "=== bench"; timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_default.log | cut -c1-600
echo "=== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_final2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_final2.log 2>&1; echo "rc=$?"
echo "=== full captures"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pair_tc_kernel" --launch-skip 31 --launch-count 3 -f -o gpurun_out/prof_pair_final \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_pair_final.log 2>&1; echo "pair rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"resblock_tc_kernel" --launch-skip 9 --launch-count 9 -f -o gpurun_out/prof_resblock_final \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_resblock_final.log 2>&1; echo "resblock rc=$?"
ls -la gpurun_out/*final*.ncu-rep
