#!/bin/bash
# Final evidence run: launch list of the default bench command + full ncu captures of the dominant kernels.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_final.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pair_tc_kernel|resblock_tc_kernel" --launch-skip 48 --launch-count 3 -f -o gpurun_out/prof_pair128 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_pair128.log 2>&1
echo "pair128 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"resblock_tc_kernel" --launch-skip 8 --launch-count 2 -f -o gpurun_out/prof_resblock \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_resblock.log 2>&1
echo "resblock rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nsf_source_kernel|convn_tc_kernel" --launch-skip 43 --launch-count 8 -f -o gpurun_out/prof_misc \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_misc.log 2>&1
echo "misc rc=$?"
ls -la gpurun_out/*.ncu-rep
