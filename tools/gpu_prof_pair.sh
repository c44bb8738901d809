#!/bin/bash
# Full ncu capture (with source-level stall sampling) of a few C=128 pair launches of the default bench step.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pair_tc_kernel" --launch-skip 11 --launch-count 5 -f -o gpurun_out/prof_pair_v2 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_pair_v2.log 2>&1
echo "rc=$?"; ls -la gpurun_out/prof_pair_v2.ncu-rep
