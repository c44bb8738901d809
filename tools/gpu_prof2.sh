#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"resblock_tc_kernel|pair_tc_kernel" --launch-skip 7 --launch-count 7 -f -o gpurun_out/prof_set2 \
    python tools/prof_kernels.py > gpurun_out/ncu_set2.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_set2.log; ls -la gpurun_out/*.ncu-rep
