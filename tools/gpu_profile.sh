#!/bin/bash
# ncu launch list of one bench step + a full capture of the dominant kernel (stage-1 ResBlock pair, C=128, k=11).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
PREC=${1:-tc}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_${PREC}.csv \
    python bench.py --precision $PREC --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_${PREC}.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_tc_kernel -s 60 -c 3 -f -o gpurun_out/prof_pair \
    python bench.py --precision tc --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out
