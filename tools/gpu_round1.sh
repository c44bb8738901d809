#!/bin/bash
# First GPU pass: descriptor probe, fp32 parity, fp32 bench, PyTorch-CUDA reference timing, then the tensor-core path.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== probe" ; timeout 120 ./so-vits-svc_b200/csrc/build/probe_tc > gpurun_out/probe.log 2>&1 ; echo "probe rc=$?" ; tail -3 gpurun_out/probe.log
echo "=== fp32 tests" ; timeout 900 python -m pytest tests -m gpu -q -s -k "nsf or flow or fp32 or error" > gpurun_out/test_fp32.log 2>&1 ; echo "rc=$?" ; tail -15 gpurun_out/test_fp32.log
echo "=== bench fp32" ; timeout 600 python bench.py --precision fp32 --steps 3 --warmup 3 > gpurun_out/bench_fp32.log 2>&1 ; echo "rc=$?" ; tail -2 gpurun_out/bench_fp32.log
echo "=== ref cuda" ; timeout 600 python tools/ref_cuda_baseline.py --steps 3 > gpurun_out/ref_cuda.log 2>&1 ; echo "rc=$?" ; tail -2 gpurun_out/ref_cuda.log
echo "=== tc tests" ; timeout 900 python -m pytest tests -m gpu -q -s -k "tc or fixture or full" > gpurun_out/test_tc.log 2>&1 ; echo "rc=$?" ; tail -30 gpurun_out/test_tc.log
echo "=== bench tc" ; timeout 600 python bench.py --precision tc --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.log 2>&1 ; echo "rc=$?" ; tail -2 gpurun_out/bench_tc.log
