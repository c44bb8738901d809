#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 ; echo "rc=$?" ; tail -3 gpurun_out/smoke.log
echo "=== tests" ; timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/test_all.log 2>&1 ; echo "rc=$?" ; grep -E "passed|failed|Error|error" gpurun_out/test_all.log | tail -5; grep -E "full" gpurun_out/test_all.log | grep parity
echo "=== bench tc (full default run incl. cpu baseline)" ; timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_default.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e'], d['clocks'], d['roofline'], d['roofline_secondary'], d['cpu_baseline'], d['gpu_launches'])"
echo "=== bench reference arm" ; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_ref.log | cut -c1-700
NG=$(nvidia-smi -L | wc -l); echo "gpus visible: $NG"
if [ "$NG" -ge 2 ]; then
echo "=== bench 2 GPUs" ; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_2gpu.log | cut -c1-400
fi
