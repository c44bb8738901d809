#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests" ; timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/test_all.log 2>&1 ; echo "rc=$?" ; grep -E "passed|failed|Error|error|snake" gpurun_out/test_all.log | tail -8
echo "=== snake bench (B=8, fp32 path + SnakeAlias)"; timeout 600 python - > gpurun_out/snake_bench.log 2>&1 <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import sovits_b200
from sovits_b200 import models, synth
from sovits_b200.config import load_config
cfg = load_config(); cfg.vocoder_name = "nsf-snake-hifigan"
kw = json.load(open(sovits_b200.DEFAULT_CONFIG))["model"]; kw["vocoder_name"] = "nsf-snake-hifigan"
dev = torch.device("cuda:0")
net = models.SynthesizerTrn(1025, 20, **kw).eval(); net.load_state_dict(synth.synth_state_dict(cfg)); net = net.to(dev)
B, T = 8, 862
c, f0, uv, sid = [t.to(dev) for t in synth.synth_inputs(cfg, B, T)]
for _ in range(2): o = net.infer(c, f0, uv, g=sid, noice_scale=0.4)[0]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): o = net.infer(c, f0, uv, g=sid, noice_scale=0.4)[0]
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(json.dumps({"config": "config4: nsf-snake-hifigan, batch 8 x 862 frames", "ms_per_step": ms, "samples_per_s": B * T * 512 / (ms * 1e-3), "finite": bool(torch.isfinite(o).all())}))
PY
echo "rc=$?"; tail -2 gpurun_out/snake_bench.log
echo "=== ncu launch list (default bench command)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_default.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
echo "=== ncu full: pair (C=128 k=11) + fused resblock (C=32) + convn"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"pair_tc_kernel|resblock_tc_kernel" --launch-skip 33 --launch-count 6 -f -o gpurun_out/prof_final \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1
echo "full rc=$?"; ls -la gpurun_out/*.ncu-rep
