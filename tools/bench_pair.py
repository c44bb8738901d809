"""Microbenchmark + unit check of the tensor-core ResBlock pair kernel for every (stage, k, dilation) and tile
variant, at BASELINE config-2 sizes.  Prints a table and writes gpurun_out/bench_pair.json.

    python tools/bench_pair.py [--batch 8 --frames 862 --iters 5 --variants 0,1]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sovits_b200  # noqa: E402,F401
from sovits_b200 import synth  # noqa: E402
from sovits_b200.config import load_config  # noqa: E402
from sovits_b200.engine import TailEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--frames", type=int, default=862)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--variants", default="0,1")
    ap.add_argument("--stages", default="0,1,2,3,4")
    ap.add_argument("--check", type=int, default=1)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = load_config()
    eng = TailEngine(cfg, dev, "tc")
    eng.load_state_dict(synth.synth_state_dict(cfg))
    variants = [int(v) for v in a.variants.split(",")]
    rows = []
    peak = 1461.8
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk)).get("bf16_tflops_sustained", peak)
    L = a.frames
    for i, u in enumerate(cfg.upsample_rates):
        L *= u
        if str(i) not in a.stages.split(","):
            continue
        C = cfg.stage_channels[i]
        g = torch.Generator(device="cpu").manual_seed(100 + i)
        x = (torch.randn((a.batch, C, L), generator=g) * 1.3).to(dev)
        for j, k in enumerate(cfg.resblock_kernel_sizes):
            for d, dil in enumerate(cfg.resblock_dilation_sizes[j]):
                ref = eng.debug_pair(i, j, d, x, -2) if a.check else None
                flops = 2 * 2 * C * C * k * L * a.batch
                for v in variants:
                    out = eng.debug_pair(i, j, d, x, v)
                    err = float((out - ref).abs().max()) if ref is not None else float("nan")
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(a.iters):
                        eng.debug_pair(i, j, d, x, v, out=out)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / a.iters
                    tf = flops / (ms * 1e-3) / 1e12
                    row = {"stage": i, "C": C, "L": L, "k": k, "dil": dil, "variant": v, "ms": ms, "tflops": tf,
                           "frac_of_peak": tf / peak, "linf_vs_fp32": err,
                           "gbs_algorithmic": 3 * C * L * a.batch * 4 / (ms * 1e-3) / 1e9}
                    rows.append(row)
                    print(f"stage{i} C={C:3d} k={k:2d} d={dil} v={v}: {ms:7.3f} ms  {tf:7.1f} TF/s ({tf / peak:5.1%})  "
                          f"{row['gbs_algorithmic']:6.0f} GB/s  linf {err:.2e}", flush=True)
        if C <= 64:
            for j, k in enumerate(cfg.resblock_kernel_sizes):
                for v in variants:
                    out = eng.debug_resblock(i, j, x, v)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(a.iters):
                        eng.debug_resblock(i, j, x, v, out=out)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / a.iters
                    flops = 3 * 2 * 2 * C * C * k * L * a.batch
                    tf = flops / (ms * 1e-3) / 1e12
                    rows.append({"stage": i, "C": C, "L": L, "k": k, "fused": True, "variant": v, "ms": ms, "tflops": tf,
                                 "frac_of_peak": tf / peak, "gbs_algorithmic": 2 * C * L * a.batch * 4 / (ms * 1e-3) / 1e9})
                    print(f"stage{i} C={C:3d} k={k:2d} FUSED-RESBLOCK v={v}: {ms:7.3f} ms  {tf:7.1f} TF/s ({tf / peak:5.1%})", flush=True)
        del x
    tot = {}
    for r in rows:
        key = ("fused" if r.get("fused") else "pair", r["variant"])
        tot[str(key)] = tot.get(str(key), 0.0) + r["ms"]
    print("sum of pair ms per step by variant:", tot)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"rows": rows, "total_ms": tot, "peak_tflops": peak}, open(os.path.join(ROOT, "gpurun_out", "bench_pair.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
