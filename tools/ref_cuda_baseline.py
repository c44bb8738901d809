"""Times the reference path as plain PyTorch on the B200 (the denominator of BASELINE.json's ">= 5x the
reference's own PyTorch-CUDA infer" target).  /root/reference is not available on the GPU box, so this runs the
oracle port — the same torch ops (F.conv1d / conv_transpose1d / cumsum / sin through cuDNN + ATen, weight-norm
recomputed every forward like the reference does) — on cuda:0, with cuDNN's default TF32 convolutions and with
fp32_precision='ieee'.  Also reports the reference's own TF32-vs-fp32 self-disagreement (T2).

    python tools/ref_cuda_baseline.py [--batch 8 --frames 862 --steps 5]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sovits_b200  # noqa: E402,F401
import svc_oracle as O  # noqa: E402
from sovits_b200 import synth  # noqa: E402
from sovits_b200.config import load_config  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--frames", type=int, default=862)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = load_config()
    sd = {k: v.to(dev) for k, v in synth.synth_state_dict(cfg).items()}
    B, T = a.batch, a.frames
    c, f0, uv, sid = [t.to(dev) for t in synth.synth_inputs(cfg, B, T)]
    N = T * cfg.hop
    torch.manual_seed(52468)
    noise = {"z_noise": torch.randn(B, cfg.inter_channels, T, device=dev), "rand_ini": torch.rand(B, cfg.n_harmonics, device=dev),
             "har_noise": torch.randn(B, N, cfg.n_harmonics, device=dev)}
    res = {}
    outs = {}
    for mode in ("tf32", "ieee"):
        torch.backends.cudnn.conv.fp32_precision = mode
        torch.backends.cudnn.benchmark = True
        for _ in range(2):
            o, _ = O.infer(sd, cfg, c, f0, uv, sid, noise, noice_scale=0.4)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            o, _ = O.infer(sd, cfg, c, f0, uv, sid, noise, noice_scale=0.4)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        res[mode] = {"ms_per_step": ms, "samples_per_s": B * N / (ms * 1e-3)}
        outs[mode] = o
    res["tf32_vs_ieee_linf"] = float((outs["tf32"] - outs["ieee"]).abs().max())
    res["config"] = {"batch": B, "frames": T, "what": "oracle port on cuda:0 = the reference's PyTorch ops (cuDNN/ATen)"}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
