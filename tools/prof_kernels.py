"""Launch a few chosen kernels a few times each (for `ncu -k regex:... -c N` captures)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sovits_b200  # noqa: E402,F401
from sovits_b200 import synth  # noqa: E402
from sovits_b200.config import load_config  # noqa: E402
from sovits_b200.engine import TailEngine  # noqa: E402

dev = torch.device("cuda:0")
cfg = load_config()
eng = TailEngine(cfg, dev, "tc")
eng.load_state_dict(synth.synth_state_dict(cfg))
B, T = 8, 862
L = T
xs = {}
for i, u in enumerate(cfg.upsample_rates):
    L *= u
    xs[i] = (torch.randn((B, cfg.stage_channels[i], L), device=dev) * 1.3, L)
# the SHIPPED variants: block-skewed fused ResBlock (variant 2), pair kernel tile variant 1 (two CTAs/SM where the tile allows)
for rep in range(2):
    eng.debug_resblock(4, 0, xs[4][0], 2)      # C=16 k=3   (epilogue-bound)
    eng.debug_resblock(4, 2, xs[4][0], 2)      # C=16 k=11
    eng.debug_resblock(3, 1, xs[3][0], 2)      # C=32 k=7
    eng.debug_resblock(2, 1, xs[2][0], 2)      # C=64 k=7
    eng.debug_pair(1, 1, 1, xs[1][0], 1)       # C=128 k=7 d=3
    eng.debug_pair(1, 2, 2, xs[1][0], 1)       # C=128 k=11 d=5
    eng.debug_pair(0, 1, 1, xs[0][0], 1)       # C=256 k=7 d=3
torch.cuda.synchronize()
print("done")
