"""Golden vector for a checkpoint WITH a volume embedding (`vol_embedding=True`: `emb_vol = nn.Linear(1, hidden)`,
models.py:398-399,513) from the UNMODIFIED reference.  Build container only:

    python tests/golden/make_golden_vol.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (sets up the reference import with stubbed audio modules)

cfg = MG.load_config()
sd = MG.synth.synth_state_dict(cfg)
gen = torch.Generator().manual_seed(515)
sd["emb_vol.weight"] = torch.randn((cfg.hidden_channels, 1), generator=gen) * 0.3
sd["emb_vol.bias"] = torch.randn((cfg.hidden_channels,), generator=gen) * 0.1
with open(MG.sovits_b200.DEFAULT_CONFIG) as f:
    model_kw = json.load(f)["model"]
model_kw["vol_embedding"] = True
net = MG.ref_models.SynthesizerTrn(2048 // 2 + 1, 10240 // 512, **model_kw).eval()
own = net.state_dict()
assert all(k in own and own[k].shape == v.shape for k, v in sd.items())
net.load_state_dict(sd, strict=False)
B, T = 2, 22
c, f0, uv, sid = MG.synth.synth_inputs(cfg, B, T, seed=777)
vol = torch.rand((B, T), generator=gen) * 0.2                      # RMS-like volumes (utils.Volume_Extractor range)
taps = {}
h = net.flow.register_forward_pre_hook(lambda m, i: taps.__setitem__("z_p", i[0].detach().clone()))
with torch.no_grad():
    o, _ = net.infer(c, f0=f0, uv=uv, g=sid, noice_scale=0.4, vol=vol)
    h.remove()
    o_novol, _ = net.infer(c, f0=f0, uv=uv, g=sid, noice_scale=0.4, vol=None)
print("vol embedding", o.shape, float(o.abs().max()), "effect of vol on the waveform:", float((o - o_novol).abs().max()))
np.savez_compressed(os.path.join(HERE, "ref_infer_vol_b2_t22.npz"), o=o.numpy(), z_p=taps["z_p"].numpy(), vol=vol.numpy(),
                    emb_vol_w=sd["emb_vol.weight"].numpy(), emb_vol_b=sd["emb_vol.bias"].numpy(),
                    c=c.numpy(), f0=f0.numpy(), uv=uv.numpy(), sid=sid.numpy(), B=B, T=T, seed=52468, noice_scale=0.4)
