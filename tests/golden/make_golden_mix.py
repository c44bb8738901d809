"""Golden vector for speaker-mix inference (`EnableCharacterMix`, models.py:456-461,505-509) from the
UNMODIFIED reference: time-varying conditioning g[1,768,T] through flow and generator.  Build container only:

    python tests/golden/make_golden_mix.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (sets up the reference import with stubbed audio modules)

cfg = MG.load_config()
sd = MG.synth.synth_state_dict(cfg)
net = MG.build_reference(cfg, sd)
net.EnableCharacterMix(cfg.n_speakers, "cpu")
T = 26
c, f0, uv, _ = MG.synth.synth_inputs(cfg, 1, T, seed=4242)
gen = torch.Generator().manual_seed(99)
mix = torch.softmax(torch.randn((T, cfg.n_speakers), generator=gen) * 1.5, dim=-1)      # [N frames, S speakers]
taps = {}


def _grab(module, args, kwargs):
    taps["z_p"] = args[0].detach().clone()
    taps["g"] = kwargs["g"].detach().clone()


h1 = net.flow.register_forward_pre_hook(_grab, with_kwargs=True)
h2 = net.flow.register_forward_hook(lambda m, i, o: taps.__setitem__("z", o.detach().clone()))
with torch.no_grad():
    o, _ = net.infer(c, f0=f0, uv=uv, g=mix, noice_scale=0.4)
h1.remove(); h2.remove()
print("speaker mix", o.shape, taps["g"].shape, float(o.abs().max()))
np.savez_compressed(os.path.join(HERE, "ref_infer_mix_t26.npz"), o=o.numpy(), z_p=taps["z_p"].numpy(), z=taps["z"].numpy(), g=taps["g"].numpy(),
                    mix=mix.numpy(), c=c.numpy(), f0=f0.numpy(), uv=uv.numpy(), T=T, seed=52468, noice_scale=0.4)
