"""Generate golden vectors by running the UNMODIFIED reference (imported from /root/reference) on
seeded synthetic weights/inputs.  Run in the build container only (the GPU box has no reference):

    python tests/golden/make_golden.py

Outputs small .npz fixtures next to this file.  Noise is drawn by the reference itself after
``torch.manual_seed(seed)`` (models.py:498-501); ``sovits_b200.synth.draw_noise`` replays the same
draws for the oracle and the CUDA path.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SOVITS_REF_DIR", "/root/reference")
sys.path.insert(0, ROOT)
for m in ("faiss", "librosa", "matplotlib", "matplotlib.pylab"):  # imported at module top, unused by infer
    sys.modules[m] = types.ModuleType(m)
sys.path.insert(0, REF)

import models as ref_models  # noqa: E402
import modules.modules as ref_modules  # noqa: E402
import sovits_b200  # noqa: E402,F401
from sovits_b200 import synth  # noqa: E402
from sovits_b200.config import load_config  # noqa: E402


def build_reference(cfg, sd):
    import json
    with open(sovits_b200.DEFAULT_CONFIG) as f:
        model_kw = json.load(f)["model"]
    model_kw["vocoder_name"] = cfg.vocoder_name
    net = ref_models.SynthesizerTrn(2048 // 2 + 1, 10240 // 512, **model_kw).eval()
    own = net.state_dict()
    missing = [k for k in sd if k not in own]
    assert not missing, missing
    for k, v in sd.items():
        assert own[k].shape == v.shape, (k, own[k].shape, v.shape)
    used = [k for k in own if not k.startswith(("enc_q.", "f0_decoder."))]
    assert sorted(used) == sorted(sd.keys()), set(used) ^ set(sd.keys())
    net.load_state_dict(sd, strict=False)
    return net


def main():
    torch.set_num_threads(8)
    cfg = load_config()
    sd = synth.synth_state_dict(cfg)
    net = build_reference(cfg, sd)
    out = {}
    for name, (B, T) in synth.GOLDEN_CASES.items():
        c, f0, uv, sid = synth.golden_inputs(cfg, name)
        taps = {}
        hooks = []
        hooks.append(net.flow.register_forward_hook(lambda m, i, o: taps.__setitem__("z", o.detach().clone())))
        hooks.append(net.flow.register_forward_pre_hook(lambda m, i: taps.__setitem__("z_p", i[0].detach().clone())))
        hooks.append(net.dec.m_source.register_forward_hook(lambda m, i, o: taps.__setitem__("har", o[0].detach().clone())))
        with torch.no_grad():
            o, f0_out = net.infer(c, f0=f0, uv=uv, g=sid, noice_scale=0.4)
        for h in hooks:
            h.remove()
        print(name, o.shape, float(o.abs().max()), float(taps["z"].abs().max()))
        np.savez_compressed(os.path.join(HERE, f"ref_infer_{name}.npz"),
                            o=o.numpy(), z_p=taps["z_p"].numpy(), z=taps["z"].numpy(),
                            har=taps["har"].transpose(1, 2).numpy(),
                            B=B, T=T, seed=52468, noice_scale=0.4,
                            f0=f0.numpy())
    # Snake-activation generator (vdecoder/hifiganwithsnake, BASELINE config 4)
    cfg_s = load_config()
    cfg_s.vocoder_name = "nsf-snake-hifigan"
    sd_s = synth.synth_state_dict(cfg_s)
    net_s = build_reference(cfg_s, sd_s)
    for name, (B, T) in synth.SNAKE_GOLDEN_CASES.items():
        c, f0, uv, sid = synth.golden_inputs(cfg_s, name)
        taps = {}
        hk = net_s.flow.register_forward_pre_hook(lambda m, i: taps.__setitem__("z_p", i[0].detach().clone()))
        with torch.no_grad():
            o, _ = net_s.infer(c, f0=f0, uv=uv, g=sid, noice_scale=0.4)
        hk.remove()
        print(name, o.shape, float(o.abs().max()))
        np.savez_compressed(os.path.join(HERE, f"ref_infer_{name}.npz"), o=o.numpy(), z_p=taps["z_p"].numpy(),
                            B=B, T=T, seed=52468, noice_scale=0.4, f0=f0.numpy())
    # mel-conditioned vocoder vdecoder/nsf_hifigan: RNG seeded with 777, draws rand(B,9) then randn(B,N,9)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    from vdecoder.nsf_hifigan import models as voc_models
    from vdecoder.nsf_hifigan.env import AttrDict
    from sovits_b200 import nsf_hifigan
    h = AttrDict(synth.VOCODER_H)
    vcfg = nsf_hifigan.cfg_from_h(h)
    vsd = synth.synth_vocoder_state_dict(vcfg)
    G = voc_models.Generator(h).eval()
    assert set(G.state_dict()) == set(vsd)
    G.load_state_dict(vsd)
    B, T = 2, 21
    mel, f0 = synth.synth_vocoder_inputs(vcfg, B, T)
    torch.manual_seed(777)
    with torch.no_grad():
        o = G(mel, f0)
    print("vocoder", o.shape, float(o.abs().max()))
    np.savez_compressed(os.path.join(HERE, "ref_vocoder_b2_t21.npz"), o=o.numpy(), B=B, T=T, seed=777)
    print("done")


if __name__ == "__main__":
    main()
