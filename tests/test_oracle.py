"""CPU tier: the oracle against the committed reference fixtures (tests/golden/make_golden.py ran the
unmodified reference in the build container), plus oracle self-consistency."""
import os

import numpy as np
import pytest
import torch

import svc_oracle as O
from sovits_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, f"ref_infer_{name}.npz"))


@pytest.mark.parametrize("name", list(synth.GOLDEN_CASES))
def test_oracle_matches_reference_fixture(cfg, sd, name):
    g = _load(name)
    B, T = synth.GOLDEN_CASES[name]
    c, f0, uv, sid = synth.golden_inputs(cfg, name)
    noise = synth.draw_noise(B, T, cfg, seed=int(g["seed"]))
    taps = {}
    o, _ = O.infer(sd, cfg, c, f0, uv, sid, noise, noice_scale=float(g["noice_scale"]), taps=taps)
    # fp32 restatement vs fp32 reference: identical op order up to the rel-pos band sums
    assert torch.allclose(taps["z_p"], torch.from_numpy(g["z_p"]), atol=2e-5)
    assert torch.allclose(taps["z"], torch.from_numpy(g["z"]), atol=5e-5)
    assert torch.allclose(taps["har"], torch.from_numpy(g["har"]), atol=1e-6)
    assert float((o - torch.from_numpy(g["o"])).abs().max()) < 2e-5   # waveform in (-1,1)


def test_snake_oracle_matches_reference_fixture():
    """vdecoder/hifiganwithsnake (SnakeAlias activations): oracle vs the reference run stored by make_golden.py."""
    from sovits_b200.config import load_config
    cfg_s = load_config()
    cfg_s.vocoder_name = "nsf-snake-hifigan"
    sd_s = synth.synth_state_dict(cfg_s)
    for name, (B, T) in synth.SNAKE_GOLDEN_CASES.items():
        g = _load(name)
        c, f0, uv, sid = synth.golden_inputs(cfg_s, name)
        noise = synth.draw_noise(B, T, cfg_s)
        o, _ = O.infer(sd_s, cfg_s, c, f0, uv, sid, noise, noice_scale=0.4)
        assert float((o - torch.from_numpy(g["o"])).abs().max()) < 2e-5


def test_vocoder_oracle_matches_reference_fixture():
    """vdecoder/nsf_hifigan Generator (mel-conditioned, frame-rate SineGen): oracle vs the stored reference run."""
    from sovits_b200 import nsf_hifigan
    vcfg = nsf_hifigan.cfg_from_h(synth.VOCODER_H)
    vsd = synth.synth_vocoder_state_dict(vcfg)
    g = np.load(os.path.join(GOLD, "ref_vocoder_b2_t21.npz"))
    B, T = int(g["B"]), int(g["T"])
    mel, f0 = synth.synth_vocoder_inputs(vcfg, B, T)
    torch.manual_seed(int(g["seed"]))
    ri, hn = torch.rand(B, 9), torch.randn(B, T * vcfg.hop, 9)
    o = O.vocoder(vsd, vcfg, mel, f0, ri, hn)
    assert float((o - torch.from_numpy(g["o"])).abs().max()) < 2e-5


def test_oracle_fp64_close_to_fp32_reference(cfg, sd):
    g = _load("b1_t33")
    c, f0, uv, sid = synth.golden_inputs(cfg, "b1_t33")
    noise = synth.draw_noise(1, 33, cfg)
    o64, _ = O.infer(sd, cfg, c, f0, uv, sid, noise, noice_scale=0.4, dtype=torch.float64)
    assert float((o64 - torch.from_numpy(g["o"]).double()).abs().max()) < 5e-5


@pytest.mark.parametrize("name", list(synth.GOLDEN_CASES))
def test_closed_form_source_matches_reference(cfg, sd, name):
    """the per-hop closed form (what the CUDA kernel implements) vs the reference's cumsum."""
    g = _load(name)
    B, T = synth.GOLDEN_CASES[name]
    _, f0, _, _ = synth.golden_inputs(cfg, name)
    noise = synth.draw_noise(B, T, cfg)
    har = O.nsf_source_closed_form(sd, f0, noise["rand_ini"], noise["har_noise"], cfg)
    assert float((har - torch.from_numpy(g["har"]).double()).abs().max()) < 2e-5


def test_closed_form_source_long_clip(cfg, sd):
    """At 10 s (862 hops): the closed form (fp32 `r` like the reference, fp64 phase) stays within 1e-4 of both
    the reference's fp32 double-cumsum and its fp64 evaluation (measured 7e-6 / 1.3e-5)."""
    T = 862
    _, f0, _, _ = synth.synth_inputs(cfg, 1, T)
    noise = synth.draw_noise(1, T, cfg)
    lit = O.nsf_source(sd, f0, noise["rand_ini"], noise["har_noise"], cfg, torch.float32)
    cf = O.nsf_source_closed_form(sd, f0, noise["rand_ini"], noise["har_noise"], cfg)
    lit64 = O.nsf_source(sd, f0, noise["rand_ini"], noise["har_noise"], cfg, torch.float64)
    assert float((cf - lit64).abs().max()) < 1e-4          # differs only by evaluating r in fp32 (as the reference does)
    assert float((cf - lit.double()).abs().max()) < 1e-4   # fp32 reference scan error at 441k samples


def test_flow_without_flip_equals_reference_order(cfg, sd):
    """the Flip-free formulation the library packs weights for."""
    torch.manual_seed(0)
    B, T, C = 2, 17, cfg.inter_channels
    half = C // 2
    z = torch.randn(B, C, T)
    g = torch.randn(B, cfg.gin_channels, 1)
    mask = torch.ones(B, 1, T)
    ref = O.flow_reverse(sd, z, mask, g, cfg, torch.float32)
    y = z.clone()
    import torch.nn.functional as F
    for fl in reversed(range(4)):
        p = f"flow.flows.{2 * fl}."
        odd = fl % 2 == 1
        pw, pb = sd[p + "pre.weight"], sd[p + "pre.bias"]
        qw, qb = sd[p + "post.weight"], sd[p + "post.bias"]
        if odd:
            x0 = y[:, half:]
            pw = torch.flip(pw, [1])
            qw, qb = torch.flip(qw, [0]), torch.flip(qb, [0])
        else:
            x0 = y[:, :half]
        h = F.conv1d(x0, pw, pb) * mask
        h = O.wn_forward(sd, p + "enc.", h, mask, g, cfg, torch.float32)
        m = F.conv1d(h, qw, qb) * mask
        if odd:
            y = torch.cat([y[:, :half] - m, y[:, half:]], 1)
        else:
            y = torch.cat([y[:, :half], y[:, half:] - m], 1)
    assert torch.allclose(y, ref, atol=1e-5)


def test_polyphase_transposed_conv(cfg, sd):
    """ConvTranspose1d(k=2s) == s interleaved 2-tap convs, the form the kernels use."""
    import torch.nn.functional as F
    for i, (s, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        w = O.wn_weight(sd, f"dec.ups.{i}", torch.float64)           # [Cin,Cout,k]
        b = sd[f"dec.ups.{i}.bias"].double()
        L = 11
        x = torch.randn(1, w.shape[0], L, dtype=torch.float64)
        p = (k - s + 1) // 2
        ref = F.conv_transpose1d(x, w, b, stride=s, padding=p)
        out = torch.zeros_like(ref)
        xp = F.pad(x, (1, 1))                                         # x[-1] = x[L] = 0
        for ph in range(s):
            for i0 in range(L + 1):
                n = i0 * s + ph - p
                if 0 <= n < L * s:
                    out[0, :, n] = b + w[:, :, ph].t() @ xp[0, :, i0 + 1] + w[:, :, ph + s].t() @ xp[0, :, i0]
        assert torch.allclose(out, ref, atol=1e-10)


def test_oracle_tail_with_speaker_mix_conditioning(cfg, sd):
    """SURVEY §8 f-4 (`EnableCharacterMix`, models.py:456-461,505-509): time-varying g[1,768,T] through flow and generator,
    against the reference's own speaker-mix run (tests/golden/make_golden_mix.py)."""
    gold = np.load(os.path.join(GOLD, "ref_infer_mix_t26.npz"))
    T = int(gold["T"])
    mix = torch.from_numpy(gold["mix"])                                   # [T, S]
    g = (mix @ sd["emb_g.weight"]).t().unsqueeze(0).contiguous()          # [1, 768, T]
    assert float((g - torch.from_numpy(gold["g"])).abs().max()) < 1e-6
    noise = synth.draw_noise(1, T, cfg, seed=int(gold["seed"]))
    f0 = torch.from_numpy(gold["f0"])
    out = O.tail(sd, cfg, torch.from_numpy(gold["z_p"]), g, f0, noise, torch.float32)
    assert float((out - torch.from_numpy(gold["o"])).abs().max()) < 2e-5
    # the stand-alone class builds the same g from the mix weights
    import json
    import sovits_b200
    from sovits_b200 import models
    with open(sovits_b200.DEFAULT_CONFIG) as f:
        kw = json.load(f)["model"]
    net = models.SynthesizerTrn(1025, 20, **kw).eval()
    net.load_state_dict(sd)
    net.EnableCharacterMix(cfg.n_speakers, "cpu")
    with torch.no_grad():
        assert float((net.mix_speakers(mix) - torch.from_numpy(gold["g"])).abs().max()) < 1e-6


def test_fp16_operand_model_explains_the_tensor_core_tolerance(cfg, sd, monkeypatch):
    """Why `precision="tc"` is held to 2e-2 and lands at about 3e-3: the tcgen05 path rounds the OPERANDS of every
    convolution (activations and weights) to fp16 and accumulates in fp32.  Emulating exactly that rounding inside the
    oracle (everything else fp32) reproduces the magnitude measured on the GPU, so the gap is operand rounding, not a
    kernel defect; the fixtures remain within the stated tolerance under this model."""
    import torch.nn.functional as F
    gold = np.load(os.path.join(GOLD, "ref_infer_b2_t24.npz"))
    B, T = synth.GOLDEN_CASES["b2_t24"]
    c, f0, uv, sid = synth.golden_inputs(cfg, "b2_t24")
    noise = synth.draw_noise(B, T, cfg, seed=int(gold["seed"]))
    g = sd["emb_g.weight"][sid].transpose(1, 2).contiguous()
    z_p = torch.from_numpy(gold["z_p"])
    exact = O.tail(sd, cfg, z_p, g, f0, noise, torch.float32)

    real_conv1d, real_convt = F.conv1d, F.conv_transpose1d

    def q(t):
        return t.half().float()

    def conv1d_q(x, w, b=None, *a, **k):
        return real_conv1d(q(x), q(w), b, *a, **k)

    def convt_q(x, w, b=None, *a, **k):
        return real_convt(q(x), q(w), b, *a, **k)

    monkeypatch.setattr(O.F, "conv1d", conv1d_q)
    monkeypatch.setattr(O.F, "conv_transpose1d", convt_q)
    rounded = O.tail(sd, cfg, z_p, g, f0, noise, torch.float32)
    # the same model with bf16 operands (8 mantissa bits instead of 11): why the kernels use kind::f16 with fp16, not bf16
    monkeypatch.setattr(O.F, "conv1d", lambda x, w, b=None, *a, **k: real_conv1d(x.bfloat16().float(), w.bfloat16().float(), b, *a, **k))
    monkeypatch.setattr(O.F, "conv_transpose1d", lambda x, w, b=None, *a, **k: real_convt(x.bfloat16().float(), w.bfloat16().float(), b, *a, **k))
    err_bf16 = float((O.tail(sd, cfg, z_p, g, f0, noise, torch.float32) - exact).abs().max())
    monkeypatch.undo()
    err_model = float((rounded - exact).abs().max())
    print(f"[parity] bf16-operand model: L-inf vs fp32 oracle {err_bf16:.3e}")
    assert err_bf16 > 3 * err_model
    err_fixture = float((rounded - torch.from_numpy(gold["o"])).abs().max())
    print(f"[parity] fp16-operand model of the tcgen05 path: L-inf vs fp32 oracle {err_model:.3e}, vs reference fixture {err_fixture:.3e}")
    assert 2e-4 < err_model < 2e-2          # same order as the measured 2.5-3.4e-3 of the CUDA path (profiles/r01/parity_final.txt)
    assert err_fixture < 2e-2


def test_oracle_and_host_class_with_volume_embedding(cfg, sd):
    """`vol_embedding=True` checkpoints (`emb_vol = nn.Linear(1, hidden)`, models.py:398-399,513): the oracle against the
    reference's own run with per-frame volumes (tests/golden/make_golden_vol.py), and the stand-alone class's conditioning
    (same state_dict keys, same volume term) against the oracle's prologue."""
    gold = np.load(os.path.join(GOLD, "ref_infer_vol_b2_t22.npz"))
    B, T = int(gold["B"]), int(gold["T"])
    sdv = dict(sd)
    sdv["emb_vol.weight"] = torch.from_numpy(gold["emb_vol_w"])
    sdv["emb_vol.bias"] = torch.from_numpy(gold["emb_vol_b"])
    c, f0, uv, sid, vol = (torch.from_numpy(gold[k]) for k in ("c", "f0", "uv", "sid", "vol"))
    noise = synth.draw_noise(B, T, cfg, seed=int(gold["seed"]))
    taps = {}
    out, _ = O.infer(sdv, cfg, c, f0, uv, sid, noise, noice_scale=float(gold["noice_scale"]), taps=taps, vol=vol)
    assert float((taps["z_p"] - torch.from_numpy(gold["z_p"])).abs().max()) < 1e-5
    assert float((out - torch.from_numpy(gold["o"])).abs().max()) < 2e-5
    # without the volumes the result must differ (the fixture's volumes move the waveform by 0.37)
    out0, _ = O.infer(sdv, cfg, c, f0, uv, sid, noise, noice_scale=float(gold["noice_scale"]))
    assert float((out0 - out).abs().max()) > 1e-2
    # host class: emb_vol in the state_dict, the same volume term as the oracle
    import json
    import sovits_b200
    from sovits_b200 import models
    with open(sovits_b200.DEFAULT_CONFIG) as f:
        kw = json.load(f)["model"]
    kw["vol_embedding"] = True
    net = models.SynthesizerTrn(1025, 20, **kw).eval()
    assert {"emb_vol.weight", "emb_vol.bias"} <= set(net.state_dict().keys())
    net.load_state_dict(sdv)
    with torch.no_grad():
        g_h, mask_h, v_h = net.conditioning(c, sid, vol)
        x_h = net.pre(c) * mask_h + net.emb_uv(uv.long()).transpose(1, 2) + v_h
    x_o, _, g_o = O.prologue(sdv, c, f0, uv, sid, cfg, torch.float32, vol=vol)
    assert float((x_h - x_o).abs().max()) < 1e-5 and float((g_h - g_o).abs().max()) == 0.0
    # a checkpoint without the embedding ignores `vol`, like the reference (models.py:513)
    kw["vol_embedding"] = False
    net0 = models.SynthesizerTrn(1025, 20, **kw).eval()
    net0.load_state_dict(sd)
    assert net0.conditioning(c, sid, vol)[2] == 0
