"""Seeded synthetic checkpoints and inputs (there is no network for real checkpoints).

Key names and shapes follow the reference ``SynthesizerTrn.state_dict()`` for the modules used by
``infer`` (``models.py:400-453``; probe list in SURVEY §8b).  Values are *non-degenerate*: the
reference's default init makes the flow an identity (``modules/modules.py:285-286``) and the
ResBlocks near-identity (``vdecoder/hifigan/models.py:48,58``), which would hide kernel bugs, so
``post`` layers and generator convs are re-drawn at O(1) gain.

All draws use a CPU ``torch.Generator`` so the same bits are produced in this container and on the
GPU box.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict

import torch

from .config import ModelCfg

N_FLOWS = 4  # ResidualCouplingBlock default n_flows (models.py:22)


def param_shapes(cfg: ModelCfg) -> "OrderedDict[str, tuple]":
    """state_dict keys -> shapes for emb_g, pre, emb_uv, enc_p, flow, dec."""
    H, C, G = cfg.hidden_channels, cfg.inter_channels, cfg.gin_channels
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["emb_g.weight"] = (cfg.n_speakers, G)
    s["pre.weight"] = (H, cfg.ssl_dim, 5)
    s["pre.bias"] = (H,)
    s["emb_uv.weight"] = (2, H)
    # enc_p (models.py:128-162, attentions.py:73-107)
    s["enc_p.proj.weight"] = (2 * C, H, 1)
    s["enc_p.proj.bias"] = (2 * C,)
    s["enc_p.f0_emb.weight"] = (256, H)
    dk = H // cfg.n_heads
    for i in range(cfg.n_layers):
        a = f"enc_p.enc_.attn_layers.{i}."
        s[a + "emb_rel_k"] = (1, 2 * cfg.enc_window + 1, dk)
        s[a + "emb_rel_v"] = (1, 2 * cfg.enc_window + 1, dk)
        for n in ("q", "k", "v", "o"):
            s[a + f"conv_{n}.weight"] = (H, H, 1)
            s[a + f"conv_{n}.bias"] = (H,)
    for i in range(cfg.n_layers):
        s[f"enc_p.enc_.norm_layers_1.{i}.gamma"] = (H,)
        s[f"enc_p.enc_.norm_layers_1.{i}.beta"] = (H,)
    for i in range(cfg.n_layers):
        f = f"enc_p.enc_.ffn_layers.{i}."
        s[f + "conv_1.weight"] = (cfg.filter_channels, H, cfg.kernel_size)
        s[f + "conv_1.bias"] = (cfg.filter_channels,)
        s[f + "conv_2.weight"] = (H, cfg.filter_channels, cfg.kernel_size)
        s[f + "conv_2.bias"] = (H,)
    for i in range(cfg.n_layers):
        s[f"enc_p.enc_.norm_layers_2.{i}.gamma"] = (H,)
        s[f"enc_p.enc_.norm_layers_2.{i}.beta"] = (H,)
    # flow (models.py:15-52, modules.py:73-108,260-286); Flip modules at odd indices have no params
    half = C // 2
    L = cfg.flow_wn_layers
    for fl in range(N_FLOWS):
        p = f"flow.flows.{2 * fl}."
        s[p + "pre.weight"] = (H, half, 1)
        s[p + "pre.bias"] = (H,)
        for i in range(L):
            s[p + f"enc.in_layers.{i}.bias"] = (2 * H,)
            s[p + f"enc.in_layers.{i}.weight_g"] = (2 * H, 1, 1)
            s[p + f"enc.in_layers.{i}.weight_v"] = (2 * H, H, cfg.flow_kernel_size)
        for i in range(L):
            co = 2 * H if i < L - 1 else H
            s[p + f"enc.res_skip_layers.{i}.bias"] = (co,)
            s[p + f"enc.res_skip_layers.{i}.weight_g"] = (co, 1, 1)
            s[p + f"enc.res_skip_layers.{i}.weight_v"] = (co, H, 1)
        s[p + "enc.cond_layer.bias"] = (2 * H * L,)
        s[p + "enc.cond_layer.weight_g"] = (2 * H * L, 1, 1)
        s[p + "enc.cond_layer.weight_v"] = (2 * H * L, G, 1)
        s[p + "post.weight"] = (half, H, 1)
        s[p + "post.bias"] = (half,)
    # dec (vdecoder/hifigan/models.py:324-360)
    U = cfg.upsample_initial_channel
    s["dec.m_source.l_linear.weight"] = (1, cfg.n_harmonics)
    s["dec.m_source.l_linear.bias"] = (1,)
    n_up = len(cfg.upsample_rates)
    for i in range(n_up):
        c_cur = U // (2 ** (i + 1))
        if i + 1 < n_up:
            stride = 1
            for u in cfg.upsample_rates[i + 1:]:
                stride *= u
            s[f"dec.noise_convs.{i}.weight"] = (c_cur, 1, 2 * stride)
        else:
            s[f"dec.noise_convs.{i}.weight"] = (c_cur, 1, 1)
        s[f"dec.noise_convs.{i}.bias"] = (c_cur,)
    s["dec.conv_pre.bias"] = (U,)
    s["dec.conv_pre.weight_g"] = (U, 1, 1)
    s["dec.conv_pre.weight_v"] = (U, C, 7)
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        cin, cout = U // (2 ** i), U // (2 ** (i + 1))
        s[f"dec.ups.{i}.bias"] = (cout,)
        s[f"dec.ups.{i}.weight_g"] = (cin, 1, 1)
        s[f"dec.ups.{i}.weight_v"] = (cin, cout, k)
    for i in range(n_up):
        ch = U // (2 ** (i + 1))
        for j, k in enumerate(cfg.resblock_kernel_sizes):
            r = f"dec.resblocks.{i * len(cfg.resblock_kernel_sizes) + j}."
            for grp in ("convs1", "convs2"):
                for d in range(len(cfg.resblock_dilation_sizes[j])):
                    s[r + f"{grp}.{d}.bias"] = (ch,)
                    s[r + f"{grp}.{d}.weight_g"] = (ch, 1, 1)
                    s[r + f"{grp}.{d}.weight_v"] = (ch, ch, k)
    if cfg.snake:
        # vdecoder/hifiganwithsnake/models.py:364-374,61-64: SnakeAlias before every ups / conv / conv_post
        def snake_keys(prefix, chn):
            s[prefix + "act.alpha"] = (chn,)
            s[prefix + "act.beta"] = (chn,)
            s[prefix + "upsample.filter"] = (1, 1, 12)
            s[prefix + "downsample.lowpass.filter"] = (1, 1, 12)
        for i in range(n_up):
            snake_keys(f"dec.snakes.{i}.", U // (2 ** i))
            chn = U // (2 ** (i + 1))
            for j in range(len(cfg.resblock_kernel_sizes)):
                for a in range(2 * len(cfg.resblock_dilation_sizes[j])):
                    snake_keys(f"dec.resblocks.{i * len(cfg.resblock_kernel_sizes) + j}.activations.{a}.", chn)
        snake_keys("dec.snake_post.", U // (2 ** n_up))
    ch = U // (2 ** n_up)
    s["dec.conv_post.bias"] = (1,)
    s["dec.conv_post.weight_g"] = (1, 1, 1)
    s["dec.conv_post.weight_v"] = (1, ch, 7)
    s["dec.cond.weight"] = (U, G, 1)
    s["dec.cond.bias"] = (U,)
    return s


def kaiser_sinc_filter12() -> torch.Tensor:
    """The fixed 12-tap anti-aliasing filter of SnakeAlias (cutoff 0.25, half-width 0.3, Kaiser window;
    vdecoder/hifiganwithsnake/alias/filter.py:29-58), normalised to unit sum."""
    ks, cutoff, half_width = 12, 0.25, 0.3
    half = ks // 2
    A = 2.285 * (half - 1) * math.pi * (4 * half_width) + 7.95
    beta = 0.1102 * (A - 8.7) if A > 50 else (0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21) if A >= 21 else 0.0)
    win = torch.kaiser_window(ks, beta=beta, periodic=False)
    t = torch.arange(-half, half, dtype=torch.float32) + 0.5
    f = 2 * cutoff * win * torch.sinc(2 * cutoff * t)
    return f / f.sum()


def synth_state_dict(cfg: ModelCfg, seed: int = 20260922, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic, non-degenerate weights in the reference's checkpoint layout."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    shapes = param_shapes(cfg)
    sd: Dict[str, torch.Tensor] = OrderedDict()

    def randn(shape, std=1.0):
        return torch.randn(shape, generator=gen, dtype=torch.float32) * std

    def rand(shape, lo, hi):
        return torch.rand(shape, generator=gen, dtype=torch.float32) * (hi - lo) + lo

    for key, shape in shapes.items():
        leaf = key.rsplit(".", 1)[1]
        if leaf == "weight_g":
            continue  # drawn together with weight_v below
        if leaf == "weight_v":
            if ".ups." in key:                      # ConvTranspose1d [Cin, Cout, k]: k/u = 2 taps per output
                fan = shape[0] * 2
            else:
                fan = shape[1] * shape[2]
            gain = 1.0
            if ".res_skip_layers." in key or ".in_layers." in key or "cond_layer" in key:
                gain = 1.0
            if ".resblocks." in key:
                gain = 0.45 if cfg.snake else 1.0   # residual branches comparable to the skip path (snake passes more energy than lrelu)
            if cfg.snake and (".ups." in key or "conv_pre" in key):
                gain = 0.7
            v = randn(shape, gain / math.sqrt(fan))
            sd[key] = v
            # weight_norm: norm over all dims but 0 (also for ConvTranspose1d)
            nrm = v.reshape(shape[0], -1).norm(dim=1).reshape(shape[0], 1, 1)
            sd[key.replace("weight_v", "weight_g")] = nrm * rand((shape[0], 1, 1), 0.5, 1.5)
        elif leaf == "filter":
            sd[key] = kaiser_sinc_filter12().reshape(shape)
        elif leaf == "alpha" and ".act." in key:
            sd[key] = randn(shape, 0.3)              # log-scale frequency of SnakeBeta
        elif leaf == "beta" and ".act." in key:
            sd[key] = randn(shape, 0.3) + 1.0        # log-scale magnitude (1/beta ~ 0.37)
        elif leaf in ("gamma",):
            sd[key] = rand(shape, 0.8, 1.2)
        elif leaf in ("beta", "bias"):
            sd[key] = randn(shape, 0.05)
        elif key in ("emb_g.weight",):
            sd[key] = randn(shape, 1.0)
        elif key in ("emb_uv.weight", "enc_p.f0_emb.weight"):
            sd[key] = randn(shape, 0.3)
        elif leaf in ("emb_rel_k", "emb_rel_v"):
            sd[key] = randn(shape, shape[-1] ** -0.5)
        elif key == "dec.m_source.l_linear.weight":
            sd[key] = rand(shape, -1.0 / 3, 1.0 / 3)
        elif ".noise_convs." in key:
            sd[key] = randn(shape, 1.0 / math.sqrt(shape[2]))
        elif ".post.weight" in key:
            sd[key] = randn(shape, 0.05)            # reference zero-inits this (identity flow)
        elif leaf == "weight":
            fan = shape[1] * (shape[2] if len(shape) > 2 else 1)
            sd[key] = randn(shape, 1.0 / math.sqrt(fan))
        else:
            raise KeyError(key)
    ordered = OrderedDict((k, sd[k].to(dtype)) for k in shapes)
    return ordered


def synth_inputs(cfg: ModelCfg, B: int, T: int, seed: int = 1234):
    """BASELINE.md §3.2 inputs: c~N(0,1), smooth f0 in 110-500 Hz with ~15% unvoiced frames."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    c = torch.randn((B, cfg.ssl_dim, T), generator=gen, dtype=torch.float32)
    t = torch.arange(T, dtype=torch.float32)[None, :]
    b = torch.arange(B, dtype=torch.float32)[:, None]
    f0 = 220.0 * torch.pow(2.0, 0.5 * torch.sin(2 * math.pi * t / 200.0 + b) + b / 12.0)
    f0 = torch.where((torch.arange(T)[None, :] % 97) < 15, torch.zeros_like(f0), f0)
    uv = (f0 > 0).float()
    sid = (torch.arange(B) % cfg.n_speakers)[:, None].long()
    return c, f0, uv, sid


def draw_noise(B: int, T: int, cfg: ModelCfg, seed: int = 52468, device="cpu"):
    """The reference's RNG draws, in its order (models.py:498-501,160;
    vdecoder/hifigan/models.py:147,266,319), after seeding like ``infer`` does."""
    torch.manual_seed(seed)
    N = T * cfg.hop
    z_noise = torch.randn((B, cfg.inter_channels, T), device=device)
    rand_ini = torch.rand((B, cfg.n_harmonics), device=device)
    har_noise = torch.randn((B, N, cfg.n_harmonics), device=device)
    _unused = torch.randn((B, N, 1), device=device)  # advances the generator like :319
    return {"z_noise": z_noise, "rand_ini": rand_ini, "har_noise": har_noise}


GOLDEN_CASES = {"b2_t24": (2, 24), "b1_t33": (1, 33)}
SNAKE_GOLDEN_CASES = {"snake_b1_t20": (1, 20)}        # vocoder_name = "nsf-snake-hifigan" (BASELINE config 4)


def golden_inputs(cfg: ModelCfg, name: str):
    """Inputs of the committed reference fixtures (tests/golden/make_golden.py)."""
    B, T = GOLDEN_CASES[name] if name in GOLDEN_CASES else SNAKE_GOLDEN_CASES[name]
    c, f0, uv, sid = synth_inputs(cfg, B, T)
    if name == "b1_t33":            # odd T, an unvoiced span with voiced<->unvoiced edges inside a clip
        f0[:, 5:12] = 0.0
        uv = (f0 > 0).float()
    return c, f0, uv, sid


def synth_vocoder_state_dict(cfg: ModelCfg, seed: int = 20260923) -> Dict[str, torch.Tensor]:
    """Seeded weights for the mel-conditioned vdecoder/nsf_hifigan Generator (cfg.num_mels > 0), reference key layout."""
    from .nsf_hifigan import vocoder_param_shapes
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sd: Dict[str, torch.Tensor] = OrderedDict()
    for key, shape in vocoder_param_shapes(cfg).items():
        leaf = key.rsplit(".", 1)[1]
        if leaf == "weight_g":
            continue
        if leaf == "weight_v":
            fan = shape[0] * 2 if key.startswith("ups.") else shape[1] * shape[2]
            v = torch.randn(shape, generator=gen) / math.sqrt(fan)
            sd[key] = v
            nrm = v.reshape(shape[0], -1).norm(dim=1).reshape(shape[0], 1, 1)
            sd[key.replace("weight_v", "weight_g")] = nrm * (torch.rand((shape[0], 1, 1), generator=gen) + 0.5)
        elif leaf == "bias":
            sd[key] = torch.randn(shape, generator=gen) * 0.05
        elif key == "m_source.l_linear.weight":
            sd[key] = torch.rand(shape, generator=gen) * (2.0 / 3) - 1.0 / 3
        elif key.startswith("noise_convs."):
            sd[key] = torch.randn(shape, generator=gen) / math.sqrt(shape[2])
        else:
            raise KeyError(key)
    from .nsf_hifigan import vocoder_param_shapes as _vps
    return OrderedDict((k, sd[k]) for k in _vps(cfg))


VOCODER_H = {"sampling_rate": 44100, "num_mels": 128, "resblock": "1", "resblock_kernel_sizes": [3, 7, 11],
             "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]], "upsample_rates": [8, 8, 2, 2, 2],
             "upsample_kernel_sizes": [16, 16, 4, 4, 4], "upsample_initial_channel": 512,
             "hop_size": 512, "n_fft": 2048, "win_size": 2048, "fmin": 40, "fmax": 16000}


def synth_vocoder_inputs(cfg: ModelCfg, B: int, T: int, seed: int = 4321):
    gen = torch.Generator(device="cpu").manual_seed(seed)
    mel = torch.randn((B, cfg.num_mels, T), generator=gen)                     # standardised log-mel
    _, f0, _, _ = synth_inputs(cfg, B, T)
    return mel, f0
