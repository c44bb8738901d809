"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.

A plain-PyTorch (CPU, fp32 or fp64) restatement of the reference's waveform-generation path
``SynthesizerTrn.infer -> ResidualCouplingBlock(reverse) -> vdecoder.hifigan Generator``.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; the product package never does (it fails loudly without its CUDA library).

Pinning: the reference ships no tests or golden vectors ("parity unpinned" by the
reference's own suite).  This restatement is therefore pinned against *outputs of the reference
itself executed in the build container* — ``tests/golden/make_golden.py`` imports
``/root/reference/models.py`` and stores its outputs; ``tests/test_oracle.py`` checks this file
against those fixtures.

Every function cites the reference lines it restates (paths relative to the reference root).
Weights arrive as a ``state_dict`` in the reference's own key layout (weight_g / weight_v pairs).
All noise is an explicit argument so CPU/GPU comparisons are deterministic.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # vdecoder/hifigan/models.py:14


# ----------------------------------------------------------------------------- helpers
def wn_weight(sd: Dict[str, torch.Tensor], prefix: str, dtype) -> torch.Tensor:
    """torch.nn.utils.weight_norm(dim=0): w = g * v / ||v|| with the norm over all dims but 0
    (also for ConvTranspose1d, where dim 0 is Cin) — vdecoder/hifigan/models.py:335,340-342."""
    g = sd[prefix + ".weight_g"].to(dtype)
    v = sd[prefix + ".weight_v"].to(dtype)
    nrm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
    return g * v / nrm


def f0_to_coarse(f0: torch.Tensor) -> torch.Tensor:
    """utils.py:69-80 (f0_bin=256, f0 50..1100 Hz on a mel scale)."""
    f0_bin, f0_max, f0_min = 256, 1100.0, 50.0
    mel_min = 1127 * math.log(1 + f0_min / 700)
    mel_max = 1127 * math.log(1 + f0_max / 700)
    f0_mel = 1127 * (1 + f0 / 700).log()
    a = (f0_bin - 2) / (mel_max - mel_min)
    b = mel_min * a - 1.0
    f0_mel = torch.where(f0_mel > 0, f0_mel * a - b, f0_mel)
    c = torch.round(f0_mel).long()
    c = c * (c > 0)
    c = c + ((c < 1) * 1)
    c = c * (c < f0_bin)
    c = c + ((c >= f0_bin) * (f0_bin - 1))
    return c


def layer_norm_c(x, gamma, beta, eps=1e-5):
    """modules/modules.py:23-35: LayerNorm over the channel dim of [B,C,T]."""
    return F.layer_norm(x.transpose(1, -1), (x.shape[1],), gamma, beta, eps).transpose(1, -1)


# ----------------------------------------------------------------------------- enc_p
def rel_attention(sd, p, x, attn_mask, n_heads, window, dtype):
    """modules/attentions.py:198-239 (MultiHeadAttention with windowed relative positions,
    heads_share=True).  The reference builds the banded terms with pad/reshape tricks
    (:275-303); here the same sums are formed by explicit band gathers."""
    W = lambda n: sd[p + n].to(dtype)
    q = F.conv1d(x, W("conv_q.weight"), W("conv_q.bias"))
    k = F.conv1d(x, W("conv_k.weight"), W("conv_k.bias"))
    v = F.conv1d(x, W("conv_v.weight"), W("conv_v.bias"))
    B, D, L = q.shape
    dk = D // n_heads
    q = q.view(B, n_heads, dk, L).transpose(2, 3)
    k = k.view(B, n_heads, dk, L).transpose(2, 3)
    v = v.view(B, n_heads, dk, L).transpose(2, 3)
    qs = q / math.sqrt(dk)
    scores = qs @ k.transpose(-2, -1)                                  # [B,h,L,L]
    ek, ev = W("emb_rel_k")[0], W("emb_rel_v")[0]                       # [2w+1, dk]
    idx = torch.arange(L, device=x.device)
    rel = idx[None, :] - idx[:, None]                                   # j - i
    band = rel.abs() <= window
    ridx = (rel + window).clamp(0, 2 * window)
    rl = qs @ ek.t()                                                    # [B,h,L,2w+1]
    local = torch.gather(rl, 3, ridx[None, None].expand(B, n_heads, L, L))
    scores = scores + local * band
    scores = scores.masked_fill(attn_mask == 0, -1e4)
    pa = F.softmax(scores, dim=-1)
    out = pa @ v
    # relative values: w[b,h,i,r] = p[b,h,i,i+r-window]
    rw = torch.zeros(B, n_heads, L, 2 * window + 1, dtype=dtype, device=x.device)
    for r in range(2 * window + 1):
        off = r - window
        d = torch.diagonal(pa, offset=off, dim1=2, dim2=3)              # p[i, i+off]
        if off >= 0:
            rw[:, :, : L - off, r] = d
        else:
            rw[:, :, -off:, r] = d
    out = out + rw @ ev
    out = out.transpose(2, 3).contiguous().view(B, D, L)
    return F.conv1d(out, W("conv_o.weight"), W("conv_o.bias"))


def text_encoder(sd, x, x_mask, f0_coarse, z_noise, noice_scale, cfg, dtype):
    """models.py:155-162 + modules/attentions.py:95-107,350-363."""
    W = lambda n: sd[n].to(dtype)
    x = x + W("enc_p.f0_emb.weight")[f0_coarse].transpose(1, 2)
    x = x * x_mask
    attn_mask = x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)
    x = x * x_mask
    ks = cfg.kernel_size
    for i in range(cfg.n_layers):
        y = rel_attention(sd, f"enc_p.enc_.attn_layers.{i}.", x, attn_mask, cfg.n_heads, cfg.enc_window, dtype)
        x = layer_norm_c(x + y, W(f"enc_p.enc_.norm_layers_1.{i}.gamma"), W(f"enc_p.enc_.norm_layers_1.{i}.beta"))
        f = f"enc_p.enc_.ffn_layers.{i}."
        pad = ((ks - 1) // 2, ks // 2)
        y = F.conv1d(F.pad(x * x_mask, pad), W(f + "conv_1.weight"), W(f + "conv_1.bias"))
        y = torch.relu(y)
        y = F.conv1d(F.pad(y * x_mask, pad), W(f + "conv_2.weight"), W(f + "conv_2.bias")) * x_mask
        x = layer_norm_c(x + y, W(f"enc_p.enc_.norm_layers_2.{i}.gamma"), W(f"enc_p.enc_.norm_layers_2.{i}.beta"))
    x = x * x_mask
    stats = F.conv1d(x, W("enc_p.proj.weight"), W("enc_p.proj.bias")) * x_mask
    m, logs = torch.split(stats, cfg.inter_channels, dim=1)
    z = (m + z_noise.to(dtype) * torch.exp(logs) * noice_scale) * x_mask
    return z, m, logs


# ----------------------------------------------------------------------------- flow
def wn_forward(sd, p, x, x_mask, g, cfg, dtype):
    """modules/modules.py:110-138 (+ commons.py:129-136 gate)."""
    H = cfg.hidden_channels
    L = cfg.flow_wn_layers
    out = torch.zeros_like(x)
    gc = F.conv1d(g, wn_weight(sd, p + "cond_layer", dtype), sd[p + "cond_layer.bias"].to(dtype))
    pad = (cfg.flow_kernel_size - 1) // 2   # dilation_rate == 1 (models.py:441)
    for i in range(L):
        x_in = F.conv1d(x, wn_weight(sd, p + f"in_layers.{i}", dtype), sd[p + f"in_layers.{i}.bias"].to(dtype), padding=pad)
        a = x_in + gc[:, 2 * H * i: 2 * H * (i + 1), :]
        acts = torch.tanh(a[:, :H]) * torch.sigmoid(a[:, H:])
        rs = F.conv1d(acts, wn_weight(sd, p + f"res_skip_layers.{i}", dtype), sd[p + f"res_skip_layers.{i}.bias"].to(dtype))
        if i < L - 1:
            x = (x + rs[:, :H]) * x_mask
            out = out + rs[:, H:]
        else:
            out = out + rs
    return out * x_mask


def flow_reverse(sd, z_p, x_mask, g, cfg, dtype):
    """models.py:45-52 (reverse branch) over [L0,Flip,L1,Flip,L2,Flip,L3,Flip];
    modules/modules.py:288-307 (mean_only coupling), :232-239 (Flip)."""
    x = z_p
    half = cfg.inter_channels // 2
    for fl in reversed(range(4)):
        x = torch.flip(x, [1])
        p = f"flow.flows.{2 * fl}."
        x0, x1 = x[:, :half], x[:, half:]
        h = F.conv1d(x0, sd[p + "pre.weight"].to(dtype), sd[p + "pre.bias"].to(dtype)) * x_mask
        h = wn_forward(sd, p + "enc.", h, x_mask, g, cfg, dtype)
        m = F.conv1d(h, sd[p + "post.weight"].to(dtype), sd[p + "post.bias"].to(dtype)) * x_mask
        x1 = (x1 - m) * x_mask
        x = torch.cat([x0, x1], 1)
    return x


# ----------------------------------------------------------------------------- NSF source
def nsf_source(sd, f0, rand_ini, har_noise, cfg, dtype):
    """vdecoder/hifigan/models.py:369-372 (f0 nearest upsample), :250-271 (SineGen.forward),
    :138-166 (_f02sine with the doubly-wrapped cumsum), :307-320 (SourceModuleHnNSF).
    Returns har_source [B,1,N].  ``rand_ini`` [B,9] (column 0 is forced to 0 like :148),
    ``har_noise`` [B,N,9] standard normal."""
    sr = float(cfg.sampling_rate)
    upp = cfg.hop
    f0u = torch.repeat_interleave(f0.to(dtype), upp, dim=1)[:, :, None]            # [B,N,1]
    harm = torch.arange(1, cfg.n_harmonics + 1, dtype=dtype, device=f0.device)[None, None, :]
    fn = f0u * harm
    rad = (fn / sr) % 1
    ri = rand_ini.to(dtype).clone()
    ri[:, 0] = 0
    rad[:, 0, :] = rad[:, 0, :] + ri
    tmp = torch.cumsum(rad, 1) % 1
    wrap = (tmp[:, 1:, :] - tmp[:, :-1, :]) < 0
    shift = torch.zeros_like(rad)
    shift[:, 1:, :] = wrap * -1.0
    sines = torch.sin(torch.cumsum(rad + shift, dim=1) * 2 * math.pi)
    sine_waves = sines * 0.1
    uv = (f0u > 0).to(dtype)
    noise_amp = uv * 0.003 + (1 - uv) * 0.1 / 3
    sine_waves = sine_waves * uv + noise_amp * har_noise.to(dtype)
    merged = torch.tanh(F.linear(sine_waves, sd["dec.m_source.l_linear.weight"].to(dtype),
                                 sd["dec.m_source.l_linear.bias"].to(dtype)))
    return merged.transpose(1, 2)


def nsf_source_closed_form(sd, f0, rand_ini, har_noise, cfg):
    """fp64 closed form of the same source: because f0 is constant inside a hop,
    phase[b, hop*F+k, h] = rand_ini + sum_{f<F} hop*r[f] + (k+1)*r[F] (mod 1), with r evaluated
    in fp32 exactly like :144 and then promoted.  This is what the CUDA kernel implements."""
    sr = cfg.sampling_rate
    upp = cfg.hop
    harm = torch.arange(1, cfg.n_harmonics + 1, dtype=torch.float32)[None, None, :]
    r32 = ((f0.float()[:, :, None] * harm) / float(sr)) % 1                        # [B,T,9] fp32 like the reference
    r = r32.double()
    ri = rand_ini.double().clone()
    ri[:, 0] = 0
    start = torch.cumsum(r * upp, dim=1) - r * upp + ri[:, None, :]                # phase before frame F
    k = torch.arange(1, upp + 1, dtype=torch.float64)[None, None, :, None]
    ph = start[:, :, None, :] + k * r[:, :, None, :]                               # [B,T,upp,9]
    ph = ph - torch.floor(ph)
    sines = torch.sin(2 * math.pi * ph).reshape(f0.shape[0], -1, cfg.n_harmonics) * 0.1
    uv = torch.repeat_interleave((f0 > 0).double(), upp, dim=1)[:, :, None]
    amp = uv * 0.003 + (1 - uv) * 0.1 / 3
    sw = sines * uv + amp * har_noise.double()
    merged = torch.tanh(F.linear(sw, sd["dec.m_source.l_linear.weight"].double(), sd["dec.m_source.l_linear.bias"].double()))
    return merged.transpose(1, 2)


# ----------------------------------------------------------------------------- generator
def snake_alias(sd, p, x, dtype):
    """vdecoder/hifiganwithsnake/alias/act.py:109-129: 2x kaiser-sinc upsample (resample.py:35-54) -> SnakeBeta with
    log-scale alpha/beta (act.py:81-92) -> 2x low-pass downsample (filter.py:93-109), replicate padding."""
    C = x.shape[1]
    f = sd[p + "upsample.filter"].to(dtype).expand(C, -1, -1)
    xp = F.pad(x, (5, 5), mode="replicate")
    u = 2 * F.conv_transpose1d(xp, f, stride=2, groups=C)[..., 15:-15]
    alpha = torch.exp(sd[p + "act.alpha"].to(dtype))[None, :, None]
    beta = torch.exp(sd[p + "act.beta"].to(dtype))[None, :, None]
    sn = u + (1.0 / (beta + 0.000000001)) * torch.sin(u * alpha) ** 2
    sp = F.pad(sn, (5, 6), mode="replicate")
    fd = sd[p + "downsample.lowpass.filter"].to(dtype).expand(C, -1, -1)
    return F.conv1d(sp, fd, stride=2, groups=C)


def resblock1(sd, p, x, k, dils, dtype, snake=False):
    """vdecoder/hifigan/models.py:60-67 (hifiganwithsnake/models.py:66-74 when snake)."""
    for j, d in enumerate(dils):
        if snake:
            xt = snake_alias(sd, p + f"activations.{2 * j}.", x, dtype)
            xt = F.conv1d(xt, wn_weight(sd, p + f"convs1.{j}", dtype), sd[p + f"convs1.{j}.bias"].to(dtype),
                          dilation=d, padding=d * (k - 1) // 2)
            xt = snake_alias(sd, p + f"activations.{2 * j + 1}.", xt, dtype)
            xt = F.conv1d(xt, wn_weight(sd, p + f"convs2.{j}", dtype), sd[p + f"convs2.{j}.bias"].to(dtype),
                          padding=(k - 1) // 2)
            x = xt + x
            continue
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, wn_weight(sd, p + f"convs1.{j}", dtype), sd[p + f"convs1.{j}.bias"].to(dtype),
                      dilation=d, padding=d * (k - 1) // 2)
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, wn_weight(sd, p + f"convs2.{j}", dtype), sd[p + f"convs2.{j}.bias"].to(dtype),
                      padding=(k - 1) // 2)
        x = xt + x
    return x


def generator(sd, z, g, har, cfg, dtype, taps: Optional[dict] = None):
    """vdecoder/hifigan/models.py:373-392 given the excitation ``har`` [B,1,N]."""
    x = F.conv1d(z, wn_weight(sd, "dec.conv_pre", dtype), sd["dec.conv_pre.bias"].to(dtype), padding=3)
    x = x + F.conv1d(g, sd["dec.cond.weight"].to(dtype), sd["dec.cond.bias"].to(dtype))
    if taps is not None:
        taps["conv_pre"] = x
    nk = len(cfg.resblock_kernel_sizes)
    snake = getattr(cfg, "snake", False)
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        x = snake_alias(sd, f"dec.snakes.{i}.", x, dtype) if snake else F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, wn_weight(sd, f"dec.ups.{i}", dtype), sd[f"dec.ups.{i}.bias"].to(dtype),
                               stride=u, padding=(k - u + 1) // 2)
        nw = sd[f"dec.noise_convs.{i}.weight"].to(dtype)
        if nw.shape[2] > 1:
            s = nw.shape[2] // 2
            xs_ = F.conv1d(har, nw, sd[f"dec.noise_convs.{i}.bias"].to(dtype), stride=s, padding=(s + 1) // 2)
        else:
            xs_ = F.conv1d(har, nw, sd[f"dec.noise_convs.{i}.bias"].to(dtype))
        x = x + xs_
        if taps is not None:
            taps[f"ups{i}"] = x
        acc = None
        for j in range(nk):
            r = resblock1(sd, f"dec.resblocks.{i * nk + j}.", x, cfg.resblock_kernel_sizes[j],
                          cfg.resblock_dilation_sizes[j], dtype, snake)
            acc = r if acc is None else acc + r
        x = acc / nk
        if taps is not None:
            taps[f"stage{i}"] = x
    x = snake_alias(sd, "dec.snake_post.", x, dtype) if snake else F.leaky_relu(x)  # default slope 0.01 (:390)
    x = F.conv1d(x, wn_weight(sd, "dec.conv_post", dtype), sd["dec.conv_post.bias"].to(dtype), padding=3)
    return torch.tanh(x)


# ----------------------------------------------------------------------------- infer
@torch.no_grad()
def prologue(sd, c, f0, uv, sid, cfg, dtype, vol=None):
    """models.py:503-520: mask (all ones), g = emb_g(sid)^T, x = pre(c)*mask + emb_uv(uv)^T (+vol)."""
    B, _, T = c.shape
    x_mask = torch.ones(B, 1, T, dtype=dtype, device=c.device)
    if sid.dim() == 1:
        sid = sid.unsqueeze(0)
    g = sd["emb_g.weight"].to(dtype)[sid].transpose(1, 2)            # [B,gin,1]
    x = F.conv1d(c.to(dtype), sd["pre.weight"].to(dtype), sd["pre.bias"].to(dtype), padding=2) * x_mask
    x = x + sd["emb_uv.weight"].to(dtype)[uv.long()].transpose(1, 2)
    if vol is not None and "emb_vol.weight" in sd:
        # models.py:513: vol = emb_vol(vol[:, :, None]).transpose(1, 2) with emb_vol = nn.Linear(1, hidden) (models.py:398-399)
        v = F.linear(vol.to(dtype)[:, :, None], sd["emb_vol.weight"].to(dtype), sd["emb_vol.bias"].to(dtype))
        x = x + v.transpose(1, 2)
    return x, x_mask, g


@torch.no_grad()
def infer(sd, cfg, c, f0, uv, sid, noise, noice_scale=0.35, dtype=torch.float32, taps: Optional[dict] = None, vol=None):
    """models.py:495-532 with predict_f0=False.  ``noise`` = {"z_noise","rand_ini","har_noise"}; ``vol`` [B,T] is used when the
    checkpoint has a volume embedding (``emb_vol.*``), exactly like the reference ignores it otherwise."""
    x, x_mask, g = prologue(sd, c, f0, uv, sid, cfg, dtype, vol=vol)
    z_p, m_p, logs_p = text_encoder(sd, x, x_mask, f0_to_coarse(f0), noise["z_noise"], noice_scale, cfg, dtype)
    z = flow_reverse(sd, z_p, x_mask, g, cfg, dtype)
    har = nsf_source(sd, f0, noise["rand_ini"], noise["har_noise"], cfg, dtype)
    if taps is not None:
        taps.update({"g": g, "z_p": z_p, "z": z, "har": har})
    o = generator(sd, z * x_mask, g, har, cfg, dtype, taps)
    return o, f0


@torch.no_grad()
def tail(sd, cfg, z_p, g, f0, noise, dtype=torch.float32, taps: Optional[dict] = None):
    """The three CUDA kernels' scope only: flow(reverse) -> NSF source -> generator."""
    B, _, T = z_p.shape
    x_mask = torch.ones(B, 1, T, dtype=dtype, device=z_p.device)
    z = flow_reverse(sd, z_p.to(dtype), x_mask, g.to(dtype), cfg, dtype)
    har = nsf_source(sd, f0, noise["rand_ini"], noise["har_noise"], cfg, dtype)
    if taps is not None:
        taps.update({"z": z, "har": har})
    return generator(sd, z, g.to(dtype), har, cfg, dtype, taps)


# ----------------------------------------------------------------------------- mel-conditioned vocoder
def vocoder_source(sd, f0, rand_ini, har_noise, cfg, dtype=torch.float32):
    """vdecoder/nsf_hifigan/models.py:138-178 (SineGen: frame-rate rad_values, fp64 cumsum, linear-interpolated wrap
    detection, fp64 sample-rate cumsum) + :215-218 (SourceModuleHnNSF).  Returns [B,1,N]."""
    upp = cfg.hop
    f0 = f0.to(dtype).unsqueeze(-1)
    fn = f0 * torch.arange(1, cfg.n_harmonics + 1, device=f0.device).reshape(1, 1, -1)
    rad = (fn / cfg.sampling_rate) % 1
    ri = rand_ini.to(dtype).clone()
    ri[:, 0] = 0
    rad[:, 0, :] = rad[:, 0, :] + ri
    tmp = torch.cumsum(rad.double(), 1).to(dtype)
    tmp = tmp * upp
    tmp = F.interpolate(tmp.transpose(2, 1), scale_factor=upp, mode="linear", align_corners=True).transpose(2, 1)
    rad_up = F.interpolate(rad.transpose(2, 1), scale_factor=upp, mode="nearest").transpose(2, 1)
    tmp = tmp % 1
    wrap = (tmp[:, 1:, :] - tmp[:, :-1, :]) < 0
    shift = torch.zeros_like(rad_up)
    shift[:, 1:, :] = wrap * -1.0
    sines = torch.sin(torch.cumsum(rad_up.double() + shift.double(), dim=1) * 2 * math.pi).to(dtype) * 0.1
    uv = (f0 > 0).to(dtype)
    uv = F.interpolate(uv.transpose(2, 1), scale_factor=upp, mode="nearest").transpose(2, 1)
    amp = uv * 0.003 + (1 - uv) * 0.1 / 3
    sw = sines * uv + amp * har_noise.to(dtype)
    merged = torch.tanh(F.linear(sw, sd["m_source.l_linear.weight"].to(dtype), sd["m_source.l_linear.bias"].to(dtype)))
    return merged.transpose(1, 2)


@torch.no_grad()
def vocoder(sd, cfg, mel, f0, rand_ini, har_noise, dtype=torch.float32):
    """vdecoder/nsf_hifigan/models.py:259-278: har = m_source(f0); conv_pre(mel); 5 x (lrelu, ups, + noise_conv, 3 ResBlocks
    averaged); lrelu; conv_post; tanh.  Paddings (k-u)//2 and stride//2 coincide with the SVC decoder's for even k-u / stride."""
    har = vocoder_source(sd, f0, rand_ini, har_noise, cfg, dtype)
    sd2 = {"dec." + k: v for k, v in sd.items()}
    U = cfg.upsample_initial_channel
    sd2["dec.cond.weight"] = torch.zeros(U, 1, 1)
    sd2["dec.cond.bias"] = torch.zeros(U)
    g = torch.zeros(mel.shape[0], 1, 1, dtype=dtype, device=mel.device)
    return generator(sd2, mel.to(dtype), g, har, cfg, dtype)
