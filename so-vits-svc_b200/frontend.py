"""Prefix of ``SynthesizerTrn.infer`` that stays on stock PyTorch:
speaker/uv embeddings, the ``pre`` conv and the ``enc_p`` relative-position transformer.

Written from the reference's behaviour (models.py:128-162,495-529; modules/attentions.py:73-107,
161-303,317-363), with the same parameter names so reference checkpoints load unchanged.  The
relative-position terms are applied on the nine diagonals directly instead of the reference's
pad/reshape skewing (attentions.py:275-303) — same sums, far less memory traffic.
"""
from __future__ import annotations

import math

import torch
from torch import nn
from torch.nn import functional as F


def f0_to_coarse(f0: torch.Tensor) -> torch.Tensor:
    """Mel-scale f0 quantisation to 1..255 (utils.py:69-80)."""
    f0_bin, f0_max, f0_min = 256, 1100.0, 50.0
    mel_min = 1127 * math.log(1 + f0_min / 700)
    mel_max = 1127 * math.log(1 + f0_max / 700)
    a = (f0_bin - 2) / (mel_max - mel_min)
    b = mel_min * a - 1.0
    mel = 1127 * (1 + f0 / 700).log()
    mel = torch.where(mel > 0, mel * a - b, mel)
    q = torch.round(mel).long()
    q = q * (q > 0)
    q = q + ((q < 1) * 1)
    q = q * (q < f0_bin)
    q = q + ((q >= f0_bin) * (f0_bin - 1))
    return q


class ChannelNorm(nn.Module):
    """LayerNorm over the channel axis of [B,C,T] (modules/modules.py:23-35)."""

    def __init__(self, channels: int, eps: float = 1e-5):
        super().__init__()
        self.channels, self.eps = channels, eps
        self.gamma = nn.Parameter(torch.ones(channels))
        self.beta = nn.Parameter(torch.zeros(channels))

    def forward(self, x):
        return F.layer_norm(x.transpose(1, 2), (self.channels,), self.gamma, self.beta, self.eps).transpose(1, 2)


class _TF32Like:
    """The reference's 1x1 convolutions run through cuDNN, whose CUDA default is TF32.  The same mixes are
    plain GEMMs here; let them follow torch's *convolution* precision switch so that `cudnn.conv.fp32_precision='ieee'`
    still gives a strict-fp32 prefix."""

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        want = False
        if torch.cuda.is_available():
            try:
                want = torch.backends.cudnn.conv.fp32_precision == "tf32"
            except Exception:
                want = bool(torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = want
        return self

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev
        return False


class WindowedRelAttention(nn.Module):
    """Multi-head self-attention with shared windowed relative-position keys/values
    (modules/attentions.py:161-239, heads_share=True, window_size=4).  Works on time-major [B,T,C] activations."""

    def __init__(self, channels: int, n_heads: int, window: int = 4):
        super().__init__()
        self.n_heads, self.window = n_heads, window
        self.dk = channels // n_heads
        self.conv_q = nn.Conv1d(channels, channels, 1)
        self.conv_k = nn.Conv1d(channels, channels, 1)
        self.conv_v = nn.Conv1d(channels, channels, 1)
        self.conv_o = nn.Conv1d(channels, channels, 1)
        self.emb_rel_k = nn.Parameter(torch.randn(1, 2 * window + 1, self.dk) * self.dk ** -0.5)
        self.emb_rel_v = nn.Parameter(torch.randn(1, 2 * window + 1, self.dk) * self.dk ** -0.5)
        self._band_cache = {}

    def _band(self, L: int, device):
        """Flat indices i*L + j of the 2w+1 band entries of each row i and their validity (0 <= j < L)."""
        key = (L, str(device))
        hit = self._band_cache.get(key)
        if hit is None:
            i = torch.arange(L, device=device)[:, None]
            j = i + torch.arange(-self.window, self.window + 1, device=device)[None, :]
            valid = (j >= 0) & (j < L)
            idx = (i * L + j.clamp(0, L - 1)).reshape(1, 1, L * (2 * self.window + 1))
            hit = (idx, valid.reshape(1, 1, L, 2 * self.window + 1))
            self._band_cache = {key: hit}
        return hit

    def forward(self, x, attn_mask=None):
        """x: [B,T,C] -> [B,T,C]."""
        B, L, D = x.shape
        h, dk, w = self.n_heads, self.dk, self.window
        wqkv = torch.cat([self.conv_q.weight[:, :, 0], self.conv_k.weight[:, :, 0], self.conv_v.weight[:, :, 0]], 0)
        bqkv = torch.cat([self.conv_q.bias, self.conv_k.bias, self.conv_v.bias], 0)
        qkv = F.linear(x, wqkv, bqkv).view(B, L, 3, h, dk).permute(2, 0, 3, 1, 4)      # [3,B,h,L,dk]
        q = qkv[0] * (1.0 / math.sqrt(dk))
        k, v = qkv[1], qkv[2]
        scores = q @ k.transpose(-2, -1)                                                  # [B,h,L,L]
        idx, valid = self._band(L, x.device)
        nb = 2 * w + 1
        rel_k = (q @ self.emb_rel_k[0].t()) * valid                                       # [B,h,L,2w+1]
        scores.view(B, h, L * L).scatter_add_(2, idx.expand(B, h, L * nb), rel_k.reshape(B, h, L * nb))
        if attn_mask is not None:
            scores = scores.masked_fill(attn_mask == 0, -1e4)
        p = F.softmax(scores, dim=-1)
        out = p @ v
        rel_w = p.view(B, h, L * L).gather(2, idx.expand(B, h, L * nb)).view(B, h, L, nb) * valid
        out = out + rel_w @ self.emb_rel_v[0]
        out = out.transpose(1, 2).reshape(B, L, D)
        return F.linear(out, self.conv_o.weight[:, :, 0], self.conv_o.bias)


class ConvFFN(nn.Module):
    """modules/attentions.py:317-363 with same padding and ReLU; the k-tap convs are run as one GEMM over the
    concatenation of the k shifted time-major views."""

    def __init__(self, channels: int, filter_channels: int, kernel_size: int):
        super().__init__()
        self.kernel_size = kernel_size
        self.conv_1 = nn.Conv1d(channels, filter_channels, kernel_size)
        self.conv_2 = nn.Conv1d(filter_channels, channels, kernel_size)

    def _conv(self, x, conv):
        k = self.kernel_size
        if k == 1:
            return F.linear(x, conv.weight[:, :, 0], conv.bias)
        L = x.shape[1]
        cout, cin = conv.weight.shape[0], conv.weight.shape[1]
        if cin > cout:
            # wide -> narrow: one GEMM against the k tap matrices stacked along the OUTPUT, then shift-and-add the k
            # narrow slices (avoids materialising the k-times wider im2col tensor)
            wst = conv.weight.permute(2, 0, 1).reshape(k * cout, cin)                     # [k*Cout, Cin], tap-major
            ya = F.pad(F.linear(x, wst), (0, 0, (k - 1) // 2, k // 2))                    # [B, L+k-1, k*Cout]
            y = ya[:, 0:L, 0:cout]
            for t in range(1, k):
                y = y + ya[:, t:t + L, t * cout:(t + 1) * cout]
            return y + conv.bias
        xp = F.pad(x, (0, 0, (k - 1) // 2, k // 2))
        cols = torch.cat([xp[:, t:t + L] for t in range(k)], dim=-1)                     # [B,T,k*C], tap-major
        wmat = conv.weight.permute(0, 2, 1).reshape(cout, -1)                             # [F, k*C]
        return F.linear(cols, wmat, conv.bias)

    def forward(self, x, x_mask=None):
        """x: [B,T,C]; x_mask: [B,T,1] or None (all ones)."""
        if x_mask is not None:
            x = x * x_mask
        y = torch.relu(self._conv(x, self.conv_1))
        if x_mask is not None:
            y = y * x_mask
        y = self._conv(y, self.conv_2)
        return y if x_mask is None else y * x_mask


class ChannelNormT(ChannelNorm):
    """Same parameters as ChannelNorm, applied to time-major [B,T,C] (no transposes)."""

    def forward(self, x):
        return F.layer_norm(x, (self.channels,), self.gamma, self.beta, self.eps)


class RelEncoder(nn.Module):
    """modules/attentions.py:73-107 (dropout omitted: inference only).  Internally time-major."""

    def __init__(self, hidden, filter_channels, n_heads, n_layers, kernel_size, window=4):
        super().__init__()
        self.attn_layers = nn.ModuleList(WindowedRelAttention(hidden, n_heads, window) for _ in range(n_layers))
        self.norm_layers_1 = nn.ModuleList(ChannelNormT(hidden) for _ in range(n_layers))
        self.ffn_layers = nn.ModuleList(ConvFFN(hidden, filter_channels, kernel_size) for _ in range(n_layers))
        self.norm_layers_2 = nn.ModuleList(ChannelNormT(hidden) for _ in range(n_layers))

    def forward(self, x, x_mask, all_ones_mask: bool = False):
        """x: [B,C,T], x_mask: [B,1,T] -> [B,C,T]."""
        mt = None if all_ones_mask else x_mask.transpose(1, 2)                            # [B,T,1]
        attn_mask = None if all_ones_mask else x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)
        x = x.transpose(1, 2).contiguous()
        if mt is not None:
            x = x * mt
        with _TF32Like():
            for attn, n1, ffn, n2 in zip(self.attn_layers, self.norm_layers_1, self.ffn_layers, self.norm_layers_2):
                x = n1(x + attn(x, attn_mask))
                x = n2(x + ffn(x, mt))
        if mt is not None:
            x = x * mt
        return x.transpose(1, 2)


class PriorEncoder(nn.Module):
    """``enc_p`` (models.py:128-162): f0 embedding + RelEncoder + projection to (m, logs)."""

    def __init__(self, out_channels, hidden, filter_channels, n_heads, n_layers, kernel_size):
        super().__init__()
        self.out_channels = out_channels
        self.proj = nn.Conv1d(hidden, out_channels * 2, 1)
        self.f0_emb = nn.Embedding(256, hidden)
        self.enc_ = RelEncoder(hidden, filter_channels, n_heads, n_layers, kernel_size)

    def forward(self, x, x_mask, f0_coarse, noice_scale=1.0, z_noise=None, all_ones_mask=False):
        x = x + self.f0_emb(f0_coarse).transpose(1, 2)
        x = self.enc_(x * x_mask, x_mask, all_ones_mask)
        stats = self.proj(x) * x_mask
        m, logs = torch.split(stats, self.out_channels, dim=1)
        if z_noise is None:
            z_noise = torch.randn_like(m)                 # RNG draw #1 (models.py:160)
        z = (m + z_noise * torch.exp(logs) * noice_scale) * x_mask
        return z, m, logs, x_mask
