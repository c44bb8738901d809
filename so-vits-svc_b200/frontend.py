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


class WindowedRelAttention(nn.Module):
    """Multi-head self-attention with shared windowed relative-position keys/values
    (modules/attentions.py:161-239, heads_share=True, window_size=4)."""

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

    def forward(self, x, attn_mask=None):
        B, D, L = x.shape
        h, dk, w = self.n_heads, self.dk, self.window
        q = self.conv_q(x).view(B, h, dk, L).transpose(2, 3) / math.sqrt(dk)
        k = self.conv_k(x).view(B, h, dk, L).transpose(2, 3)
        v = self.conv_v(x).view(B, h, dk, L).transpose(2, 3)
        scores = q @ k.transpose(-2, -1)                               # [B,h,L,L]
        rel_k = q @ self.emb_rel_k[0].t()                              # [B,h,L,2w+1]
        for r in range(2 * w + 1):
            off = r - w
            if abs(off) >= L:
                continue
            i0, i1 = max(0, -off), min(L, L - off)
            scores.diagonal(off, -2, -1).add_(rel_k[:, :, i0:i1, r])
        if attn_mask is not None:
            scores = scores.masked_fill(attn_mask == 0, -1e4)
        p = F.softmax(scores, dim=-1)
        out = p @ v
        rel_w = p.new_zeros(B, h, L, 2 * w + 1)
        for r in range(2 * w + 1):
            off = r - w
            if abs(off) >= L:
                continue
            i0, i1 = max(0, -off), min(L, L - off)
            rel_w[:, :, i0:i1, r] = p.diagonal(off, -2, -1)
        out = out + rel_w @ self.emb_rel_v[0]
        out = out.transpose(2, 3).reshape(B, D, L)
        return self.conv_o(out)


class ConvFFN(nn.Module):
    """modules/attentions.py:317-363 with same padding and ReLU."""

    def __init__(self, channels: int, filter_channels: int, kernel_size: int):
        super().__init__()
        self.kernel_size = kernel_size
        self.conv_1 = nn.Conv1d(channels, filter_channels, kernel_size)
        self.conv_2 = nn.Conv1d(filter_channels, channels, kernel_size)

    def _pad(self, x):
        k = self.kernel_size
        return x if k == 1 else F.pad(x, ((k - 1) // 2, k // 2))

    def forward(self, x, x_mask):
        x = torch.relu(self.conv_1(self._pad(x * x_mask)))
        return self.conv_2(self._pad(x * x_mask)) * x_mask


class RelEncoder(nn.Module):
    """modules/attentions.py:73-107 (dropout omitted: inference only)."""

    def __init__(self, hidden, filter_channels, n_heads, n_layers, kernel_size, window=4):
        super().__init__()
        self.attn_layers = nn.ModuleList(WindowedRelAttention(hidden, n_heads, window) for _ in range(n_layers))
        self.norm_layers_1 = nn.ModuleList(ChannelNorm(hidden) for _ in range(n_layers))
        self.ffn_layers = nn.ModuleList(ConvFFN(hidden, filter_channels, kernel_size) for _ in range(n_layers))
        self.norm_layers_2 = nn.ModuleList(ChannelNorm(hidden) for _ in range(n_layers))

    def forward(self, x, x_mask, all_ones_mask: bool = False):
        attn_mask = None if all_ones_mask else x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)
        x = x * x_mask
        for attn, n1, ffn, n2 in zip(self.attn_layers, self.norm_layers_1, self.ffn_layers, self.norm_layers_2):
            x = n1(x + attn(x, attn_mask))
            x = n2(x + ffn(x, x_mask))
        return x * x_mask


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
