"""Prefix of ``SynthesizerTrn.infer`` that stays on stock PyTorch:
speaker/uv embeddings, the ``pre`` conv and the ``enc_p`` relative-position transformer.

Written from the reference's behaviour (models.py:128-162,495-529; modules/attentions.py:73-107,
161-303,317-363), with the same parameter names so reference checkpoints load unchanged.  The
relative-position terms are applied on the nine diagonals directly instead of the reference's
pad/reshape skewing (attentions.py:275-303) — same sums, far less memory traffic.
"""
from __future__ import annotations

import math
import os

import torch
from torch import nn
from torch.nn import functional as F


def _cached(mod: nn.Module, name: str, srcs, fn):
    """Derived weight matrices (tap-stacked conv weights, the fused q/k/v projection) are rebuilt only when a source
    parameter was re-allocated or written (load_state_dict, .to(), .half())."""
    key = tuple((t.data_ptr(), t._version, t.dtype) for t in srcs)
    cache = mod.__dict__.setdefault("_svb_cache", {})
    hit = cache.get(name)
    if hit is None or hit[0] != key:
        with torch.no_grad():
            val = fn()
        cache[name] = (key, val)
        return val
    return hit[1]


def _fused_tails(x: torch.Tensor):
    """libsovits_b200 handle when the fused layer tails (csrc/kernels_prefix.cu) apply: CUDA fp32 inference."""
    if not x.is_cuda or x.dtype != torch.float32 or torch.is_grad_enabled() or os.environ.get("SVB_PREFIX_FUSED", "1") == "0":
        return None
    from .lib import load_library
    return load_library()


def _add_ln_im2col(lib, x, r, norm, k):
    """y = LayerNorm(x + r) and the k shifted copies of y side by side ([B,L,k*C], zero padded) in one kernel."""
    B, L, C = x.shape
    x, r = x.contiguous(), r.contiguous()
    y = torch.empty_like(x)
    cols = torch.empty((B, L, k * C), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):     # raw launches go to the CURRENT device: make it the tensor's device
        rc = lib.svb_prefix_add_ln_im2col(x.data_ptr(), r.data_ptr(), norm.gamma.data_ptr(), norm.beta.data_ptr(), float(norm.eps),
                                          y.data_ptr(), cols.data_ptr(), B, L, C, k, torch.cuda.current_stream(x.device).cuda_stream)
    if rc != 0:
        raise RuntimeError(f"svb_prefix_add_ln_im2col failed: {lib.svb_strerror(rc).decode()}")
    return y, cols


def _ffn_tail(lib, ya, x, bias, norm, k):
    """LayerNorm(x + bias + shift-and-add of the k tap slices of ya) in one kernel."""
    B, L, C = x.shape
    ya, x = ya.contiguous(), x.contiguous()
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = lib.svb_prefix_ffn_tail(ya.data_ptr(), x.data_ptr(), bias.data_ptr(), norm.gamma.data_ptr(), norm.beta.data_ptr(), float(norm.eps),
                                     y.data_ptr(), B, L, C, k, torch.cuda.current_stream(x.device).cuda_stream)
    if rc != 0:
        raise RuntimeError(f"svb_prefix_ffn_tail failed: {lib.svb_strerror(rc).decode()}")
    return y


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
    """The reference's 1x1 / k-tap convolutions run through cuDNN, whose CUDA default is TF32.  The same mixes
    are plain GEMMs here; let exactly THOSE follow torch's *convolution* precision switch (so that
    `cudnn.conv.fp32_precision='ieee'` still gives a strict-fp32 prefix).  The attention products q@k^T, p@v and the
    relative-position terms are `torch.matmul` in the reference (modules/attentions.py:243-262): they stay under torch's
    matmul default (IEEE fp32) and are never wrapped in this context."""

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
        scale = 1.0 / math.sqrt(dk)
        srcs = (self.conv_q.weight, self.conv_k.weight, self.conv_v.weight, self.conv_q.bias, self.conv_k.bias, self.conv_v.bias)
        if torch.is_grad_enabled():
            wqkv = torch.cat([self.conv_q.weight[:, :, 0] * scale, self.conv_k.weight[:, :, 0], self.conv_v.weight[:, :, 0]], 0)
            bqkv = torch.cat([self.conv_q.bias * scale, self.conv_k.bias, self.conv_v.bias], 0)
        else:       # one projection for q, k, v with the 1/sqrt(dk) of attentions.py:243 folded into the q rows
            wqkv, bqkv = _cached(self, "qkv", srcs, lambda: (
                torch.cat([self.conv_q.weight[:, :, 0] * scale, self.conv_k.weight[:, :, 0], self.conv_v.weight[:, :, 0]], 0).contiguous(),
                torch.cat([self.conv_q.bias * scale, self.conv_k.bias, self.conv_v.bias], 0).contiguous()))
        with _TF32Like():                                                                 # conv_q / conv_k / conv_v
            qkv = F.linear(x, wqkv, bqkv).view(B, L, 3, h, dk).permute(2, 0, 3, 1, 4)      # [3,B,h,L,dk]
        q, k, v = qkv[0], qkv[1], qkv[2]
        scores = q @ k.transpose(-2, -1)                                                  # [B,h,L,L]
        lib = _fused_tails(x) if (attn_mask is None and L <= 12000) else None   # rel_softmax keeps one score row in shared memory
        if lib is not None:
            # equal-length batch on the GPU: band add + softmax + band extraction, then p@v + relative values + head merge,
            # as two fused kernels (csrc/kernels_prefix.cu) around the cuBLAS GEMMs
            nb_ = 2 * w + 1
            relk = (q @ self.emb_rel_k[0].t()).contiguous()                               # [B,h,L,2w+1]
            pband = torch.empty_like(relk)
            stream = torch.cuda.current_stream(x.device).cuda_stream
            with torch.cuda.device(x.device):
                rc = lib.svb_prefix_rel_softmax(scores.data_ptr(), relk.data_ptr(), pband.data_ptr(), B * h * L, L, w, stream)
            if rc != 0:
                raise RuntimeError(f"svb_prefix_rel_softmax failed: {lib.svb_strerror(rc).decode()}")
            pv = (scores @ v).contiguous()                                                # [B,h,L,dk]
            merged = torch.empty((B, L, D), dtype=x.dtype, device=x.device)
            embv = self.emb_rel_v[0].contiguous()
            with torch.cuda.device(x.device):
                rc = lib.svb_prefix_attn_merge(pv.data_ptr(), pband.data_ptr(), embv.data_ptr(), merged.data_ptr(), B, h, L, dk, nb_, stream)
            if rc != 0:
                raise RuntimeError(f"svb_prefix_attn_merge failed: {lib.svb_strerror(rc).decode()}")
            with _TF32Like():
                return F.linear(merged, self.conv_o.weight[:, :, 0], self.conv_o.bias)
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
        with _TF32Like():
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
        with _TF32Like():
            return self._conv_impl(x, conv)

    def _conv_impl(self, x, conv):
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

    def hidden_from_cols(self, cols):
        """relu(conv_1) from the k shifted copies of the input laid side by side ([B,L,k*C], tap-major)."""
        conv = self.conv_1
        wmat = _cached(self, "w1", (conv.weight,), lambda: conv.weight.permute(0, 2, 1).reshape(conv.weight.shape[0], -1).contiguous())
        B, L, KC = cols.shape
        with _TF32Like():
            if hasattr(torch, "_addmm_activation"):        # bias + ReLU in the GEMM epilogue (cuBLASLt)
                return torch._addmm_activation(conv.bias, cols.view(B * L, KC), wmat.t(), use_gelu=False).view(B, L, -1)
            return torch.relu_(F.linear(cols, wmat, conv.bias))

    def stacked_out(self, hid):
        """conv_2 as ONE GEMM against the k tap matrices stacked along the output: [B,L,k*Cout], tap-major (no bias)."""
        conv = self.conv_2
        wst = _cached(self, "w2", (conv.weight,), lambda: conv.weight.permute(2, 0, 1).reshape(-1, conv.weight.shape[1]).contiguous())
        with _TF32Like():
            return F.linear(hid, wst)

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
        lib = _fused_tails(x) if (mt is None and x.shape[-1] <= 256 and self.ffn_layers[0].kernel_size <= 7) else None
        for attn, n1, ffn, n2 in zip(self.attn_layers, self.norm_layers_1, self.ffn_layers, self.norm_layers_2):
            if lib is not None:
                # equal-length batch on the GPU: the element-wise tails of the layer run as two fused kernels
                # (csrc/kernels_prefix.cu), the GEMMs stay on cuBLAS
                x1, cols = _add_ln_im2col(lib, x, attn(x, None), n1, ffn.kernel_size)
                x = _ffn_tail(lib, ffn.stacked_out(ffn.hidden_from_cols(cols)), x1, ffn.conv_2.bias, n2, ffn.kernel_size)
                continue
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
