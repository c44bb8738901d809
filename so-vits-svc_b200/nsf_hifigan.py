"""Drop-in for ``vdecoder.nsf_hifigan.models`` (the mel-conditioned NSF-HiFiGAN used by the enhancer and the shallow-
diffusion vocoder): ``load_model(model_path, device) -> (generator, h)`` and
``generator(mel[B,num_mels,T], f0[B,T]) -> wav[B,1,T*hop]`` (vdecoder/nsf_hifigan/models.py:17-27,221-287; callers
modules/enhancer.py:87,106 and diffusion/vocoder.py:47-95).

The module holds the checkpoint's parameters under the reference's own names; the arithmetic is the same sm_100a
library as the SVC decoder (``svb_vocoder``): fp64 closed-form harmonic source with the frame-rate ``rand_ini``
convention of this vocoder's SineGen (:146-148), tcgen05 generator, no speaker conditioning.
"""
from __future__ import annotations

import json
import os
from collections import OrderedDict

import torch
from torch import nn

from .config import ModelCfg
from .models import _ParamTree


class AttrDict(dict):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.__dict__ = self


def cfg_from_h(h) -> ModelCfg:
    g = (lambda k, d=None: h[k] if isinstance(h, dict) and k in h else getattr(h, k, d))
    cfg = ModelCfg()
    cfg.num_mels = int(g("num_mels"))
    cfg.sampling_rate = int(g("sampling_rate"))
    cfg.resblock = str(g("resblock", "1"))
    cfg.resblock_kernel_sizes = list(g("resblock_kernel_sizes"))
    cfg.resblock_dilation_sizes = [list(d) for d in g("resblock_dilation_sizes")]
    cfg.upsample_rates = list(g("upsample_rates"))
    cfg.upsample_kernel_sizes = list(g("upsample_kernel_sizes"))
    cfg.upsample_initial_channel = int(g("upsample_initial_channel"))
    cfg.vocoder_name = "nsf-hifigan"
    cfg.gin_channels = 1
    return cfg


def vocoder_param_shapes(cfg: ModelCfg) -> "OrderedDict[str, tuple]":
    """state_dict keys of vdecoder.nsf_hifigan Generator before remove_weight_norm (models.py:221-258)."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    U = cfg.upsample_initial_channel
    n_up = len(cfg.upsample_rates)
    s["m_source.l_linear.weight"] = (1, cfg.n_harmonics)
    s["m_source.l_linear.bias"] = (1,)
    for i in range(n_up):
        c_cur = U // (2 ** (i + 1))
        stride = 1
        for u in cfg.upsample_rates[i + 1:]:
            stride *= u
        s[f"noise_convs.{i}.weight"] = (c_cur, 1, 2 * stride if i + 1 < n_up else 1)
        s[f"noise_convs.{i}.bias"] = (c_cur,)
    s["conv_pre.bias"] = (U,)
    s["conv_pre.weight_g"] = (U, 1, 1)
    s["conv_pre.weight_v"] = (U, cfg.num_mels, 7)
    for i, k in enumerate(cfg.upsample_kernel_sizes):
        cin, cout = U // (2 ** i), U // (2 ** (i + 1))
        s[f"ups.{i}.bias"] = (cout,)
        s[f"ups.{i}.weight_g"] = (cin, 1, 1)
        s[f"ups.{i}.weight_v"] = (cin, cout, k)
    nk = len(cfg.resblock_kernel_sizes)
    for i in range(n_up):
        ch = U // (2 ** (i + 1))
        for j, k in enumerate(cfg.resblock_kernel_sizes):
            for grp in ("convs1", "convs2"):
                for d in range(len(cfg.resblock_dilation_sizes[j])):
                    r = f"resblocks.{i * nk + j}.{grp}.{d}."
                    s[r + "bias"] = (ch,)
                    s[r + "weight_g"] = (ch, 1, 1)
                    s[r + "weight_v"] = (ch, ch, k)
    s["conv_post.bias"] = (1,)
    s["conv_post.weight_g"] = (1, 1, 1)
    s["conv_post.weight_v"] = (1, U // (2 ** n_up), 7)
    return s


class Generator(nn.Module):
    """``Generator(h)`` with ``forward(x, f0)`` like vdecoder/nsf_hifigan/models.py:221-278."""

    def __init__(self, h, precision: str = "tc"):
        super().__init__()
        self.h = h
        self.cfg = cfg_from_h(h)
        self.cfg.check_cuda_tail_supported()
        self.upp = self.cfg.hop
        self.tree = _ParamTree(vocoder_param_shapes(self.cfg))
        self.precision = precision
        self._engine_obj = None
        self._dirty = True

    # state_dict keys must not carry the "tree." prefix: expose the tree's keys directly
    def state_dict(self, *a, **k):
        return OrderedDict((key[len("tree."):], v) for key, v in super().state_dict(*a, **k).items())

    def load_state_dict(self, sd, strict: bool = True, **kw):
        r = super().load_state_dict({"tree." + k: v for k, v in sd.items()}, strict=strict, **kw)
        self._dirty = True
        return r

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._dirty = True
        return r

    def remove_weight_norm(self):
        """The reference folds weight-norm here (models.py:280-287); the library folds at load, nothing to do."""
        return None

    def _engine(self, device):
        from .engine import TailEngine
        if self._engine_obj is None or self._engine_obj.device != device:
            self._engine_obj = TailEngine(self.cfg, device, self.precision)
            self._dirty = True
        if self._dirty:
            self._engine_obj.load_state_dict(self.state_dict())
            self._dirty = False
        return self._engine_obj

    @torch.no_grad()
    def forward(self, x, f0):
        if x.device.type != "cuda":
            raise RuntimeError("sovits_b200: the vocoder needs CUDA tensors on a B200; there is no CPU fallback")
        B, _, T = x.shape
        N = T * self.upp
        # SineGen's draws in order (models.py:146,175): rand(B,9) then randn_like(sine_waves[B,N,9])
        rand_ini = torch.rand(B, self.cfg.n_harmonics, device=x.device)
        har_noise = torch.randn(B, N, self.cfg.n_harmonics, device=x.device)
        return self._engine(x.device).vocoder(x, f0, rand_ini, har_noise).to(x.dtype)


def load_config(model_path):
    with open(os.path.join(os.path.split(model_path)[0], "config.json")) as f:
        return AttrDict(json.load(f))


def load_model(model_path, device="cuda"):
    """vdecoder/nsf_hifigan/models.py:17-27."""
    h = load_config(model_path)
    generator = Generator(h).to(device)
    cp_dict = torch.load(model_path, map_location="cpu")
    generator.load_state_dict(cp_dict["generator"])
    generator.eval()
    generator.remove_weight_norm()
    return generator, h
