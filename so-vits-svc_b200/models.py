"""Drop-in ``SynthesizerTrn`` for inference (reference surface: models.py:339-532).

``SynthesizerTrn`` here is an ``nn.Module`` with the reference's constructor signature, the
reference's ``state_dict`` key layout for every module ``infer`` uses (so ``utils.load_checkpoint``
+ ``load_state_dict`` work unchanged, utils.py:155-187), the module verbs ``Svc`` calls
(``.half() .eval() .to(dev) .EnableCharacterMix()``, inference/infer_tool.py:196-202) and the
``infer`` call (``:297``).  ``pre``/``enc_p`` run on stock PyTorch; ``flow`` and ``dec`` hold only
parameters — their arithmetic is the sm_100a library behind ``TailEngine``.

``patch_reference(models_module)`` instead subclasses the *reference's own* class (when the reference
tree is importable) and overrides only the tail of ``infer``; that is the zero-edit integration
described in INTEGRATION.md.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn

from .config import ModelCfg, model_cfg_from_dict
from .frontend import PriorEncoder, f0_to_coarse
from .synth import param_shapes


class _ParamTree(nn.Module):
    """Holds parameters under dotted names (e.g. ``flows.0.enc.in_layers.1.weight_v``) so that
    ``state_dict()`` reproduces the reference's keys.  Has no ``forward``: the CUDA library computes."""

    def __init__(self, shapes: Dict[str, tuple]):
        super().__init__()
        for name, shape in shapes.items():
            mod = self
            parts = name.split(".")
            for p in parts[:-1]:
                if p not in mod._modules:
                    mod.add_module(p, _ParamTree({}))
                mod = mod._modules[p]
            mod.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape), requires_grad=False))

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("flow/dec are executed by libsovits_b200 (TailEngine), not by PyTorch")


class _TailMixin:
    """Shared by the standalone class and the patched reference subclass: engine lifecycle, RNG
    draws in the reference's order and the CUDA tail call."""

    precision = "tc"
    noise_mode = "torch"   # "torch": the reference's own RNG stream (bit-comparable); "philox": harmonic noise drawn inside the
                           # source kernel (no [B,N,9] tensor written and re-read; statistically equivalent)
    own_prefix = True      # precision "tc": pre + enc_p through svb_pre_conv / svb_enc_p instead of cuBLAS / ATen

    def _tail_init(self, cfg: ModelCfg):
        self._b200_cfg = cfg
        self._b200_engine = None
        self._b200_dirty = True

    def set_precision(self, precision: str):
        assert precision in ("fp32", "tc")
        self.precision = precision
        if self._b200_engine is not None:
            self._b200_engine.set_precision(precision)

    def _engine(self, device: torch.device):
        from .engine import TailEngine  # raises ImportError if libsovits_b200.so is missing (no fallback)
        eng = self._b200_engine
        if eng is None or eng.device != device:
            if eng is not None:
                eng.close()
            eng = TailEngine(self._b200_cfg, device, self.precision)
            self._b200_engine = eng
            self._b200_dirty = True
        if self._b200_dirty:
            eng.load_state_dict(self.state_dict())
            self._b200_dirty = False
        return eng

    def _mark_dirty(self):
        self._b200_dirty = True

    def _run_tail(self, z_p, c_mask, g, f0, lengths=None):
        """models.py:530-531 + the RNG draws of vdecoder/hifigan/models.py:147,266,319."""
        dev = z_p.device
        if dev.type != "cuda":
            raise RuntimeError("sovits_b200: infer needs a CUDA (B200) device; there is no CPU fallback")
        cfg = self._b200_cfg
        B, _, T = z_p.shape
        N = T * cfg.hop
        rand_ini = torch.rand(B, cfg.n_harmonics, device=dev)
        philox = self.noise_mode == "philox"
        har_noise = None if philox else torch.randn(B, N, cfg.n_harmonics, device=dev)
        # Draw #4 (`randn_like(uv)`, vdecoder/hifigan/models.py:319) is discarded by Generator.forward (:371) and is the LAST
        # draw of a call whose first act is to re-seed (models.py:498-501): it can never influence an output, so it is not made.
        eng = self._engine(dev)
        if philox != getattr(eng, "_philox_on", False):
            eng.set_option("philox_noise", int(philox))
            eng._philox_on = philox
        # `infer` always builds an all-ones mask (c_lengths = ones * T, models.py:503,515), so no length vector is needed and
        # the mask is NOT inspected on the device (that cost a device->host sync right before the ~100 tail launches).
        o = eng.infer_tail(z_p, g, f0, rand_ini, har_noise, lengths)
        return o.to(z_p.dtype)


class SynthesizerTrn(_TailMixin, nn.Module):
    """Standalone replacement (no reference import needed).  Constructor signature: models.py:344-372."""

    def __init__(self, spec_channels, segment_size, inter_channels, hidden_channels, filter_channels,
                 n_heads, n_layers, kernel_size, p_dropout, resblock, resblock_kernel_sizes,
                 resblock_dilation_sizes, upsample_rates, upsample_initial_channel, upsample_kernel_sizes,
                 gin_channels, ssl_dim, n_speakers, sampling_rate=44100, vol_embedding=False,
                 vocoder_name="nsf-hifigan", use_depthwise_conv=False, use_automatic_f0_prediction=True,
                 flow_share_parameter=False, n_flow_layer=4, n_layers_trans_flow=3,
                 use_transformer_flow=False, **kwargs):
        super().__init__()
        cfg = model_cfg_from_dict(dict(
            inter_channels=inter_channels, hidden_channels=hidden_channels, filter_channels=filter_channels,
            n_heads=n_heads, n_layers=n_layers, kernel_size=kernel_size, p_dropout=p_dropout, resblock=resblock,
            resblock_kernel_sizes=list(resblock_kernel_sizes),
            resblock_dilation_sizes=[list(d) for d in resblock_dilation_sizes],
            upsample_rates=list(upsample_rates), upsample_initial_channel=upsample_initial_channel,
            upsample_kernel_sizes=list(upsample_kernel_sizes), gin_channels=gin_channels, ssl_dim=ssl_dim,
            n_speakers=n_speakers, vol_embedding=vol_embedding, vocoder_name=vocoder_name,
            use_depthwise_conv=use_depthwise_conv, flow_share_parameter=flow_share_parameter,
            use_automatic_f0_prediction=use_automatic_f0_prediction, n_flow_layer=n_flow_layer,
            use_transformer_flow=use_transformer_flow), sampling_rate)
        if flow_share_parameter:
            raise NotImplementedError("flow_share_parameter (config_tiny) is not implemented in the CUDA tail")
        cfg.check_cuda_tail_supported()
        self.cfg = cfg
        self.spec_channels, self.segment_size = spec_channels, segment_size
        self.gin_channels, self.ssl_dim, self.vol_embedding = gin_channels, ssl_dim, vol_embedding
        self.use_automatic_f0_prediction = use_automatic_f0_prediction
        self.emb_g = nn.Embedding(n_speakers, gin_channels)
        if vol_embedding:
            self.emb_vol = nn.Linear(1, hidden_channels)
        self.pre = nn.Conv1d(ssl_dim, hidden_channels, kernel_size=5, padding=2)
        self.enc_p = PriorEncoder(inter_channels, hidden_channels, filter_channels, n_heads, n_layers, kernel_size)
        self.emb_uv = nn.Embedding(2, hidden_channels)
        shapes = param_shapes(cfg)
        self.flow = _ParamTree({k[len("flow."):]: v for k, v in shapes.items() if k.startswith("flow.")})
        self.dec = _ParamTree({k[len("dec."):]: v for k, v in shapes.items() if k.startswith("dec.")})
        self.character_mix = False
        self._tail_init(cfg)

    # ---- nn.Module plumbing: repack lazily whenever weights may have changed
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        own = self.state_dict().keys()
        filtered = {k: v for k, v in state_dict.items() if k in own}     # enc_q.* / f0_decoder.* are not used by infer
        r = super().load_state_dict(filtered, strict=strict, **kw)
        self._mark_dirty()
        return r

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._mark_dirty()
        return r

    def EnableCharacterMix(self, n_speakers_map, device):
        """models.py:456-461: table of speaker embeddings for time-varying mixes."""
        idx = torch.arange(n_speakers_map, device=device)
        self.speaker_map = self.emb_g.to(device)(idx).reshape(1, n_speakers_map, 1, 1, self.gin_channels)
        self.character_mix = True

    def mix_speakers(self, weights):
        """[T,S] per-frame speaker weights -> time-varying conditioning g [1,gin,T] (models.py:505-509)."""
        gm = weights.reshape(weights.shape[0], weights.shape[1], 1, 1, 1) * self.speaker_map
        gm = gm.sum(dim=1)                              # [T,1,1,gin]
        return gm.transpose(0, -1).transpose(0, -2).squeeze(0)

    def conditioning(self, c, g, vol=None):
        """models.py:505-513: speaker conditioning g [B,gin,1] (or [1,gin,T] for a [T,S] speaker mix), the all-ones frame mask
        and the volume embedding (0 when the checkpoint has none or no ``vol`` is given).  Plain torch indexing, any device."""
        B, _, T = c.shape
        if self.character_mix and len(g) > 1:           # [T,S] mix weights -> g [1,gin,T] (models.py:505-509)
            g = self.mix_speakers(g)
        else:
            if g.dim() == 1:
                g = g.unsqueeze(0)
            g = self.emb_g(g).transpose(1, 2)
        x_mask = torch.ones(B, 1, T, dtype=c.dtype, device=c.device)
        v = self.emb_vol(vol[:, :, None]).transpose(1, 2) if (vol is not None and self.vol_embedding) else 0
        return g, x_mask, v

    @torch.no_grad()
    def infer(self, c, f0, uv, g=None, noice_scale=0.35, seed=52468, predict_f0=False, vol=None):
        """models.py:495-532.  ``c`` [B,ssl,T], ``f0``/``uv`` [B,T], ``g`` [B,1] int64 (or [T,S] mix)."""
        if c.device.type != "cuda":
            raise RuntimeError("sovits_b200: infer needs CUDA tensors on a B200; there is no CPU fallback")
        if c.device == torch.device("cuda"):          # same quirk as models.py:498-501
            torch.cuda.manual_seed_all(seed)
        else:
            torch.manual_seed(seed)
        B, _, T = c.shape
        g, x_mask, v = self.conditioning(c, g, vol)
        if self.use_automatic_f0_prediction and predict_f0:
            raise NotImplementedError("predict_f0 needs the reference f0_decoder: use patch_reference() (INTEGRATION.md)")
        eng = self._engine(c.device)
        if self.precision == "tc" and self.own_prefix and getattr(eng, "has_prefix", False) and c.dtype == torch.float32:
            # pre + enc_p on the library's own kernels (svb_pre_conv / svb_enc_p): tcgen05 GEMMs + the fused attention kernel.
            # The embedding gathers stay torch indexing ("a1 stays PyTorch"); RNG draw #1 keeps its place in the stream.
            x = eng.pre_conv(c) + self.emb_uv(uv.long()).transpose(1, 2) + v
            x = x + self.enc_p.f0_emb(f0_to_coarse(f0)).transpose(1, 2)
            z_noise = torch.randn(B, self.cfg.inter_channels, T, dtype=torch.float32, device=c.device)   # models.py:160
            z_p = eng.enc_p(x, z_noise, noice_scale)
            c_mask = x_mask
        else:
            x = self.pre(c) * x_mask + self.emb_uv(uv.long()).transpose(1, 2) + v
            z_p, m_p, logs_p, c_mask = self.enc_p(x, x_mask, f0_to_coarse(f0), noice_scale=noice_scale, all_ones_mask=True)
        o = self._run_tail(z_p, c_mask, g, f0)
        return o, f0


def patch_reference(ref_models, precision: str = "tc"):
    """Replace ``ref_models.SynthesizerTrn`` (the reference's models.py, already imported) by a subclass
    whose ``infer`` keeps the reference's own prefix modules and runs flow+dec through libsovits_b200.
    Everything else (constructor, state_dict, forward for training, EnableCharacterMix) is inherited."""
    Base = ref_models.SynthesizerTrn
    if getattr(Base, "_b200_patched", False):
        return Base
    import utils as ref_utils  # the reference's utils (f0_to_coarse, normalize_f0)

    class SynthesizerTrnB200(_TailMixin, Base):
        _b200_patched = True

        def __init__(self, *a, **kw):
            Base.__init__(self, *a, **kw)
            names = Base.__init__.__code__.co_varnames[3:Base.__init__.__code__.co_argcount]
            merged = dict(zip(names, a[2:]))
            merged.update(kw)
            cfg = model_cfg_from_dict(merged, merged.get("sampling_rate", 44100))
            cfg.check_cuda_tail_supported()
            if merged.get("flow_share_parameter", False):
                raise NotImplementedError("flow_share_parameter is not implemented in the CUDA tail")
            self._tail_init(cfg)
            self.precision = precision

        def load_state_dict(self, *a, **k):
            r = Base.load_state_dict(self, *a, **k)
            self._mark_dirty()
            return r

        def _apply(self, fn, *a, **k):
            r = Base._apply(self, fn, *a, **k)
            self._mark_dirty()
            return r

        @torch.no_grad()
        def infer(self, c, f0, uv, g=None, noice_scale=0.35, seed=52468, predict_f0=False, vol=None):
            if c.device == torch.device("cuda"):
                torch.cuda.manual_seed_all(seed)
            else:
                torch.manual_seed(seed)
            c_lengths = (torch.ones(c.size(0)) * c.size(-1)).to(c.device)
            if self.character_mix and len(g) > 1:
                g = g.reshape((g.shape[0], g.shape[1], 1, 1, 1)) * self.speaker_map
                g = torch.sum(g, dim=1).transpose(0, -1).transpose(0, -2).squeeze(0)
            else:
                if g.dim() == 1:
                    g = g.unsqueeze(0)
                g = self.emb_g(g).transpose(1, 2)
            x_mask = torch.unsqueeze(ref_models.commons.sequence_mask(c_lengths, c.size(2)), 1).to(c.dtype)
            v = self.emb_vol(vol[:, :, None]).transpose(1, 2) if vol is not None and self.vol_embedding else 0
            x = self.pre(c) * x_mask + self.emb_uv(uv.long()).transpose(1, 2) + v
            if self.use_automatic_f0_prediction and predict_f0:
                lf0 = 2595. * torch.log10(1. + f0.unsqueeze(1) / 700.) / 500
                norm_lf0 = ref_utils.normalize_f0(lf0, x_mask, uv, random_scale=False)
                pred_lf0 = self.f0_decoder(x, norm_lf0, x_mask, spk_emb=g)
                f0 = (700 * (torch.pow(10, pred_lf0 * 500 / 2595) - 1)).squeeze(1)
            z_p, m_p, logs_p, c_mask = self.enc_p(x, x_mask, f0=ref_utils.f0_to_coarse(f0), noice_scale=noice_scale)
            o = self._run_tail(z_p, c_mask, g, f0)
            return o, f0

    SynthesizerTrnB200.__name__ = "SynthesizerTrn"
    ref_models.SynthesizerTrn = SynthesizerTrnB200
    return SynthesizerTrnB200
