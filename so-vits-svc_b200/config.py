"""Hyper-parameters of the waveform-generation path.

The reference keeps them in ``config.json`` (generated from
``configs_template/config_template.json:42-71`` by ``preprocess_flist_config.py:87-117``)
and passes ``hps.model`` as keyword arguments to ``SynthesizerTrn`` (``models.py:344-372``).
This module holds the subset the hot path needs and validates what the CUDA tail supports.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import List

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_CONFIG = os.path.join(_HERE, "configs", "config_44k_vec768.json")


@dataclass
class ModelCfg:
    inter_channels: int = 192
    hidden_channels: int = 192
    filter_channels: int = 768
    n_heads: int = 2
    n_layers: int = 6
    kernel_size: int = 3
    p_dropout: float = 0.1
    resblock: str = "1"
    resblock_kernel_sizes: List[int] = field(default_factory=lambda: [3, 7, 11])
    resblock_dilation_sizes: List[List[int]] = field(default_factory=lambda: [[1, 3, 5]] * 3)
    upsample_rates: List[int] = field(default_factory=lambda: [8, 8, 2, 2, 2])
    upsample_initial_channel: int = 512
    upsample_kernel_sizes: List[int] = field(default_factory=lambda: [16, 16, 4, 4, 4])
    n_flow_layer: int = 4
    gin_channels: int = 768
    ssl_dim: int = 768
    n_speakers: int = 8
    sampling_rate: int = 44100
    vocoder_name: str = "nsf-hifigan"
    vol_embedding: bool = False
    use_depthwise_conv: bool = False
    flow_share_parameter: bool = False
    use_automatic_f0_prediction: bool = True
    use_transformer_flow: bool = False
    # WN inside each coupling layer (models.py:441: kernel 5, dilation_rate 1, n_layers = n_flow_layer)
    flow_kernel_size: int = 5
    flow_wn_layers: int = 4
    enc_window: int = 4          # attentions.py:74 default window_size
    n_harmonics: int = 9         # hifigan/models.py:332 harmonic_num=8 -> dim 9
    num_mels: int = 0            # > 0: mel-conditioned vocoder vdecoder/nsf_hifigan (no flow / speaker conditioning)

    @property
    def hop(self) -> int:
        p = 1
        for u in self.upsample_rates:
            p *= u
        return p

    @property
    def snake(self) -> bool:
        """vdecoder/hifiganwithsnake: every LeakyReLU replaced by anti-aliased SnakeBeta (models.py:426-431)."""
        return self.vocoder_name == "nsf-snake-hifigan"

    @property
    def stage_channels(self) -> List[int]:
        return [self.upsample_initial_channel // (2 ** (i + 1)) for i in range(len(self.upsample_rates))]

    def check_cuda_tail_supported(self) -> None:
        """Raise for configurations the sm_100a tail does not implement (no silent fallback)."""
        if self.use_depthwise_conv:
            raise NotImplementedError("use_depthwise_conv (config_tiny) is CPU-reference-only (SURVEY §2 row 7)")
        if self.use_transformer_flow:
            raise NotImplementedError("use_transformer_flow is outside the hot path (SURVEY §8)")
        if self.resblock != "1":
            raise NotImplementedError("only ResBlock1 generators are implemented")
        if self.vocoder_name not in ("nsf-hifigan", "nsf-snake-hifigan"):
            raise NotImplementedError(f"vocoder {self.vocoder_name!r} not implemented in the CUDA tail")
        if len(self.resblock_kernel_sizes) != 3:
            raise NotImplementedError("generator expects three ResBlock branches")


def model_cfg_from_dict(model: dict, sampling_rate: int = 44100) -> ModelCfg:
    known = {k: model[k] for k in ModelCfg.__dataclass_fields__ if k in model}
    cfg = ModelCfg(**known)
    cfg.sampling_rate = model.get("sampling_rate", sampling_rate)
    cfg.flow_wn_layers = cfg.n_flow_layer
    return cfg


def load_config(path: str = DEFAULT_CONFIG) -> ModelCfg:
    with open(path) as f:
        d = json.load(f)
    return model_cfg_from_dict(d["model"], d.get("data", {}).get("sampling_rate", 44100))
