"""sovits_b200 — B200-native (sm_100a) waveform-generation tail for so-vits-svc.

Scope: ``SynthesizerTrn.infer`` -> reverse ``ResidualCouplingBlock`` flow ->
NSF source -> HiFiGAN ``Generator``, behind the reference's ``models.SynthesizerTrn`` surface.
The compute lives in ``libsovits_b200.so`` (hand-written CUDA behind a C ABI, ``include/sovits_b200.h``);
this package is the Python host side.  There is no CPU fallback: importing ``engine`` without the
built library raises.
"""
from .config import ModelCfg, load_config, model_cfg_from_dict, DEFAULT_CONFIG  # noqa: F401

__version__ = "0.1.0"
