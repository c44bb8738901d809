"""Batched ``Svc.slice_inference``: the reference's own slicing / padding / cross-fade loop
(inference/infer_tool.py:446-496) with the per-slice ``net_g_ms.infer`` calls (``:470 -> :297``) collected and run as
length-bucketed batches on the B200 (``batching.infer_slices``).

Nothing of the reference loop is re-implemented.  ``patch_svc`` wraps ``Svc.slice_inference`` so that it runs TWICE:

* pass 1 - with ``net_g_ms`` replaced by a recorder that stores the arguments of every ``infer`` call and returns silence of
  the right length, and ``get_unit_f0`` (ContentVec + f0 extraction, upstream of the hot path) memoised;
* the recorded slices are synthesised in batches: every item with its own replayed noise, exactly what its serial call
  would have drawn (``infer`` re-seeds on every call, models.py:498-501);
* pass 2 - the reference loop again, ``get_unit_f0`` served from the memo and ``net_g_ms.infer`` returning the precomputed
  waveforms in call order, so cropping of the 0.5 s pads, ``pad_array`` and the linear cross-fades are the reference's code.

Configurations whose per-slice post-processing feeds the audio back into a model (shallow diffusion, the NSF-HiFiGAN
enhancer) or that use speaker mixing keep the serial path.
"""
from __future__ import annotations

from typing import List

import torch

from . import batching


class _Recorder:
    """Stands in for ``net_g_ms`` during pass 1."""

    def __init__(self, net):
        self._net = net
        self.items: List[dict] = []
        self.kwargs: List[dict] = []

    def __getattr__(self, name):          # anything but infer (dtype probes, .parameters(), ...) goes to the real model
        return getattr(self._net, name)

    @torch.no_grad()
    def infer(self, c, f0, uv, g=None, noice_scale=0.35, seed=52468, predict_f0=False, vol=None):
        if predict_f0 or c.shape[0] != 1 or g is None or g.numel() != 1:
            raise _Unsupported("only single-speaker, non-f0-predicting slices are batched")
        self.items.append(dict(c=c[0], f0=f0[0], uv=uv[0], sid=int(g.reshape(-1)[0]), vol=None if vol is None else vol.reshape(-1)))
        self.kwargs.append(dict(noice_scale=float(noice_scale), seed=int(seed)))
        T = f0.shape[-1]
        return torch.zeros(1, 1, T * self._net.cfg.hop, dtype=c.dtype, device=c.device), f0


class _Replayer:
    """Stands in for ``net_g_ms`` during pass 2: hands out the precomputed waveforms in call order."""

    def __init__(self, net, outs):
        self._net, self._outs, self._i = net, outs, 0

    def __getattr__(self, name):
        return getattr(self._net, name)

    def infer(self, c, f0, uv, g=None, noice_scale=0.35, seed=52468, predict_f0=False, vol=None):
        o = self._outs[self._i]
        self._i += 1
        return o.reshape(1, 1, -1).to(c.dtype), f0


class _Memo:
    def __init__(self, fn):
        self.fn, self.vals, self.i = fn, [], 0

    def record(self, *a, **k):
        v = self.fn(*a, **k)
        self.vals.append(v)
        return v

    def replay(self, *a, **k):
        v = self.vals[self.i]
        self.i += 1
        return v


class _Unsupported(Exception):
    pass


def _batchable(svc) -> bool:
    net = getattr(svc, "net_g_ms", None)
    if net is None or not hasattr(net, "_b200_cfg") or getattr(net, "character_mix", False):
        return False
    return not (getattr(svc, "shallow_diffusion", False) or getattr(svc, "only_diffusion", False) or getattr(svc, "nsf_hifigan_enhance", False))


def batched_slice_inference(svc, orig_slice_inference, *args, max_batch: int = 8, max_pad_ratio: float = 1.25, **kwargs):
    """Run ``orig_slice_inference(svc, *args, **kwargs)`` with its per-slice model calls batched (see module docstring)."""
    if not _batchable(svc) or kwargs.get("use_spk_mix", False):
        return orig_slice_inference(svc, *args, **kwargs)
    net, get_unit_f0 = svc.net_g_ms, svc.get_unit_f0
    rec, memo = _Recorder(net), _Memo(get_unit_f0)
    svc.net_g_ms, svc.get_unit_f0 = rec, memo.record
    try:
        try:
            orig_slice_inference(svc, *args, **kwargs)
        except _Unsupported:
            svc.net_g_ms, svc.get_unit_f0 = net, get_unit_f0
            return orig_slice_inference(svc, *args, **kwargs)
    finally:
        svc.net_g_ms, svc.get_unit_f0 = net, get_unit_f0
    if not rec.items:
        return orig_slice_inference(svc, *args, **kwargs)
    ns, seed = rec.kwargs[0]["noice_scale"], rec.kwargs[0]["seed"]
    outs = batching.infer_slices(net, rec.items, noice_scale=ns, seed=seed, max_batch=max_batch, max_pad_ratio=max_pad_ratio)
    svc.net_g_ms, svc.get_unit_f0 = _Replayer(net, outs), memo.replay
    try:
        return orig_slice_inference(svc, *args, **kwargs)
    finally:
        svc.net_g_ms, svc.get_unit_f0 = net, get_unit_f0


def patch_svc(svc_class, max_batch: int = 8, max_pad_ratio: float = 1.25):
    """Replace ``svc_class.slice_inference`` (the reference's ``inference.infer_tool.Svc``) by the batched wrapper.
    Idempotent; returns the class."""
    if getattr(svc_class, "_b200_batched", False):
        return svc_class
    orig = svc_class.slice_inference

    def slice_inference(self, *args, **kwargs):
        return batched_slice_inference(self, orig, *args, max_batch=max_batch, max_pad_ratio=max_pad_ratio, **kwargs)

    slice_inference.__doc__ = orig.__doc__
    svc_class.slice_inference = slice_inference
    svc_class._b200_batched = True
    return svc_class
