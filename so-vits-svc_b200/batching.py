"""Batched slice inference.

The reference's ``Svc.slice_inference`` (inference/infer_tool.py:446-496) cuts the input at silences, pads every slice
with ``pad_seconds`` (0.5 s) of silence on both sides, runs ``SynthesizerTrn.infer`` on ONE slice at a time (B = 1, with a
device->host sync and a Python list append per slice) and crops the pad from each result.  Here the slices of similar
length are padded to a common length and run as one batch:

* every item draws its noise exactly like its own serial call would (``infer`` re-seeds with the same seed at every call,
  models.py:498-501, so each slice consumes a prefix of the same random stream): the per-item draws are replayed and
  zero-padded;
* ``enc_p`` and the flow run with the per-item length mask (the reference's own mask plumbing, commons.sequence_mask,
  modules/modules.py:134,138), so frames of an item never see the padding of the batch;
* the generator has no mask in the reference either; its receptive field is about 13 frames (conv_pre 3 + stage-0
  ResBlocks 7.5 + ups / later stages < 3), so only the last ~13 frames of an item can differ from the serial run - they lie
  inside the 0.5 s (43 frames) of slice padding that the caller crops anyway.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

from .frontend import f0_to_coarse


def plan_batches(lengths: Sequence[int], max_batch: int = 8, max_pad_ratio: float = 1.25) -> List[List[int]]:
    """Group item indices into batches of similar length: sort by length (descending), open a new batch when the batch is
    full or the next item is shorter than ``longest / max_pad_ratio``.  Every index appears exactly once."""
    order = sorted(range(len(lengths)), key=lambda i: -int(lengths[i]))
    batches: List[List[int]] = []
    cur: List[int] = []
    for i in order:
        if cur and (len(cur) >= max_batch or int(lengths[cur[0]]) > max_pad_ratio * int(lengths[i])):
            batches.append(cur)
            cur = []
        cur.append(i)
    if cur:
        batches.append(cur)
    return batches


def replay_item_noise(T: int, cfg, device, seed: int = 52468) -> Dict[str, torch.Tensor]:
    """The three tensors a serial ``infer`` call on a slice of T frames draws (models.py:160; vdecoder/hifigan/models.py:
    147,266; the fourth draw, :319, is discarded by the reference), in its order, after its re-seed."""
    if device == torch.device("cuda"):                  # same quirk as models.py:498-501
        torch.cuda.manual_seed_all(seed)
    else:
        torch.manual_seed(seed)
    N = T * cfg.hop
    return {"z_noise": torch.randn(1, cfg.inter_channels, T, device=device),
            "rand_ini": torch.rand(1, cfg.n_harmonics, device=device),
            "har_noise": torch.randn(1, N, cfg.n_harmonics, device=device)}


def pad_batch(items: Sequence[dict], idx: Sequence[int], cfg, device, seed: int = 52468):
    """Zero-pad the items ``idx`` to their longest length and replay their noise.  Item = dict(c [ssl,T], f0 [T], uv [T],
    sid int[, vol [T]])."""
    Ts = [int(items[i]["f0"].shape[-1]) for i in idx]
    Tm, B = max(Ts), len(idx)
    ssl = items[idx[0]]["c"].shape[0]
    c = torch.zeros(B, ssl, Tm, device=device)
    f0 = torch.zeros(B, Tm, device=device)
    uv = torch.zeros(B, Tm, device=device)
    vol = torch.zeros(B, Tm, device=device) if any("vol" in items[i] and items[i]["vol"] is not None for i in idx) else None
    zn = torch.zeros(B, cfg.inter_channels, Tm, device=device)
    ri = torch.zeros(B, cfg.n_harmonics, device=device)
    hn = torch.zeros(B, Tm * cfg.hop, cfg.n_harmonics, device=device)
    for b, (i, T) in enumerate(zip(idx, Ts)):
        it = items[i]
        c[b, :, :T], f0[b, :T], uv[b, :T] = it["c"].to(device).float(), it["f0"].to(device).float(), it["uv"].to(device).float()
        if vol is not None and it.get("vol") is not None:
            vol[b, :T] = it["vol"].to(device).float()
        nz = replay_item_noise(T, cfg, device, seed)
        zn[b, :, :T], ri[b], hn[b, :T * cfg.hop] = nz["z_noise"][0], nz["rand_ini"][0], nz["har_noise"][0]
    sid = torch.tensor([[int(items[i]["sid"])] for i in idx], dtype=torch.long, device=device)
    lengths = torch.tensor(Ts, device=device)
    return c, f0, uv, sid, lengths, {"z_noise": zn, "rand_ini": ri, "har_noise": hn}, vol


@torch.no_grad()
def infer_slices(net, items: Sequence[dict], noice_scale: float = 0.4, seed: int = 52468, max_batch: int = 8,
                 max_pad_ratio: float = 1.25) -> List[torch.Tensor]:
    """Run ``sovits_b200.models.SynthesizerTrn`` on a list of slices; returns the waveforms [512*T_i] in input order."""
    cfg = net.cfg
    dev = next(net.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("sovits_b200: infer needs a CUDA (B200) device; there is no CPU fallback")
    out: List[torch.Tensor] = [None] * len(items)       # type: ignore[list-item]
    lens = [int(it["f0"].shape[-1]) for it in items]
    for idx in plan_batches(lens, max_batch, max_pad_ratio):
        c, f0, uv, sid, lengths, nz, vol = pad_batch(items, idx, cfg, dev, seed)
        B, _, Tm = c.shape
        x_mask = (torch.arange(Tm, device=dev)[None, :] < lengths[:, None]).to(c.dtype)[:, None, :]
        g = net.emb_g(sid).transpose(1, 2)
        x = net.pre(c) * x_mask + net.emb_uv(uv.long()).transpose(1, 2)
        if vol is not None:
            if not getattr(net, "vol_embedding", False):
                raise RuntimeError("items carry `vol` but the model has no volume embedding (vol_embedding=False)")
            x = x + net.emb_vol(vol[:, :, None]).transpose(1, 2)          # models.py:518-520
        all_ones = all(lens[i] == Tm for i in idx)          # host-side: no device sync
        z_p, _, _, _ = net.enc_p(x, x_mask, f0_to_coarse(f0), noice_scale=noice_scale, z_noise=nz["z_noise"], all_ones_mask=all_ones)
        eng = net._engine(dev)
        o = eng.infer_tail(z_p, g, f0, nz["rand_ini"], nz["har_noise"], None if all_ones else lengths)
        for b, i in enumerate(idx):
            out[i] = o[b, 0, :lens[i] * cfg.hop].clone()
    return out
