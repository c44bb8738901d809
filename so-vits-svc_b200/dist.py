"""Batch-of-utterances sharding over the GPUs of one box.

Utterances are independent (no cross-item op anywhere in ``SynthesizerTrn.infer``, models.py:495-532),
so the only collective is ONE broadcast of the checkpoint at start-up; steady state has no communication.
One process per GPU, ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in CPU tests).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition: rank r gets items [lo, hi); the first n_items % world ranks get one extra."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def broadcast_state_dict(sd: Optional[Dict[str, torch.Tensor]], shapes: "OrderedDict[str, tuple]", src: int = 0,
                         device: torch.device = torch.device("cpu")) -> Dict[str, torch.Tensor]:
    """Rank ``src`` holds ``sd``; every rank returns the same fp32 state_dict (CPU tensors).
    The tensors travel as one flat buffer so a single collective (ncclBroadcast) is issued."""
    total = sum(int(torch.Size(s).numel()) for s in shapes.values())
    flat = torch.empty(total, dtype=torch.float32, device=device)
    if dist.get_rank() == src:
        assert sd is not None
        off = 0
        for k, s in shapes.items():
            n = int(torch.Size(s).numel())
            flat[off:off + n].copy_(sd[k].reshape(-1).to(torch.float32))
            off += n
    dist.broadcast(flat, src=src)
    out = OrderedDict()
    host = flat.cpu()
    off = 0
    for k, s in shapes.items():
        n = int(torch.Size(s).numel())
        out[k] = host[off:off + n].reshape(s).clone()
        off += n
    return out


def gather_waveforms(local: torch.Tensor, world: int) -> Optional[torch.Tensor]:
    """Optional all-gather of equal-sized per-rank outputs [B_local,1,N] -> [B_local*world,1,N]."""
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous())
    return torch.cat(parts, dim=0)
