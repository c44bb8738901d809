"""SURVEY §8(e) / BASELINE config 3 on the GPU: one process per rank, the checkpoint broadcast once, every rank infers its
contiguous block of the global batch with its own re-seeded RNG - and rank r's output must equal, bit for bit, a
single-process run on items [r*B, (r+1)*B) (the 64-item single-process run is NOT the oracle: `infer` re-seeds per call).
Two ranks; NCCL when the box has two GPUs, otherwise gloo for the broadcast with both ranks computing on cuda:0."""
import json
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

B_LOCAL, T = 2, 40


def _build_net(sd, dev):
    import sovits_b200
    from sovits_b200 import models
    with open(sovits_b200.DEFAULT_CONFIG) as f:
        kw = json.load(f)["model"]
    net = models.SynthesizerTrn(1025, 20, **kw).eval()
    net.load_state_dict(sd)
    net = net.to(dev)
    net.set_precision("tc")
    return net


def _worker(rank, world, port, use_nccl, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import sovits_b200  # noqa: F401
    from sovits_b200 import dist as sdist, synth
    from sovits_b200.config import load_config
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device(f"cuda:{rank}" if use_nccl else "cuda:0")
    torch.cuda.set_device(dev)
    if use_nccl:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = load_config()
        shapes = synth.param_shapes(cfg)
        sd = synth.synth_state_dict(cfg) if rank == 0 else None
        sd = sdist.broadcast_state_dict(sd, shapes, src=0, device=dev if use_nccl else torch.device("cpu"))
        net = _build_net(sd, dev)
        c, f0, uv, sid = synth.synth_inputs(cfg, B_LOCAL * world, T)
        lo, hi = sdist.shard_range(B_LOCAL * world, rank, world)
        o, _ = net.infer(c[lo:hi].to(dev), f0[lo:hi].to(dev), uv[lo:hi].to(dev), g=sid[lo:hi].to(dev), noice_scale=0.4)
        q.put((rank, lo, hi, o.cpu()))
    finally:
        dist.destroy_process_group()


def test_rank_shards_equal_single_process_runs():
    from sovits_b200 import synth
    from sovits_b200.config import load_config
    world = 2
    use_nccl = torch.cuda.device_count() >= 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, use_nccl, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    cfg = load_config()
    sd = synth.synth_state_dict(cfg)
    c, f0, uv, sid = synth.synth_inputs(cfg, B_LOCAL * world, T)
    for rank, lo, hi, o in res:
        assert (lo, hi) == (rank * B_LOCAL, (rank + 1) * B_LOCAL)
        dev = torch.device(f"cuda:{rank}" if use_nccl else "cuda:0")
        net = _build_net(sd, dev)
        want, _ = net.infer(c[lo:hi].to(dev), f0[lo:hi].to(dev), uv[lo:hi].to(dev), g=sid[lo:hi].to(dev), noice_scale=0.4)
        assert torch.equal(o, want.cpu()), f"rank {rank}: shard output differs from the single-process run on items [{lo},{hi})"
    assert not torch.equal(res[0][3], res[1][3])
    print(f"[parity] config-3 shards: {world} ranks ({'nccl' if use_nccl else 'gloo, shared cuda:0'}) bit-identical to single-process runs")
