"""CPU tier: the N>1 host logic (shard + one weight broadcast) on gloo, world_size 2."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sovits_b200 import dist as sdist


def test_shard_range_partitions():
    for n in (1, 7, 8, 64, 65):
        for w in (1, 2, 4, 8):
            spans = [sdist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from collections import OrderedDict
        shapes = OrderedDict([("a.weight", (3, 4, 5)), ("b.bias", (7,)), ("c.weight_g", (2, 1, 1))])
        sd = None
        if rank == 0:
            g = torch.Generator().manual_seed(5)
            sd = {k: torch.randn(s, generator=g) for k, s in shapes.items()}
        got = sdist.broadcast_state_dict(sd, shapes, src=0)
        lo, hi = sdist.shard_range(5, rank, world)
        local = torch.arange(lo, hi, dtype=torch.float32)
        q.put((rank, {k: v.clone() for k, v in got.items()}, (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_broadcast_and_shard_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda t: t[0])
    g = torch.Generator().manual_seed(5)
    for k, s in [("a.weight", (3, 4, 5)), ("b.bias", (7,)), ("c.weight_g", (2, 1, 1))]:
        want = torch.randn(s, generator=g)
        assert torch.equal(res[0][1][k], want) and torch.equal(res[1][1][k], want)
    assert res[0][2] == (0, 3) and res[1][2] == (3, 5)
