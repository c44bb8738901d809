"""SURVEY §8 f-2 host logic: batch planning and per-item noise replay (CPU)."""
import torch

import sovits_b200  # noqa: F401
from sovits_b200 import batching
from sovits_b200.config import load_config


def test_plan_batches_partitions_and_bounds_padding():
    lens = [100, 37, 98, 860, 862, 400, 97, 36, 35, 34, 33, 32, 31, 30, 410]
    plan = batching.plan_batches(lens, max_batch=4, max_pad_ratio=1.25)
    flat = sorted(i for b in plan for i in b)
    assert flat == list(range(len(lens)))
    for b in plan:
        assert 1 <= len(b) <= 4
        ls = [lens[i] for i in b]
        assert max(ls) <= 1.25 * min(ls)
    assert batching.plan_batches([], 8) == []
    assert batching.plan_batches([5], 8) == [[0]]


def test_replayed_noise_is_what_a_serial_call_draws():
    cfg = load_config()
    dev = torch.device("cpu")
    T = 7
    nz = batching.replay_item_noise(T, cfg, dev, seed=52468)
    torch.manual_seed(52468)
    a = torch.randn(1, cfg.inter_channels, T)
    b = torch.rand(1, cfg.n_harmonics)
    c = torch.randn(1, T * cfg.hop, cfg.n_harmonics)
    assert torch.equal(nz["z_noise"], a) and torch.equal(nz["rand_ini"], b) and torch.equal(nz["har_noise"], c)


def test_pad_batch_layout():
    cfg = load_config()
    dev = torch.device("cpu")
    items = [dict(c=torch.randn(cfg.ssl_dim, T), f0=torch.rand(T) * 300, uv=torch.ones(T), sid=s) for T, s in ((5, 0), (9, 1), (7, 2))]
    c, f0, uv, sid, lengths, nz = batching.pad_batch(items, [1, 2, 0], cfg, dev)
    assert c.shape == (3, cfg.ssl_dim, 9) and lengths.tolist() == [9, 7, 5] and sid[:, 0].tolist() == [1, 2, 0]
    assert torch.equal(c[2, :, :5], items[0]["c"]) and float(c[2, :, 5:].abs().max()) == 0
    assert float(nz["z_noise"][2, :, 5:].abs().max()) == 0 and float(nz["har_noise"][1, 7 * cfg.hop:].abs().max()) == 0
    one = batching.replay_item_noise(7, cfg, dev)
    assert torch.equal(nz["z_noise"][1, :, :7], one["z_noise"][0])
