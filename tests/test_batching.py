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
    c, f0, uv, sid, lengths, nz, vol = batching.pad_batch(items, [1, 2, 0], cfg, dev)
    assert vol is None
    assert c.shape == (3, cfg.ssl_dim, 9) and lengths.tolist() == [9, 7, 5] and sid[:, 0].tolist() == [1, 2, 0]
    assert torch.equal(c[2, :, :5], items[0]["c"]) and float(c[2, :, 5:].abs().max()) == 0
    assert float(nz["z_noise"][2, :, 5:].abs().max()) == 0 and float(nz["har_noise"][1, 7 * cfg.hop:].abs().max()) == 0
    one = batching.replay_item_noise(7, cfg, dev)
    assert torch.equal(nz["z_noise"][1, :, :7], one["z_noise"][0])


class _FakeNet:
    """A deterministic stand-in with the SynthesizerTrn surface the Svc patch touches (CPU test of the two-pass wiring)."""

    class _Cfg:
        hop = 4

    def __init__(self):
        self.cfg = self._Cfg()
        self._b200_cfg = self.cfg
        self.calls = 0

    def infer(self, c, f0, uv, g=None, noice_scale=0.35, seed=52468, predict_f0=False, vol=None):
        self.calls += 1
        T = f0.shape[-1]
        return (f0.repeat_interleave(self.cfg.hop, dim=-1) * 0.01 + float(g.reshape(-1)[0]))[:, None, :], f0


class _FakeSvc:
    """The shape of inference/infer_tool.Svc.slice_inference: per slice -> get_unit_f0 -> net_g_ms.infer -> crop -> concatenate."""

    def __init__(self):
        self.net_g_ms = _FakeNet()
        self.unit_calls = 0

    def get_unit_f0(self, wav):
        self.unit_calls += 1
        T = len(wav)
        return torch.ones(1, 3, T), torch.tensor(wav, dtype=torch.float32)[None, :], torch.ones(1, T)

    def slice_inference(self, slices, spk):
        audio = []
        for wav in slices:
            c, f0, uv = self.get_unit_f0(wav)
            o, _ = self.net_g_ms.infer(c, f0=f0, g=torch.tensor([[spk]]), uv=uv, noice_scale=0.4)
            audio.extend(o[0, 0, 2:-2].tolist())            # "crop the pads"
        return audio


def test_svc_patch_two_pass_wiring(monkeypatch):
    """patch_svc: features extracted once per slice, model calls batched, stitching by the ORIGINAL loop - the result equals
    the unpatched serial run."""
    from sovits_b200 import svc_batch
    slices = [[100.0 + i for i in range(n)] for n in (5, 9, 7)]
    want = _FakeSvc().slice_inference(slices, 3)
    seen = {}

    def fake_infer_slices(net, items, noice_scale=0.4, seed=52468, max_batch=8, max_pad_ratio=1.25):
        seen["n"], seen["ns"] = len(items), noice_scale
        return [net.infer(it["c"][None], it["f0"][None], it["uv"][None], g=torch.tensor([[it["sid"]]]))[0][0, 0] for it in items]

    monkeypatch.setattr(svc_batch.batching, "infer_slices", fake_infer_slices)

    class Svc(_FakeSvc):
        pass

    svc_batch.patch_svc(Svc)
    svc_batch.patch_svc(Svc)                                  # idempotent
    s = Svc()
    got = s.slice_inference(slices, 3)
    assert got == want
    assert seen == {"n": 3, "ns": 0.4} and s.unit_calls == 3          # one feature extraction per slice, one batch of 3
    assert isinstance(s.net_g_ms, _FakeNet)                             # the real model is back in place
