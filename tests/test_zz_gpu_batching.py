"""SURVEY §8 f-2 on the GPU: batched slice inference against the serial per-slice calls it replaces."""
import json

import pytest
import torch

import sovits_b200
from sovits_b200 import batching, models, synth
from sovits_b200.config import load_config

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def test_infer_slices_matches_serial_calls():
    cfg = load_config()
    sd = synth.synth_state_dict(cfg)
    with open(sovits_b200.DEFAULT_CONFIG) as f:
        kw = json.load(f)["model"]
    net = models.SynthesizerTrn(1025, 20, **kw).eval()
    net.load_state_dict(sd)
    net = net.to(DEV)
    net.set_precision("fp32")
    prev = torch.backends.cudnn.conv.fp32_precision
    torch.backends.cudnn.conv.fp32_precision = "ieee"
    try:
        items = []
        for n, T in enumerate((60, 52, 47, 20)):
            c, f0, uv, sid = synth.synth_inputs(cfg, 1, T, seed=100 + n)
            items.append(dict(c=c[0], f0=f0[0], uv=uv[0], sid=int(sid[0, 0])))
        got = batching.infer_slices(net, items, noice_scale=0.4, max_batch=3, max_pad_ratio=1.3)
        assert len(got) == 4
        guard = 16 * cfg.hop                       # generator receptive field at the end of an item (module docstring)
        for it, o in zip(items, got):
            T = it["f0"].shape[-1]
            ref, _ = net.infer(it["c"][None].to(DEV), it["f0"][None].to(DEV), it["uv"][None].to(DEV),
                               g=torch.tensor([[it["sid"]]], device=DEV), noice_scale=0.4)
            assert o.shape == (T * cfg.hop,)
            err = float((o[:-guard] - ref[0, 0, :-guard]).abs().max())
            print(f"[parity] batched slice T={T}: L-inf vs serial call (excluding the last 16 frames) = {err:.3e}")
            assert err < 2e-4
    finally:
        torch.backends.cudnn.conv.fp32_precision = prev


def test_speaker_mix_matches_reference_fixture():
    """time-varying conditioning g[1,768,T] (EnableCharacterMix) through the CUDA flow and generator against the
    reference's own speaker-mix run (tests/golden/make_golden_mix.py), in both precisions.  On the tensor-core path the
    conditioning is a per-(frame, column) bias of the gate epilogue (flow) and of conv_pre (generator): no FFMA fallback."""
    import os
    import numpy as np
    from sovits_b200.engine import TailEngine
    cfg = load_config()
    sd = synth.synth_state_dict(cfg)
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_infer_mix_t26.npz"))
    T = int(gold["T"])
    noise = synth.draw_noise(1, T, cfg, seed=int(gold["seed"]))
    eng = TailEngine(cfg, DEV, "fp32")
    eng.load_state_dict(sd)
    for precision, tol in (("fp32", 1e-4), ("tc", 5e-3)):
        eng.set_precision(precision)
        got = eng.infer_tail(torch.from_numpy(gold["z_p"]).to(DEV), torch.from_numpy(gold["g"]).to(DEV), torch.from_numpy(gold["f0"]).to(DEV),
                             noise["rand_ini"].to(DEV), noise["har_noise"].to(DEV)).cpu()
        err = float((got - torch.from_numpy(gold["o"])).abs().max())
        print(f"[parity] speaker mix (time-varying g) {precision}: L-inf vs reference waveform = {err:.3e}")
        assert err < tol
    assert eng.fallback_count == 0
    eng.close()
