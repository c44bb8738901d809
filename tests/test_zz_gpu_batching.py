"""SURVEY §8 f-2 on the GPU: batched slice inference against the ORACLE run slice by slice (the serial semantics of
inference/infer_tool.py:446-496), through the Svc.slice_inference patch, with a slices/s figure against the serial calls."""
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


class _SvcLike:
    """The part of inference/infer_tool.Svc that slice_inference exercises around the model call (:446-496): per slice
    get_unit_f0 -> net_g_ms.infer -> D2H -> crop the pads -> extend a Python list."""

    def __init__(self, net, cfg, slices, pad_frames=16):      # the reference pads 0.5 s = 43 frames per side
        self.net_g_ms, self.cfg, self.slices, self.pad = net, cfg, slices, pad_frames
        self.shallow_diffusion = self.only_diffusion = self.nsf_hifigan_enhance = False

    def get_unit_f0(self, i):
        it = self.slices[i]
        return it["c"][None].to(DEV), it["f0"][None].to(DEV), it["uv"][None].to(DEV)

    def slice_inference(self, sid):
        audio = []
        for i in range(len(self.slices)):
            c, f0, uv = self.get_unit_f0(i)
            o, _ = self.net_g_ms.infer(c, f0=f0, g=torch.tensor([[sid]], device=DEV), uv=uv, noice_scale=0.4)
            a = o[0, 0].data.float().cpu().numpy()
            p = self.pad * self.cfg.hop
            audio.extend(list(a[p:-p]))
        return audio


def test_svc_slice_inference_batched_vs_oracle_and_serial():
    """patch_svc on a Svc-shaped driver with the real model: (a) every slice of the batched result matches the fp32 ORACLE run
    on that slice alone (its own replayed noise), (b) the stitched output equals the unpatched serial loop inside the pad crop,
    (c) slices/s of both."""
    import time
    import numpy as np
    import svc_oracle as O
    from sovits_b200 import svc_batch
    cfg = load_config()
    sd = synth.synth_state_dict(cfg)
    with open(sovits_b200.DEFAULT_CONFIG) as f:
        kw = json.load(f)["model"]
    net = models.SynthesizerTrn(1025, 20, **kw).eval()
    net.load_state_dict(sd)
    net = net.to(DEV)
    net.set_precision("fp32")
    lens = [172, 160, 150, 96, 90, 88, 64, 60, 40, 172, 130, 120]          # frames (0.5 - 2 s slices incl. pads)
    items = []
    for n, T in enumerate(lens):
        c, f0, uv, sid = synth.synth_inputs(cfg, 1, T, seed=300 + n)
        items.append(dict(c=c[0], f0=f0[0], uv=uv[0], sid=2))
    prev = torch.backends.cudnn.conv.fp32_precision
    torch.backends.cudnn.conv.fp32_precision = "ieee"
    try:
        # (a) per slice against the oracle
        got = batching.infer_slices(net, items[:5], noice_scale=0.4, max_batch=8, max_pad_ratio=1.3)
        guard = 16 * cfg.hop
        for it, o in zip(items[:5], got):
            T = it["f0"].shape[-1]
            nz = {k: v.cpu() for k, v in batching.replay_item_noise(T, cfg, DEV).items()}
            ref, _ = O.infer(sd, cfg, it["c"][None], it["f0"][None], it["uv"][None], torch.tensor([[2]]), nz, noice_scale=0.4)
            err = float((o.cpu()[:-guard] - ref[0, 0, :-guard]).abs().max())
            print(f"[parity] batched slice T={T} vs ORACLE on the slice alone (excluding the last 16 frames): L-inf = {err:.3e}")
            assert err < 2e-4
        # (b) + (c) through the Svc patch
        class Svc(_SvcLike):
            pass
        serial = _SvcLike(net, cfg, items)
        serial.slice_inference(2)                                   # warm-up
        torch.cuda.synchronize(); t0 = time.perf_counter()
        want = np.array(serial.slice_inference(2)); torch.cuda.synchronize(); t_serial = time.perf_counter() - t0
        svc_batch.patch_svc(Svc)
        batched = Svc(net, cfg, items)
        batched.slice_inference(2)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        gotb = np.array(batched.slice_inference(2)); torch.cuda.synchronize(); t_batched = time.perf_counter() - t0
        assert gotb.shape == want.shape
        err = float(np.abs(gotb - want).max())
        print(f"[parity] Svc.slice_inference patched vs serial loop ({len(lens)} slices, pads cropped): L-inf = {err:.3e}; "
              f"serial {len(lens) / t_serial:.1f} slices/s, batched {len(lens) / t_batched:.1f} slices/s (fp32 precision)")
        assert err < 2e-4
        net.set_precision("tc")
        batched.slice_inference(2); serial.slice_inference(2)
        torch.cuda.synchronize(); t0 = time.perf_counter(); serial.slice_inference(2); torch.cuda.synchronize(); ts = time.perf_counter() - t0
        torch.cuda.synchronize(); t0 = time.perf_counter(); batched.slice_inference(2); torch.cuda.synchronize(); tb = time.perf_counter() - t0
        print(f"[perf] Svc.slice_inference, precision tc: serial {len(lens) / ts:.1f} slices/s, batched {len(lens) / tb:.1f} slices/s")
    finally:
        torch.backends.cudnn.conv.fp32_precision = prev
