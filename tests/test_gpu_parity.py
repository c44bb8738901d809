"""GPU tier (B200): the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the same
seeded inputs, against the committed reference fixtures, and — at BASELINE's full size — through
size-independent properties.  Tolerances (waveform in (-1,1), activations O(1..10)):
  * fp32 path  vs oracle/reference fixture:  L-inf <= 1e-4   (T1)
  * tensor-core path (fp16 operands, fp32 accumulate) vs the same: L-inf <= TC_TOL = 5e-3 (frozen; T3 <= T2 of SURVEY §8d:
    the reference's own cuDNN-TF32 vs IEEE-fp32 self-disagreement measured on this path at config 2 is 4.4e-3,
    profiles/r01/ref_pytorch_cuda_baseline.json); measured values are printed.
  * flow alone on the tensor-core path: L-inf <= FLOW_TC_REL * |z|max (z is not bounded like the waveform).
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import svc_oracle as O
from sovits_b200 import synth

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
TC_TOL = 5e-3
FLOW_TC_REL = 1e-3
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = torch.device("cuda:0")


@pytest.fixture(scope="module")
def eng(cfg, sd):
    from sovits_b200.engine import TailEngine
    e = TailEngine(cfg, DEV, "fp32")
    e.load_state_dict(sd)
    yield e
    e.close()


def _case(cfg, sd, B, T, seed=1234, zero_f0=False):
    c, f0, uv, sid = synth.synth_inputs(cfg, B, T, seed)
    if zero_f0:
        f0 = torch.zeros_like(f0)
    noise = synth.draw_noise(B, T, cfg)
    g = sd["emb_g.weight"][sid].transpose(1, 2).contiguous()
    gen = torch.Generator().manual_seed(seed + 1)
    z_p = torch.randn((B, cfg.inter_channels, T), generator=gen) * 1.4
    return z_p, g, f0, noise


@pytest.mark.parametrize("B,T", [(2, 24), (1, 33), (3, 130)])
def test_nsf_source_matches_oracle(cfg, sd, eng, B, T):
    _, _, f0, noise = _case(cfg, sd, B, T)
    f0[:, 3:7] = 0.0
    ref = O.nsf_source_closed_form(sd, f0, noise["rand_ini"], noise["har_noise"], cfg)[:, 0]
    got = eng.nsf_source(f0.to(DEV), noise["rand_ini"].to(DEV), noise["har_noise"].to(DEV)).cpu().double()
    assert float((got - ref).abs().max()) < 5e-6
    lit = O.nsf_source(sd, f0, noise["rand_ini"], noise["har_noise"], cfg, torch.float32)[:, 0]
    assert float((got - lit.double()).abs().max()) < 1e-4      # vs the reference's own fp32 double-cumsum
    got0 = eng.nsf_source(f0.to(DEV), noise["rand_ini"].to(DEV), None).cpu().double()
    ref0 = O.nsf_source_closed_form(sd, f0, noise["rand_ini"], torch.zeros_like(noise["har_noise"]), cfg)[:, 0]
    assert float((got0 - ref0).abs().max()) < 5e-6


def test_nsf_source_ten_seconds_and_unvoiced(cfg, sd, eng):
    B, T = 2, 862
    _, _, f0, noise = _case(cfg, sd, B, T)
    ref = O.nsf_source_closed_form(sd, f0, noise["rand_ini"], noise["har_noise"], cfg)[:, 0]
    got = eng.nsf_source(f0.to(DEV), noise["rand_ini"].to(DEV), noise["har_noise"].to(DEV)).cpu().double()
    assert float((got - ref).abs().max()) < 5e-6
    f00 = torch.zeros_like(f0)
    got = eng.nsf_source(f00.to(DEV), noise["rand_ini"].to(DEV), noise["har_noise"].to(DEV)).cpu().double()
    ref = O.nsf_source_closed_form(sd, f00, noise["rand_ini"], noise["har_noise"], cfg)[:, 0]
    assert torch.isfinite(got).all() and float((got - ref).abs().max()) < 5e-6


@pytest.mark.parametrize("B,T", [(2, 24), (1, 33), (2, 200)])
def test_flow_reverse_matches_oracle(cfg, sd, eng, B, T):
    z_p, g, _, _ = _case(cfg, sd, B, T)
    ref = O.flow_reverse(sd, z_p, torch.ones(B, 1, T), g, cfg, torch.float32)
    got = eng.flow_reverse(z_p.to(DEV), g.to(DEV)).cpu()
    assert float((got - ref).abs().max()) < FP32_TOL


def test_flow_reverse_ragged_lengths_and_time_varying_g(cfg, sd, eng):
    B, T = 3, 40
    z_p, g, _, _ = _case(cfg, sd, B, T)
    lengths = torch.tensor([40, 17, 1])
    mask = (torch.arange(T)[None, :] < lengths[:, None]).float()[:, None, :]
    z_p = z_p * mask
    ref = O.flow_reverse(sd, z_p, mask, g, cfg, torch.float32)
    got = eng.flow_reverse(z_p.to(DEV), g.to(DEV), lengths.to(DEV)).cpu()
    assert float((got - ref).abs().max()) < FP32_TOL
    gt = g.expand(B, cfg.gin_channels, T).clone() + 0.1 * torch.randn(B, cfg.gin_channels, T)   # speaker-mix style g
    ref = O.flow_reverse(sd, z_p, mask, gt, cfg, torch.float32)
    got = eng.flow_reverse(z_p.to(DEV), gt.to(DEV), lengths.to(DEV)).cpu()
    assert float((got - ref).abs().max()) < FP32_TOL


@pytest.mark.parametrize("B,T", [(2, 24), (1, 33), (2, 200)])
def test_flow_reverse_tensor_core(cfg, sd, eng, B, T):
    """The flow on the tcgen05 conv-as-GEMM kernel (gate fused) incl. ragged lengths; z has |max| ~ 15."""
    z_p, g, _, _ = _case(cfg, sd, B, T)
    lengths = torch.tensor([T, max(1, T // 3), 1][:B])
    mask = (torch.arange(T)[None, :] < lengths[:, None]).float()[:, None, :]
    z_p = z_p * mask
    ref = O.flow_reverse(sd, z_p, mask, g, cfg, torch.float32)
    eng.set_precision("tc")
    got = eng.flow_reverse(z_p.to(DEV), g.to(DEV), lengths.to(DEV)).cpu()
    eng.set_precision("fp32")
    err = float((got - ref).abs().max())
    print(f"[parity] flow tc B={B} T={T}: L-inf = {err:.3e} (|z|max {float(ref.abs().max()):.1f})")
    assert err < FLOW_TC_REL * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("B,T", [(2, 24), (1, 33)])
def test_generator_fp32_stagewise(cfg, sd, eng, B, T):
    z_p, g, f0, noise = _case(cfg, sd, B, T)
    taps = {}
    ref = O.tail(sd, cfg, z_p, g, f0, noise, torch.float32, taps)
    eng.set_precision("fp32")
    eng.debug_enable(True)
    got = eng.infer_tail(z_p.to(DEV), g.to(DEV), f0.to(DEV), noise["rand_ini"].to(DEV), noise["har_noise"].to(DEV)).cpu()
    for name in ["z", "conv_pre"] + [f"{k}{i}" for i in range(5) for k in ("ups", "stage")]:
        t = taps[name]
        d = eng.debug_fetch(name, tuple(t.shape)).cpu()
        err = float((d - t).abs().max())
        assert err < FP32_TOL * max(1.0, float(t.abs().max())), (name, err)
    eng.debug_enable(False)
    assert float((got - ref).abs().max()) < FP32_TOL


@pytest.mark.parametrize("name", list(synth.GOLDEN_CASES))
@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_tail_matches_reference_fixture(cfg, sd, eng, name, precision):
    """flow + source + generator on the reference's own z_p, against the reference's own waveform."""
    gold = np.load(os.path.join(GOLD, f"ref_infer_{name}.npz"))
    B, T = synth.GOLDEN_CASES[name]
    c, f0, uv, sid = synth.golden_inputs(cfg, name)
    noise = synth.draw_noise(B, T, cfg, seed=int(gold["seed"]))
    g = sd["emb_g.weight"][sid].transpose(1, 2).contiguous()
    eng.set_precision(precision)
    got = eng.infer_tail(torch.from_numpy(gold["z_p"]).to(DEV), g.to(DEV), f0.to(DEV), noise["rand_ini"].to(DEV),
                         noise["har_noise"].to(DEV)).cpu()
    eng.set_precision("fp32")
    err = float((got - torch.from_numpy(gold["o"])).abs().max())
    print(f"[parity] {name} {precision}: L-inf vs reference waveform = {err:.3e}")
    assert err < (FP32_TOL if precision == "fp32" else TC_TOL)


def test_tc_stagewise_against_oracle(cfg, sd, eng):
    """Every (C, k, dilation) of the tensor-core pair kernel, boundary tiles included (T*hop is not a
    multiple of any tile size)."""
    B, T = 2, 37
    z_p, g, f0, noise = _case(cfg, sd, B, T)
    taps = {}
    ref = O.tail(sd, cfg, z_p, g, f0, noise, torch.float32, taps)
    eng.set_precision("tc")
    eng.debug_enable(True)
    got = eng.infer_tail(z_p.to(DEV), g.to(DEV), f0.to(DEV), noise["rand_ini"].to(DEV), noise["har_noise"].to(DEV)).cpu()
    worst = {}
    for name in ["z", "conv_pre"] + [f"ups{i}" for i in range(5)]:
        t = taps[name]
        d = eng.debug_fetch(name, tuple(t.shape)).cpu()
        rel = float((d - t).abs().max()) / float(t.abs().max())
        print(f"[parity] tc {name}: relative L-inf {rel:.2e}")
        assert rel < 1e-2, (name, rel)
    for i in range(5):
        t = taps[f"stage{i}"]
        d = eng.debug_fetch(f"stage{i}", tuple(t.shape)).cpu()
        worst[i] = float((d - t).abs().max()) / float(t.abs().max())
    eng.debug_enable(False)
    eng.set_precision("fp32")
    print("[parity] tc stage-wise relative L-inf:", {k: f"{v:.2e}" for k, v in worst.items()})
    for i, v in worst.items():
        assert v < 1e-2, (i, v)
    assert float((got - ref).abs().max()) < TC_TOL


@pytest.mark.parametrize("stage", [2, 3, 4])
def test_fused_resblock_matches_pair_chain(cfg, sd, eng, stage):
    """The fused ResBlock kernel (residual stream in TMEM, halo recompute) against three fp32 FFMA pairs, for every
    branch and both tile variants; L is not a multiple of any tile size and alpha/beta accumulate like the generator."""
    C = cfg.stage_channels[stage]
    B, Ln = 2, 5000 + 37 * stage
    g = torch.Generator().manual_seed(7 + stage)
    x = (torch.randn((B, C, Ln), generator=g) * 1.3).to(DEV)
    old = (torch.randn((B, C, Ln), generator=g)).to(DEV)
    for j in range(3):
        ref = x
        for d in range(3):
            ref = eng.debug_pair(stage, j, d, ref, -2)
        want = ref / 3.0 + old
        for variant in (0, 1, 2):                 # 2 = block-skewed hand-off (C <= 32; falls back to 1 for C = 64)
            out = old.clone()
            eng.debug_resblock(stage, j, x, variant, out=out, alpha=1.0 / 3.0, beta=1.0)
            rel = float((out - want).abs().max()) / float(want.abs().max())
            print(f"[parity] fused resblock stage{stage} C={C} k={cfg.resblock_kernel_sizes[j]} v{variant}: rel L-inf {rel:.2e}")
            assert rel < 1e-2, (stage, j, variant, rel)


def test_full_infer_against_oracle_with_replayed_rng(cfg, sd):
    """SynthesizerTrn.infer end to end on the GPU; the oracle gets the very noise tensors torch's CUDA generator
    produced (same seed, same order: SURVEY §9.9)."""
    import json
    import sovits_b200
    from sovits_b200 import models
    with open(sovits_b200.DEFAULT_CONFIG) as f:
        kw = json.load(f)["model"]
    net = models.SynthesizerTrn(1025, 20, **kw).eval()
    net.load_state_dict(sd)
    net = net.to(DEV)
    B, T = 2, 31
    c, f0, uv, sid = synth.synth_inputs(cfg, B, T)
    N = T * cfg.hop
    torch.manual_seed(52468)
    noise = {"z_noise": torch.randn(B, cfg.inter_channels, T, device=DEV).cpu(),
             "rand_ini": torch.rand(B, cfg.n_harmonics, device=DEV).cpu(),
             "har_noise": torch.randn(B, N, cfg.n_harmonics, device=DEV).cpu()}
    ref, _ = O.infer(sd, cfg, c, f0, uv, sid, noise, noice_scale=0.4)
    # the PyTorch prefix (pre/enc_p) follows torch's CUDA default of TF32 convolutions; pin it to
    # IEEE fp32 here so that the strict fp32 comparison measures the library, not cuDNN's TF32.
    old = torch.backends.cudnn.conv.fp32_precision
    torch.backends.cudnn.conv.fp32_precision = "ieee"
    for precision, tol, own in (("fp32", 2e-4, False), ("tc", TC_TOL, False), ("tc", 2 * TC_TOL, True)):
        # own = the prior encoder on the library's kernels too (fp16-operand GEMMs + fused attention): a second reduced-precision
        # stage in front of the tail, bounded like the bench-mode test
        net.set_precision(precision)
        net.own_prefix = own
        o, f0_out = net.infer(c.to(DEV), f0.to(DEV), uv.to(DEV), g=sid.to(DEV), noice_scale=0.4)
        assert o.shape == (B, 1, N) and torch.equal(f0_out.cpu(), f0)
        err = float((o.cpu() - ref).abs().max())
        print(f"[parity] full infer {precision}{' + own prefix' if own else ''}: L-inf = {err:.3e}")
        assert err < tol
        o2, _ = net.infer(c.to(DEV), f0.to(DEV), uv.to(DEV), g=sid.to(DEV), noice_scale=0.4)
        assert torch.equal(o, o2)                       # same seed -> bit-identical (reference behaviour)
    torch.backends.cudnn.conv.fp32_precision = old
    net.own_prefix = True


def test_host_pipeline_is_bit_identical_to_direct_calls(cfg, sd):
    """sovits_b200.pipeline.HostPipeline (bench.py's e2e entry point): overlapped upload / kernels / read-back must return
    exactly what `infer` on device tensors returns, for every submission, with result buffers recycled."""
    import json
    import sovits_b200
    from sovits_b200 import models
    from sovits_b200.pipeline import HostPipeline
    with open(sovits_b200.DEFAULT_CONFIG) as f:
        kw = json.load(f)["model"]
    net = models.SynthesizerTrn(1025, 20, **kw).eval()
    net.load_state_dict(sd)
    net = net.to(DEV)
    batches = []
    for n in range(5):
        c, f0, uv, sid = synth.synth_inputs(cfg, 2, 50 + 7 * (n % 2), seed=40 + n)
        batches.append([t.pin_memory() for t in (c, f0, uv, sid)])
    want = [net.infer(*[t.to(DEV) for t in b[:3]], g=b[3].to(DEV), noice_scale=0.4)[0].cpu() for b in batches]
    got = HostPipeline(net, DEV, depth=2).run(batches, noice_scale=0.4)
    assert len(got) == len(want)
    for w, g_ in zip(want, got):
        assert torch.equal(w, g_)


def test_own_prior_encoder_matches_oracle(cfg, sd):
    """`pre` and `enc_p` on the library's own kernels (svb_pre_conv, svb_enc_p: tcgen05 conv-as-GEMM launches, the
    fused relative-position attention kernel, channel-major LayerNorm) against the oracle's fp32 restatement of
    models.py:155-162 / modules/attentions.py, same noise.  T is not a multiple of the 128-row tiles; the band terms, the
    sequence end inside a key tile and both heads are exercised."""
    import torch.nn.functional as F
    from sovits_b200.engine import TailEngine
    e = TailEngine(cfg, DEV, "tc")
    e.load_state_dict(sd)
    assert e.has_prefix
    for B, T in ((2, 150), (1, 300)):
        c, f0, uv, sid = synth.synth_inputs(cfg, B, T)
        gen = torch.Generator().manual_seed(5)
        z_noise = torch.randn((B, cfg.inter_channels, T), generator=gen)
        x_ref, x_mask, _ = O.prologue(sd, c, f0, uv, sid, cfg, torch.float32)
        pre_ref = F.conv1d(c, sd["pre.weight"], sd["pre.bias"], padding=2)
        pre_got = e.pre_conv(c.to(DEV)).cpu()
        rel = float((pre_got - pre_ref).abs().max()) / float(pre_ref.abs().max())
        print(f"[parity] own pre conv B={B} T={T}: relative L-inf {rel:.2e}")
        assert rel < 5e-3
        z_ref, m_ref, logs_ref = O.text_encoder(sd, x_ref, x_mask, O.f0_to_coarse(f0), z_noise, 0.4, cfg, torch.float32)
        x_in = x_ref + sd["enc_p.f0_emb.weight"][O.f0_to_coarse(f0)].transpose(1, 2)
        z, m, lg = e.enc_p(x_in.to(DEV), z_noise.to(DEV), 0.4, want_stats=True)
        errs = {n: float((a.cpu() - r).abs().max()) / max(1.0, float(r.abs().max())) for n, a, r in (("z_p", z, z_ref), ("m", m, m_ref), ("logs", lg, logs_ref))}
        print(f"[parity] own enc_p B={B} T={T}: relative L-inf {errs} (|z_p|max {float(z_ref.abs().max()):.1f})")
        assert max(errs.values()) < 1e-2, errs
    e.close()


def test_full_size_properties(cfg, sd, eng):
    """BASELINE config 2 size (8 x 862 frames): batch independence (item i of a batch == item i alone, given its
    own noise slice), run-to-run bit reproducibility, finiteness, tc-vs-fp32 agreement, host-buffer entry point."""
    B, T = 8, 862
    z_p, g, f0, noise = _case(cfg, sd, B, T)
    dv = lambda t: t.to(DEV)
    outs = {}
    for precision in ("fp32", "tc"):
        eng.set_precision(precision)
        full = eng.infer_tail(dv(z_p), dv(g), dv(f0), dv(noise["rand_ini"]), dv(noise["har_noise"]))
        assert torch.isfinite(full).all() and float(full.abs().max()) <= 1.0
        i = 5
        one = eng.infer_tail(dv(z_p[i:i + 1]), dv(g[i:i + 1]), dv(f0[i:i + 1]), dv(noise["rand_ini"][i:i + 1]),
                             dv(noise["har_noise"][i:i + 1]))
        assert torch.equal(one[0], full[i]), precision
        # run-to-run reproducibility: the branch-merged launches reduce into the stage output in an order-independent way
        again = eng.infer_tail(dv(z_p), dv(g), dv(f0), dv(noise["rand_ini"]), dv(noise["har_noise"]))
        assert torch.equal(again, full), precision
        outs[precision] = full
    eng.set_precision("fp32")
    err = float((outs["tc"] - outs["fp32"]).abs().max())
    print(f"[parity] full size tc vs fp32 L-inf = {err:.3e}")
    assert err < TC_TOL
    i = 2
    host = eng.infer_tail_host(z_p[i:i + 1], g[i:i + 1], f0[i:i + 1], noise["rand_ini"][i:i + 1], noise["har_noise"][i:i + 1])
    assert torch.equal(host[0], outs["fp32"][i].cpu())


def test_schedule_options_are_equivalent(cfg, sd, eng):
    """TMA-fed pair kernels and the pair-vs-fused ResBlock schedules compute the same function."""
    B, T = 2, 150
    z_p, g, f0, noise = _case(cfg, sd, B, T)
    args = [t.to(DEV) for t in (z_p, g, f0, noise["rand_ini"], noise["har_noise"])]
    eng.set_precision("tc")

    def run():
        n0 = eng.launch_count
        out = eng.infer_tail(*args)
        return out, eng.launch_count - n0

    base, n_base = run()                          # defaults: thread-staged 128-bit loader, fused ResBlocks for C <= 64
    eng.set_option("tma", 1)                      # pairs 2 and 3 of a ResBlock load their operand tile by TMA
    tma, n_tma = run()
    eng.set_option("fuse_resblock", 0)
    pairs_tma, n_pairs_tma = run()
    eng.set_option("tma", 0)
    pairs, n_pairs = run()
    eng.set_option("fuse_resblock", 1)
    eng.set_option("fuse_maxc", 32)
    fused32, n_f32 = run()
    eng.set_option("fuse_maxc", 64)
    eng.set_option("fuse_flow", 0)                # flow as 10 conv-as-GEMM launches per coupling layer instead of one kernel
    unfused_flow, n_uf = run()
    eng.set_option("fuse_flow", 1)
    eng.set_option("merge_branches", 0)           # the wide stages' three branches as separate pair launches: 9 per stage instead
    unmerged, n_um = run()                        # of 4 (two merged launches, then two branches merged + the third alone so that
    eng.set_option("merge_branches", 1)           # the reduction into the stage output stays order-independent)
    assert n_um == n_base + 2 * 5, (n_um, n_base)
    # the options really select different schedules: 9 pair launches replace each fused ResBlock launch of 3
    assert n_pairs == n_pairs_tma and n_pairs > n_f32 > n_base, (n_base, n_tma, n_pairs, n_f32)
    assert n_uf == n_base + 4 * 11 - 5, (n_uf, n_base)   # 4 x 11 launches -> one conditioning GEMV + 4 coupling-layer kernels
    eng.set_precision("fp32")
    ref = eng.infer_tail(*args)
    for name, o in (("default", base), ("tma", tma), ("pairs-only", pairs), ("pairs-only+tma", pairs_tma), ("fused<=32", fused32),
                    ("unfused flow", unfused_flow), ("unmerged branches", unmerged)):
        err = float((o - ref).abs().max())
        print(f"[parity] schedule {name}: L-inf vs fp32 path = {err:.3e}")
        assert err < TC_TOL


def test_snake_variant_matches_reference_fixture():
    """BASELINE config 4 (vdecoder/hifiganwithsnake) vs the reference's own waveform: fp32 = SnakeAlias kernel + FFMA
    convolutions; tc = every convolution on tcgen05 with the SnakeAlias activation computed by its loader (no FFMA launch)."""
    from sovits_b200.config import load_config
    from sovits_b200.engine import TailEngine
    cfg_s = load_config()
    cfg_s.vocoder_name = "nsf-snake-hifigan"
    sd_s = synth.synth_state_dict(cfg_s)
    e = TailEngine(cfg_s, DEV, "fp32")
    e.load_state_dict(sd_s)
    for name, (B, T) in synth.SNAKE_GOLDEN_CASES.items():
        gold = np.load(os.path.join(GOLD, f"ref_infer_{name}.npz"))
        c, f0, uv, sid = synth.golden_inputs(cfg_s, name)
        noise = synth.draw_noise(B, T, cfg_s)
        g = sd_s["emb_g.weight"][sid].transpose(1, 2).contiguous()
        outs = []
        for precision in ("fp32", "tc"):
            e.set_precision(precision)
            fb0 = e.fallback_count
            got = e.infer_tail(torch.from_numpy(gold["z_p"]).to(DEV), g.to(DEV), f0.to(DEV), noise["rand_ini"].to(DEV),
                               noise["har_noise"].to(DEV)).cpu()
            err = float((got - torch.from_numpy(gold["o"])).abs().max())
            print(f"[parity] snake {name} {precision}: L-inf vs reference waveform = {err:.3e}")
            assert err < (FP32_TOL if precision == "fp32" else TC_TOL)
            assert e.fallback_count == fb0
            outs.append(got)
        assert not torch.equal(outs[0], outs[1])          # the two precisions really are different code paths
    # a longer ragged-tile case against the oracle (tile edges, all five stages, both ends of the sequence)
    import svc_oracle as OO
    B, T = 2, 45
    c, f0, uv, sid = synth.synth_inputs(cfg_s, B, T)
    noise = synth.draw_noise(B, T, cfg_s)
    g = sd_s["emb_g.weight"][sid].transpose(1, 2).contiguous()
    gen = torch.Generator().manual_seed(77)
    z_p = torch.randn((B, cfg_s.inter_channels, T), generator=gen) * 1.4
    taps = {}
    ref = OO.tail(sd_s, cfg_s, z_p, g, f0, noise, torch.float32, taps)
    e.set_precision("tc")
    e.debug_enable(True)
    got = e.infer_tail(z_p.to(DEV), g.to(DEV), f0.to(DEV), noise["rand_ini"].to(DEV), noise["har_noise"].to(DEV)).cpu()
    for name in [f"{k}{i}" for i in range(5) for k in ("ups", "stage")]:
        t = taps[name]
        d = e.debug_fetch(name, tuple(t.shape)).cpu()
        rel = float((d - t).abs().max()) / float(t.abs().max())
        print(f"[parity] snake tc {name}: relative L-inf {rel:.2e}")
        assert rel < 1e-2, (name, rel)
    err = float((got - ref).abs().max())
    print(f"[parity] snake tc B={B} T={T} vs oracle: L-inf = {err:.3e}")
    assert err < TC_TOL
    e.close()


def test_mel_vocoder_matches_reference_fixture():
    """vdecoder/nsf_hifigan Generator(mel, f0) through svb_vocoder vs the reference's own waveform, and the
    drop-in module end to end (replaying its RNG draws for the oracle)."""
    from sovits_b200 import nsf_hifigan
    from sovits_b200.engine import TailEngine
    vcfg = nsf_hifigan.cfg_from_h(synth.VOCODER_H)
    vsd = synth.synth_vocoder_state_dict(vcfg)
    gold = np.load(os.path.join(GOLD, "ref_vocoder_b2_t21.npz"))
    B, T = int(gold["B"]), int(gold["T"])
    mel, f0 = synth.synth_vocoder_inputs(vcfg, B, T)
    torch.manual_seed(int(gold["seed"]))
    ri, hn = torch.rand(B, 9), torch.randn(B, T * vcfg.hop, 9)
    e = TailEngine(vcfg, DEV, "fp32")
    e.load_state_dict(vsd)
    for precision, tol in (("fp32", FP32_TOL), ("tc", TC_TOL)):
        e.set_precision(precision)
        got = e.vocoder(mel.to(DEV), f0.to(DEV), ri.to(DEV), hn.to(DEV)).cpu()
        err = float((got - torch.from_numpy(gold["o"])).abs().max())
        print(f"[parity] mel vocoder {precision}: L-inf vs reference waveform = {err:.3e}")
        assert err < tol
    e.close()
    gen = nsf_hifigan.Generator(synth.VOCODER_H)
    gen.load_state_dict(vsd)
    gen = gen.to(DEV).eval()
    torch.manual_seed(99)
    o = gen(mel.to(DEV), f0.to(DEV))
    torch.manual_seed(99)
    ri2 = torch.rand(B, 9, device=DEV).cpu()
    hn2 = torch.randn(B, T * vcfg.hop, 9, device=DEV).cpu()
    ref = O.vocoder(vsd, vcfg, mel, f0, ri2, hn2)
    err = float((o.cpu() - ref).abs().max())
    print(f"[parity] mel vocoder module (tc): L-inf vs oracle = {err:.3e}")
    assert o.shape == (B, 1, T * 512) and err < TC_TOL


def test_error_paths(cfg, sd):
    from sovits_b200 import lib as L
    lib = L.load_library()
    ctx = C.c_void_p()
    assert lib.svb_create(0, C.byref(ctx)) == 0
    buf = torch.zeros(16, device=DEV)
    rc = lib.svb_nsf_source(ctx, buf.data_ptr(), buf.data_ptr(), None, buf.data_ptr(), 1, 1, None)
    assert rc == -3 and b"svb_load_weights" in lib.svb_last_error(ctx)          # SVB_ERR_NOT_LOADED
    assert lib.svb_create(99, C.byref(C.c_void_p())) != 0
    lib.svb_destroy(ctx)
    from sovits_b200.engine import TailEngine
    e = TailEngine(cfg, DEV, "fp32")
    bad = {k: v for k, v in sd.items() if k != "dec.ups.2.weight_g"}
    with pytest.raises(L.SvbError, match="dec.ups.2.weight_g"):
        e.load_state_dict(bad)
    e.load_state_dict(sd)
    z = torch.zeros(1, cfg.inter_channels, 8, device=DEV)
    g_bad = torch.zeros(1, cfg.gin_channels, 3, device=DEV)
    with pytest.raises(L.SvbError):
        e.flow_reverse(z, g_bad)
    e.close()


def test_prefix_fused_tails_match_torch(monkeypatch):
    """the fused element-wise tails of an enc_p layer (csrc/kernels_prefix.cu) against the torch ops they
    replace, and the whole RelEncoder with and without them (odd lengths, boundaries of the k-tap shifts)."""
    import torch.nn.functional as F
    from sovits_b200 import frontend as fe
    from sovits_b200.lib import load_library
    lib = load_library()
    g = torch.Generator().manual_seed(11)
    B, L, C, k = 3, 37, 192, 3
    x = torch.randn((B, L, C), generator=g).to(DEV)
    r = torch.randn((B, L, C), generator=g).to(DEV)
    norm = fe.ChannelNormT(C).to(DEV)
    with torch.no_grad():
        norm.gamma.copy_(torch.randn(C, generator=g).to(DEV)); norm.beta.copy_(torch.randn(C, generator=g).to(DEV))
        y, cols = fe._add_ln_im2col(lib, x, r, norm, k)
        want = norm(x + r)
        wp = F.pad(want, (0, 0, (k - 1) // 2, k // 2))
        wcols = torch.cat([wp[:, t:t + L] for t in range(k)], dim=-1)
        assert float((y - want).abs().max()) < 2e-5
        assert float((cols - wcols).abs().max()) < 2e-5
        ya = torch.randn((B, L, k * C), generator=g).to(DEV)
        bias = torch.randn(C, generator=g).to(DEV)
        got = fe._ffn_tail(lib, ya, x, bias, norm, k)
        yap = F.pad(ya, (0, 0, (k - 1) // 2, k // 2))
        acc = sum(yap[:, t:t + L, t * C:(t + 1) * C] for t in range(k))
        want2 = norm(x + acc + bias)
        assert float((got - want2).abs().max()) < 5e-5
        enc = fe.RelEncoder(192, 768, 2, 3, 3).to(DEV).eval()
        xin = torch.randn((2, 192, 75), generator=g).to(DEV)
        mask = torch.ones(2, 1, 75, device=DEV)
        prev = torch.backends.cudnn.conv.fp32_precision
        torch.backends.cudnn.conv.fp32_precision = "ieee"
        try:
            fused = enc(xin, mask, True)
            monkeypatch.setenv("SVB_PREFIX_FUSED", "0")
            plain = enc(xin, mask, True)
        finally:
            torch.backends.cudnn.conv.fp32_precision = prev
        err = float((fused - plain).abs().max())
        print(f"[parity] enc_p fused tails vs torch ops: L-inf {err:.2e}")
        assert err < 1e-4


def test_fp16_checkpoint_equals_upcast_weights(cfg, sd):
    """`compress_model.py --half` checkpoints (compress_model.py:37-41) store fp16 tensors.  The library
    converts them to fp32 before folding weight norm, so the result must be bit-identical to loading the same values
    upcast on the host."""
    from sovits_b200.engine import TailEngine
    sd16 = {k: (v.half() if v.is_floating_point() else v) for k, v in sd.items()}
    sd32 = {k: (v.float() if v.is_floating_point() else v) for k, v in sd16.items()}
    z_p, g, f0, noise = _case(cfg, sd, 2, 40)
    args = [t.to(DEV) for t in (z_p, sd32["emb_g.weight"][torch.tensor([[0], [1]])].transpose(1, 2).contiguous(), f0,
                                noise["rand_ini"], noise["har_noise"])]
    outs = []
    for weights in (sd16, sd32):
        e = TailEngine(cfg, DEV, "tc")
        e.load_state_dict(weights)
        outs.append(e.infer_tail(*args).cpu())
        e.close()
    assert torch.equal(outs[0], outs[1])
    assert torch.isfinite(outs[0]).all() and float(outs[0].abs().max()) > 1e-3


def test_flow_only_config5(cfg, sd, eng):
    """SURVEY §8(d) config 5 (flow-only microbench): z_p[1,192,100000] ~ N(0,1), g[1,768,1] ~ N(0,1), mask = ones.  The
    oracle runs on the GPU through the reference's own ops (strict fp32 convolutions) for this size."""
    gen = torch.Generator().manual_seed(1234)
    z_p = torch.randn((1, cfg.inter_channels, 100_000), generator=gen)
    g = torch.randn((1, cfg.gin_channels, 1), generator=gen)
    prev = torch.backends.cudnn.conv.fp32_precision
    torch.backends.cudnn.conv.fp32_precision = "ieee"
    try:
        sd_dev = {k: v.to(DEV) for k, v in sd.items() if k.startswith("flow.")}
        ref = O.flow_reverse(sd_dev, z_p.to(DEV), torch.ones(1, 1, 100_000, device=DEV), g.to(DEV), cfg, torch.float32).cpu()
    finally:
        torch.backends.cudnn.conv.fp32_precision = prev
    eng.set_precision("fp32")
    got = eng.flow_reverse(z_p.to(DEV), g.to(DEV)).cpu()
    e32 = float((got - ref).abs().max())
    eng.set_precision("tc")
    got_tc = eng.flow_reverse(z_p.to(DEV), g.to(DEV)).cpu()
    eng.set_precision("fp32")
    etc = float((got_tc - ref).abs().max())
    print(f"[parity] config5 flow-only T=100000: fp32 L-inf {e32:.3e}, tc L-inf {etc:.3e} (|z|max {float(ref.abs().max()):.1f})")
    assert e32 < 5e-4 and etc < FLOW_TC_REL * float(ref.abs().max())


def test_full_size_config2_against_oracle(cfg, sd, eng):
    """BASELINE config 2 at its own size (8 x 862 frames) against the ORACLE: the reference ops run on the GPU with strict
    fp32 convolutions (cudnn.conv.fp32_precision='ieee'), the excitation from the fp64 closed form; compared with the CUDA
    path in both precisions, waveform and every stage tap at full length (full-length tiles, the 1.5-wave stage-0 grid,
    first-wave de-phasing)."""
    B, T = 8, 862
    z_p, g, f0, noise = _case(cfg, sd, B, T)
    har = O.nsf_source_closed_form(sd, f0, noise["rand_ini"], noise["har_noise"], cfg).float()
    sd_dev = {k: v.to(DEV) for k, v in sd.items() if k.startswith(("flow.", "dec."))}
    prev = torch.backends.cudnn.conv.fp32_precision
    torch.backends.cudnn.conv.fp32_precision = "ieee"
    taps = {}
    try:
        z_ref = O.flow_reverse(sd_dev, z_p.to(DEV), torch.ones(B, 1, T, device=DEV), g.to(DEV), cfg, torch.float32)
        o_ref = O.generator(sd_dev, z_ref, g.to(DEV), har.to(DEV), cfg, torch.float32, taps)
    finally:
        torch.backends.cudnn.conv.fp32_precision = prev
    taps["z"] = z_ref
    dv = lambda t: t.to(DEV)
    for precision, tol, rel_tol in (("fp32", FP32_TOL, 1e-4), ("tc", TC_TOL, 1e-2)):
        eng.set_precision(precision)
        eng.debug_enable(True)
        got = eng.infer_tail(dv(z_p), dv(g), dv(f0), dv(noise["rand_ini"]), dv(noise["har_noise"]))
        worst = {}
        for name in ["z", "conv_pre"] + [f"{k}{i}" for i in range(5) for k in ("ups", "stage")]:
            t = taps[name]
            d = eng.debug_fetch(name, tuple(t.shape))
            worst[name] = float((d - t).abs().max()) / max(1.0, float(t.abs().max()))
            assert worst[name] < rel_tol, (precision, name, worst[name])
        eng.debug_enable(False)
        err = float((got - o_ref).abs().max())
        print(f"[parity] config2 full size {precision}: waveform L-inf vs oracle = {err:.3e}; worst tap {max(worst, key=worst.get)} "
              f"rel {max(worst.values()):.2e}")
        assert err < tol, (precision, err)
    eng.set_precision("fp32")


def test_bench_mode_end_to_end_parity(cfg, sd):
    """The configuration bench.py times: torch defaults for the PyTorch prefix (cuDNN-style TF32 for the conv-equivalent
    GEMMs, IEEE fp32 for q@k^T / p@v exactly like the reference's CUDA path) + the tensor-core tail, end to end
    through SynthesizerTrn.infer against the fp32 CPU oracle.  Two independent reduced-precision sources (prefix TF32, tail
    fp16 operands), each within the reference's own self-disagreement band: bound 2 x TC_TOL."""
    import json
    import sovits_b200
    from sovits_b200 import models
    with open(sovits_b200.DEFAULT_CONFIG) as f:
        kw = json.load(f)["model"]
    net = models.SynthesizerTrn(1025, 20, **kw).eval()
    net.load_state_dict(sd)
    net = net.to(DEV)
    net.set_precision("tc")
    B, T = 2, 120
    c, f0, uv, sid = synth.synth_inputs(cfg, B, T)
    N = T * cfg.hop
    torch.manual_seed(52468)
    noise = {"z_noise": torch.randn(B, cfg.inter_channels, T, device=DEV).cpu(),
             "rand_ini": torch.rand(B, cfg.n_harmonics, device=DEV).cpu(),
             "har_noise": torch.randn(B, N, cfg.n_harmonics, device=DEV).cpu()}
    ref, _ = O.infer(sd, cfg, c, f0, uv, sid, noise, noice_scale=0.4)
    assert torch.backends.cudnn.conv.fp32_precision != "ieee"      # torch's CUDA default, as in bench.py
    o, _ = net.infer(c.to(DEV), f0.to(DEV), uv.to(DEV), g=sid.to(DEV), noice_scale=0.4)
    err = float((o.cpu() - ref).abs().max())
    print(f"[parity] bench mode (TF32-like prefix + tc tail) B={B} T={T}: L-inf vs fp32 oracle = {err:.3e}")
    assert err < 2 * TC_TOL
    assert net._b200_engine.fallback_count == 0


def test_fp16_range_of_the_tensor_core_path(cfg, sd):
    """fp16 operands have a 5-bit exponent (TF32, the reference's CUDA arithmetic, has 8).  LeakyReLU is positively
    homogeneous, so scaling every ResBlock's first convolution (weight_g and bias) by s and its second convolution's weight_g
    by 1/s leaves the network function unchanged - but moves weights and mid activations by s = 1e-4 / 1e+3 in magnitude.
    The tensor-core path must stay inside TC_TOL of the unscaled reference fixture (power-of-two range normalisation of the
    weight images, docs in DESIGN.md)."""
    from sovits_b200.engine import TailEngine
    gold_name = "b2_t24"
    gold = np.load(os.path.join(GOLD, f"ref_infer_{gold_name}.npz"))
    B, T = synth.GOLDEN_CASES[gold_name]
    c, f0, uv, sid = synth.golden_inputs(cfg, gold_name)
    noise = synth.draw_noise(B, T, cfg, seed=int(gold["seed"]))
    g = sd["emb_g.weight"][sid].transpose(1, 2).contiguous()
    for s in (1e-4, 1e3):
        sd2 = dict(sd)
        for k, v in sd.items():
            if ".convs1." in k and (k.endswith("weight_g") or k.endswith("bias")):
                sd2[k] = v * s
            elif ".convs2." in k and k.endswith("weight_g"):
                sd2[k] = v / s
        e = TailEngine(cfg, DEV, "tc")
        e.load_state_dict(sd2)
        got = e.infer_tail(torch.from_numpy(gold["z_p"]).to(DEV), g.to(DEV), f0.to(DEV), noise["rand_ini"].to(DEV),
                           noise["har_noise"].to(DEV)).cpu()
        e.close()
        err = float((got - torch.from_numpy(gold["o"])).abs().max())
        print(f"[parity] range test, ResBlock conv1 x{s:g} / conv2 x{1 / s:g} (tc): L-inf vs reference waveform = {err:.3e}")
        assert torch.isfinite(got).all() and err < TC_TOL


def test_nsf_source_philox_mode(cfg, sd, eng):
    """SURVEY §2a "throughput mode": harmonic noise drawn in-kernel (Philox4x32-10 + Box-Muller) instead of a [B,N,9] torch
    tensor.  Deterministic per seed, different across seeds, and on unvoiced frames (where the excitation is pure noise,
    vdecoder/hifigan/models.py:262-270) distributed like the torch-noise run: same mean / std, matching quantiles, no
    sample-to-sample correlation.  With the option off, noise = NULL stays the noiseless path (bit-exact tests above)."""
    B, T = 2, 600
    _, _, f0, noise = _case(cfg, sd, B, T)
    f0[:, 100:400] = 0.0
    args = (f0.to(DEV), noise["rand_ini"].to(DEV))
    ref = eng.nsf_source(*args, noise["har_noise"].to(DEV))
    eng.set_option("philox_seed", 1234)
    eng.set_option("philox_noise", 1)
    try:
        a = eng.nsf_source(*args, None)
        b = eng.nsf_source(*args, None)
        eng.set_option("philox_seed", 99)
        c = eng.nsf_source(*args, None)
    finally:
        eng.set_option("philox_noise", 0)
    quiet = eng.nsf_source(*args, None)
    assert torch.equal(a, b) and not torch.equal(a, c) and not torch.equal(a, quiet)
    hop = cfg.hop
    seg = slice(100 * hop, 400 * hop)
    xa, xr = a[:, seg].flatten().double(), ref[:, seg].flatten().double()
    assert abs(float(xa.mean() - xr.mean())) < 3e-4 and abs(float(xa.std() / xr.std()) - 1.0) < 0.02
    qs = torch.linspace(0.001, 0.999, 999, dtype=torch.float64, device=DEV)
    ks = float((torch.quantile(xa[:200000], qs) - torch.quantile(xr[:200000], qs)).abs().max() / xr.std())
    xc = xa - xa.mean()
    rho = float((xc[1:] * xc[:-1]).mean() / xc.var())
    print(f"[parity] philox source noise: std ratio {float(xa.std() / xr.std()):.4f}, max quantile gap {ks:.4f} sigma, lag-1 autocorrelation {rho:.4f}")
    assert ks < 0.05 and abs(rho) < 0.01
    voiced = slice(0, 100 * hop)
    assert abs(float(a[:, voiced].double().std() / ref[:, voiced].double().std()) - 1.0) < 0.02
