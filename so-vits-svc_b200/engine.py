"""TailEngine: torch-tensor wrapper over the C ABI for flow(reverse) -> NSF source -> generator.

PyTorch is plumbing here (device memory, streams); all arithmetic of the tail happens in
``libsovits_b200.so``.  Mirrors ``ResidualCouplingBlock.forward(reverse=True)`` (models.py:45-52) and
``Generator.forward`` (vdecoder/hifigan/models.py:366-394) of the reference.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import lib as L
from .config import ModelCfg

TAIL_PREFIXES = ("flow.", "dec.", "pre.", "enc_p.")      # pre./enc_p. feed the library's own prior encoder (svb_enc_p)


def _cfg_struct(cfg: ModelCfg) -> L.svb_model_cfg:
    cfg.check_cuda_tail_supported()
    s = L.svb_model_cfg()
    s.inter_channels = cfg.inter_channels
    s.hidden_channels = cfg.hidden_channels
    s.gin_channels = cfg.gin_channels
    s.n_flows = 4
    s.flow_wn_layers = cfg.flow_wn_layers
    s.flow_kernel_size = cfg.flow_kernel_size
    s.upsample_initial_channel = cfg.upsample_initial_channel
    s.n_upsamples = len(cfg.upsample_rates)
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        s.upsample_rates[i] = u
        s.upsample_kernel_sizes[i] = k
    s.n_resblock_kernels = len(cfg.resblock_kernel_sizes)
    for j, k in enumerate(cfg.resblock_kernel_sizes):
        s.resblock_kernel_sizes[j] = k
        for d, dil in enumerate(cfg.resblock_dilation_sizes[j]):
            s.resblock_dilations[j][d] = dil
    s.sampling_rate = cfg.sampling_rate
    s.n_harmonics = cfg.n_harmonics
    s.snake = 1 if cfg.snake else 0
    s.num_mels = cfg.num_mels
    if cfg.num_mels == 0:
        s.ssl_dim = cfg.ssl_dim
        s.enc_layers = cfg.n_layers
        s.enc_heads = cfg.n_heads
        s.enc_filter = cfg.filter_channels
        s.enc_kernel = cfg.kernel_size
        s.enc_window = cfg.enc_window
    return s


class TailEngine:
    """One context per device (``svb_create``).  Not re-entrant, like the reference's ``Svc``."""

    def __init__(self, cfg: ModelCfg, device: torch.device, precision: str = "tc"):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("sovits_b200 has no CPU path: the tail only runs on a B200 (sm_100a) device")
        self.lib = L.load_library()
        self.cfg = cfg
        self.device = device
        idx = device.index if device.index is not None else torch.cuda.current_device()
        self._ctx = C.c_void_p()
        L.check(self.lib, None, self.lib.svb_create(idx, C.byref(self._ctx)), "svb_create")
        self._cfg_struct = _cfg_struct(cfg)
        self.loaded = False
        self.set_precision(precision)

    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.svb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """Hand the reference-layout tensors (flow.*, dec.*) to ``svb_load_weights``."""
        keep = []
        arr_t = []
        for k, v in sd.items():
            if self.cfg.num_mels == 0 and not k.startswith(TAIL_PREFIXES):
                continue
            t = v.detach()
            if t.dtype not in (torch.float32, torch.float16):
                t = t.float()
            t = t.to("cpu").contiguous()
            keep.append((k.encode(), t))
        arr = (L.svb_tensor * len(keep))()
        for i, (name, t) in enumerate(keep):
            arr[i].name = name
            arr[i].data = t.data_ptr()
            arr[i].dtype = 0 if t.dtype == torch.float32 else 1
            arr[i].ndim = t.dim()
            for d in range(t.dim()):
                arr[i].shape[d] = t.shape[d]
            arr_t.append(t)
        with torch.cuda.device(self.device):
            rc = self.lib.svb_load_weights(self._ctx, arr, len(keep), C.byref(self._cfg_struct))
        L.check(self.lib, self._ctx, rc, "svb_load_weights")
        self.loaded = True
        self.has_prefix = any(k.startswith(b"enc_p.proj") for k, _ in keep) and any(k.startswith(b"pre.weight") for k, _ in keep)

    def set_precision(self, precision: str) -> None:
        code = {"fp32": L.PREC_FP32, "tc": L.PREC_TC}[precision]
        L.check(self.lib, self._ctx, self.lib.svb_set_precision(self._ctx, code), "svb_set_precision")
        self.precision = precision

    def set_option(self, name: str, value: int) -> None:
        L.check(self.lib, self._ctx, self.lib.svb_set_option(self._ctx, name.encode(), int(value)), "svb_set_option")

    def debug_enable(self, on: bool = True) -> None:
        L.check(self.lib, self._ctx, self.lib.svb_debug_enable(self._ctx, int(on)), "svb_debug_enable")

    def debug_fetch(self, name: str, shape) -> torch.Tensor:
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        rc = self.lib.svb_debug_fetch(self._ctx, name.encode(), out.data_ptr(), out.numel(), self._stream())
        L.check(self.lib, self._ctx, rc, "svb_debug_fetch")
        return out

    def debug_pair(self, stage: int, j: int, d: int, x: torch.Tensor, variant: int, out: Optional[torch.Tensor] = None,
                   alpha: float = 1.0, beta: float = 0.0) -> torch.Tensor:
        """One ResBlock pair on x [B,C,L]; variant >= 0: tensor-core tile variant, -2: fp32 FFMA convs."""
        x = self._f32(x, "x")
        B, Cc, Ln = x.shape
        if out is None:
            out = torch.empty_like(x)
        scratch = torch.empty_like(x) if variant < 0 else None
        rc = self.lib.svb_debug_pair(self._ctx, stage, j, d, x.data_ptr(), out.data_ptr(),
                                     scratch.data_ptr() if scratch is not None else None, B, Ln, variant,
                                     alpha, beta, self._stream())
        L.check(self.lib, self._ctx, rc, "svb_debug_pair")
        return out

    def debug_resblock(self, stage: int, j: int, x: torch.Tensor, variant: int, out: Optional[torch.Tensor] = None,
                       alpha: float = 1.0, beta: float = 0.0) -> torch.Tensor:
        """One whole ResBlock branch (3 pairs) on x [B,C,L] through the fused tensor-core kernel."""
        x = self._f32(x, "x")
        B, Cc, Ln = x.shape
        if out is None:
            out = torch.empty_like(x)
        rc = self.lib.svb_debug_resblock(self._ctx, stage, j, x.data_ptr(), out.data_ptr(), B, Ln, variant, alpha, beta,
                                         self._stream())
        L.check(self.lib, self._ctx, rc, "svb_debug_resblock")
        return out

    def profile_enable(self, on: bool = True) -> None:
        L.check(self.lib, self._ctx, self.lib.svb_profile_enable(self._ctx, int(on)), "svb_profile_enable")

    def profile_read(self, name: str):
        """-> dict(ms, count, flops, bytes) summed over the launches since profile_enable(True), or None."""
        ms, cnt, fl, by = C.c_double(), C.c_int64(), C.c_double(), C.c_double()
        rc = self.lib.svb_profile_read(self._ctx, name.encode(), C.byref(ms), C.byref(cnt), C.byref(fl), C.byref(by))
        if rc != 0:
            return None
        return {"ms": ms.value, "count": cnt.value, "flops": fl.value, "bytes": by.value}

    @property
    def launch_count(self) -> int:
        return int(self.lib.svb_launch_count(self._ctx))

    @property
    def fallback_count(self) -> int:
        """How often a precision="tc" call ran FFMA kernels instead (see svb_fallback_count)."""
        return int(self.lib.svb_fallback_count(self._ctx))

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _f32(self, t: torch.Tensor, name: str) -> torch.Tensor:
        if t.device != self.device:
            if t.device.type != "cuda":
                raise RuntimeError(f"{name} must live on {self.device} (got {t.device}); there is no CPU path")
            t = t.to(self.device)
        return t.to(torch.float32).contiguous()

    def _lengths(self, lengths: Optional[torch.Tensor]):
        if lengths is None:
            return None, None
        l32 = lengths.to(self.device, torch.int32).contiguous()
        return l32, l32.data_ptr()

    # ------------------------------------------------------------------ the three kernels + fused tail
    @torch.no_grad()
    def flow_reverse(self, z_p: torch.Tensor, g: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        z_p = self._f32(z_p, "z_p"); g = self._f32(g, "g")
        B, Cc, T = z_p.shape
        out = torch.empty_like(z_p)
        keep, lp = self._lengths(lengths)
        rc = self.lib.svb_flow_reverse(self._ctx, z_p.data_ptr(), g.data_ptr(), g.shape[2], lp, out.data_ptr(), B, T,
                                       None, 0, self._stream())
        L.check(self.lib, self._ctx, rc, "svb_flow_reverse")
        return out

    @torch.no_grad()
    def nsf_source(self, f0: torch.Tensor, rand_ini: torch.Tensor, har_noise: Optional[torch.Tensor]) -> torch.Tensor:
        f0 = self._f32(f0, "f0"); rand_ini = self._f32(rand_ini, "rand_ini")
        B, T = f0.shape
        N = T * self.cfg.hop
        nz = None
        if har_noise is not None:
            har_noise = self._f32(har_noise, "har_noise")
            assert har_noise.shape == (B, N, self.cfg.n_harmonics), har_noise.shape
            nz = har_noise.data_ptr()
        har = torch.empty((B, N), dtype=torch.float32, device=self.device)
        rc = self.lib.svb_nsf_source(self._ctx, f0.data_ptr(), rand_ini.data_ptr(), nz, har.data_ptr(), B, T, self._stream())
        L.check(self.lib, self._ctx, rc, "svb_nsf_source")
        return har

    @torch.no_grad()
    def pre_conv(self, c: torch.Tensor) -> torch.Tensor:
        """``self.pre(c)`` (models.py:400,518) on the library's tcgen05 conv kernel: [B,ssl,T] -> [B,hidden,T]."""
        c = self._f32(c, "c")
        B, _, T = c.shape
        x = torch.empty((B, self.cfg.hidden_channels, T), dtype=torch.float32, device=self.device)
        rc = self.lib.svb_pre_conv(self._ctx, c.data_ptr(), x.data_ptr(), B, T, self._stream())
        L.check(self.lib, self._ctx, rc, "svb_pre_conv")
        return x

    @torch.no_grad()
    def enc_p(self, x_in: torch.Tensor, z_noise: torch.Tensor, noice_scale: float, want_stats: bool = False):
        """``TextEncoder.forward`` after the f0-embedding add, all-ones mask (models.py:155-162) -> z_p [B,inter,T]
        (and m_p, logs_p when ``want_stats``)."""
        x_in = self._f32(x_in, "x_in"); z_noise = self._f32(z_noise, "z_noise")
        B, _, T = x_in.shape
        z = torch.empty((B, self.cfg.inter_channels, T), dtype=torch.float32, device=self.device)
        m = torch.empty_like(z) if want_stats else None
        lg = torch.empty_like(z) if want_stats else None
        rc = self.lib.svb_enc_p(self._ctx, x_in.data_ptr(), z_noise.data_ptr(), float(noice_scale), z.data_ptr(),
                                m.data_ptr() if want_stats else None, lg.data_ptr() if want_stats else None, B, T, self._stream())
        L.check(self.lib, self._ctx, rc, "svb_enc_p")
        return (z, m, lg) if want_stats else z

    @torch.no_grad()
    def generator(self, z: torch.Tensor, g: torch.Tensor, har: torch.Tensor) -> torch.Tensor:
        z = self._f32(z, "z"); g = self._f32(g, "g"); har = self._f32(har, "har")
        B, _, T = z.shape
        N = T * self.cfg.hop
        wav = torch.empty((B, 1, N), dtype=torch.float32, device=self.device)
        rc = self.lib.svb_generator(self._ctx, z.data_ptr(), g.data_ptr(), g.shape[2], har.data_ptr(), wav.data_ptr(),
                                    B, T, None, 0, self._stream())
        L.check(self.lib, self._ctx, rc, "svb_generator")
        return wav

    @torch.no_grad()
    def vocoder(self, mel: torch.Tensor, f0: torch.Tensor, rand_ini: torch.Tensor, har_noise: Optional[torch.Tensor]) -> torch.Tensor:
        """vdecoder.nsf_hifigan Generator.forward(mel, f0) (vdecoder/nsf_hifigan/models.py:259-278) -> [B,1,N]."""
        mel = self._f32(mel, "mel"); f0 = self._f32(f0, "f0"); rand_ini = self._f32(rand_ini, "rand_ini")
        B, _, T = mel.shape
        N = T * self.cfg.hop
        nz = None
        if har_noise is not None:
            har_noise = self._f32(har_noise, "har_noise")
            assert har_noise.shape == (B, N, self.cfg.n_harmonics), har_noise.shape
            nz = har_noise.data_ptr()
        wav = torch.empty((B, 1, N), dtype=torch.float32, device=self.device)
        rc = self.lib.svb_vocoder(self._ctx, mel.data_ptr(), f0.data_ptr(), rand_ini.data_ptr(), nz, wav.data_ptr(), B, T,
                                  None, 0, self._stream())
        L.check(self.lib, self._ctx, rc, "svb_vocoder")
        return wav

    @torch.no_grad()
    def infer_tail(self, z_p, g, f0, rand_ini, har_noise=None, lengths=None) -> torch.Tensor:
        """z = flow(z_p, reverse); o = dec(z, g, f0)  (models.py:530-531).  Returns [B,1,N]."""
        z_p = self._f32(z_p, "z_p"); g = self._f32(g, "g"); f0 = self._f32(f0, "f0")
        rand_ini = self._f32(rand_ini, "rand_ini")
        B, _, T = z_p.shape
        N = T * self.cfg.hop
        nz = None
        if har_noise is not None:
            har_noise = self._f32(har_noise, "har_noise")
            assert har_noise.shape == (B, N, self.cfg.n_harmonics), har_noise.shape
            nz = har_noise.data_ptr()
        keep, lp = self._lengths(lengths)
        wav = torch.empty((B, 1, N), dtype=torch.float32, device=self.device)
        rc = self.lib.svb_infer_tail(self._ctx, z_p.data_ptr(), g.data_ptr(), g.shape[2], lp, f0.data_ptr(),
                                     rand_ini.data_ptr(), nz, wav.data_ptr(), B, T, None, 0, self._stream())
        L.check(self.lib, self._ctx, rc, "svb_infer_tail")
        return wav

    @torch.no_grad()
    def infer_tail_host(self, z_p, g, f0, rand_ini, har_noise=None) -> torch.Tensor:
        """Same through ``svb_infer_tail_host``: HOST tensors in, HOST waveform out."""
        hs = [t.detach().to("cpu", torch.float32).contiguous() for t in (z_p, g, f0, rand_ini)]
        B, _, T = hs[0].shape
        N = T * self.cfg.hop
        nz = None
        if har_noise is not None:
            har_noise = har_noise.detach().to("cpu", torch.float32).contiguous()
            nz = har_noise.data_ptr()
        wav = torch.empty((B, 1, N), dtype=torch.float32, pin_memory=True)
        with torch.cuda.device(self.device):
            rc = self.lib.svb_infer_tail_host(self._ctx, hs[0].data_ptr(), hs[1].data_ptr(), hs[1].shape[2],
                                              hs[2].data_ptr(), hs[3].data_ptr(), nz, wav.data_ptr(), B, T)
        L.check(self.lib, self._ctx, rc, "svb_infer_tail_host")
        return wav
