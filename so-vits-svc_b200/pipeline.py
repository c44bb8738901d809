"""Host-buffer serving loop: ``SynthesizerTrn.infer`` with HOST inputs and a HOST waveform, copies overlapped with compute.

``Svc.infer`` (inference/infer_tool.py:289-298) uploads the features of a slice, runs ``net_g_ms.infer`` and reads the
waveform back on ONE stream, so every call pays H2D + compute + D2H in series (21 MB up and 14 MB down per 8 x 10 s batch:
0.6 ms next to 10.9 ms of kernels).  ``HostPipeline`` keeps the reference's call semantics - inputs in, waveform out, same
seed handling inside ``infer`` - but issues the upload on a copy stream, the kernels on the caller's stream and the read-back
on a second copy stream, ordered by events.  Back-to-back calls therefore overlap call i+1's upload and call i-1's read-back
with call i's kernels.  Results are bit-identical to calling ``infer`` directly (tests/test_gpu_parity.py).
"""
from __future__ import annotations

from typing import List, Sequence

import torch


class HostPipeline:
    def __init__(self, net, device: torch.device, depth: int = 2):
        self.net, self.device, self.depth = net, torch.device(device), depth
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self._slot = 0
        self._out_host: List[torch.Tensor] = [None] * depth          # pinned result buffers, reused round-robin
        self._done: List[torch.cuda.Event] = [None] * depth
        # device input buffers are owned by the pipeline (one set per slot): the host may run many steps ahead of the GPU, and
        # per-step allocations on the copy stream would then miss the caching allocator and fall into cudaMalloc (a device-wide
        # synchronisation) - measured as 16-22 ms steps instead of 10.9
        self._in_dev: List[list] = [None] * depth
        self._computed: List[torch.cuda.Event] = [None] * depth
        self._keep: list = [None] * depth

    @torch.no_grad()
    def submit(self, c: torch.Tensor, f0: torch.Tensor, uv: torch.Tensor, g: torch.Tensor, **kw):
        """Pinned HOST tensors in; returns (pinned HOST waveform [B,1,N], event).  The waveform is valid once the event has
        completed (``event.synchronize()``) and until ``depth`` further submissions have been made."""
        dev = self.device
        cur = torch.cuda.current_stream(dev)
        k = self._slot
        self._slot = (k + 1) % self.depth
        host = (c, f0, uv, g)
        if self._in_dev[k] is None or any(d.shape != h.shape or d.dtype != h.dtype for d, h in zip(self._in_dev[k], host)):
            if self._computed[k] is not None:
                self._computed[k].synchronize()                       # (rare) shape change: the old buffers are still being read
            with torch.cuda.stream(self.s_in):                        # blocks come from the COPY stream's pool: a block recycled from
                self._in_dev[k] = [torch.empty(h.shape, dtype=h.dtype, device=dev) for h in host]   # the compute stream could still
            for t in self._in_dev[k]:                                 # be in use by a kernel in flight there
                t.record_stream(cur)
            self._computed[k] = None
        if self._computed[k] is not None:
            self.s_in.wait_event(self._computed[k])                   # the kernels that read this slot's inputs have finished
        with torch.cuda.stream(self.s_in):
            for d, h in zip(self._in_dev[k], host):
                d.copy_(h, non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record(self.s_in)
        cur.wait_event(ev_in)
        ins = self._in_dev[k]
        o, _ = self.net.infer(ins[0], ins[1], ins[2], g=ins[3], **kw)
        ev_c = torch.cuda.Event()
        ev_c.record(cur)
        self._computed[k] = ev_c
        if self._done[k] is not None:
            self._done[k].synchronize()                              # the buffer's previous read-back must have been consumed
        if self._out_host[k] is None or self._out_host[k].shape != o.shape:
            self._out_host[k] = torch.empty(o.shape, dtype=torch.float32).pin_memory()
        self.s_out.wait_event(ev_c)
        with torch.cuda.stream(self.s_out):
            self._out_host[k].copy_(o, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.s_out)
        self._done[k] = ev
        self._keep[k] = o            # the device waveform stays referenced until this slot's read-back has been waited for (no
                                     # record_stream: deferred frees make the caching allocator fall back to cudaMalloc)
        return self._out_host[k], ev

    def run(self, batches: Sequence[Sequence[torch.Tensor]], **kw) -> List[torch.Tensor]:
        """Convenience: a list of (c, f0, uv, g) host batches -> list of host waveforms (copies), fully pipelined."""
        outs = []
        pending = []
        for b in batches:
            pending.append(self.submit(*b, **kw))
            if len(pending) == self.depth:
                o, ev = pending.pop(0)
                ev.synchronize()
                outs.append(o.clone())
        for o, ev in pending:
            ev.synchronize()
            outs.append(o.clone())
        return outs
