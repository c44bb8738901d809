"""Timeline of the HostPipeline e2e loop: per step, when the upload finished, when the kernels started / ended and when the
read-back finished (CUDA events, ms from the first submit), plus the host time spent inside submit().  Diagnostic only."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sovits_b200
from sovits_b200 import models, synth
from sovits_b200.config import load_config
from sovits_b200.pipeline import HostPipeline

dev = torch.device("cuda:0")
cfg = load_config()
sd = synth.synth_state_dict(cfg)
kw = json.load(open(sovits_b200.DEFAULT_CONFIG))["model"]
net = models.SynthesizerTrn(1025, 20, **kw).eval()
net.load_state_dict(sd)
net = net.to(dev)
net.set_precision("tc")
B, T = 8, 862
c, f0, uv, sid = synth.synth_inputs(cfg, B, T)
host = [t.contiguous().pin_memory() for t in (c, f0, uv, sid)]
depth = int(os.environ.get("DEPTH", "2"))
pipe = HostPipeline(net, dev, depth=depth)
for _ in range(4):
    o, ev = pipe.submit(*host, noice_scale=0.4)
ev.synchronize(); torch.cuda.synchronize()

class Probe:
    def __init__(self): self.rows = []
probe = Probe()
orig_infer = net.infer
def traced_infer(*a, **k):
    s = torch.cuda.current_stream()
    e0 = torch.cuda.Event(enable_timing=True); e0.record(s)
    r = orig_infer(*a, **k)
    e1 = torch.cuda.Event(enable_timing=True); e1.record(s)
    probe.rows.append([e0, e1])
    return r
net.infer = traced_infer
t0e = torch.cuda.Event(enable_timing=True); t0e.record()
hosts = []
evs = []
tw = time.perf_counter()
for i in range(12):
    th = time.perf_counter()
    o, ev = pipe.submit(*host, noice_scale=0.4)
    hosts.append((th - tw, time.perf_counter() - th))
    evs.append(ev)
ev.synchronize(); torch.cuda.synchronize()
wall = time.perf_counter() - tw
print(f"depth {depth}: 12 steps in {wall*1e3:.2f} ms wall = {wall*1e3/12:.3f} ms/step")
for i, (e0, e1) in enumerate(probe.rows):
    print(f"step {i:2d}: host enter {hosts[i][0]*1e3:8.3f} ms, in submit {hosts[i][1]*1e3:7.3f} ms | kernels start {t0e.elapsed_time(e0):8.3f} end {t0e.elapsed_time(e1):8.3f} (dur {e0.elapsed_time(e1):.3f})")
# the same loop without the pipeline: device-resident inputs
net.infer = orig_infer
devin = [t.to(dev) for t in host]
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(12): net.infer(devin[0], devin[1], devin[2], g=devin[3], noice_scale=0.4)
e1.record(); torch.cuda.synchronize()
print(f"device-resident: {e0.elapsed_time(e1)/12:.3f} ms/step")
