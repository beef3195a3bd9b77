"""Developer tool (GPU box): device timeline of a few frames of fl_match via torch.profiler (CUPTI activity records):
every kernel / memcpy / memset with start offset, duration and the gap since the previous activity ended."""
import os, sys, json, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
import fealess_b200 as fb
from fealess_b200 import synth

W, H, T = 640, 480, (5, 8)
NT = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
DEVICE_INPUT = len(sys.argv) > 2 and sys.argv[2] == "device"
torch.cuda.init()
frames = [synth.make_frame(W, H, i) for i in range(4)]
h = fb.Handle(T, (0, 1), W, H)
h.upload_templates(synth.make_templates(0))
rc, _, q = h.match(frames[0][0], frames[0][1], 75.0, want_quantized=True)
ts = synth.make_templates(NT, W, H, T, seed=1, quantized=q, planted_fraction=0.01)
h.upload_templates(ts)
dev = torch.device("cuda", 0)
d_frames = [(torch.from_numpy(b).to(dev), torch.from_numpy(d.view(np.int16)).to(dev)) for b, d in frames]
torch.cuda.synchronize()

def one(i):
    if DEVICE_INPUT:
        tb, td = d_frames[i % 4]
        h.match_device(tb.data_ptr(), td.data_ptr(), W, H, 75.0)
    else:
        h.match(frames[i % 4][0], frames[i % 4][1], 75.0, capacity=4096)

for i in range(8):
    one(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(6):
        one(i)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "trace.json")
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
gpu = [e for e in ev if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
gpu.sort(key=lambda e: e["ts"])
cpu = [e for e in ev if e.get("ph") == "X" and e.get("cat") in ("cuda_runtime", "cuda_driver")]
cpu.sort(key=lambda e: e["ts"])
t0 = gpu[0]["ts"]
prev_end = None
print("== device activities ==")
for e in gpu:
    gap = (e["ts"] - prev_end) if prev_end is not None else 0.0
    print("%9.1f us  dur %7.1f  gap %7.1f  %s" % (e["ts"] - t0, e["dur"], gap, e["name"][:70]))
    prev_end = e["ts"] + e["dur"]
print("== host API calls (first two frames) ==")
c0 = cpu[0]["ts"] if cpu else 0
for e in cpu[:80]:
    print("%9.1f us  dur %7.1f  %s" % (e["ts"] - t0, e["dur"], e["name"][:60]))
