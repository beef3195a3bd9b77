import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import fealess_b200 as fb
from fealess_b200 import synth
W, H, T = 640, 480, (5, 8)
frames = [synth.make_frame(W, H, i) for i in range(4)]
h = fb.Handle(T, (0, 1), W, H)
h.upload_templates(synth.make_templates(0))
rc, _, q = h.match(frames[0][0], frames[0][1], 75.0, want_quantized=True)
ts = synth.make_templates(8000, W, H, T, seed=1, quantized=q, planted_fraction=0.01)
h.upload_templates(ts)
pinned = [(torch.from_numpy(b).pin_memory().numpy(), torch.from_numpy(d.view(np.int16)).pin_memory().numpy().view(np.uint16)) for b, d in frames]
for name, fr in (("pageable", frames), ("pinned", pinned), ("pageable", frames), ("pinned", pinned)):
    for i in range(5): h.match(*fr[i % 4], 75.0, capacity=4096)
    t0 = time.perf_counter()
    for i in range(200): h.match(*fr[i % 4], 75.0, capacity=4096)
    print(name, "us/frame", (time.perf_counter() - t0) / 200 * 1e6)
