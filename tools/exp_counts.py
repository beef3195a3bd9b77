import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np
import fealess_b200 as fb
from fealess_b200 import synth
W, H, T = 640, 480, (5, 8)
frames = [synth.make_frame(W, H, i) for i in range(8)]
h = fb.Handle(T, (0, 1), W, H)
h.upload_templates(synth.make_templates(0))
rc, _, q = h.match(frames[0][0], frames[0][1], 75.0, want_quantized=True)
ts = synth.make_templates(8000, W, H, T, seed=1, quantized=q, planted_fraction=0.01)
h.upload_templates(ts)
for i in range(8):
    rc, m = h.match(*frames[i], 75.0)
    c = np.zeros(16, np.int32)
    fb.lib().fl_debug_get(h._h, 5, 0, 0, 0, C.c_void_p(c.ctypes.data), C.c_size_t(64))
    print("frame", i, "matches", c[0], "live after refine", c[1], "big", c[2], "raw candidates", c[3])
