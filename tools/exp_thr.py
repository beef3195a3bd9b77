"""Developer experiment: similarity stage time vs threshold / planted templates (does candidate emission cost time?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import fealess_b200 as fb
from fealess_b200 import synth
W, H, T = 640, 480, (5, 8)
frames = [synth.make_frame(W, H, i) for i in range(4)]
h = fb.Handle(T, (0, 1), W, H)
h.upload_templates(synth.make_templates(0))
rc, _, q = h.match(frames[0][0], frames[0][1], 75.0, want_quantized=True)
for planted in (0.01, 0.0):
    ts = synth.make_templates(8000, W, H, T, seed=1, quantized=q, planted_fraction=planted)
    h.upload_templates(ts)
    h.profile(True)
    for thr in (75.0, 90.0, 99.9):
        st = np.zeros(4); n = 0
        for it in range(44):
            rc, m = h.match(*frames[it % 4], thr)
            if it >= 4:
                st += h.last_stage_ms(); n += 1
        st /= n
        print("planted %.2f thr %.1f matches(last) %d | stage us: fe %.1f sim %.1f refine %.1f sort %.1f" % (planted, thr, len(m), *(1e3 * st)), flush=True)
