"""Developer experiment (GPU box): where the per-stage dead time comes from (memset between kernels, shared-memory carveout changes)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import fealess_b200 as fb
from fealess_b200 import synth

W, H, T = 640, 480, (5, 8)
os.environ["FL_SS_CLUSTER"] = "1"
os.environ["FL_SS_PHASE_KB"] = "100"
frames = [synth.make_frame(W, H, i) for i in range(4)]
h0 = fb.Handle(T, (0, 1), W, H)
h0.upload_templates(synth.make_templates(0))
rc, _, q = h0.match(frames[0][0], frames[0][1], 75.0, want_quantized=True)
ts = synth.make_templates(8000, W, H, T, seed=1, quantized=q, planted_fraction=0.01)
h0.close()
for carve in (0, 1):
    for nomemset in (0, 1):
        os.environ.pop("FL_CARVEOUT", None); os.environ.pop("FL_NO_MEMSET", None)
        if carve: os.environ["FL_CARVEOUT"] = "1"
        if nomemset: os.environ["FL_NO_MEMSET"] = "1"
        h = fb.Handle(T, (0, 1), W, H)
        h.upload_templates(ts)
        h.profile(True)
        st = np.zeros(4); n = 0
        for it in range(44):
            b, d = frames[it % 4]
            rc, m = h.match(b, d, 75.0)
            if it >= 4:
                st += h.last_stage_ms(); n += 1
        st /= n
        print("carveout %d no_memset %d | matches %d | stage us: fe %.1f sim %.1f refine %.1f sort %.1f total %.1f" % (carve, nomemset, len(m), *(1e3 * st), 1e3 * st.sum()), flush=True)
        h.close()
