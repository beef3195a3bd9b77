"""Developer experiment: per-warp end times inside the staged kernel (fl_debug_option FL_OPT_TRACE)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np
import fealess_b200 as fb
from fealess_b200 import synth
W, H, T = 640, 480, (5, 8)
frames = [synth.make_frame(W, H, i) for i in range(4)]
h = fb.Handle(T, (0, 1), W, H)
h.debug_option(fb.FL_OPT_TRACE, 1)
h.upload_templates(synth.make_templates(0))
rc, _, q = h.match(frames[0][0], frames[0][1], 75.0, want_quantized=True)
ts = synth.make_templates(8000, W, H, T, seed=1, quantized=q, planted_fraction=0.01)
h.upload_templates(ts)
for it in range(8):
    rc, m = h.match(*frames[it % 4], 75.0)
buf = np.zeros(1024 * 136 + 8, np.uint64)
n = fb.lib().fl_debug_get(h._h, 4, 0, 0, 0, C.c_void_p(buf.ctypes.data), C.c_size_t(buf.nbytes))
tr = buf[:n * 8].reshape(n, 8).astype(np.int64)
pw = buf[n * 8 + 8:n * 8 + 8 + n * 128].reshape(n, 32, 4).astype(np.int64)
t0 = tr[:, 0].min()
nw = int((pw[:, :, 0] > 0).sum(1).max())
le = (pw[:, :nw, 0] - t0) / 1e3
fe = (pw[:, :nw, 1] - t0) / 1e3
print("matches", len(m), "n_cta", n)
print("per-warp loop end (us since first CTA start): min %.1f mean %.1f max %.1f" % (le.min(), le.mean(), le.max()))
print("per-warp final end: min %.1f mean %.1f max %.1f" % (fe.min(), fe.mean(), fe.max()))
print("per-CTA (max - min) of warp loop end: mean %.1f max %.1f" % ((le.max(1) - le.min(1)).mean(), (le.max(1) - le.min(1)).max()))
print("emission time per warp: mean %.2f max %.2f; warps with > 1 us: %d" % ((fe - le).mean(), (fe - le).max(), int(((fe - le) > 1).sum())))
worst = np.argsort(-le.max(1))[:5]
for c in worst:
    print("cta", c, "smid", tr[c, 5], "warp loop ends", np.round(le[c], 1).tolist())
print("per-CTA max loop end sorted (last 10):", np.round(np.sort(le.max(1))[-10:], 1).tolist())
print("per-CTA max loop end sorted (first 10):", np.round(np.sort(le.max(1))[:10], 1).tolist())
rel = (tr[:, :5] - t0) / 1e3
print("per-CTA timeline (us since first CTA start), mean / max: start %.1f/%.1f ready %.1f/%.1f first data %.1f/%.1f loop end %.1f/%.1f end %.1f/%.1f"
      % tuple(x for k in range(5) for x in (rel[:, k].mean(), rel[:, k].max())))
st = buf[n * 8:n * 8 + 2].astype(np.int64)
print("stamp before launch -> first CTA start %.1f us; last CTA end -> stamp after %.1f us; whole %.1f us" % ((t0 - st[0]) / 1e3, (st[1] - tr[:, 4].max()) / 1e3, (st[1] - st[0]) / 1e3))
