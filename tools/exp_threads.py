"""Developer experiment (GPU box): K host threads, one handle each, on ONE GPU - how do match-only, ICP-only and match + ICP
frames overlap?  Prints frames/s per mode and K."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import fealess_b200 as fb
from fealess_b200 import synth
W, H, T = (int(sys.argv[1]), int(sys.argv[2]), (5, 8)) if len(sys.argv) > 2 else (640, 480, (5, 8))
NT = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
NCLS = int(sys.argv[4]) if len(sys.argv) > 4 else 1
frames = [synth.make_frame(W, H, i) for i in range(2)]
h0 = fb.Handle(T, (0, 1), W, H)
fb_Handle = lambda: fb.Handle(T, (0, 1), W, H)
h0.upload_templates(synth.make_templates(0))
rc, _, q = h0.match(frames[0][0], frames[0][1], 75.0, want_quantized=True)
ts = synth.make_templates(NT, W, H, T, n_classes=NCLS, seed=1 if NCLS > 1 else 7, quantized=q, planted_fraction=0.004 if NCLS > 1 else 0.01)
hdr0 = ts.headers.reshape(NT, 4, 7)[:, 0, :]
rects = np.stack([hdr0[:, 2], hdr0[:, 3], hdr0[:, 0], hdr0[:, 1]], axis=1).astype(np.int32)
K_ = (608.0 * W / 640, 608.0 * W / 640, W / 2.0, H / 2.0)
dev = [(torch.from_numpy(b).cuda(), torch.from_numpy(d.view(np.int16)).cuda()) for b, d in frames]
KMAX = 8 if NT <= 4000 else 4
hs = []
for k in range(KMAX):
    h = fb.Handle(T, (0, 1), W, H); h.upload_templates(ts); h.upload_model_depths([frames[0][1]] * NT, rects); hs.append(h)
h = hs[0]
h.match_device(dev[0][0].data_ptr(), dev[0][1].data_ptr(), W, H, 75.0)
top = h.match_fetch()[:5]
cf = np.array([int(np.argmax(ts.class_of == c)) for c in range(NCLS)])
gidx = (cf[top["class_idx"]] + top["template_id"]).astype(np.int32)
rr = np.stack([top["x"], top["y"], rects[gidx, 2], rects[gidx, 3]], axis=1).astype(np.int32)
P = ts.pose13[gidx][:, :12].reshape(len(top), 3, 4)
Rm, tm = np.ascontiguousarray(P[:, :, :3]), np.ascontiguousarray(P[:, :, 3])
print("hypotheses", len(top), "points", (rects[gidx, 2] * rects[gidx, 3]).tolist(), flush=True)

def work(mode, hk, n):
    torch.cuda.set_device(0)
    for i in range(n):
        if mode in ("match", "both"):
            hk.match_device(dev[0][0].data_ptr(), dev[0][1].data_ptr(), W, H, 75.0); hk.match_fetch()
        if mode in ("icp", "both"):
            hk.detection_batch_resident_device(dev[0][1].data_ptr(), W, H, K_, gidx, rr, Rm, tm)

for mode in ("match", "icp", "both"):
    for K in [k for k in (1, 2, 4, 8) if k <= KMAX]:
        n = 200 if NT <= 4000 else 40
        for rep in range(2):
            ths = [threading.Thread(target=work, args=(mode, hs[k], n if rep else 10)) for k in range(K)]
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for t in ths: t.start()
            for t in ths: t.join()
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("%-5s K=%d: %.0f frames/s (%.3f ms per frame per thread)" % (mode, K, K * n / dt, dt / n * 1e3), flush=True)
