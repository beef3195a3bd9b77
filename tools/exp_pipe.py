"""Developer experiment (GPU box): K handles on one GPU, frames dealt round-robin, K frames in flight - steady-state frame time
(device-resident inputs; host wall clock over many frames, which at steady state equals the device rate) against one handle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import fealess_b200 as fb
from fealess_b200 import synth
W, H, T = 640, 480, (5, 8)
NT = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
frames = [synth.make_frame(W, H, i) for i in range(4)]
h0 = fb.Handle(T, (0, 1), W, H)
h0.upload_templates(synth.make_templates(0))
rc, _, q = h0.match(frames[0][0], frames[0][1], 75.0, want_quantized=True)
ts = synth.make_templates(NT, W, H, T, seed=1, quantized=q, planted_fraction=0.01)
dev = [(torch.from_numpy(b).cuda(), torch.from_numpy(d.view(np.int16)).cuda()) for b, d in frames]
want = None
for K in (1, 2, 4, 6, 8):
    hs = [fb.Handle(T, (0, 1), W, H) for _ in range(K)]
    for h in hs: h.upload_templates(ts)
    N = 2000
    res = []
    def run(n):
        for i in range(n + K):
            if i >= K:
                hh = hs[i % K]; hh.match_wait()
                if i < 2 * K + 4: res.append(hh.match_fetch(4096))
            if i < n:
                b, d = dev[i % 4]
                hs[i % K].match_device_async(b.data_ptr(), d.data_ptr(), W, H, 75.0)
    run(50); res.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); run(N); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if want is None: want = [r.copy() for r in res[:4]]
    ok = all(np.array_equal(res[i], want[i]) for i in range(4))
    print("K=%d frames in flight: %.1f us/frame (%.0f frames/s) lists equal %s" % (K, dt / N * 1e6, N / dt, ok), flush=True)
    for h in hs: h.close()
