"""One small pass over every kernel of the hot path, meant to run under `compute-sanitizer --tool racecheck` (shared-memory hazards)
and `--tool memcheck`: a VGA match with 320 templates (single-launch front end, staged similarity kernel, fused refinement + sort),
the same frame with per-wave launches and separate refinement, a masked frame, two ICP hypotheses, NMS.  Results are checked
against the oracle so that a tool-induced time-out cannot pass silently."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import fealess_b200 as fb
from fealess_b200 import synth
import fl_oracle_py as F
W, H, T = 640, 480, (5, 8)
b, d = synth.make_frame(W, H, 0)
det = F.Detector(T); det.process(b, d)
q = [det.quantized(l, m) for l in range(2) for m in range(2)]
ts = synth.make_templates(320, W, H, T, n_classes=2, seed=3, quantized=q, planted_fraction=0.05)
det.set_templates(ts)
want = det.match(70.0)
h = fb.Handle(T, (0, 1), W, H)
h.upload_templates(ts)
rc, got = h.match(b, d, 70.0)
assert rc == 0 and h.uses_staged() and np.array_equal(got, want), "default path"
h.debug_option(fb.FL_OPT_FE_WAVES, 1); h.debug_option(fb.FL_OPT_SPLIT_REFINE, 1)
rc, got = h.match(b, d, 70.0)
assert rc == 0 and np.array_equal(got, want), "wave launches"
h.debug_option(fb.FL_OPT_FE_WAVES, 0); h.debug_option(fb.FL_OPT_SPLIT_REFINE, 0)
m0 = np.zeros((H, W), np.uint8); m0[40:440, 60:580] = 255
det.process(b, d, [m0, None])
rc, got = h.match(b, d, 70.0, masks=[m0, None])
assert rc == 0 and np.array_equal(got, det.match(70.0)), "masked frame"
Kc = (608.0, 608.0, 320.0, 240.0)
for seed in (0, 1):
    md, rf, rm, rr, p = synth.make_icp_pair(W, H, seed=seed, max_rot_deg=5, max_shift_mm=8)
    R0, t0 = p[:12].reshape(3, 4)[:, :3], p[:12].reshape(3, 4)[:, 3]
    g = h.detection_batch(rf, Kc, [md, md], [rm, rm], [rr, rr], [R0, R0], [t0, t0])
    o = F.detection(md, rf, Kc, rm, rr, r_match=R0, t_match=t0)
    assert np.array_equal(g[0]["T"], o["T"]) and np.array_equal(g[1]["R"].reshape(3, 3), o["R"]), "icp"
t3 = np.random.default_rng(1).uniform(-50, 50, (9, 3)).astype(np.float32)
keep = h.nms(t3, np.full(9, 5000, np.int32), np.linspace(0.1, 2, 9).astype(np.float32), 30.0)
assert len(keep) >= 1
h.close()
print("RACECHECK CASE PASS", flush=True)
