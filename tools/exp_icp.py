"""Developer timeline of the fused ICP kernel on the C3 workload (GPU box): per-phase SM cycles from fl_debug_icp_trace.
usage: python tools/exp_icp.py [n_hyp]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import fealess_b200 as fb
from fealess_b200 import synth
import bench

n_hyp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ref, mds, rms, rrs, Rs, ts = bench.make_icp_workload(synth, n_hyp)
K = (608.0, 608.0, 320.0, 240.0)
h = fb.Handle((5, 8), (0, 1), 640, 480)
h.detection_batch(ref, K, mds, rms, rrs, Rs, ts)
h.profile(True)
dev = []
for _ in range(5):
    res = h.detection_batch(ref, K, mds, rms, rrs, Rs, ts)
    dev.append(h.last_icp_ms())
tr = h.icp_trace(n_hyp).astype(np.float64)
names = ["pairing", "centroid sums", "shift+grid", "distances", "dist sum(+NN)", "correspond", "cov sums", "svd", "transform", "TOTAL"]
mhz = 1965.0
print("n_hyp %d device_ms min %.3f  iterations total %d max %d  points mean %d" % (n_hyp, min(dev), res["iterations"].sum(), res["iterations"].max(), res["n_points"].mean()))
print("sum over hypotheses of TOTAL: %.1f us -> / 148 SMs = %.1f us" % (tr[:, 9].sum() / mhz, tr[:, 9].sum() / mhz / min(148, n_hyp)))
print("%-16s %10s %10s %12s" % ("phase", "mean us", "max us", "us/iteration"))
its = np.maximum(tr[:, 10], 1)
for p, nm in enumerate(names):
    print("%-16s %10.1f %10.1f %12.1f" % (nm, tr[:, p].mean() / mhz, tr[:, p].max() / mhz, (tr[:, p] / its).mean() / mhz))
worst = int(np.argmax(tr[:, 9]))
print("chain stalls (us, mean): waiting for records %.1f, waiting for distances %.1f; distance sum alone %.1f (of %.1f with the search)" % (tr[:, 12].mean() / mhz, tr[:, 13].mean() / mhz, tr[:, 14].mean() / mhz, tr[:, 4].mean() / mhz))
print("inside the additions (us, mean): record sums %.1f of %.1f, distance sums %.1f of %.1f" % (tr[:, 15].mean() / mhz, (tr[:, 1] + tr[:, 6]).mean() / mhz, tr[:, 16].mean() / mhz, tr[:, 14].mean() / mhz))
print("searches redone after the distance sum (total): %d" % tr[:, 17].sum())
print("slowest hypothesis %d: %d iterations, %.1f us:" % (worst, tr[worst, 10], tr[worst, 9] / mhz), " ".join("%s=%.1f" % (n, tr[worst, p] / mhz) for p, n in enumerate(names[:9])))
by_it = {}
for i in range(n_hyp):
    by_it.setdefault(int(tr[i, 10]), []).append(tr[i, 9] / mhz)
print("total us by iteration count:", {k: round(float(np.mean(v)), 1) for k, v in sorted(by_it.items())}, "counts", {k: len(v) for k, v in sorted(by_it.items())})

if n_hyp > 148:
    first, second = tr[:148], tr[148:]
    print("first 148 tickets vs the rest (mean us):", " ".join("%s=%.1f/%.1f" % (n, first[:, p].mean() / mhz, second[:, p].mean() / mhz) for p, n in enumerate(names[:9])))
    print("pairing us, every 16th hypothesis:", [round(float(v) / mhz, 1) for v in tr[::16, 0]])
if n_hyp > 148 and len(sys.argv) > 2:
    print("pairing us by hypothesis:", [int(round(float(v) / mhz)) for v in tr[:, 0]])
    print("valid pairs by hypothesis:", [int(v) for v in tr[:48, 11]])

if len(sys.argv) > 2:
    print("pairing pass 1 / scan (us) for hypotheses 0..23:", [(int(round(float(a) / mhz)), int(round(float(b) / mhz))) for a, b in zip(tr[:24, 18], tr[:24, 19])])
