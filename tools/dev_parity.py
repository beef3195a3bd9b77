"""Developer smoke script (GPU box): run every stage of the CUDA path against the C oracle and print mismatch counts."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import fealess_b200 as fb
from fealess_b200 import synth
import fl_oracle_py as F

def rot_err(Ra, Rb):
    D = Ra.astype(np.float64).T @ Rb.astype(np.float64)
    return float(np.linalg.norm([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]]) / 2)

def main():
    W, H, T = 640, 480, (5, 8)
    b, d = synth.make_frame(W, H, 0)
    det = F.Detector(T); det.process(b, d)
    q = [det.quantized(l, m) for l in range(2) for m in range(2)]
    ts = synth.make_templates(300, W, H, T, quantized=q, planted_fraction=0.1, seed=2, n_classes=3)
    det.set_templates(ts)
    h = fb.Handle(T, (0, 1), W, H)
    h.upload_templates(ts); h.keep_spread(True)
    t0 = time.time(); rc, m, qg = h.match(b, d, 75.0, want_quantized=True); print("fl_match rc", rc, "n", len(m), "t", time.time() - t0)
    for l in range(2):
        for mo in range(2):
            print("quantized L%d M%d mismatches" % (l, mo), int((qg[l * 2 + mo] != q[l * 2 + mo]).sum()))
            print("   spread mismatches", int((h.debug_quantized(l, mo, W, H, spread=True) != det.spread(l, mo)).sum()))
            print("   lm mismatches", [int((h.debug_lm(l, mo, lab, W, H) != det.lm(l, mo, lab)).sum()) for lab in range(8)])
    bad = 0
    for t in range(0, ts.n_templates, 7):
        bad += int((h.debug_similarity(t, W, H) != det.similarity(t)).sum())
    print("similarity map mismatches", bad)
    for thr in (75.0, 60.0, 52.0):
        rc, m = h.match(b, d, thr)
        o = det.match(thr)
        same = len(m) == len(o) and bool((m == o).all())
        print("threshold", thr, "rc", rc, "gpu", len(m), "oracle", len(o), "identical", same)
        if not same:
            k = min(len(m), len(o))
            diff = np.nonzero(m[:k] != o[:k])[0]
            print("  first diff at", diff[:5], m[diff[:3]] if len(diff) else None, o[diff[:3]] if len(diff) else None)
    # ICP
    cases = [synth.make_icp_pair(W, H, seed=s, max_rot_deg=r, max_shift_mm=sh) for s, (r, sh) in enumerate([(3, 6), (8, 10), (12, 15), (5, 3), (15, 20), (2, 2), (10, 5), (6, 12)])]
    K = (608.0, 608.0, 320.0, 240.0)
    ref = cases[0][1]
    for i, (md, rf, rm, rr, p) in enumerate(cases):
        R0 = p[:12].reshape(3, 4)[:, :3]; t0_ = p[:12].reshape(3, 4)[:, 3]
        o = F.detection(md, rf, K, rm, rr, r_match=R0, t_match=t0_)
        g = h.detection_batch(rf, K, [md], [rm], [rr], [R0], [t0_])[0]
        print("icp case", i, "iters", o["iterations"], g["iterations"], "n", o["n_points"], g["n_points"], "dm", o["dist_mean"], g["dist_mean"],
              "rot_err", rot_err(o["R"], g["R"].reshape(3, 3)), "dT", float(np.abs(o["T"] - g["T"]).max()), "bitexact", bool((o["R"].ravel() == g["R"]).all() and (o["T"] == g["T"]).all()))
    print("launches", h.launch_count())

if __name__ == "__main__":
    main()
