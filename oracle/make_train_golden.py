"""TEST INFRASTRUCTURE.  Golden vectors for template training: the reference's own Detector::addTemplate (oracle/_ref, i.e.
/root/reference/linemod/linemod.cpp:1579-1615 compiled unmodified) on synthetic views -> tests/golden/train_vga.npz.
Run in the build container (needs oracle/_ref):  python oracle/make_train_golden.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import fl_ref_py as R
from fealess_b200 import synth

W, H, T = 640, 480, (5, 8)
CASES = [(0, (320, 240, 110, 80), 255), (0, (300, 250, 120, 90), 1), (2, None, 0), (3, (320, 240, 6, 5), 255), (1, (200, 300, 60, 140), 255)]   # (frame, ellipse cx cy a b | None, mask value)


def ellipse(cx, cy, a, b, value):
    yy, xx = np.mgrid[0:H, 0:W]
    return (((xx - cx) / a) ** 2 + ((yy - cy) / b) ** 2 <= 1.0).astype(np.uint8) * value


def main():
    assert R.available(), "oracle/_ref/libfl_ref.so is missing: python oracle/build_ref.py"
    out = {"cases": np.array([[c[0]] + list(c[1] or (0, 0, 0, 0)) + [c[2]] for c in CASES], np.int32)}
    det = R.Detector(T)
    for i, (frame, ell, val) in enumerate(CASES):
        b, d = synth.make_frame(W, H, frame)
        mask = ellipse(*ell, val) if ell else None
        rc, hdr, ft, bb = R.add_template(det, b, d, mask)
        out["rc%d" % i] = np.int32(rc); out["hdr%d" % i] = hdr; out["ft%d" % i] = ft; out["bb%d" % i] = bb
        print("case", i, "rc", rc, "features", len(ft), "bbox", bb.tolist())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "train_vga.npz"), **out)


if __name__ == "__main__":
    main()
