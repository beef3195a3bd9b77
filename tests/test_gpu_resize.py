"""GPU (-m gpu): the device rescale in front of the matcher (CObjRecoLmICP::PrepareInputData, obj_reco_lmicp.cpp:38-45, 216-259:
cv::resize INTER_LINEAR of the 8UC3 colour and 16UC1 depth frame to 640 columns) through the C ABI, bit-exact against the oracle
(oracle/resize_oracle.py, pinned on OpenCV's own code by tests/test_resize_oracle.py)."""
import os
import sys

import numpy as np
import pytest

import fealess_b200 as fb
import fl_oracle_py as F
from fealess_b200 import synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import resize_oracle as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sw,sh,dw,dh", [(1280, 960, 640, 480), (1024, 768, 640, 480), (320, 240, 640, 480), (848, 480, 640, 362),
                                         (1280, 961, 640, 480), (641, 481, 640, 480), (1920, 1440, 640, 480), (7, 5, 640, 480)])
def test_device_resize_bit_exact(sw, sh, dw, dh):
    h = fb.Handle()
    rng = np.random.default_rng(sw + 3 * sh)
    a = rng.integers(0, 256, (sh, sw, 3)).astype(np.uint8)
    d = rng.integers(0, 65536, (sh, sw)).astype(np.uint16)
    assert np.array_equal(h.resize_linear(a, dw, dh), R.resize_linear(a, dw, dh))
    assert np.array_equal(h.resize_linear(d, dw, dh), R.resize_linear(d, dw, dh))
    # strided source rows (a view into a wider image)
    wide = np.zeros((sh, sw + 9, 3), np.uint8); wide[:, :sw] = a
    assert np.array_equal(h.resize_linear(wide[:, :sw], dw, dh), R.resize_linear(a, dw, dh))


def test_device_resize_equals_committed_golden(golden_dir):
    """The fixtures cv2.resize (IPP off) produced when oracle/make_golden.py ran (tests/golden/resize_small.npz)."""
    z = np.load(os.path.join(golden_dir, "resize_small.npz"))
    h = fb.Handle()
    for i, (sw, sh, dw, dh) in enumerate(z["sizes"]):
        assert np.array_equal(h.resize_linear(z["bgr_%d" % i], int(dw), int(dh)), z["bgr_out_%d" % i])
        assert np.array_equal(h.resize_linear(z["depth_%d" % i], int(dw), int(dh)), z["depth_out_%d" % i])


@pytest.mark.parametrize("sw,sh", [(1024, 768), (1280, 960), (320, 240)])
def test_match_rescaled_equals_match_of_rescaled_frame(sw, sh):
    """fl_match_rescaled(frame at its own size) == fl_match(oracle-rescaled frame) == the CPU oracle's match list; the rescaled depth
    frame stays on the device as the reference frame of the resident ICP path."""
    import cv2
    W, H, T = 640, 480, (5, 8)
    b, d = synth.make_frame(W, H, 5)
    big_b = cv2.resize(b, (sw, sh), interpolation=cv2.INTER_CUBIC)
    big_d = cv2.resize(d, (sw, sh), interpolation=cv2.INTER_NEAREST)
    rb, rd = R.resize_linear(big_b, W, H), R.resize_linear(big_d, W, H)
    det = F.Detector(T)
    assert det.process(rb, rd) == 0
    q = [det.quantized(l, m) for l in range(2) for m in range(2)]
    ts = synth.make_templates(300, W, H, T, n_classes=2, seed=17, quantized=q, planted_fraction=0.05)
    det.set_templates(ts)
    want = det.match(70.0)
    assert len(want) > 0
    h = fb.Handle()
    h.upload_templates(ts)
    rc, got, dep = h.match_rescaled(big_b, big_d, W, H, 70.0, want_depth=True)
    assert rc == 0 and np.array_equal(dep, rd)
    rc2, plain = h.match(rb, rd, 70.0)
    assert rc2 == 0 and got.tobytes() == plain.tobytes()
    assert len(got) == len(want) and np.array_equal(got, want)
    # same size: plain fl_match
    rc3, same = h.match_rescaled(rb, rd, W, H, 70.0)
    assert rc3 == 0 and same.tobytes() == plain.tobytes()
    # ICP against the frame left on the device by match_rescaled == ICP against the rescaled frame passed from the host
    rc, _ = h.match_rescaled(big_b, big_d, W, H, 70.0)
    K = (608.0, 608.0, 320.0, 240.0)
    rects = [(100, 100, 90, 80), (300, 200, 64, 64)]
    h.upload_model_depths([rd, rd], rects)
    on_dev = h.detection_batch_resident(None, K, [0, 1], [(104, 98, 90, 80), (297, 203, 64, 64)], frame_size=(W, H))
    from_host = h.detection_batch(rd, K, [rd, rd], rects, [(104, 98, 90, 80), (297, 203, 64, 64)])
    assert on_dev.tobytes() == from_host.tobytes()


def test_recognition_rescales_on_the_device(tmp_path):
    """Recognition on a 1280x960 frame (PrepareInputData -> 640x480) == Recognition on the oracle-rescaled 640x480 frame."""
    import cv2
    from fealess_b200 import reco
    W, H, T = 640, 480, (5, 8)
    b, d = synth.make_frame(W, H, 0)
    big_b = cv2.resize(b, (1280, 960), interpolation=cv2.INTER_CUBIC)
    big_d = cv2.resize(d, (1280, 960), interpolation=cv2.INTER_NEAREST)
    rb, rd = R.resize_linear(big_b, W, H), R.resize_linear(big_d, W, H)
    det = F.Detector(T)
    assert det.process(rb, rd) == 0
    q = [det.quantized(l, m) for l in range(2) for m in range(2)]
    ts = synth.make_templates(24, W, H, T, n_classes=1, seed=61, quantized=q, planted_fraction=0.5)
    results = []
    for frame, Kd in (((big_b, big_d), dict(fx=608.0, fy=608.0, cx=320.0, cy=240.0, width=1280, height=960)),
                      ((rb, rd), dict(fx=608.0, fy=608.0, cx=320.0, cy=240.0, width=640, height=480))):
        D = fb.Detector()
        D.add_template_set(ts)
        r = reco.ObjRecoLmICP()
        r.add_detector(D, {(None, tid): rd for tid in range(ts.n_templates)})
        rc, res = r.Recognition(frame[0], frame[1], Kd, top_k=3)
        assert rc == 0 and len(res) >= 1 and r.last_icp_path == "resident"
        results.append(res)
    assert len(results[0]) == len(results[1])
    for a, c in zip(*results):
        assert a["template_id"] == c["template_id"] and a["similarity"] == c["similarity"]
        assert a["tWorld2Cam"].tobytes() == c["tWorld2Cam"].tobytes() and a["icp"].tobytes() == c["icp"].tobytes()
