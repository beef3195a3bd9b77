"""CPU: the numpy restatement of cv::resize INTER_LINEAR (oracle/resize_oracle.py) against OpenCV itself (cv2 4.13, IPP off).
The reference rescales every frame to 640 columns with it (CadReco/obj_reco_lmicp.cpp:38-45, 255-256)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import resize_oracle as R

SIZES = [(1280, 960, 640, 480), (1280, 800, 640, 400), (320, 240, 640, 480), (800, 600, 640, 480), (1920, 1440, 640, 480),
         (1024, 768, 640, 480), (848, 480, 640, 362), (641, 481, 640, 480), (333, 250, 640, 480), (4, 3, 640, 480), (1280, 961, 640, 480)]


@pytest.mark.parametrize("sw,sh,dw,dh", SIZES)
def test_oracle_equals_opencv(sw, sh, dw, dh):
    import cv2
    was = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)                                         # OpenCV's own code, not Intel's closed 16U routine
    try:
        rng = np.random.default_rng(sw * 7 + sh)
        a = rng.integers(0, 256, (sh, sw, 3)).astype(np.uint8)
        d = rng.integers(0, 65536, (sh, sw)).astype(np.uint16)
        assert np.array_equal(R.resize_linear(a, dw, dh), cv2.resize(a, (dw, dh), interpolation=cv2.INTER_LINEAR))
        assert np.array_equal(R.resize_linear(d, dw, dh), cv2.resize(d, (dw, dh), interpolation=cv2.INTER_LINEAR))
    finally:
        cv2.ipp.setUseIPP(was)


def test_depth_like_input_and_identity():
    import cv2
    cv2.ipp.setUseIPP(False)
    try:
        from fealess_b200 import synth
        b, d = synth.make_frame(640, 480, 3)
        big_b = cv2.resize(b, (1024, 768), interpolation=cv2.INTER_CUBIC)
        big_d = cv2.resize(d, (1024, 768), interpolation=cv2.INTER_NEAREST)          # keeps the zero holes
        assert np.array_equal(R.resize_linear(big_b, 640, 480), cv2.resize(big_b, (640, 480), interpolation=cv2.INTER_LINEAR))
        assert np.array_equal(R.resize_linear(big_d, 640, 480), cv2.resize(big_d, (640, 480), interpolation=cv2.INTER_LINEAR))
        assert np.array_equal(R.resize_linear(d, 640, 480), d)
    finally:
        cv2.ipp.setUseIPP(True)


def test_oracle_equals_committed_golden(golden_dir):
    """tests/golden/resize_small.npz (written by oracle/make_golden.py from cv2.resize with IPP off) - no cv2 needed to check."""
    z = np.load(os.path.join(golden_dir, "resize_small.npz"))
    for i, (sw, sh, dw, dh) in enumerate(z["sizes"]):
        assert z["bgr_%d" % i].shape == (sh, sw, 3)
        assert np.array_equal(R.resize_linear(z["bgr_%d" % i], int(dw), int(dh)), z["bgr_out_%d" % i])
        assert np.array_equal(R.resize_linear(z["depth_%d" % i], int(dw), int(dh)), z["depth_out_%d" % i])
