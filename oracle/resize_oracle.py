"""TEST INFRASTRUCTURE - CPU oracle of ``cv::resize(src, dst, Size(w, h), 0, 0, INTER_LINEAR)`` for 8UC3 and 16UC1 images.

The reference rescales every input frame to 640 columns with it (CadReco/obj_reco_lmicp.cpp:38-45 ``TImage2Mat`` with
``interpolation = true`` = INTER_LINEAR, called at :255-256).  The arithmetic lives in OpenCV, an un-vendored dependency whose version
the reference does not pin (CMakeLists.txt:13-16); this file restates OpenCV's own implementation (modules/imgproc/src/resize.cpp:
the xofs / alpha / yofs / beta table loops, ``HResizeLinear``, ``VResizeLinear`` with ``FixedPtCast<int, uchar, 22>`` for 8U and
``Cast<float, ushort>`` for 16U, and the ``INTER_LINEAR -> INTER_AREA`` switch for an exact 2x decimation) in numpy.

PARITY STATUS: parity unpinned by the reference itself (it holds no tests or vectors for this step); the anchor is
OpenCV.  Pinned by tests/test_resize_oracle.py against ``cv2.resize`` 4.13 of this image with IPP switched off (``cv2.ipp.setUseIPP(False)``):
bit-exact for both types over down-scales, up-scales, odd sizes and the 2x case.  With IPP on, OpenCV hands 16U linear resizing to
Intel's closed routine, whose results differ from OpenCV's own code by up to 3 counts on ~25 % of the pixels; the oracle (and the
CUDA kernel) follow OpenCV's code.  Only tests/, smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np


def axis_tables(ssize: int, dsize: int, clamp_weights: bool):
    """Index of the first tap and the (1 - f, f) weights per destination coordinate.  Columns clamp index and weight at the image
    border; rows keep their weight and have only the index clipped when the row is fetched."""
    scale = 1.0 / (dsize / ssize)                                   # double, like inv_scale_x -> scale_x
    ofs = np.zeros(dsize, np.int32)
    w = np.zeros((dsize, 2), np.float32)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if clamp_weights and s < 0:
            f, s = np.float32(0), 0
        if clamp_weights and s >= ssize - 1:
            f, s = np.float32(0), ssize - 1
        ofs[d] = s
        w[d] = (np.float32(1.0) - f, f)
    return ofs, w


def _half(a: np.ndarray, dw: int, dh: int) -> np.ndarray:
    x = a.astype(np.int64)
    return ((x[0::2, 0::2] + x[0::2, 1::2] + x[1::2, 0::2] + x[1::2, 1::2] + 2) >> 2).astype(a.dtype)   # ResizeAreaFast, scale 2


def resize_linear(a: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """``a``: H x W x 3 uint8 or H x W uint16."""
    H, W = a.shape[:2]
    if (W, H) == (dw, dh):
        return a.copy()
    if W == 2 * dw and H == 2 * dh:
        return _half(a, dw, dh)
    xo, xw = axis_tables(W, dw, True)
    yo, yw = axis_tables(H, dh, False)
    x1 = np.minimum(xo + 1, W - 1)
    y0, y1 = np.clip(yo, 0, H - 1), np.clip(yo + 1, 0, H - 1)
    if a.dtype == np.uint8:
        ia = np.rint(xw * np.float32(2048)).astype(np.int32)        # saturate_cast<short>(w * INTER_RESIZE_COEF_SCALE)
        ib = np.rint(yw * np.float32(2048)).astype(np.int32)
        s = a.astype(np.int32)
        rows = s[:, xo] * ia[:, 0][None, :, None] + s[:, x1] * ia[:, 1][None, :, None]
        b0, b1 = ib[:, 0][:, None, None], ib[:, 1][:, None, None]
        out = (((b0 * (rows[y0] >> 4)) >> 16) + ((b1 * (rows[y1] >> 4)) >> 16) + 2) >> 2
        return out.astype(np.uint8)
    if a.dtype == np.uint16:
        s = a.astype(np.float32)
        rows = s[:, xo] * xw[:, 0][None, :] + s[:, x1] * xw[:, 1][None, :]      # fp32 multiply, fp32 add, no contraction
        out = rows[y0] * yw[:, 0][:, None] + rows[y1] * yw[:, 1][:, None]
        return np.clip(np.rint(out), 0, 65535).astype(np.uint16)
    raise TypeError("resize_linear: uint8 (3 channels) or uint16 (1 channel)")
