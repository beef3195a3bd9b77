"""CPU: live cross-check of the C oracle against the cv2-based restatement (needs cv2; the committed fixtures cover the
same ground when cv2 is absent)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

import fl_oracle_py as F            # noqa: E402
import oracle_cv2 as O              # noqa: E402
from fealess_b200 import synth      # noqa: E402
from helpers import rot_err         # noqa: E402


def test_phase_bins_exhaustive_over_the_sobel_domain():
    """Every (dx, dy) a 3x3 Sobel on 8-bit data can produce: dx, dy in [-1020, 1020] (4.17 M pairs)."""
    v = np.arange(-1020, 1021, dtype=np.float32)
    bad = 0
    for y0 in range(0, len(v), 256):
        dy, dx = np.meshgrid(v[y0:y0 + 256], v, indexing="ij")
        ref = cv2.convertScaleAbs(cv2.phase(np.ascontiguousarray(dx), np.ascontiguousarray(dy), angleInDegrees=True), alpha=16.0 / 360.0)
        bad += int((F.phase_q16(dx, dy) != ref).sum())
    assert bad == 0


@pytest.mark.parametrize("shape", [(480, 640), (135, 241), (64, 48)])
def test_opencv_primitives(shape):
    H, W = shape
    rng = np.random.default_rng(H * 1000 + W)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    sm = cv2.GaussianBlur(img, (7, 7), 0, 0, borderType=cv2.BORDER_REPLICATE)
    assert np.array_equal(F.gaussian7_bgr(img), sm)
    dx, dy = F.sobel3_bgr(sm)
    assert np.array_equal(dx, cv2.Sobel(sm, cv2.CV_16S, 1, 0, ksize=3, borderType=cv2.BORDER_REPLICATE))
    assert np.array_equal(dy, cv2.Sobel(sm, cv2.CV_16S, 0, 1, ksize=3, borderType=cv2.BORDER_REPLICATE))
    assert np.array_equal(F.pyrdown_bgr(img), O.pyr_down_bgr(img))
    g = rng.integers(0, 256, (H, W), dtype=np.uint8)
    assert np.array_equal(F.median5(g), cv2.medianBlur(g, 5))
    assert np.array_equal(F.resize_nn_half(g), O.resize_nn(g))
    onehot = (1 << rng.integers(0, 9, (H, W))).astype(np.uint16)
    onehot = np.where(onehot == 256, 0, onehot).astype(np.uint8)
    assert np.array_equal(F.median5(onehot), cv2.medianBlur(onehot, 5))


@pytest.mark.parametrize("frame_idx", [2, 3])
def test_full_front_end_and_match(frame_idx):
    b, d = synth.make_frame(640, 480, frame_idx)
    fe = O.FrontEnd(b, d, (5, 8))
    det = F.Detector((5, 8))
    assert det.process(b, d) == 0
    for l in range(2):
        for m in range(2):
            assert np.array_equal(det.quantized(l, m), fe.quantized[l * 2 + m])
            assert np.array_equal(det.spread(l, m), fe.spread[l * 2 + m])
            for lab in range(8):
                assert np.array_equal(det.lm(l, m, lab), fe.lm[(l * 2 + m) * 8 + lab])
    ts = synth.make_templates(50, quantized=fe.quantized, planted_fraction=0.2, seed=frame_idx, n_classes=2)
    det.set_templates(ts)
    raw, fin = O.match(fe, ts, 60.0)
    got = det.match(60.0)
    assert len(got) == len(fin) > 0
    for a, b_ in zip(fin, got):
        assert (a[0], a[1], a[2], a[3], a[4]) == (b_["x"], b_["y"], b_["similarity"], b_["class_idx"], b_["template_id"])


def test_three_levels_and_single_modality():
    b, d = synth.make_frame(640, 480, 5)
    T = (5, 8, 5)                                     # 640x480, 320x240, 160x120
    fe = O.FrontEnd(b, d, T)
    ts = synth.make_templates(30, 640, 480, T, quantized=fe.quantized, planted_fraction=0.3, seed=9, max_size=64, min_size=32)
    det = F.Detector(T)
    det.set_templates(ts)
    assert det.process(b, d) == 0
    raw, fin = O.match(fe, ts, 65.0)
    got = det.match(65.0)
    assert len(got) == len(fin)
    assert all((a[0], a[1], a[2], a[4]) == (g["x"], g["y"], g["similarity"], g["template_id"]) for a, g in zip(fin, got))


def test_svd_rotation_agrees_with_cv2():
    rng = np.random.default_rng(5)
    for _ in range(50):
        A = rng.normal(size=(3, 3)).astype(np.float32) * 1e6 + np.outer(rng.normal(size=3), rng.normal(size=3)).astype(np.float32) * 1e8
        w, u, vt = cv2.SVDecomp(A)
        assert rot_err(F.svd3_rot(A), vt.T @ u.T) < 5e-6


def test_icp_against_cv2_flann_path():
    for seed, (rot, sh) in enumerate([(3, 6), (8, 10), (12, 15)]):
        m, r, rm, rr, p = synth.make_icp_pair(seed=seed, max_rot_deg=rot, max_shift_mm=sh)
        a = O.detection(m, r, (608.0, 608.0, 320.0, 240.0), rm, rr)
        b = F.detection(m, r, (608.0, 608.0, 320.0, 240.0), rm, rr)
        assert a["iterations"] == b["iterations"] and a["n_points"] == b["n_points"]
        assert rot_err(a["R"], b["R"]) < 5e-5 and np.abs(a["T"] - b["T"]).max() < 5e-3
