"""GPU (-m gpu): the CUDA path, through the C ABI, against the REFERENCE'S OWN CODE (oracle/_ref/libfl_ref.so =
/root/reference/linemod/linemod.cpp + ICP/*.cpp compiled unmodified, oracle/build_ref.py) - no restatement in between.
The library is prebuilt in the build container and travels to the GPU box; without it these tests skip (the same comparisons
against the C oracle, which tests/test_oracle_ref.py pins on the reference bit for bit, run in test_gpu_match / test_gpu_icp)."""
import numpy as np
import pytest

import fealess_b200 as fb
import fl_ref_py as R
from fealess_b200 import synth
from helpers import canonical, rot_err

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.available(), reason="oracle/_ref/libfl_ref.so is not in this checkout")]


def _setup(W, H, T, n, n_classes, frame_idx=0, planted=0.04, seed=77):
    b, d = synth.make_frame(W, H, frame_idx)
    ref = R.Detector(T)
    assert ref.process(b, d) == 0
    L = len(T)
    q = [ref.quantized(l, m) for l in range(L) for m in range(2)]
    ts = synth.make_templates(n, W, H, T, n_classes=n_classes, seed=seed, quantized=q, planted_fraction=planted)
    ref.set_templates(ts)
    assert ref.process(b, d) == 0
    h = fb.Handle(T, (0, 1), W, H)
    h.upload_templates(ts)
    h.keep_spread(True)
    return b, d, ts, ref, h


@pytest.mark.parametrize("cfg", [dict(W=640, H=480, T=(5, 8), n=900, classes=3), dict(W=1280, H=720, T=(5, 8), n=500, classes=15),
                                 dict(W=640, H=480, T=(4, 8, 8), n=300, classes=2)])
def test_match_equals_reference_code(cfg):
    W, H, T = cfg["W"], cfg["H"], cfg["T"]
    b, d, ts, ref, h = _setup(W, H, T, cfg["n"], cfg["classes"])
    L = len(T)
    seen = 0
    for thr in (75.0, 60.0):
        rc, got, quant = h.match(b, d, thr, want_quantized=True)
        assert rc == 0
        # the real Detector::match (linemod.cpp:1356-1441): same matches as a set, same best match, same `quantized_images`
        rrc, full = ref.match_full(b, d, thr)
        assert rrc == 0
        assert set(map(tuple, full.tolist())) == set(map(tuple, got.tolist()))
        if len(got):
            assert tuple(full[0])[:4] == tuple(got[0])[:4]
        for i in range(L * 2):
            assert np.array_equal(quant[i], ref.match_quantized(i // 2, i % 2)), "quantized image %d" % i
        # matchClass' pre-sort list, put in the canonical order, is the CUDA list record for record
        assert np.array_equal(canonical(ref.match(thr)), got)
        seen += len(got)
    assert seen > 10
    # every by-product of the front end: spread images and all 8 x M x L linear memories (spread / computeResponseMaps / linearize)
    for l in range(L):
        for m in range(2):
            assert np.array_equal(h.debug_quantized(l, m, W, H, spread=True), ref.spread(l, m))
            for lab in range(8):
                assert np.array_equal(h.debug_lm(l, m, lab, W, H), ref.lm(l, m, lab)), (l, m, lab)
    # similarity + addSimilarities maps of a few templates (linemod.cpp:1130-1214, 1322-1338)
    for t in list(range(0, ts.n_templates, max(ts.n_templates // 12, 1)))[:12]:
        assert np.array_equal(h.debug_similarity(t, W, H), ref.similarity(t)), t
    # class filter
    cf = [cfg["classes"] - 1]
    rc, got = h.match(b, d, 60.0, class_filter=cf)
    assert rc == 0 and np.array_equal(canonical(ref.match(60.0, class_filter=cf)), got)
    h.close()


def test_masks_equal_reference_code():
    W, H, T = 640, 480, (5, 8)
    b, d, ts, ref, h = _setup(W, H, T, 300, 2, frame_idx=2, planted=0.1)
    rng = np.random.default_rng(3)
    m0 = np.zeros((H, W), np.uint8); m0[60:420, 80:560] = 255
    m1 = (rng.random((H, W)) < 0.9).astype(np.uint8) * 255
    assert ref.process(b, d, masks=[m0, m1]) == 0
    rc, got, quant = h.match(b, d, 60.0, masks=[m0, m1], want_quantized=True)
    assert rc == 0
    for i in range(4):
        assert np.array_equal(quant[i], ref.quantized(i // 2, i % 2))
    assert np.array_equal(canonical(ref.match(60.0)), got)
    h.close()


def test_geometry_error_equals_reference_code():
    T = (5, 8)
    b, d = synth.make_frame(640, 488, 1)                     # 244 % 8 != 0 -> CV_Assert in linearize (linemod.cpp:1062-1063)
    ref = R.Detector(T)
    assert ref.match_full(b, d)[0] == -2
    h = fb.Handle(T, (0, 1), 640, 488)
    assert h.match(b, d, 75.0)[0] == fb.FL_ERR_GEOMETRY
    h.close()


def test_icp_equals_reference_code():
    W, H = 640, 480
    h = fb.Handle((5, 8), (0, 1), W, H)
    K = (608.0, 608.0, 320.0, 240.0)
    worst_r = worst_t = 0.0
    iters = []
    for seed in range(6):
        md, rf, rm, rr, p = synth.make_icp_pair(W, H, seed=seed, max_rot_deg=4 + 2 * seed, max_shift_mm=5 + 3 * seed)
        R0, t0 = p[:12].reshape(3, 4)[:, :3], p[:12].reshape(3, 4)[:, 3]
        r = R.detection(md, rf, K, rm, rr, r_match=R0, t_match=t0)          # detection(), detection.cpp:11-254
        assert r["rc"] == 0
        g = h.detection_batch(rf, K, [md], [rm], [rr], [R0], [t0])[0]
        worst_r = max(worst_r, rot_err(g["R"].reshape(3, 3), r["R"]))
        worst_t = max(worst_t, float(np.abs(g["T"] - r["T"]).max()))
        iters.append(int(g["iterations"]))
    assert max(iters) >= 3
    # tolerance of the task: 1e-4 rad, 1e-4 m = 0.1 mm; the kernel reproduces the reference's fp32 operation order, so the
    # poses are expected to be bit-identical
    assert worst_r < 1e-4 and worst_t < 0.1, (worst_r, worst_t)
    assert worst_r == 0.0 and worst_t == 0.0, (worst_r, worst_t)
    # back-projection and NMS
    _, depth = synth.make_frame(W, H, 2)
    mm = h.depth_to_3d(depth, K) * np.float32(1000)                      # depthTo3d (metres) then scale_mat_vec3f(.., 1000)
    want = R.depth_to_3d_mm(depth, *K)
    hole = depth == 0                                                    # NaN points (payload / sign of a NaN is not part of parity)
    assert np.isnan(mm[hole]).all() and np.isnan(want[hole]).all()
    assert np.array_equal(mm[~hole].view(np.uint32), want[~hole].view(np.uint32))
    rng = np.random.default_rng(4)
    for n in (1, 7, 120):
        t3 = rng.uniform(-60, 60, (n, 3)).astype(np.float32)
        nm = rng.integers(50, 12000, n).astype(np.int32)
        dd = rng.uniform(0, 3, n).astype(np.float32)
        for th in (5.0, 30.0):
            assert np.array_equal(h.nms(t3, nm, dd, th), R.nms(t3, nm, dd, th))
    h.close()
