"""GPU (-m gpu): ICP / back-projection / NMS parity of the CUDA path through the C ABI.
Tolerance from BASELINE.json north_star: 1e-4 rad rotation, 1e-4 m (= 0.1 mm) translation, same iteration count."""
import os

import numpy as np
import pytest

import fealess_b200 as fb
import fl_oracle_py as F
from fealess_b200 import synth
from helpers import rot_err

pytestmark = pytest.mark.gpu

ROT_TOL = 1e-4      # rad
T_TOL = 0.1         # mm  (= 1e-4 m)
K = (608.0, 608.0, 320.0, 240.0)


@pytest.fixture(scope="module")
def h():
    return fb.Handle()


def _rt(p):
    return p[:12].reshape(3, 4)[:, :3].copy(), p[:12].reshape(3, 4)[:, 3].copy()


def test_detection_matches_oracle_and_golden(h, golden_dir):
    z = np.load(os.path.join(golden_dir, "icp_small.npz"))
    Kz = tuple(float(v) for v in z["K"])
    for i in range(int(z["n_cases"])):
        R0, t0 = _rt(z["rt_match_%d" % i])
        rm, rr = z["rects_%d" % i]
        g = h.detection_batch(z["ref_%d" % i], Kz, [z["model_%d" % i]], [rm], [rr], [R0], [t0])[0]
        o = F.detection(z["model_%d" % i], z["ref_%d" % i], Kz, rm, rr, r_match=R0, t_match=t0)
        dm, ratio, it, npts = z["scalars_%d" % i]
        assert g["status"] == 0 and g["n_points"] == int(npts) == o["n_points"]
        assert g["iterations"] == int(it) == o["iterations"], "case %d" % i
        # against the C oracle the CUDA path is designed to be bit-identical (same fp32 chains, same SVD)
        assert rot_err(g["R"], o["R"]) < 1e-6 and np.abs(g["T"] - o["T"]).max() < 1e-3
        # against the cv2-evaluated golden: the task tolerance
        assert rot_err(g["R"], z["R_%d" % i]) < ROT_TOL
        assert np.abs(g["T"] - z["T_%d" % i]).max() < T_TOL
        assert abs(float(g["dist_mean"]) - dm) < 1e-3 and abs(float(g["inlier_ratio"]) - ratio) < 1e-4


def test_batched_hypotheses_equal_single_calls(h):
    cases = [synth.make_icp_pair(seed=s, max_rot_deg=r, max_shift_mm=sh, rect_wh=wh)
             for s, (r, sh, wh) in enumerate([(3, 6, (100, 100)), (8, 10, (80, 120)), (12, 15, (64, 64)), (5, 3, (120, 90)), (2, 2, (50, 70)),
                                              (10, 5, (100, 100)), (15, 20, (90, 90)), (6, 12, (110, 60))])]
    ref = cases[0][1]
    # all hypotheses against ONE reference frame (the batch API's contract): re-use frame 0's reference depth
    mds, rms, rrs, Rs, ts = [], [], [], [], []
    for md, rf, rm, rr, p in cases:
        R0, t0 = _rt(p)
        mds.append(md); rms.append(rm); rrs.append(rr); Rs.append(R0); ts.append(t0)
    batch = h.detection_batch(ref, K, mds, rms, rrs, Rs, ts)
    for i in range(len(cases)):
        o = F.detection(mds[i], ref, K, rms[i], rrs[i], r_match=Rs[i], t_match=ts[i])
        g = batch[i]
        assert g["iterations"] == o["iterations"] and g["n_points"] == o["n_points"]
        assert rot_err(g["R"], o["R"]) < ROT_TOL and np.abs(g["T"] - o["T"]).max() < T_TOL
        assert abs(float(g["dist_mean"]) - float(o["dist_mean"])) < 1e-3 or (np.isnan(g["dist_mean"]) and np.isnan(o["dist_mean"]))
        single = h.detection_batch(ref, K, [mds[i]], [rms[i]], [rrs[i]], [Rs[i]], [ts[i]])[0]
        assert single["R"].tobytes() == g["R"].tobytes() and single["T"].tobytes() == g["T"].tobytes()   # batching does not change a result


def test_reference_signature_wrappers(h):
    md, rf, rm, rr, p = synth.make_icp_pair(seed=3, max_rot_deg=8, max_shift_mm=10)
    R0, t0 = _rt(p)
    T_final, R_final = fb.detection(md, rf, K, rm, rr, 10, 0.5, 0.01, R0, t0, float(p[12]), handle=h)
    o = F.detection(md, rf, K, rm, rr, r_match=R0, t_match=t0)
    assert rot_err(R_final, o["R"]) < ROT_TOL and np.abs(T_final - o["T"]).max() < T_TOL
    with pytest.raises(fb.FealessError) as e:                       # cv::Mat ROI throw in the reference (detection.cpp:43-44)
        fb.detection(md, rf, K, (600, 10, 100, 100), rr, 10, 0.5, 0.01, R0, t0, handle=h)
    assert e.value.rc == fb.FL_ERR_ROI


def test_cloud_api_matches_oracle(h):
    rng = np.random.default_rng(1)
    for n, ang, noise in [(5000, 0.03, 0.2), (1200, 0.08, 0.5), (20000, 0.02, 0.1)]:
        g = int(np.sqrt(n))
        yy, xx = np.mgrid[0:g, 0:g].astype(np.float32)
        ref = np.stack([xx.ravel() * 1.1 - 50, yy.ravel() * 1.1 - 40, 600 + 15 * np.sin(xx.ravel() / 9) * np.cos(yy.ravel() / 7)], 1).astype(np.float32)
        c, s = np.cos(ang), np.sin(ang)
        Rz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], np.float32)
        mod = ((ref - ref.mean(0)) @ Rz.T + ref.mean(0) + rng.normal(0, noise, ref.shape) + [1.5, -1.0, 0.8]).astype(np.float32)
        o = F.icp_cloud_to_cloud_ex(ref, mod, 10, 0.5, 0.01, want_trace=True)
        dist_mean, R, T, ratio = fb.icpCloudToCloud_Ex(ref, mod, 10, 0.5, 0.01, handle=h)
        r = h.icp_cloud_to_cloud_ex(ref, mod, 10, 0.5, 0.01)
        assert r["iterations"] == o["iterations"] >= 1
        assert rot_err(R, o["R"]) < ROT_TOL and np.abs(T - o["T"]).max() < T_TOL
        assert abs(dist_mean - float(o["dist_mean"])) < 1e-3 and abs(ratio - float(o["inlier_ratio"])) < 1e-4


def test_cloud_api_edge_cases(h):
    r = h.icp_cloud_to_cloud_ex(np.zeros((2, 3), np.float32), np.zeros((2, 3), np.float32))
    assert r["dist_mean"] == -1 and r["iterations"] == 0 and not r["R"].any()          # ICP.cpp:633-638
    with pytest.raises(fb.FealessError):                                                 # lock-step loops need n_ref >= n_model
        h.icp_cloud_to_cloud_ex(np.zeros((5, 3), np.float32), np.zeros((9, 3), np.float32))
    # points beyond the 900 mm validity limit are ignored by the paired statistics but still queried (ICP.cpp:193-279)
    rng = np.random.default_rng(2)
    ref = rng.uniform(-40, 40, (3000, 3)).astype(np.float32) + [0, 0, 850]
    mod = (ref + rng.normal(0, 0.4, ref.shape) + [0.8, 0.5, 0.3]).astype(np.float32)
    o = F.icp_cloud_to_cloud_ex(ref, mod, 10, 0.5, 0.01)
    r = h.icp_cloud_to_cloud_ex(ref, mod, 10, 0.5, 0.01)
    assert r["iterations"] == o["iterations"]
    assert rot_err(r["R"], o["R"]) < ROT_TOL and np.abs(r["T"] - o["T"]).max() < T_TOL
    # identical clouds: dist_mean 0 -> loop never runs, R = I
    r = h.icp_cloud_to_cloud_ex(ref[:100], ref[:100], 10, 0.5, 0.01)
    assert r["iterations"] == 0 and np.array_equal(r["R"].reshape(3, 3), np.eye(3, dtype=np.float32)) and r["dist_mean"] == 0


def test_resident_model_crops_equal_per_call_upload(h):
    """fl_upload_model_depths + fl_detection_batch_resident (template depth crops kept on the device) must give
    byte-identical results to fl_detection_batch, with a host reference frame and with the frame the last match left on the device."""
    cases = [synth.make_icp_pair(seed=20 + s, max_rot_deg=r, max_shift_mm=sh, rect_wh=wh)
             for s, (r, sh, wh) in enumerate([(3, 6, (100, 100)), (8, 10, (80, 120)), (12, 15, (64, 64)), (5, 3, (120, 90)), (2, 2, (50, 70))])]
    ref = cases[0][1]
    mds = [c[0] for c in cases]; rms = [c[2] for c in cases]; rrs = [c[3] for c in cases]
    Rs = [_rt(c[4])[0] for c in cases]; ts = [_rt(c[4])[1] for c in cases]
    want = h.detection_batch(ref, K, mds, rms, rrs, Rs, ts)
    o = F.detection(mds[1], ref, K, rms[1], rrs[1], r_match=Rs[1], t_match=ts[1])
    assert want[1]["iterations"] == o["iterations"] and rot_err(want[1]["R"], o["R"]) < 1e-6
    with pytest.raises(fb.FealessError):                                                 # nothing uploaded yet
        fb.Handle().detection_batch_resident(ref, K, [0], [rrs[0]])
    h.upload_model_depths(mds, rms)
    order = [3, 0, 4, 1, 2, 1]                                                           # any subset, any order, repeats
    got = h.detection_batch_resident(ref, K, order, [rrs[i] for i in order], [Rs[i] for i in order], [ts[i] for i in order])
    assert got.tobytes() == want[order].tobytes()
    # the reference frame of the last host-input match is still on the device: ref_depth = NULL reads it
    b, _ = synth.make_frame(640, 480, 0)
    hm = fb.Handle()
    hm.upload_templates(synth.make_templates(8, 640, 480, (5, 8), seed=1))
    with pytest.raises(fb.FealessError):                                                 # no frame yet on this handle
        hm.upload_model_depths(mds, rms); hm.detection_batch_resident(None, K, [0], [rrs[0]], frame_size=(640, 480))
    rc, _ = hm.match(b, ref, 90.0)
    assert rc == 0
    got = hm.detection_batch_resident(None, K, order, [rrs[i] for i in order], [Rs[i] for i in order], [ts[i] for i in order], frame_size=(640, 480))
    assert got.tobytes() == want[order].tobytes()
    # a reference rect outside the frame is a per-hypothesis ROI error, like the per-call path (detection.cpp:43-44)
    bad = h.detection_batch_resident(ref, K, [0, 1], [(600, 400, 100, 100), rrs[1]], [Rs[0], Rs[1]], [ts[0], ts[1]])
    assert bad[0]["status"] == fb.FL_ERR_ROI and bad[1].tobytes() == want[1].tobytes()
    with pytest.raises(fb.FealessError):                                                 # crop index out of range
        h.detection_batch_resident(ref, K, [len(mds)], [rrs[0]])
    with pytest.raises(fb.FealessError):                                                 # a model rect outside its image is refused at upload
        h.upload_model_depths(mds[:1], [(600, 400, 100, 100)])


def test_depth_to_3d_bit_exact(h):
    _, d = synth.make_frame(640, 480, 2)
    for Kc in (K, (525.3, 531.7, 311.2, 247.9)):
        got = fb.depthTo3d(d, Kc, handle=h)
        want_mm = F.depth_to_3d_mm(d, *Kc)
        # the oracle exports millimetres (depthTo3d followed by scale_mat_vec3f); metres * 1000 must reproduce it bit for bit
        assert np.array_equal(np.isnan(got[..., 2]), d == 0)
        assert np.array_equal((got * np.float32(1000.0))[d > 0], want_mm[d > 0])
    K3 = np.array([[608.0, 0, 320.0], [0, 608.0, 240.0], [0, 0, 1]])
    assert np.array_equal(np.nan_to_num(fb.depthTo3d(d, K3, handle=h)), np.nan_to_num(fb.depthTo3d(d, K, handle=h)))


def test_nms_matches_oracle_and_golden(h, golden_dir):
    z = np.load(os.path.join(golden_dir, "icp_small.npz"))
    assert np.array_equal(h.nms(z["nms_t3"], z["nms_n"], z["nms_dist"], 25.0), z["nms_out_th25"])
    assert np.array_equal(h.nms(z["nms_t3"], z["nms_n"], z["nms_dist"], 8.0), z["nms_out_th8"])
    rng = np.random.default_rng(11)
    for n in (1, 2, 50, 300):
        t3 = rng.uniform(0, 100, (n, 3)).astype(np.float32)
        nm = rng.integers(100, 2000, n).astype(np.int32)
        dd = rng.uniform(0.1, 5, n).astype(np.float32)
        for th in (0.0, 15.0, 1e9):
            assert np.array_equal(h.nms(t3, nm, dd, th), F.nms(t3, nm, dd, th))
    objs = [dict(match_class=i % 3, match_sim=90.0 - i, r=np.eye(3), t=z["nms_t3"][i], pts_model=int(z["nms_n"][i]), icp_dist=float(z["nms_dist"][i])) for i in range(24)]
    res = fb.nonMaximumSuppression(objs, 25.0, handle=h)
    assert [r["index"] for r in res] == list(z["nms_out_th25"])


def test_recognition_mirror_end_to_end(tmp_path):
    """CObjRecoLmICP mirror (obj_reco_lmicp.cpp:67-204): AddObj reads linemod_templates.yml + every depth/<id>.png once, Recognition =
    match at 75 % -> ICP on matches[0] with the template box as model rect and the box at the match position as reference rect ->
    4x4 pose.  Checked against the oracle's match list and the oracle's detection() on the same rects; top_k + NMS wire the batch path."""
    import cv2
    from fealess_b200 import linemod_io, reco
    W, H, T = 640, 480, (5, 8)
    b, d = synth.make_frame(W, H, 0)
    det = F.Detector(T)
    assert det.process(b, d) == 0
    q = [det.quantized(l, m) for l in range(2) for m in range(2)]
    ts = synth.make_templates(24, W, H, T, n_classes=1, seed=61, quantized=q, planted_fraction=0.5)
    det.set_templates(ts)
    want = det.match(75.0)
    assert len(want) > 0
    D0 = fb.Detector()
    D0.add_template_set(ts)
    linemod_io.write_linemod(D0, str(tmp_path / "linemod_templates.yml"))
    (tmp_path / "depth").mkdir()
    for tid in range(ts.n_templates):                                   # rendered model depth in 0.1 mm units, as the reference stores it
        cv2.imwrite(str(tmp_path / "depth" / ("%d.png" % tid)), (d.astype(np.uint32) * 10).clip(0, 65535).astype(np.uint16))
    r = reco.ObjRecoLmICP()
    assert r.AddObj(str(tmp_path)) == 0
    K = dict(fx=608.0, fy=608.0, cx=320.0, cy=240.0, width=W, height=H)
    rc, res = r.Recognition(b, d, K)
    assert rc == 0 and len(res) == 1 and r.last_icp_path == "resident"      # template depth crops on the device, frame from match()
    top = want[0]
    assert res[0]["strObjTag"] == "obj%02d" % top["class_idx"] and res[0]["template_id"] == top["template_id"]
    hdr, _ = ts.template(int(top["template_id"]), 0, 0)
    rect_model = (int(hdr[2]), int(hdr[3]), int(hdr[0]), int(hdr[1]))
    rect_ref = (int(top["x"]), int(top["y"]), int(hdr[0]), int(hdr[1]))
    P = ts.pose13[int(top["template_id"])][:12].reshape(3, 4)
    md = reco.model_depth_to_mm((d.astype(np.uint32) * 10).clip(0, 65535).astype(np.uint16))
    o = F.detection(md, d, (608.0, 608.0, 320.0, 240.0), rect_model, rect_ref, r_match=P[:, :3], t_match=P[:, 3], d_match=float(ts.pose13[int(top["template_id"])][12]))
    if o["rc"] == 0:
        pose = res[0]["tWorld2Cam"]
        assert rot_err(o["R"], pose[:3, :3]) < 1e-4 and float(np.abs(o["T"] - pose[:3, 3]).max()) < 0.1
        assert pose[3].tolist() == [0.0, 0.0, 0.0, 1.0]
    rc, many = r.Recognition(b, d, K, top_k=5, th_obj_dist=1e9)           # everything within th of the first -> one object survives NMS
    assert rc == 0 and len(many) == 1
    rc, many = r.Recognition(b, d, K, top_k=5)
    assert rc == 0 and 1 <= len(many) <= 5
