"""CPU: argument / status behaviour of the CObjRecoLmICP mirror (reference CadReco/obj_reco_lmicp.cpp:67-74, 216-259) that needs no GPU."""
import numpy as np

from fealess_b200 import reco


def test_status_codes_and_unit_conversion(tmp_path):
    r = reco.ObjRecoLmICP()
    assert r.AddObj(str(tmp_path / "nowhere")) == reco.ERROR_OPEN_FILE_FAILED          # no linemod_templates.yml (:71-72)
    K = dict(fx=600.0, fy=600.0, cx=320.0, cy=240.0, width=640, height=480)
    rgb, dep = np.zeros((480, 640, 3), np.uint8), np.zeros((480, 640), np.uint16)
    assert r.Recognition(rgb, dep, K)[0] == reco.ERROR_INVALID_PARAM                   # no object added
    r.m_lm_detector = object()                                                         # the checks below fail before the detector is touched
    assert r.Recognition(rgb, dep[:100], K)[0] == reco.ERROR_INVALID_PARAM             # depth size != rgb size (:223-227)
    assert r.Recognition(rgb, dep, dict(K, width=320))[0] == reco.ERROR_INVALID_PARAM  # intrinsics for another image size
    # convertTo(CV_16UC1, 0.1): fp32 scale, round half to even, saturate
    assert reco.model_depth_to_mm(np.array([[0, 4, 5, 15, 25, 6000, 65535]], np.uint16)).tolist() == [[0, 0, 0, 2, 2, 600, 6554]]
    # PrepareInputData zooms the stored intrinsics to the 640-column processing size (:238-246)
    r2 = reco.ObjRecoLmICP()
    big = np.zeros((960, 1280, 3), np.uint8)
    m_rgb, m_dep = r2._prepare(big, np.zeros((960, 1280), np.uint16), dict(fx=1200.0, fy=1200.0, cx=640.0, cy=480.0, width=1280, height=960))
    # the frame itself stays at its own size on the host: the INTER_LINEAR rescale runs on the device (fl_match_rescaled)
    assert m_rgb.shape == (960, 1280, 3) and m_dep.shape == (960, 1280) and r2.m_cam["fx"] == 600.0 and r2.m_cam["cx"] == 320.0
    assert (r2.m_cam["width"], r2.m_cam["height"]) == (640, 480)


class _FakeHandle:
    """Records which C-ABI entry points Recognition drives (no GPU): the host logic of the Recognition mirror."""
    def __init__(self):
        self.calls = []

    def upload_model_depths(self, images, rects):
        self.calls.append(("upload_model_depths", len(images), [tuple(r) for r in rects], images[0].shape))

    def _results(self, n):
        import fealess_b200 as fb
        out = np.zeros(n, fb.ICP_RESULT_DTYPE)
        out["R"] = np.eye(3, dtype=np.float32).reshape(-1)
        return out

    def detection_batch_resident(self, ref, K, idx, rects_ref, r, t, it, mean, diff, frame_size=None):
        self.calls.append(("resident", ref is None, list(idx), [tuple(x) for x in rects_ref], frame_size))
        return self._results(len(idx))

    def detection_batch(self, ref, K, mds, rms, rrs, r, t, it, mean, diff):
        self.calls.append(("per-call", ref.shape, len(mds)))
        return self._results(len(mds))

    def resize_linear(self, img, W, H):
        self.calls.append(("resize_linear", W, H))
        return np.zeros((H, W), img.dtype)


class _FakeDetector:
    def __init__(self):
        import fealess_b200 as fb
        self._handle = _FakeHandle()
        self.seen = []
        self._matches = [fb.Match(50, 60, 91.0, "obj00", 1), fb.Match(10, 20, 88.0, "obj00", 0)]

    def getModalities(self):
        return ["ColorGradient", "DepthNormal"]

    def classIds(self):
        return ["obj00"]

    def numTemplates(self, cid=None):
        return 2

    def getTemplates(self, cid, tid):
        return [(100 + tid, 80 + tid, 40, 30, 0, np.zeros((0, 3), np.int32))]

    def getPoseInfo(self, tid, cid=None):
        return np.arange(13, dtype=np.float32)

    def match(self, sources, thr):
        self.seen.append(("match", sources[0].shape))
        return 0, self._matches

    def match_rescaled(self, sources, W, H, thr):
        self.seen.append(("match_rescaled", sources[0].shape, W, H))
        return 0, self._matches


def test_recognition_host_logic_takes_the_resident_path():
    from fealess_b200 import reco
    K = dict(fx=608.0, fy=608.0, cx=320.0, cy=240.0, width=640, height=480)
    md = np.full((480, 640), 700, np.uint16)
    for (W, H) in ((640, 480), (1280, 960)):
        det = _FakeDetector()
        r = reco.ObjRecoLmICP()
        r.add_detector(det, {(None, 0): md, (None, 1): md})
        rgb, dep = np.zeros((H, W, 3), np.uint8), np.zeros((H, W), np.uint16)
        rc, res = r.Recognition(rgb, dep, dict(K, width=W, height=H), top_k=2)
        assert rc == 0 and len(res) == 2 and r.last_icp_path == "resident"
        assert det.seen == ([("match", (480, 640, 3))] if W == 640 else [("match_rescaled", (960, 1280, 3), 640, 480)])
        up, icp = det._handle.calls
        assert up == ("upload_model_depths", 2, [(40, 30, 100, 80), (40, 30, 101, 81)], (480, 640))     # template boxes, once
        # hypothesis 0 is matches[0] (template 1): crop index 1, box moved to the match position; reference frame = the one on the device
        assert icp == ("resident", True, [1, 0], [(50, 60, 101, 81), (10, 20, 100, 80)], (640, 480))
        rc, res = r.Recognition(rgb, dep, dict(K, width=W, height=H))
        assert rc == 0 and len(res) == 1 and len(det._handle.calls) == 3 and det._handle.calls[2][0] == "resident"   # no second upload
    # a template without a depth image of the frame's size falls back to the per-call upload (rescaled depth fetched for it)
    det = _FakeDetector()
    r = reco.ObjRecoLmICP()
    r.add_detector(det, {(None, 0): md, (None, 1): np.zeros((240, 320), np.uint16)})
    rc, res = r.Recognition(np.zeros((960, 1280, 3), np.uint8), np.zeros((960, 1280), np.uint16), dict(K, width=1280, height=960), top_k=2)
    assert rc == 0 and r.last_icp_path == "per-call"
    assert [c[0] for c in det._handle.calls] == ["upload_model_depths", "resize_linear", "per-call"] and det._handle.calls[2][1] == (480, 640)
