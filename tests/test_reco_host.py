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
    assert m_rgb.shape == (480, 640, 3) and m_dep.shape == (480, 640) and r2.m_cam["fx"] == 600.0 and r2.m_cam["cx"] == 320.0
