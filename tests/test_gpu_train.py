"""GPU (-m gpu): fl_add_template = Detector::addTemplate against the REFERENCE'S OWN addTemplate
(oracle/_ref: linemod.cpp:1579-1615 compiled unmodified; its erode / distanceTransform stand-ins are pinned on cv2 by
tests/test_oracle_ref.py).  The template pyramid - every feature, the cropped boxes, the bounding box - has to be identical."""
import numpy as np
import pytest

import fealess_b200 as fb
import fl_ref_py as R
from fealess_b200 import synth

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libfl_ref.so is not in this checkout")


@pytest.mark.parametrize("case", [0, 1, 2, 3, 5])
def test_add_template_equals_the_c_oracle(case):
    """Against oracle/fl_oracle.c (flo_add_template, which tests/test_oracle_ref.py shows equal to the reference's addTemplate): runs with
    or without oracle/_ref on the box."""
    import fl_oracle_py as F
    c = CASES[case]
    W, H, T = c["W"], c["H"], c["T"]
    b, d = synth.make_frame(W, H, c["frame"])
    mask = c["mask"](W, H) if c["mask"] else None
    orc, ohdr, oft, obb = F.add_template(F.Detector(T), b, d, mask)
    h = fb.Handle(T, (0, 1), W, H)
    rc, hdr, ft, bb = h.add_template(b, d, mask)
    assert (rc == 0) == (orc == 0), (rc, orc)                    # (case 5, an object on the image border, has too few colour candidates: both say so)
    if orc == 0:
        assert np.array_equal(hdr, ohdr) and np.array_equal(ft, oft) and np.array_equal(bb, obb)
    else:
        assert rc == fb.FL_ERR_TRAIN
    h.close()


def test_add_template_equals_the_committed_fixture():
    """The fixture tests/golden/train_vga.npz holds the reference's own addTemplate results (oracle/make_train_golden.py; pinned by
    tests/test_oracle_ref.py); this comparison needs no oracle/_ref on the GPU box."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_vga.npz"))
    W, H = 640, 480
    h = fb.Handle((5, 8), (0, 1), W, H)
    for i, c in enumerate(g["cases"]):
        b, d = synth.make_frame(W, H, int(c[0]))
        mask = None if c[3] == 0 else _ellipse(W, H, float(c[1]), float(c[2]), float(c[3]), float(c[4]), value=int(c[5]))
        rc, hdr, ft, bb = h.add_template(b, d, mask)
        want_rc = int(g["rc%d" % i])
        assert (rc == 0) == (want_rc >= 0), (i, rc, want_rc)
        if want_rc >= 0:
            assert np.array_equal(hdr, g["hdr%d" % i]) and np.array_equal(ft, g["ft%d" % i]) and np.array_equal(bb, g["bb%d" % i]), i
        else:
            assert rc == fb.FL_ERR_TRAIN
    h.close()


def _ellipse(W, H, cx, cy, a, b, value=255):
    yy, xx = np.mgrid[0:H, 0:W]
    return (((xx - cx) / a) ** 2 + ((yy - cy) / b) ** 2 <= 1.0).astype(np.uint8) * value


CASES = [
    dict(W=640, H=480, T=(5, 8), mask=lambda W, H: _ellipse(W, H, 320, 240, 110, 80), frame=0),
    dict(W=640, H=480, T=(5, 8), mask=lambda W, H: _ellipse(W, H, 300, 250, 120, 90, value=1), frame=0),      # mask values other than 255
    dict(W=640, H=480, T=(5, 8), mask=None, frame=2),                                                        # no mask: the whole view
    dict(W=640, H=480, T=(4, 8, 8), mask=lambda W, H: _ellipse(W, H, 400, 200, 150, 120), frame=3),          # three levels
    dict(W=1280, H=720, T=(5, 8), mask=lambda W, H: _ellipse(W, H, 640, 360, 300, 200), frame=0),
    dict(W=640, H=480, T=(5, 8), mask=lambda W, H: np.pad(np.full((200, 260), 255, np.uint8), ((0, 280), (0, 380))), frame=1),   # object touching the image border
]


@needs_ref
@pytest.mark.parametrize("case", CASES)
def test_add_template_equals_reference(case):
    W, H, T = case["W"], case["H"], case["T"]
    b, d = synth.make_frame(W, H, case["frame"])
    mask = case["mask"](W, H) if case["mask"] else None
    ref = R.Detector(T)
    rrc, rhdr, rft, rbb = R.add_template(ref, b, d, mask)
    h = fb.Handle(T, (0, 1), W, H)
    rc, hdr, ft, bb = h.add_template(b, d, mask)
    assert (rc == 0) == (rrc >= 0), (rc, rrc)
    if rrc >= 0:
        assert np.array_equal(hdr, rhdr), (hdr, rhdr)
        assert np.array_equal(ft, rft)
        assert np.array_equal(bb, rbb)
        assert len(ft) == sum(63 >> l for l in range(len(T))) * 2
    h.close()


@needs_ref
def test_too_few_candidates_and_trained_template_matches_its_own_view():
    """addTemplate's failure mode, and the trained pyramid going straight back into fl_upload_templates."""
    W, H, T = 640, 480, (5, 8)
    b, d = synth.make_frame(W, H, 0)
    h = fb.Handle(T, (0, 1), W, H)
    ref = R.Detector(T)
    tiny = _ellipse(W, H, 320, 240, 6, 5)                                   # too small for 63 features: both return "no template"
    assert R.add_template(ref, b, d, tiny)[0] == -1
    assert h.add_template(b, d, tiny)[0] == fb.FL_ERR_TRAIN
    # a trained template, uploaded as it comes back, matches the view it was cut from at 100 % at its own position
    mask = _ellipse(W, H, 320, 240, 110, 80)
    rc, hdr, ft, bb = h.add_template(b, d, mask)
    assert rc == 0
    ts = synth.TemplateSet(n_levels=2, n_modalities=2, T=T, class_names=["obj"], headers=hdr.copy(), features=ft.copy(), class_of=np.zeros(1, np.int32),
                           pose13=np.zeros((1, 13), np.float32))
    h.upload_templates(ts)
    rc, got = h.match(b, d, 90.0)
    assert rc == 0 and len(got) >= 1
    # (positions live on the T = 5 grid of level 0, so the best match sits within one cell of the template's box corner)
    assert abs(int(got[0]["x"]) - int(bb[0])) <= 5 and abs(int(got[0]["y"]) - int(bb[1])) <= 5 and got[0]["similarity"] >= 90.0
    h.close()


def test_python_mirror_add_template_then_match():
    """cup_linemod::Detector mirror: addTemplate -> match on the same view finds the new class; a failed addTemplate adds nothing."""
    W, H = 640, 480
    b, d = synth.make_frame(W, H, 0)
    det = fb.Detector()
    tid, bb = det.addTemplate([b, d], "obj", _ellipse(W, H, 320, 240, 6, 5))
    assert tid == -1 and det.numTemplates() == 0
    tid, bb = det.addTemplate([b, d], "obj", _ellipse(W, H, 320, 240, 110, 80), pose_info=np.arange(13, dtype=np.float32))
    assert tid == 0 and det.numTemplates("obj") == 1 and bb[0] % 2 == 0 and bb[1] % 2 == 0
    tid2, _ = det.addTemplate([b, d], "obj", _ellipse(W, H, 200, 300, 60, 140))
    assert tid2 == 1
    rc, ms = det.match([b, d], 90.0)
    assert rc == 0 and len(ms) >= 2 and {m.template_id for m in ms[:4]} >= {0, 1} and ms[0].class_id == "obj"
