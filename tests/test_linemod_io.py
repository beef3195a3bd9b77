"""CPU: readLinemod / writeLinemod round trip in the reference's file layout (linemod_if.cpp:36-66, linemod.cpp:98-129, 1681-1786)."""
import os

import numpy as np
import pytest

import fealess_b200 as fb
from fealess_b200 import linemod_io, synth


def _detector(n=12, n_classes=3):
    ts = synth.make_templates(n, 640, 480, (5, 8), n_classes=n_classes, seed=3)
    det = fb.Detector()
    det.add_template_set(ts)
    return det, ts


@pytest.mark.parametrize("ext", ["yml", "xml"])
def test_round_trip(tmp_path, ext):
    det, ts = _detector()
    path = str(tmp_path / ("linemod_templates." + ext))
    linemod_io.write_linemod(det, path)
    back = linemod_io.read_linemod(path)
    assert back.getModalities() == det.getModalities() and back.T_at_level == det.T_at_level
    assert back.classIds() == det.classIds() and back.numTemplates() == det.numTemplates() == ts.n_templates
    for cid in det.classIds():
        assert back.numTemplates(cid) == det.numTemplates(cid)
        for tid in range(det.numTemplates(cid)):
            a, b = det.getTemplates(cid, tid), back.getTemplates(cid, tid)
            assert len(a) == len(b) == 4
            for (w, h, ox, oy, lvl, f), (w2, h2, ox2, oy2, lvl2, f2) in zip(a, b):
                assert (w, h, ox, oy, lvl) == (w2, h2, ox2, oy2, lvl2) and np.array_equal(np.asarray(f, np.int32), f2)
            assert np.allclose(det.getPoseInfo(tid, cid), back.getPoseInfo(tid, cid), rtol=0, atol=1e-5)


def test_file_has_the_reference_layout(tmp_path):
    det, _ = _detector(4, 2)
    path = str(tmp_path / "t.yml")
    linemod_io.write_linemod(det, path)
    text = open(path).read()
    for key in ("pyramid_levels: 2", "T: [ 5, 8 ]", "type: ColorGradient", "type: DepthNormal", "weak_threshold", "distance_threshold",
                "classes:", "class_id:", "template_pyramids:", "template_id: 0", "template_pose:", "templates:", "offset_x", "pyramid_level", "features:"):
        assert key in text, key


def test_reader_rejects_what_the_reference_asserts(tmp_path):
    det, _ = _detector(4, 1)
    path = str(tmp_path / "t.yml")
    linemod_io.write_linemod(det, path)
    text = open(path).read()
    bad = str(tmp_path / "bad.yml")
    open(bad, "w").write(text.replace("template_id: 1", "template_id: 7", 1))          # ids must be consecutive (:1746)
    with pytest.raises(ValueError):
        linemod_io.read_linemod(bad)
    open(bad, "w").write(text.replace("pyramid_levels: 2", "pyramid_levels: 3", 1))    # header says 3 levels, T has 2
    with pytest.raises(ValueError):
        linemod_io.read_linemod(bad)
    with pytest.raises(IOError):
        linemod_io.read_linemod(str(tmp_path / "missing.yml"))
