"""CPU: the C-ABI library loads without a GPU, exports every symbol include/fealess_b200.h declares, fails loudly instead of
falling back, and the host-side mirror behaves like the reference for argument errors."""
import ctypes
import os
import re

import numpy as np
import pytest

import fealess_b200 as fb
from fealess_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "fealess_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fl_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = fb.lib()
    decl = _declared_symbols()
    assert len(decl) >= 25
    for name in decl:
        assert hasattr(L, name), "libfealess_b200.so does not export %s" % name
    assert sorted(fb.EXPORTED_SYMBOLS) == decl
    assert b"sm_100a" in L.fl_version()


def test_struct_layouts_match_the_header():
    assert fb.MATCH_DTYPE.itemsize == 20                      # fl_match_t
    assert fb.ICP_RESULT_DTYPE.itemsize == 9 * 4 + 3 * 4 + 5 * 4
    assert ctypes.sizeof(fb.Params) == 4 * (1 + 8 + 1 + 4 + 1 + 1 + 1 + 2 + 1 + 1)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fb.FealessError) as e:
        fb.Handle()
    assert e.value.rc == fb.FL_ERR_CUDA
    with pytest.raises(fb.FealessError):
        fb.icpCloudToCloud_Ex(np.zeros((10, 3), np.float32), np.zeros((10, 3), np.float32))
    with pytest.raises(fb.FealessError) as e:                 # frames in flight / several GPUs: the same, no device -> no object
        fb.Pipe(2)
    assert e.value.rc == fb.FL_ERR_CUDA
    with pytest.raises(fb.FealessError) as e:
        fb.Group([0, 0])
    assert e.value.rc == fb.FL_ERR_CUDA


def test_product_package_does_not_import_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "fealess_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(root, f), errors="replace").read()
                assert "fl_oracle" not in txt and "oracle_cv2" not in txt and "fl_ref_py" not in txt and "libfl_ref" not in txt, "%s references the oracle" % f


def test_detector_argument_errors_mirror_the_reference():
    det = fb.Detector()
    b, d = synth.make_frame(64, 48, 0)
    # sources.size() != modalities.size() -> return -1 before anything else (linemod.cpp:1364-1367); no GPU needed
    assert det.match([b], 75.0) == (-1, [])
    # masks.size() != modalities.size() -> -1 (linemod.cpp:1373-1377)
    assert det.match([b, d], 75.0, masks=[np.ones((48, 64), np.uint8)]) == (-1, [])
    assert det.pyramidLevels() == 2 and det.getT(0) == 5 and det.getT(1) == 8 and det.numClasses() == 0


def test_detector_template_bookkeeping():
    det = fb.Detector()
    ts = synth.make_templates(6, n_classes=2, seed=3)
    det.add_template_set(ts)
    assert det.numTemplates() == 6 and det.numClasses() == 2 and det.classIds() == ["obj00", "obj01"]
    assert det.numTemplates("obj01") == 3
    tp = det.getTemplates("obj00", 1)
    assert len(tp) == 4 and tp[0][4] == 0 and tp[2][4] == 1 and len(tp[0][5]) == 63 and len(tp[3][5]) == 31
    assert np.allclose(det.getPoseInfo(2, "obj01"), ts.pose13[5])
    tid = det.addSyntheticTemplate(tp, "zzz")
    assert tid == 0 and det.classIds()[-1] == "zzz"


def test_synth_is_deterministic_and_in_range():
    b1, d1 = synth.make_frame(640, 480, 3)
    b2, d2 = synth.make_frame(640, 480, 3)
    assert np.array_equal(b1, b2) and np.array_equal(d1, d2)
    nz = d1[d1 > 0]
    assert nz.min() >= 400 and nz.max() <= 899 and 0.01 < (d1 == 0).mean() < 0.06
    ts = synth.make_templates(50, seed=2)
    for e, h in enumerate(ts.headers):
        f = ts.features[h[5]:h[5] + h[6]]
        assert h[6] == (63 if h[4] == 0 else 31)
        assert f[:, 0].min() >= 0 and f[:, 0].max() <= h[0] and f[:, 1].max() <= h[1] and f[:, 2].max() <= 7
        assert len(np.unique(f[:, 1] * 10000 + f[:, 0])) == len(f)
