"""The C++ mirror of the reference's product API for this path: CObjRecoCAD::Create / AddObj / Recognition
(include/fealess_b200/obj_reco.hpp; reference CadReco/obj_reco_temp.h:6-30, obj_reco_lmicp.cpp:47-259).

CPU: the PNG decoder against cv2.imread, AddObj (template file + depth images -> mm) and the argument checks of Recognition.
GPU (-m gpu): Create -> AddObj -> Recognition from C++ on a 640x480 and on a 1280x960 frame, compared bit for bit with the Python
mirror (fealess_b200/reco.py), whose result tests/test_gpu_icp.py and tests/test_gpu_resize.py check against the CPU oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

import fealess_b200 as fb
from fealess_b200 import build as fbuild
from fealess_b200 import linemod_io, reco, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT_DIR = os.path.join(ROOT, "tests", "cpp", "_build")
EXE = os.path.join(OUT_DIR, "reco_test")


@pytest.fixture(scope="module")
def exe():
    fbuild.build()
    os.makedirs(OUT_DIR, exist_ok=True)
    libdir = os.path.dirname(fb.library_path())
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++11", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "reco_test.cpp"),
           "-o", EXE, fb.library_path(), "-lz", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return EXE


def _fnv(a: np.ndarray) -> str:
    h = 1469598103934665603
    for b in np.ascontiguousarray(a).tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def _feature_dir(tmp_path, ts, depth_mm, n_depth=None):
    import cv2
    D0 = fb.Detector()
    D0.add_template_set(ts)
    linemod_io.write_linemod(D0, str(tmp_path / "linemod_templates.yml"))
    (tmp_path / "depth").mkdir()
    png = (depth_mm.astype(np.uint32) * 10).clip(0, 65535).astype(np.uint16)       # the reference stores 0.1 mm units
    for tid in range(ts.n_templates if n_depth is None else n_depth):
        cv2.imwrite(str(tmp_path / "depth" / ("%d.png" % tid)), png)
    return png


def test_png_decoder_equals_imread(exe, tmp_path):
    import cv2
    rng = np.random.default_rng(5)
    smooth = (np.add.outer(np.arange(97), np.arange(131)) * 37 % 65536).astype(np.uint16)      # exercises the Sub / Up / Paeth filters
    for name, img in (("noise16", rng.integers(0, 65536, (53, 71)).astype(np.uint16)), ("smooth16", smooth),
                      ("gray8", rng.integers(0, 256, (40, 33)).astype(np.uint8)), ("frame", synth.make_frame(640, 480, 1)[1])):
        path = str(tmp_path / (name + ".png"))
        assert cv2.imwrite(path, img)
        back = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        r = subprocess.run([exe, "png", path], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout
        w, h, digest = r.stdout.split()
        assert (int(w), int(h)) == (img.shape[1], img.shape[0]) and digest == _fnv(back.astype(np.uint16))
    cv2.imwrite(str(tmp_path / "colour.png"), np.zeros((8, 8, 3), np.uint8))
    assert subprocess.run([exe, "png", str(tmp_path / "colour.png")], capture_output=True, text=True).returncode == 1   # not a grey image
    assert subprocess.run([exe, "png", str(tmp_path / "missing.png")], capture_output=True, text=True).returncode == 1


def test_addobj_and_argument_checks(exe, tmp_path):
    _, d = synth.make_frame(640, 480, 0)
    ts = synth.make_templates(6, 640, 480, (5, 8), n_classes=1, seed=61)
    png = _feature_dir(tmp_path, ts, d, n_depth=4)
    r = subprocess.run([exe, "addobj", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    lines = r.stdout.strip().split("\n")
    assert lines[0] == "status 00000000"
    # convertTo(CV_16UC1, 0.1) like the Python mirror; 4 of the 6 depth images exist
    assert lines[1] == "classes 1 depths 4 640 480 %s" % _fnv(reco.model_depth_to_mm(png))
    r = subprocess.run([exe, "addobj", str(tmp_path / "nowhere")], capture_output=True, text=True)
    assert r.stdout.strip() == "status %08x" % reco.ERROR_OPEN_FILE_FAILED                       # obj_reco_lmicp.cpp:71-72
    r = subprocess.run([exe, "badparams", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    inv = "%08x" % reco.ERROR_INVALID_PARAM
    assert r.stdout.strip().split("\n") == ["no object " + inv, "addobj 00000000", "depth size " + inv, "intrinsics size " + inv, "null image " + inv,
                                            "negative timestamp " + inv, "unsupported type null", "misc 0 0 0"]


def _write_frame(path, bgr, depth, K):
    with open(path, "wb") as f:
        f.write(struct.pack("<ii4d", bgr.shape[1], bgr.shape[0], *K))
        f.write(np.ascontiguousarray(bgr, np.uint8).tobytes())
        f.write(np.ascontiguousarray(depth, np.uint16).tobytes())


def test_recognition_fails_loudly_without_a_gpu(exe, tmp_path):
    """There is no CPU path behind the product API: on a machine without a CUDA device Recognition raises (fl_create -> FL_ERR_CUDA)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    b, d = synth.make_frame(640, 480, 0)
    ts = synth.make_templates(4, 640, 480, (5, 8), n_classes=1, seed=61)
    _feature_dir(tmp_path, ts, d)
    _write_frame(str(tmp_path / "frame.bin"), b, d, (608.0, 608.0, 320.0, 240.0))
    r = subprocess.run([exe, "run", str(tmp_path), str(tmp_path / "frame.bin")], capture_output=True, text=True)
    assert r.returncode == 4 and "addobj 00000000" in r.stdout and "no usable CUDA device" in r.stdout


@pytest.mark.gpu
def test_cpp_recognition_equals_python_mirror(exe, tmp_path):
    import cv2
    import fl_oracle_py as F
    W, H, T = 640, 480, (5, 8)
    b, d = synth.make_frame(W, H, 0)
    det = F.Detector(T)
    assert det.process(b, d) == 0
    q = [det.quantized(l, m) for l in range(2) for m in range(2)]
    ts = synth.make_templates(24, W, H, T, n_classes=1, seed=61, quantized=q, planted_fraction=0.5)
    _feature_dir(tmp_path, ts, d)
    big_b = cv2.resize(b, (1280, 960), interpolation=cv2.INTER_CUBIC)
    big_d = cv2.resize(d, (1280, 960), interpolation=cv2.INTER_NEAREST)
    K = (608.0, 608.0, 320.0, 240.0)
    frames = [(b, d), (big_b, big_d)]
    for i, (fb_, fd_) in enumerate(frames):
        _write_frame(str(tmp_path / ("frame%d.bin" % i)), fb_, fd_, K)
    r = subprocess.run([exe, "run", str(tmp_path), str(tmp_path / "frame0.bin"), str(tmp_path / "frame1.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().split("\n")
    assert lines[0] == "addobj 00000000" and len(lines) == 5
    py = reco.ObjRecoLmICP()
    assert py.AddObj(str(tmp_path)) == 0
    for i, (fb_, fd_) in enumerate(frames):
        rc, res = py.Recognition(fb_, fd_, dict(fx=K[0], fy=K[1], cx=K[2], cy=K[3], width=fb_.shape[1], height=fb_.shape[0]))
        assert rc == 0 and len(res) == 1 and py.last_icp_path == "resident"
        want = "frame %d status 00000000 results 1 %s %s" % (i, res[0]["strObjTag"], " ".join("%08x" % v for v in res[0]["tWorld2Cam"].reshape(-1).view(np.uint32)))
        assert lines[1 + 2 * i] == want and lines[2 + 2 * i] == want


@pytest.mark.gpu
@pytest.mark.parametrize("top_k,per_class,th", [(4, False, 0.0), (1, True, 0.0), (2, True, 25.0), (6, False, 40.0)])
def test_cpp_multi_hypothesis_equals_python_mirror(exe, tmp_path, top_k, per_class, th):
    """top-K / best-per-class selection (test/linemod_acq.cpp:165-184) + nonMaximumSuppression wiring
    (ICP/NMS.cpp:6-39), C++ CObjRecoLmICP::SetHypotheses == the Python mirror, pose for pose and bit for bit."""
    import fl_oracle_py as F
    W, H, T = 640, 480, (5, 8)
    b, d = synth.make_frame(W, H, 0)
    det = F.Detector(T)
    assert det.process(b, d) == 0
    q = [det.quantized(l, m) for l in range(2) for m in range(2)]
    ts = synth.make_templates(36, W, H, T, n_classes=3, seed=62, quantized=q, planted_fraction=0.5)
    _feature_dir(tmp_path, ts, d)
    K = (608.0, 608.0, 320.0, 240.0)
    _write_frame(str(tmp_path / "frame0.bin"), b, d, K)
    r = subprocess.run([exe, "hyp", str(tmp_path), str(top_k), "1" if per_class else "0", str(th), str(tmp_path / "frame0.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().split("\n")
    py = reco.ObjRecoLmICP()
    assert py.AddObj(str(tmp_path)) == 0
    rc, res = py.Recognition(b, d, dict(fx=K[0], fy=K[1], cx=K[2], cy=K[3], width=W, height=H), top_k=top_k, per_class=per_class,
                             th_obj_dist=th if th > 0 else None)
    assert rc == 0 and len(res) >= 1
    if per_class and th == 0:
        tags = [x["strObjTag"] for x in res]
        assert all(tags.count(t) <= top_k for t in set(tags)) and len(set(tags)) > 1
    want = "frame 0 status 00000000 results %d" % len(res) + "".join(
        " %s %s" % (x["strObjTag"], " ".join("%08x" % v for v in x["tWorld2Cam"].reshape(-1).view(np.uint32))) for x in res)
    assert lines[1] == want and lines[2] == want
