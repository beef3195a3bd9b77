"""CPU: readLinemod / writeLinemod of the C++ mirror (include/fealess_b200/linemod_io.hpp) against OpenCV's own FileStorage.

The reference parses its template database with cv::FileStorage (linemod_if.cpp:36-66, linemod.cpp:98-129, 1681-1786); the C++
mirror carries its own reader / writer for that YAML dialect.  Here files written by cv2.FileStorage are read by the C++ reader
and files written by the C++ writer are read back by cv2.FileStorage (through fealess_b200.linemod_io, which is pinned on cv2)."""
import os
import subprocess

import numpy as np
import pytest

import fealess_b200 as fb
from fealess_b200 import build as fbuild
from fealess_b200 import linemod_io, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "io_test.cpp")
OUT_DIR = os.path.join(ROOT, "tests", "cpp", "_build")
EXE = os.path.join(OUT_DIR, "io_test")


@pytest.fixture(scope="module")
def exe():
    fbuild.build()
    os.makedirs(OUT_DIR, exist_ok=True)
    libdir = os.path.dirname(fb.library_path())
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++11", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
           fb.library_path(), "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return EXE


def _detector(n, n_classes, seed=3, class_names=None):
    ts = synth.make_templates(n, 640, 480, (5, 8), n_classes=n_classes, seed=seed)
    if class_names:
        ts.class_names = list(class_names)
    det = fb.Detector()
    det.add_template_set(ts)
    return det


def _dump_of(det):
    """The text io_test prints for a detector, computed from the Python mirror (classes in sorted order = std::map order; poses in
    FILE order, which write_linemod emits in classIds() order)."""
    out = ["levels %d" % det.pyramidLevels(), "T " + " ".join(str(t) for t in det.T_at_level)]
    for name in det.getModalities():
        out.append("modality ColorGradient 10 63 55" if name == "ColorGradient" else "modality DepthNormal 2000 50 63 2")
    n_all = det.numTemplates()
    out.append("classes %d templates %d poses %d" % (det.numClasses(), n_all, n_all))
    for cid in sorted(det.classIds()):
        out.append("class %s %d" % (cid, det.numTemplates(cid)))
        for tid in range(det.numTemplates(cid)):
            for j, (w, h, ox, oy, lvl, feats) in enumerate(det.getTemplates(cid, tid)):
                f = np.asarray(feats, np.int32).reshape(-1)
                out.append(("t %d %d %d %d %d %d %d %d " % (tid, j, w, h, ox, oy, lvl, len(f) // 3) + " ".join(str(int(v)) for v in f)).rstrip())
    i = 0
    for cid in det.classIds():
        for tid in range(det.numTemplates(cid)):
            out.append("pose %d " % i + " ".join("%.9g" % v for v in np.asarray(det.getPoseInfo(tid, cid), np.float32).reshape(-1)))
            i += 1
    return out


def _run(exe, *args):
    return subprocess.run([exe] + list(args), capture_output=True, text=True)


def test_cpp_reader_agrees_with_filestorage(exe, tmp_path):
    det = _detector(14, 3)
    path = str(tmp_path / "linemod_templates.yml")
    linemod_io.write_linemod(det, path)                              # written by cv2.FileStorage
    r = _run(exe, "dump", path)
    assert r.returncode == 0, r.stdout + r.stderr
    got = [l.rstrip() for l in r.stdout.strip().split("\n")]
    assert got == _dump_of(det)


def test_cpp_writer_is_read_by_filestorage(exe, tmp_path):
    det = _detector(9, 1, seed=5)                                    # one class: the reference's flat pose list is exact (A.6 iv)
    a, b = str(tmp_path / "a.yml"), str(tmp_path / "b.yml")
    linemod_io.write_linemod(det, a)
    r = _run(exe, "rewrite", a, b)
    assert r.returncode == 0, r.stdout + r.stderr
    back = linemod_io.read_linemod(b)                                # parsed by cv2.FileStorage
    assert _dump_of(back) == _dump_of(det)
    text = open(b).read()
    for key in ("%YAML:1.0", "pyramid_levels: 2", "T: [ 5, 8 ]", "type: ColorGradient", "weak_threshold: 10.", "template_id: 0", "features:"):
        assert key in text, key
    # and the C++ reader reads its own output identically
    r2 = _run(exe, "dump", b)
    assert r2.returncode == 0 and [l.rstrip() for l in r2.stdout.strip().split("\n")] == _dump_of(det)


def test_cpp_writer_multi_class_templates(exe, tmp_path):
    det = _detector(10, 3, seed=7)
    a, b = str(tmp_path / "a.yml"), str(tmp_path / "b.yml")
    linemod_io.write_linemod(det, a)
    assert _run(exe, "rewrite", a, b).returncode == 0
    back = linemod_io.read_linemod(b)
    keep = lambda lines: [l for l in lines if not l.startswith("pose ")]   # writeClass writes TemplatePoseInfo[i] for every class (:1776)
    assert keep(_dump_of(back)) == keep(_dump_of(det))


def test_cpp_reader_error_behaviour(exe, tmp_path):
    # unreadable file: an EMPTY detector (AddObj then returns ERROR_OPEN_FILE_FAILED, obj_reco_lmicp.cpp:67-74)
    r = _run(exe, "dump", str(tmp_path / "missing.yml"))
    assert r.returncode == 0 and "classes 0 templates 0" in r.stdout
    det = _detector(4, 1)
    path, bad = str(tmp_path / "t.yml"), str(tmp_path / "bad.yml")
    linemod_io.write_linemod(det, path)
    text = open(path).read()
    open(bad, "w").write(text.replace("template_id: 1", "template_id: 7", 1))          # CV_Assert(template_id == expected_id), :1746
    r = _run(exe, "dump", bad)
    assert r.returncode == 3 and "expected_id" in r.stdout
    open(bad, "w").write(text.replace("modalities: [ ColorGradient, DepthNormal ]", "modalities: [ DepthNormal, ColorGradient ]", 1))   # :1716-1720
    assert _run(exe, "dump", bad).returncode == 3
    open(bad, "w").write(text.replace("      pyramid_levels: 2", "      pyramid_levels: 3", 1))   # the class's own entry, :1721
    assert _run(exe, "dump", bad).returncode == 3


def test_cpp_reader_other_yaml_spellings(exe, tmp_path):
    """Quoted class ids, '- key: value' elements, sequences at their key's indentation, '[:' compact sequences, comments."""
    path = str(tmp_path / "hand.yml")
    open(path, "w").write("""%YAML:1.0
---
# hand-written
pyramid_levels: 1
T: [:4]
modalities:
- type: ColorGradient
  weak_threshold: 12.5
  num_features: 31
  strong_threshold: 5.5e+01
classes:
- class_id: "my obj: 1"
  modalities: [ "ColorGradient" ]
  pyramid_levels: 1
  template_pyramids:
  - template_id: 0
    template_pose: [ 1., 0., 0., 10.5, 0., 1., 0., -2.25,
        0., 0., 1., 3., 700. ]
    templates:
    - width: 20
      height: 12
      offset_x: 4
      offset_y: 6
      pyramid_level: 0
      features:
      - [ 1, 2, 3 ]
      - [:4, 5, 6]
""")
    r = _run(exe, "dump", path)
    assert r.returncode == 0, r.stdout
    assert r.stdout.strip().split("\n") == [
        "levels 1", "T 4", "modality ColorGradient 12.5 31 55", "classes 1 templates 1 poses 1", "class my obj: 1 1",
        "t 0 0 20 12 4 6 0 2 1 2 3 4 5 6", "pose 0 1 0 0 10.5 0 1 0 -2.25 0 0 1 3 700"]
    out = str(tmp_path / "hand_out.yml")
    assert _run(exe, "rewrite", path, out).returncode == 0
    back = linemod_io.read_linemod(out)                              # cv2 parses the quoted class id the writer produced
    assert back.classIds() == ["my obj: 1"] and back.T_at_level == [4]
    assert back.modality_params["ColorGradient"]["weak_threshold"] == 12.5
