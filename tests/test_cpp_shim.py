"""The C++ host-side mirror of the reference interface (include/fealess_b200/linemod.hpp, icp.hpp).

CPU: the mirror + its driver compile and link against the C-ABI library with -Wall -Wextra (no OpenCV in this image, so
against the cv_min.hpp stand-in).  GPU (-m gpu): tests/cpp/shim_test.cpp runs Detector::match / detection / depthTo3d /
nonMaximumSuppression the way CadReco's Recognition does and compares with expectations computed here by the CPU oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

import fealess_b200 as fb
from fealess_b200 import build as fbuild
from fealess_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "shim_test.cpp")
OUT_DIR = os.path.join(ROOT, "tests", "cpp", "_build")
EXE = os.path.join(OUT_DIR, "shim_test")


def _compile():
    fbuild.build()
    os.makedirs(OUT_DIR, exist_ok=True)
    libdir = os.path.dirname(fb.library_path())
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++11", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
           fb.library_path(), "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return EXE


def test_cpp_mirror_compiles_and_links():
    exe = _compile()
    r = subprocess.run([exe], capture_output=True, text=True)          # no case file: prints usage, touches no GPU
    assert r.returncode == 2 and "usage" in r.stdout


def test_cpp_headers_are_self_contained():
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    for hdr in ("fealess_b200/linemod.hpp", "fealess_b200/linemod_io.hpp", "fealess_b200/obj_reco.hpp", "fealess_b200/png16.hpp", "fealess_b200_compat/linemod_if.h", "fealess_b200/icp.hpp", "fealess_b200/cv_min.hpp", "fealess_b200.h"):
        r = subprocess.run([cxx, "-std=c++11", "-fsyntax-only", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), "-x", "c++", "-include", hdr, os.devnull],
                           capture_output=True, text=True)
        assert r.returncode == 0, hdr + "\n" + r.stderr
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"), "-x", "c", "-include", "fealess_b200.h", os.devnull],
                       capture_output=True, text=True)                 # the C ABI header is plain C
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle(tmp_path):
    import fl_oracle_py as F
    W, H, T = 640, 480, (5, 8)
    b, d = synth.make_frame(W, H, 2)
    det = F.Detector(T)
    assert det.process(b, d) == 0
    q = [det.quantized(l, m) for l in range(2) for m in range(2)]
    ts = synth.make_templates(400, W, H, T, n_classes=3, seed=33, quantized=q, planted_fraction=0.05)
    det.set_templates(ts)
    thr = 70.0
    exp = det.match(thr)
    assert len(exp) > 0
    md, rf, rm, rr, p = synth.make_icp_pair(W, H, seed=5, max_rot_deg=6, max_shift_mm=8)
    R0 = p[:12].reshape(3, 4)[:, :3].astype(np.float32)
    t0 = p[:12].reshape(3, 4)[:, 3].astype(np.float32)
    o = F.detection(md, rf, (608.0, 608.0, 320.0, 240.0), rm, rr, r_match=R0, t_match=t0)
    rng = np.random.default_rng(7)
    n_obj = 40
    t3 = (rng.uniform(0, 150, (n_obj, 3))).astype(np.float32)
    npts = rng.integers(500, 1500, n_obj).astype(np.int32)
    dist = rng.uniform(0.2, 3.0, n_obj).astype(np.float32)
    keep = np.asarray(F.nms(t3, npts, dist, 30.0), np.int32)
    blob = b"".join([
        struct.pack("<ii", W, H), np.ascontiguousarray(b, np.uint8).tobytes(), np.ascontiguousarray(d, np.uint16).tobytes(),
        struct.pack("<ii", 2, 2), np.asarray(T, np.int32).tobytes(),
        struct.pack("<i", ts.n_templates), np.ascontiguousarray(ts.headers, np.int32).tobytes(),
        struct.pack("<i", len(ts.features)), np.ascontiguousarray(ts.features, np.int32).tobytes(),
        np.ascontiguousarray(ts.class_of, np.int32).tobytes(),
        struct.pack("<fi", thr, len(exp)), np.ascontiguousarray(exp).tobytes(),
        np.ascontiguousarray(md, np.uint16).tobytes(), np.ascontiguousarray(rf, np.uint16).tobytes(),
        np.asarray(rm, np.int32).tobytes(), np.asarray(rr, np.int32).tobytes(), R0.tobytes(), t0.tobytes(),
        np.asarray(o["T"], np.float32).tobytes(), np.asarray(o["R"], np.float32).tobytes(),
        struct.pack("<i", n_obj), t3.tobytes(), npts.tobytes(), dist.tobytes(), struct.pack("<i", len(keep)), keep.tobytes(),
    ])
    case = tmp_path / "case.bin"
    case.write_bytes(blob)
    exe = _compile()
    r = subprocess.run([exe, str(case)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout


def _case_file(tmp_path, n_templates=500, seed=33):
    import fl_oracle_py as F
    W, H, T = 640, 480, (5, 8)
    b, d = synth.make_frame(W, H, 2)
    det = F.Detector(T)
    assert det.process(b, d) == 0
    q = [det.quantized(l, m) for l in range(2) for m in range(2)]
    ts = synth.make_templates(n_templates, W, H, T, n_classes=3, seed=seed, quantized=q, planted_fraction=0.05)
    det.set_templates(ts)
    thr = 70.0
    exp = det.match(thr)
    assert len(exp) > 0
    blob = b"".join([
        struct.pack("<ii", W, H), np.ascontiguousarray(b, np.uint8).tobytes(), np.ascontiguousarray(d, np.uint16).tobytes(),
        struct.pack("<ii", 2, 2), np.asarray(T, np.int32).tobytes(),
        struct.pack("<i", ts.n_templates), np.ascontiguousarray(ts.headers, np.int32).tobytes(),
        struct.pack("<i", len(ts.features)), np.ascontiguousarray(ts.features, np.int32).tobytes(),
        np.ascontiguousarray(ts.class_of, np.int32).tobytes(),
        struct.pack("<fi", thr, len(exp)), np.ascontiguousarray(exp).tobytes()])
    case = tmp_path / "group_case.bin"
    case.write_bytes(blob)
    return case, b, d, ts, exp, thr


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["0,0", "0,0,0", "0,1", "0,1,2,3"])
def test_cpp_detector_over_several_gpus_of_one_process(tmp_path, devices):
    """cup_linemod::Detector::useDevices -> fl_group_*: template-sharded match from C++ with no Python / torch / NCCL in the process
    (CadReco/obj_reco_lmicp.cpp:86-204 must be able to use N GPUs).  "0,0": two handles on one GPU (runs on a one-GPU box); the
    multi-device variants need that many GPUs."""
    import torch
    need = max(int(x) for x in devices.split(",")) + 1
    if torch.cuda.device_count() < need:
        pytest.skip("%d GPU(s) visible, the case needs %d" % (torch.cuda.device_count(), need))
    fbuild.build()
    os.makedirs(OUT_DIR, exist_ok=True)
    exe = os.path.join(OUT_DIR, "group_test")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-std=c++11", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "group_test.cpp"),
                        "-o", exe, fb.library_path(), "-Wl,-rpath," + os.path.dirname(fb.library_path())], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    case = _case_file(tmp_path)[0]
    r = subprocess.run([exe, str(case), devices], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout


@pytest.mark.gpu
def test_cpp_add_template_equals_the_c_abi(tmp_path):
    """cup_linemod::Detector::addTemplate of the C++ mirror (reference signature) returns the pyramid fl_add_template returns through the
    ctypes binding (tests/test_gpu_train.py ties that one to the reference's own addTemplate), -1 and no side effect for a view with too
    few candidates, and the trained template matches its own view."""
    W, H = 640, 480
    b, d = synth.make_frame(W, H, 0)
    yy, xx = np.mgrid[0:H, 0:W]
    mask = ((((xx - 320) / 110.0) ** 2 + ((yy - 240) / 80.0) ** 2) <= 1.0).astype(np.uint8) * 255
    tiny = ((((xx - 320) / 6.0) ** 2 + ((yy - 240) / 5.0) ** 2) <= 1.0).astype(np.uint8) * 255
    h = fb.Handle((5, 8), (0, 1), W, H)
    rc, hdr, ft, bb = h.add_template(b, d, mask)
    assert rc == 0
    h.close()
    case = tmp_path / "train_case.bin"
    with open(case, "wb") as f:
        f.write(struct.pack("<ii", W, H)); f.write(b.tobytes()); f.write(d.tobytes()); f.write(mask.tobytes()); f.write(tiny.tobytes())
        f.write(struct.pack("<i", len(hdr))); f.write(np.ascontiguousarray(hdr, np.int32).tobytes())
        f.write(struct.pack("<i", len(ft))); f.write(np.ascontiguousarray(ft, np.int32).tobytes()); f.write(np.ascontiguousarray(bb, np.int32).tobytes())
    fbuild.build()
    os.makedirs(OUT_DIR, exist_ok=True)
    exe = os.path.join(OUT_DIR, "train_test")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-std=c++11", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "train_test.cpp"),
                        "-o", exe, fb.library_path(), "-Wl,-rpath," + os.path.dirname(fb.library_path())], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe, str(case)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "all checks passed" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_group_binding_matches_oracle(tmp_path):
    """fl_group_* through the ctypes binding, three handles on one GPU; an exchange block that is too small is reported."""
    _, b, d, ts, exp, thr = _case_file(tmp_path, n_templates=700, seed=34)
    g = fb.Group([0, 0, 0], exchange_capacity=1024)
    g.upload_templates(ts)
    for _ in range(3):
        rc, got = g.match(b, d, thr)
        assert rc == 0 and np.array_equal(got, exp)
    rc, got = g.match(b, d, thr, class_filter=[2])
    assert rc == 0 and np.array_equal(got, exp[exp["class_idx"] == 2])
    # one member's single-launch front end gives up on an in-grid dependency: the group resubmits the frame (that member on
    # per-wave launches from then on) and the caller sees the same list
    g.member_option(1, fb.FL_OPT_FE_DEP_TIMEOUT_TEST, 1)
    rc, got = g.match(b, d, thr)
    assert rc == 0 and np.array_equal(got, exp)
    assert g.member_option(1, fb.FL_OPT_FE_FORCED_WAVES) == 1 and g.member_option(0, fb.FL_OPT_FE_FORCED_WAVES) == 0
    rc, got = g.match(b, d, thr)
    assert rc == 0 and np.array_equal(got, exp)
    g.close()
    g = fb.Group([0, 0], exchange_capacity=4)
    g.upload_templates(ts)
    rc, got = g.match(b, d, 55.0)
    assert rc == fb.FL_ERR_CAPACITY
    g.close()
    # eight members and a threshold low enough for more than 1,024 candidates in the union: the exchange kernel hands the merge
    # over to the stand-alone sort (8 lists), whose per-list counts must not read as an overflow
    h = fb.Handle((5, 8), (0, 1), 640, 480, max_candidates=1 << 17)
    h.upload_templates(ts)
    g = fb.Group([0] * 8, exchange_capacity=4096)
    g.upload_templates(ts)
    sizes = []
    for t in (thr, 30.0, 20.0, 12.0):
        rc1, want = h.match(b, d, t, capacity=1 << 16)
        rc, got = g.match(b, d, t)
        assert rc1 == 0 and rc == 0 and np.array_equal(got, want), (t, rc1, rc, len(got), len(want))
        sizes.append(len(want))
    assert max(sizes) > 1024, sizes
    g.close(); h.close()


def test_compat_headers_forward_to_the_mirror():
    """`#include "linemod_if.h"` / "detection.h" / "NMS.h" as CadReco writes them resolve to the mirror."""
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    src = '#include "linemod_if.h"\n#include "detection.h"\n#include "NMS.h"\n#include "ICP.h"\n#include "depth_to_3d.h"\n' \
          'int main() { cv::Ptr<cup_linemod::Detector> d = cup_linemod::getDefaultLINEMOD(); std::vector<obj_data> o; std::vector<PoseResult> p; ' \
          'return d->numClasses() + (int)o.size() + (int)p.size(); }\n'
    r = subprocess.run([cxx, "-std=c++11", "-fsyntax-only", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include", "fealess_b200_compat"),
                        "-I", os.path.join(ROOT, "include"), "-x", "c++", "-"], input=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
