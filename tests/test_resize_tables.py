"""CPU: the host half of the device rescale - the cv::resize INTER_LINEAR index / weight tables computed by
fealess_b200/csrc/resize_tables.h - equals the oracle's tables bit for bit (float weights compared by their bit patterns)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import resize_oracle as R

OUT_DIR = os.path.join(ROOT, "tests", "cpp", "_build")
EXE = os.path.join(OUT_DIR, "resize_tables_test")


@pytest.fixture(scope="module")
def exe():
    os.makedirs(OUT_DIR, exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "tests", "cpp", "resize_tables_test.cpp"), "-o", EXE],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


@pytest.mark.parametrize("ssize,dsize", [(1280, 640), (1024, 640), (768, 480), (848, 640), (480, 362), (320, 640), (240, 480), (641, 640), (333, 640), (3, 480), (1920, 640)])
@pytest.mark.parametrize("clamp", [True, False])
def test_tables_equal_oracle(exe, ssize, dsize, clamp):
    r = subprocess.run([exe, str(ssize), str(dsize), "1" if clamp else "0"], capture_output=True, text=True)
    assert r.returncode == 0
    rows = [l.split() for l in r.stdout.strip().split("\n")]
    ofs, w = R.axis_tables(ssize, dsize, clamp)
    assert [int(x[0]) for x in rows] == ofs.tolist()
    bits = w.view(np.uint32)
    assert [int(x[1], 16) for x in rows] == bits[:, 0].tolist() and [int(x[2], 16) for x in rows] == bits[:, 1].tolist()
    iw = np.rint(w * np.float32(2048)).astype(np.int32)
    assert [int(x[3]) for x in rows] == iw[:, 0].tolist() and [int(x[4]) for x in rows] == iw[:, 1].tolist()
