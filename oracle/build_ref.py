"""TEST INFRASTRUCTURE - recipe for oracle/_ref/libfl_ref.so: the REFERENCE's own hot-path sources, compiled unmodified.

    python oracle/build_ref.py            # or: make -C oracle ref

What is compiled, from where it lies under /root/reference (nothing is copied into the repository):
    linemod/linemod.cpp   (through oracle/ref_glue_linemod.cpp, which #includes it to reach its file-static functions)
    ICP/ICP.cpp  ICP/NMS.cpp  ICP/common.cpp  ICP/detection.cpp  ICP/depth_to_3d.cpp     (g++ -c on the files directly)
against oracle/ref_shim/ (a stand-in for the OpenCV C++ API: the image has no OpenCV SDK), plus oracle/ref_glue_icp.cpp
(C entry points).  The reference's own build system (cmake + OpenCV + Eigen + librealsense2) is not run.
NOT compiled (out of reach here): CadReco/obj_reco_lmicp.cpp (needs Eigen), linemod/linemod_if.cpp + test/*.cpp (viewers,
RealSense, highgui), kcf_tracker/ (out of scope).

Flags: -O2 -msse4.2 -ffp-contract=off: the SSE2/SSE3/SSSE3 branches of the reference are the ones compiled (CV_SSE2 = 1 ...),
no FMA contraction (a stock x86-64 OpenCV 3.x build has none either).

The outputs go to oracle/_ref/ only (git-ignored; NOT gpurun-ignored, so the built .so travels to the GPU box, where
/root/reference does not exist and this script only checks that the prebuilt library is there)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FEALESS_REFERENCE", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
SO = os.path.join(OUT_DIR, "libfl_ref.so")

REF_SOURCES = ["ICP/ICP.cpp", "ICP/NMS.cpp", "ICP/common.cpp", "ICP/detection.cpp", "ICP/depth_to_3d.cpp"]
OWN_SOURCES = ["ref_glue_linemod.cpp", "ref_glue_icp.cpp", "ref_shim/cvshim_impl.cpp", "ref_shim/ref_filestorage.cpp"]
HEADERS = ["ref_shim/cvshim.hpp"]
CXXFLAGS = ["-O2", "-msse4.2", "-ffp-contract=off", "-fno-math-errno", "-std=c++14", "-fPIC", "-w"]


def reference_present() -> bool:
    return os.path.isfile(os.path.join(REF, "linemod", "linemod.cpp"))


def _inputs():
    files = [os.path.join(HERE, f) for f in OWN_SOURCES + HEADERS] + [os.path.abspath(__file__)]
    if reference_present():
        files += [os.path.join(REF, f) for f in REF_SOURCES]
        files += [os.path.join(REF, "linemod", f) for f in ("linemod.cpp", "linemod.hpp", "normal_lut.i")]
    return files


def build(force: bool = False) -> str | None:
    """Build (if the reference is present and anything changed) and return the path of the library, or None if it can
    neither be built nor found."""
    if not reference_present():
        return SO if os.path.exists(SO) else None
    if not force and os.path.exists(SO) and all(os.path.getmtime(f) <= os.path.getmtime(SO) for f in _inputs()):
        return SO
    os.makedirs(OUT_DIR, exist_ok=True)
    cxx = "/usr/bin/g++" if os.access("/usr/bin/g++", os.X_OK) else "g++"
    inc = ["-I" + os.path.join(HERE, "ref_shim"), "-I" + os.path.join(REF, "linemod"), "-I" + os.path.join(REF, "ICP"),
           "-I" + os.path.join(REF, "CadReco")]
    objs = []
    jobs = [(os.path.join(REF, s), "ref_" + os.path.basename(s)[:-4] + ".o") for s in REF_SOURCES]
    jobs += [(os.path.join(HERE, s), os.path.basename(s)[:-4] + ".o") for s in OWN_SOURCES]
    procs = []
    for src, obj in jobs:
        o = os.path.join(OUT_DIR, obj)
        objs.append(o)
        procs.append((src, subprocess.Popen([cxx] + CXXFLAGS + inc + ["-c", src, "-o", o], stderr=subprocess.PIPE, text=True)))
    for src, p in procs:
        _, err = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("build_ref: %s failed:\n%s" % (src, err[-4000:]))
    subprocess.check_call([cxx, "-shared", "-o", SO] + objs + ["-lm"])
    with open(os.path.join(OUT_DIR, "BUILD_INFO.txt"), "w") as f:
        f.write("libfl_ref.so: reference sources compiled in place from %s\n" % REF)
        for s in ["linemod/linemod.cpp (via ref_glue_linemod.cpp)"] + REF_SOURCES:
            f.write("  " + s + "\n")
        f.write("flags: " + " ".join(CXXFLAGS) + "\n")
    return SO


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p if p else "oracle/_ref: reference absent and no prebuilt library")
    sys.exit(0 if p else 1)
