"""``readLinemod`` / ``writeLinemod``: the reference's single-file template database (YAML or XML through OpenCV's FileStorage).

Mirrors, field for field, what the reference writes and reads (paths relative to /root/reference):

* file layout ........ ``writeLinemod`` / ``readLinemod``            linemod/linemod_if.cpp:36-66
* detector header .... ``Detector::write`` / ``read``                linemod/linemod.cpp:1681-1708
                       (``pyramid_levels``, ``T``, ``modalities`` = list of maps written by ``ColorGradient::write`` :552-558
                       and ``DepthNormal::write`` :869-876)
* one class .......... ``Detector::writeClass`` / ``readClass``      linemod/linemod.cpp:1710-1786
                       (``class_id``, ``modalities`` names, ``pyramid_levels``, ``template_pyramids`` = list of
                       {``template_id``, ``template_pose`` (13 floats, :1617-1634), ``templates`` = list of Template maps})
* one template ....... ``Template::write`` / ``read``                linemod/linemod.cpp:98-129
                       (``width``, ``height``, ``offset_x``, ``offset_y``, ``pyramid_level``, ``features`` = flat list of
                       [x, y, label] triples, ``Feature::write`` :38-41)

The parser is OpenCV's own ``cv2.FileStorage`` (the reference uses the C++ class of the same library), so files written by the
reference load here and files written here load in the reference.  ``read_linemod`` returns the host-side ``Detector`` mirror
(fealess_b200.Detector) with every class registered; the templates go to the device as one packed structure-of-arrays on the
first ``match`` (``Detector._upload`` -> ``fl_upload_templates``).

Deviation kept from the reference on purpose: the reference stores ``template_pose`` in ONE flat list shared by all classes and
indexes it by the per-class template id (wrong for more than one class, SURVEY.md A.6 iv); here poses are kept per class.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

_CG, _DN = "ColorGradient", "DepthNormal"
# defaults of getDefaultLINEMOD (linemod.cpp:1829-1835) / the modality constructors (:515-520, :827-833)
_DEFAULT_PARAMS = {_CG: {"weak_threshold": 10.0, "num_features": 63, "strong_threshold": 55.0},
                   _DN: {"distance_threshold": 2000, "difference_threshold": 50, "num_features": 63, "extract_threshold": 2}}


def _cv2():
    import cv2
    return cv2


def write_linemod(detector, filename: str, modality_params: Dict[str, dict] = None) -> None:
    """``writeLinemod(detector, filename)`` (linemod_if.cpp:49-66).  ``detector``: fealess_b200.Detector."""
    cv2 = _cv2()
    fs = cv2.FileStorage(filename, cv2.FILE_STORAGE_WRITE)
    if not fs.isOpened():
        raise IOError("cannot open %s for writing" % filename)
    params = dict(_DEFAULT_PARAMS)
    params.update(modality_params or {})
    # Detector::write
    fs.write("pyramid_levels", int(detector.pyramidLevels()))
    fs.startWriteStruct("T", cv2.FileNode_SEQ | cv2.FileNode_FLOW)
    for t in detector.T_at_level:
        fs.write("", int(t))
    fs.endWriteStruct()
    fs.startWriteStruct("modalities", cv2.FileNode_SEQ)
    for name in detector.getModalities():
        fs.startWriteStruct("", cv2.FileNode_MAP)
        fs.write("type", name)
        for k, v in params[name].items():
            fs.write(k, float(v) if isinstance(v, float) else int(v))
        fs.endWriteStruct()
    fs.endWriteStruct()
    # classes
    fs.startWriteStruct("classes", cv2.FileNode_SEQ)
    for cid in detector.classIds():
        fs.startWriteStruct("", cv2.FileNode_MAP)
        fs.write("class_id", cid)
        fs.startWriteStruct("modalities", cv2.FileNode_SEQ | cv2.FileNode_FLOW)
        for name in detector.getModalities():
            fs.write("", name)
        fs.endWriteStruct()
        fs.write("pyramid_levels", int(detector.pyramidLevels()))
        fs.startWriteStruct("template_pyramids", cv2.FileNode_SEQ)
        for tid in range(detector.numTemplates(cid)):
            fs.startWriteStruct("", cv2.FileNode_MAP)
            fs.write("template_id", int(tid))
            fs.startWriteStruct("template_pose", cv2.FileNode_SEQ | cv2.FileNode_FLOW)
            for v in np.asarray(detector.getPoseInfo(tid, cid), np.float32).reshape(-1):
                fs.write("", float(v))
            fs.endWriteStruct()
            fs.startWriteStruct("templates", cv2.FileNode_SEQ)
            for (w, h, ox, oy, lvl, feats) in detector.getTemplates(cid, tid):
                fs.startWriteStruct("", cv2.FileNode_MAP)
                fs.write("width", int(w)); fs.write("height", int(h))
                fs.write("offset_x", int(ox)); fs.write("offset_y", int(oy))
                fs.write("pyramid_level", int(lvl))
                fs.startWriteStruct("features", cv2.FileNode_SEQ)
                for x, y, lab in np.asarray(feats, np.int32).reshape(-1, 3):
                    fs.startWriteStruct("", cv2.FileNode_SEQ | cv2.FileNode_FLOW)
                    fs.write("", int(x)); fs.write("", int(y)); fs.write("", int(lab))
                    fs.endWriteStruct()
                fs.endWriteStruct()
                fs.endWriteStruct()
            fs.endWriteStruct()
            fs.endWriteStruct()
        fs.endWriteStruct()
        fs.endWriteStruct()
    fs.endWriteStruct()
    fs.release()


def _seq(node) -> List:
    return [node.at(i) for i in range(node.size())]


def read_linemod(filename: str, max_width: int = 640, max_height: int = 480, device: int = 0, max_candidates: int = 1 << 16):
    """``readLinemod(filename)`` (linemod_if.cpp:36-47): Detector::read on the root, then readClass for every entry of
    ``classes``.  Raises ``ValueError`` where the reference CV_Asserts (modality / pyramid mismatch, duplicate class,
    template ids out of order, > 63 features)."""
    from . import Detector
    cv2 = _cv2()
    fs = cv2.FileStorage(filename, cv2.FILE_STORAGE_READ)
    if not fs.isOpened():
        raise IOError("cannot open %s" % filename)
    root = fs.root()
    levels = int(root.getNode("pyramid_levels").real())
    T = [int(n.real()) for n in _seq(root.getNode("T"))]
    modalities, mod_params = [], {}
    for m in _seq(root.getNode("modalities")):
        name = m.getNode("type").string()
        if name not in (_CG, _DN):
            raise ValueError("unknown modality %r" % name)          # Modality::create returns an empty Ptr (linemod.cpp:208-216)
        modalities.append(name)
        mod_params[name] = {k: m.getNode(k).real() for k in m.keys() if k != "type"}
    if len(T) != levels:
        raise ValueError("T has %d entries for %d pyramid levels" % (len(T), levels))
    det = Detector(modalities, T, max_width=max_width, max_height=max_height, device=device, max_candidates=max_candidates)
    det.modality_params = mod_params
    for c in _seq(root.getNode("classes")):
        names = [n.string() for n in _seq(c.getNode("modalities"))]
        if names != modalities:                                      # CV_Assert(modalities[i]->name() == *mod_it), :1716-1720
            raise ValueError("class modalities %r do not match the detector's %r" % (names, modalities))
        if int(c.getNode("pyramid_levels").real()) != levels:        # :1721
            raise ValueError("class pyramid_levels differs from the detector's")
        cid = c.getNode("class_id").string()
        if cid in det.classIds():                                    # :1728
            raise ValueError("class %r is already registered" % cid)
        for expected, tp in enumerate(_seq(c.getNode("template_pyramids"))):
            if int(tp.getNode("template_id").real()) != expected:    # :1746
                raise ValueError("template ids of class %r are not consecutive" % cid)
            pose_node = tp.getNode("template_pose")
            pose = np.array([n.real() for n in _seq(pose_node)], np.float32) if not pose_node.empty() else np.zeros(13, np.float32)
            pyr = []
            for t in _seq(tp.getNode("templates")):
                feats = np.array([[int(v.real()) for v in _seq(f)] for f in _seq(t.getNode("features"))], np.int32).reshape(-1, 3)
                if len(feats) > 63:                                   # CV_Assert(features.size() <= 63), :1137
                    raise ValueError("template with %d features (> 63)" % len(feats))
                pyr.append((int(t.getNode("width").real()), int(t.getNode("height").real()), int(t.getNode("offset_x").real()),
                            int(t.getNode("offset_y").real()), int(t.getNode("pyramid_level").real()), feats))
            if len(pyr) != levels * len(modalities):
                raise ValueError("template pyramid with %d templates, expected %d" % (len(pyr), levels * len(modalities)))
            det.addSyntheticTemplate(pyr, cid, pose if pose.size == 13 else np.resize(pose, 13))
    fs.release()
    return det
