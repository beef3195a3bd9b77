"""TEST INFRASTRUCTURE - ctypes binding of oracle/_build/libfl_oracle.so (the C restatement of the
reference's CPU path).  Imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
reference arm; never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfl_oracle.so")


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("fl_oracle.c", "fl_oracle.h", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class Match(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("similarity", C.c_float),
                ("class_idx", C.c_int32), ("template_id", C.c_int32)]


MATCH_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("similarity", "<f4"), ("class_idx", "<i4"), ("template_id", "<i4")])

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.flo_detector_create.restype = C.c_void_p
        _lib.flo_detector_quantized.restype = C.c_void_p
        _lib.flo_detector_spread.restype = C.c_void_p
        _lib.flo_detector_lm.restype = C.c_void_p
        _lib.flo_icp_cloud_to_cloud_ex.restype = C.c_float
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def normal_lut():
    out = np.zeros(8000, np.uint8)
    lib().flo_normal_lut(_p(out))
    return out


def similarity_lut():
    out = np.zeros(256, np.uint8)
    lib().flo_similarity_lut(_p(out))
    return out


def gaussian7_bgr(bgr):
    H, W = bgr.shape[:2]
    out = np.empty_like(bgr)
    lib().flo_gaussian7_bgr(_p(np.ascontiguousarray(bgr)), W, H, _p(out))
    return out


def sobel3_bgr(bgr):
    H, W = bgr.shape[:2]
    dx = np.empty((H, W, 3), np.int16)
    dy = np.empty((H, W, 3), np.int16)
    lib().flo_sobel3_bgr(_p(np.ascontiguousarray(bgr)), W, H, _p(dx), _p(dy))
    return dx, dy


def phase_q16(dx, dy):
    dx = np.ascontiguousarray(dx, np.float32)
    dy = np.ascontiguousarray(dy, np.float32)
    q = np.empty(dx.size, np.uint8)
    lib().flo_phase_q16(_p(dx), _p(dy), dx.size, _p(q))
    return q.reshape(dx.shape)


def color_quantize(bgr, weak_thr=10.0, want_mag=False):
    H, W = bgr.shape[:2]
    q = np.empty((H, W), np.uint8)
    mag = np.empty((H, W), np.float32) if want_mag else None
    lib().flo_color_quantize(_p(np.ascontiguousarray(bgr)), W, H, C.c_float(weak_thr), _p(q), _p(mag) if want_mag else None)
    return (q, mag) if want_mag else q


def pyrdown_bgr(bgr):
    H, W = bgr.shape[:2]
    out = np.empty((H // 2, W // 2, 3), np.uint8)
    lib().flo_pyrdown_bgr(_p(np.ascontiguousarray(bgr)), W, H, _p(out))
    return out


def resize_nn_half(img):
    H, W = img.shape
    out = np.empty((H // 2, W // 2), np.uint8)
    lib().flo_resize_nn_half_u8(_p(np.ascontiguousarray(img)), W, H, _p(out))
    return out


def median5(img):
    H, W = img.shape
    out = np.empty((H, W), np.uint8)
    lib().flo_median5_u8(_p(np.ascontiguousarray(img)), W, H, _p(out))
    return out


def depth_quantize(depth, dist_thr=2000, diff_thr=50):
    H, W = depth.shape
    out = np.empty((H, W), np.uint8)
    lib().flo_depth_quantize(_p(np.ascontiguousarray(depth)), W, H, dist_thr, diff_thr, _p(out))
    return out


def spread(q, T):
    H, W = q.shape
    out = np.empty((H, W), np.uint8)
    lib().flo_spread(_p(np.ascontiguousarray(q)), W, H, T, _p(out))
    return out


def response_maps(sp):
    out = np.empty((8,) + sp.shape, np.uint8)
    lib().flo_response_maps(_p(np.ascontiguousarray(sp)), sp.size, _p(out))
    return out


def linearize(resp, T):
    H, W = resp.shape
    out = np.empty((T * T, (W // T) * (H // T)), np.uint8)
    lib().flo_linearize(_p(np.ascontiguousarray(resp)), W, H, T, _p(out))
    return out


class Detector:
    """flo_detector wrapper: ``process`` = front half of Detector::match, ``match`` = matchClass loop."""

    def __init__(self, T: Sequence[int] = (5, 8), modality_kind: Sequence[int] = (0, 1), weak_threshold=10.0,
                 distance_threshold=2000, difference_threshold=50):
        self.T = list(T)
        self.L = len(T)
        self.M = len(modality_kind)
        Tarr = (C.c_int * self.L)(*T)
        karr = (C.c_int * self.M)(*modality_kind)
        self._h = C.c_void_p(lib().flo_detector_create(self.L, Tarr, self.M, karr, C.c_float(weak_threshold),
                                                       distance_threshold, difference_threshold))
        if not self._h:
            raise ValueError("flo_detector_create failed")
        self.n_templates = 0

    def __del__(self):
        if getattr(self, "_h", None):
            lib().flo_detector_destroy(self._h)
            self._h = None

    def set_templates(self, tset):
        hdr = np.ascontiguousarray(tset.headers, np.int32)
        ft = np.ascontiguousarray(tset.features, np.int32)
        co = np.ascontiguousarray(tset.class_of, np.int32)
        rc = lib().flo_detector_set_templates(self._h, tset.n_templates, _p(hdr), _p(ft), ft.shape[0], _p(co))
        if rc != 0:
            raise ValueError("flo_detector_set_templates rc=%d" % rc)
        self.n_templates = tset.n_templates

    def process(self, bgr, depth, masks: Optional[Sequence[Optional[np.ndarray]]] = None) -> int:
        H, W = depth.shape
        self._keep = (np.ascontiguousarray(bgr), np.ascontiguousarray(depth))
        marr = None
        if masks:
            self._mk = [None if m is None else np.ascontiguousarray(m, np.uint8) for m in masks]
            marr = (C.c_void_p * self.M)(*[None if m is None else m.ctypes.data for m in self._mk])
        return lib().flo_detector_process(self._h, _p(self._keep[0]), _p(self._keep[1]), W, H, marr)

    def quantized(self, l, m):
        W, H = C.c_int(), C.c_int()
        p = lib().flo_detector_quantized(self._h, l, m, C.byref(W), C.byref(H))
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (H.value, W.value)).copy()

    def spread(self, l, m):
        W, H = C.c_int(), C.c_int()
        lib().flo_detector_quantized(self._h, l, m, C.byref(W), C.byref(H))
        p = lib().flo_detector_spread(self._h, l, m)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (H.value, W.value)).copy()

    def lm(self, l, m, label):
        r, c = C.c_int(), C.c_int()
        p = lib().flo_detector_lm(self._h, l, m, label, C.byref(r), C.byref(c))
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (r.value, c.value)).copy()

    def similarity(self, t):
        W, H = C.c_int(), C.c_int()
        lib().flo_detector_quantized(self._h, self.L - 1, 0, C.byref(W), C.byref(H))
        T = self.T[-1]
        out = np.zeros((H.value // T) * (W.value // T), np.uint16)
        lib().flo_detector_similarity(self._h, t, _p(out))
        return out.reshape(H.value // T, W.value // T)

    def match(self, threshold=75.0, class_filter: Optional[Sequence[int]] = None, canonical=True, n_threads=1,
              cap=1 << 20):
        out = np.zeros(cap, MATCH_DTYPE)
        n_total = C.c_int()
        cf = np.ascontiguousarray(class_filter if class_filter is not None else [], np.int32)
        n = lib().flo_detector_match_templates(self._h, C.c_float(threshold), _p(cf), cf.size, int(canonical),
                                               int(n_threads), _p(out), cap, C.byref(n_total))
        return out[:n].copy()


def depth_to_3d_mm(depth, fx, fy, cx, cy):
    H, W = depth.shape
    out = np.empty((H, W, 3), np.float32)
    lib().flo_depth_to_3d_mm(_p(np.ascontiguousarray(depth)), W, H, C.c_float(fx), C.c_float(fy), C.c_float(cx),
                             C.c_float(cy), _p(out))
    return out


def icp_cloud_to_cloud_ex(pts_ref, pts_model, icp_it_thr=10, dist_mean_thr=0.5, dist_diff_thr=0.01, want_trace=False):
    pr = np.ascontiguousarray(pts_ref, np.float32)
    pm = np.ascontiguousarray(pts_model, np.float32)
    R = np.zeros(9, np.float32)
    T = np.zeros(3, np.float32)
    ratio = C.c_float()
    it = C.c_int()
    trace = np.zeros(3 * max(icp_it_thr, 1), np.float32)
    dm = lib().flo_icp_cloud_to_cloud_ex(_p(pr), pr.shape[0], _p(pm), pm.shape[0], _p(R), _p(T), C.byref(ratio), icp_it_thr,
                                         C.c_float(dist_mean_thr), C.c_float(dist_diff_thr), C.byref(it), _p(trace))
    res = dict(dist_mean=np.float32(dm), R=R.reshape(3, 3), T=T, inlier_ratio=np.float32(ratio.value), iterations=it.value)
    if want_trace:
        res["trace"] = trace.reshape(-1, 3)[:it.value]
    return res


def add_template(det: "Detector", bgr, depth, mask=None, num_features=(63, 63, 63, 63), strong_threshold=55.0, extract_threshold=2,
                 feature_cap: int = 4096):
    """flo_add_template = Detector::addTemplate (linemod.cpp:1579-1615) on one view.  Returns (rc, headers[L*M, 7], features[n, 3],
    bbox[4]); rc = -1 when a pyramid level has too few candidates."""
    H, W = depth.shape[:2]
    b = np.ascontiguousarray(bgr, np.uint8); d = np.ascontiguousarray(depth, np.uint16)
    m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
    hdr = np.zeros((det.L * det.M, 7), np.int32)
    ft = np.zeros((feature_cap, 3), np.int32)
    nf = C.c_int(0)
    bb = np.zeros(4, np.int32)
    nfm = (C.c_int * 4)(*[int(v) for v in num_features][:4])
    f = lib().flo_add_template
    f.restype = C.c_int
    rc = f(det._h, C.c_void_p(b.ctypes.data), C.c_void_p(d.ctypes.data), C.c_void_p(m.ctypes.data) if m is not None else None, W, H, nfm,
           C.c_float(strong_threshold), int(extract_threshold), C.c_void_p(hdr.ctypes.data), C.c_void_p(ft.ctypes.data), feature_cap, C.byref(nf),
           C.c_void_p(bb.ctypes.data))
    return rc, hdr, ft[:nf.value].copy(), bb


def erode3(img, iterations=1):
    H, W = img.shape
    src = np.ascontiguousarray(img, np.uint8)
    out = np.zeros((H, W), np.uint8)
    lib().flo_erode3_u8(C.c_void_p(src.ctypes.data), W, H, int(iterations), C.c_void_p(out.ctypes.data))
    return out


def distance_c3(img):
    H, W = img.shape
    src = np.ascontiguousarray(img, np.uint8)
    out = np.zeros((H, W), np.float32)
    lib().flo_distance_c3(C.c_void_p(src.ctypes.data), W, H, C.c_void_p(out.ctypes.data))
    return out


def detection(model_depth, ref_depth, K_ref, rect_model, rect_ref, icp_it_thr=10, dist_mean_thr=0.5, dist_diff_thr=0.01,
              r_match=None, t_match=None, d_match=0.0):
    H, W = ref_depth.shape
    r_match = np.ascontiguousarray(np.eye(3) if r_match is None else r_match, np.float32).reshape(9)
    t_match = np.ascontiguousarray(np.zeros(3) if t_match is None else t_match, np.float32)
    K = np.ascontiguousarray(K_ref, np.float32)
    rm = np.ascontiguousarray(rect_model, np.int32)
    rr = np.ascontiguousarray(rect_ref, np.int32)
    Tf = np.zeros(3, np.float32)
    Rf = np.zeros(9, np.float32)
    dm, ratio, it, npnt = C.c_float(), C.c_float(), C.c_int(), C.c_int()
    rc = lib().flo_detection(_p(np.ascontiguousarray(model_depth)), _p(np.ascontiguousarray(ref_depth)), W, H, _p(K), _p(rm),
                             _p(rr), icp_it_thr, C.c_float(dist_mean_thr), C.c_float(dist_diff_thr), _p(r_match), _p(t_match),
                             C.c_float(d_match), _p(Tf), _p(Rf), C.byref(dm), C.byref(ratio), C.byref(it), C.byref(npnt))
    return dict(rc=rc, R=Rf.reshape(3, 3), T=Tf, dist_mean=np.float32(dm.value), inlier_ratio=np.float32(ratio.value),
                iterations=it.value, n_points=npnt.value)


def nms(t3, n_model_pts, icp_dist, th):
    t3 = np.ascontiguousarray(t3, np.float32)
    nm = np.ascontiguousarray(n_model_pts, np.int32)
    dd = np.ascontiguousarray(icp_dist, np.float32)
    out = np.zeros(max(len(nm), 1), np.int32)
    n = lib().flo_nms(_p(t3), _p(nm), _p(dd), len(nm), C.c_float(th), _p(out))
    return out[:n].copy()


def svd3_rot(cov):
    c = np.ascontiguousarray(cov, np.float32).reshape(9)
    R = np.zeros(9, np.float32)
    lib().flo_svd3_rot(_p(c), _p(R))
    return R.reshape(3, 3)
