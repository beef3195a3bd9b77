"""TEST INFRASTRUCTURE - ctypes binding of oracle/_ref/libfl_ref.so: the REFERENCE's own LINE-MOD / ICP / NMS sources
(/root/reference/linemod/linemod.cpp, ICP/*.cpp) compiled unmodified against oracle/ref_shim (recipe: oracle/build_ref.py).
Imported only by tests/, __graft_entry__.smoke() and bench.py's reference arm / cpu_baseline leg; never by the product.

The interface mirrors oracle/fl_oracle_py.py so that a test can run the same call on both and compare."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

import build_ref

MATCH_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("similarity", "<f4"), ("class_idx", "<i4"), ("template_id", "<i4")])
HOOK = C.CFUNCTYPE(None, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int))
OP_GAUSSIAN7, OP_SOBEL_DX, OP_SOBEL_DY, OP_PHASE_DEG, OP_MEDIAN5, OP_PYRDOWN, OP_RESIZE_NN, OP_SVD3, OP_KNN1 = range(9)

_lib = None


def available() -> bool:
    return build_ref.build() is not None


def lib():
    global _lib
    if _lib is None:
        so = build_ref.build()
        if so is None:
            raise RuntimeError("oracle/_ref/libfl_ref.so is neither built nor buildable (no /root/reference)")
        _lib = C.CDLL(so)
        for n in ("flr_detector_create", "flr_detector_quantized", "flr_detector_spread", "flr_detector_lm",
                  "flr_detector_match_quantized"):
            getattr(_lib, n).restype = C.c_void_p
        _lib.flr_detector_last_error.restype = C.c_char_p
        _lib.flr_version.restype = C.c_char_p
        _lib.flr_icp_cloud_to_cloud_ex.restype = C.c_float
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _arr(p, shape, dtype=np.uint8):
    n = int(np.prod(shape))
    ct = {np.uint8: C.c_uint8, np.uint16: C.c_uint16}[dtype]
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), (n,)).reshape(shape).copy()


# ---- stage functions (the reference's file-static functions) ----
def color_quantize(bgr, weak_thr=10.0, want_mag=False):
    H, W = bgr.shape[:2]
    q = np.empty((H, W), np.uint8)
    mag = np.empty((H, W), np.float32) if want_mag else None
    rc = lib().flr_color_quantize(_p(np.ascontiguousarray(bgr)), W, H, C.c_float(weak_thr), _p(q), _p(mag) if want_mag else None)
    assert rc == 0, rc
    return (q, mag) if want_mag else q


def depth_quantize(depth, dist_thr=2000, diff_thr=50):
    H, W = depth.shape
    out = np.empty((H, W), np.uint8)
    rc = lib().flr_depth_quantize(_p(np.ascontiguousarray(depth)), W, H, dist_thr, diff_thr, _p(out))
    assert rc == 0, rc
    return out


def spread(q, T):
    H, W = q.shape
    out = np.empty((H, W), np.uint8)
    rc = lib().flr_spread(_p(np.ascontiguousarray(q)), W, H, T, _p(out))
    assert rc == 0, rc
    return out


def response_maps(sp):
    H, W = sp.shape
    out = np.empty((8, H, W), np.uint8)
    rc = lib().flr_response_maps(_p(np.ascontiguousarray(sp)), W, H, _p(out))
    if rc != 0:
        raise ValueError("computeResponseMaps: CV_Assert (rc=%d)" % rc)
    return out


def linearize(resp, T):
    H, W = resp.shape
    if W % T or H % T:
        out = np.empty((1,), np.uint8)
    else:
        out = np.empty((T * T, (W // T) * (H // T)), np.uint8)
    rc = lib().flr_linearize(_p(np.ascontiguousarray(resp)), W, H, T, _p(out))
    if rc != 0:
        raise ValueError("linearize: CV_Assert (rc=%d)" % rc)
    return out


class Detector:
    """cup_linemod::Detector built through its own constructor + addSyntheticTemplate.  ``match_full`` is the real
    Detector::match; ``process`` / ``match`` expose the by-products and the pre-sort list (Probe in ref_glue_linemod.cpp)."""

    def __init__(self, T: Sequence[int] = (5, 8), modality_kind: Sequence[int] = (0, 1), weak_threshold=10.0,
                 distance_threshold=2000, difference_threshold=50):
        self.T = list(T)
        self.L = len(T)
        self.M = len(modality_kind)
        Tarr = (C.c_int * self.L)(*T)
        karr = (C.c_int * self.M)(*modality_kind)
        self._h = C.c_void_p(lib().flr_detector_create(self.L, Tarr, self.M, karr, C.c_float(weak_threshold), distance_threshold,
                                                       difference_threshold))
        self.n_templates = 0

    def __del__(self):
        if getattr(self, "_h", None):
            lib().flr_detector_destroy(self._h)
            self._h = None

    def last_error(self) -> str:
        return lib().flr_detector_last_error(self._h).decode(errors="replace")

    def set_templates(self, tset):
        hdr = np.ascontiguousarray(tset.headers, np.int32)
        ft = np.ascontiguousarray(tset.features, np.int32)
        co = np.ascontiguousarray(tset.class_of, np.int32)
        rc = lib().flr_detector_set_templates(self._h, tset.n_templates, _p(hdr), _p(ft), ft.shape[0], _p(co))
        if rc != 0:
            raise ValueError("flr_detector_set_templates rc=%d" % rc)
        self.n_templates = tset.n_templates

    def _masks(self, masks):
        if not masks:
            return None
        self._mk = [None if m is None else np.ascontiguousarray(m, np.uint8) for m in masks]
        return (C.c_void_p * self.M)(*[None if m is None else m.ctypes.data for m in self._mk])

    def match_full(self, bgr, depth, threshold=75.0, class_filter: Optional[Sequence[int]] = None, masks=None, cap=1 << 20):
        """The real Detector::match (linemod.cpp:1356-1441): (rc, matches).  rc: 0, -1 (reference's -1), -2 (CV_Assert)."""
        H, W = depth.shape
        b, d = np.ascontiguousarray(bgr), np.ascontiguousarray(depth)
        out = np.zeros(cap, MATCH_DTYPE)
        n_total = C.c_int()
        cf = np.ascontiguousarray(class_filter if class_filter is not None else [], np.int32)
        n = lib().flr_detector_match(self._h, _p(b), _p(d), W, H, self._masks(masks), C.c_float(threshold), _p(cf), cf.size, _p(out),
                                     cap, C.byref(n_total))
        if n < 0:
            return n, out[:0].copy()
        assert n_total.value == n, "cap too small"
        return 0, out[:n].copy()

    def match_quantized(self, l, m):
        W, H = C.c_int(), C.c_int()
        p = lib().flr_detector_match_quantized(self._h, l, m, C.byref(W), C.byref(H))
        return _arr(p, (H.value, W.value))

    def process(self, bgr, depth, masks=None) -> int:
        H, W = depth.shape
        self._keep = (np.ascontiguousarray(bgr), np.ascontiguousarray(depth))
        return lib().flr_detector_process(self._h, _p(self._keep[0]), _p(self._keep[1]), W, H, self._masks(masks))

    def quantized(self, l, m):
        W, H = C.c_int(), C.c_int()
        p = lib().flr_detector_quantized(self._h, l, m, C.byref(W), C.byref(H))
        return _arr(p, (H.value, W.value))

    def spread(self, l, m):
        W, H = C.c_int(), C.c_int()
        lib().flr_detector_quantized(self._h, l, m, C.byref(W), C.byref(H))
        return _arr(lib().flr_detector_spread(self._h, l, m), (H.value, W.value))

    def lm(self, l, m, label):
        r, c = C.c_int(), C.c_int()
        p = lib().flr_detector_lm(self._h, l, m, label, C.byref(r), C.byref(c))
        return _arr(p, (r.value, c.value))

    def similarity(self, t):
        W, H = C.c_int(), C.c_int()
        lib().flr_detector_quantized(self._h, self.L - 1, 0, C.byref(W), C.byref(H))
        T = self.T[-1]
        out = np.zeros((H.value // T) * (W.value // T), np.uint16)
        rc = lib().flr_detector_similarity(self._h, t, _p(out))
        assert rc == 0, rc
        return out.reshape(H.value // T, W.value // T)

    def similarity_local(self, t, level, x, y):
        out = np.zeros(256, np.uint16)
        rc = lib().flr_detector_similarity_local(self._h, t, level, x, y, _p(out))
        assert rc == 0, rc
        return out.reshape(16, 16)

    def match(self, threshold=75.0, class_filter: Optional[Sequence[int]] = None, cap=1 << 20):
        """Detector::matchClass over the processed frame, raw emission order (no sort / unique)."""
        out = np.zeros(cap, MATCH_DTYPE)
        n_total = C.c_int()
        cf = np.ascontiguousarray(class_filter if class_filter is not None else [], np.int32)
        n = lib().flr_detector_match_raw(self._h, C.c_float(threshold), _p(cf), cf.size, _p(out), cap, C.byref(n_total))
        if n < 0:
            raise ValueError("matchClass: rc=%d %s" % (n, self.last_error()))
        assert n_total.value == n, "cap too small"
        return out[:n].copy()


# ---- ICP side ----
def depth_to_3d_mm(depth, fx, fy, cx, cy):
    H, W = depth.shape
    out = np.empty((H, W, 3), np.float32)
    rc = lib().flr_depth_to_3d_mm(_p(np.ascontiguousarray(depth)), W, H, C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), _p(out))
    assert rc == 0, rc
    return out


def pair_points(ref3, mod3, rect_ref, rect_mod):
    H, W = ref3.shape[:2]
    rr = np.ascontiguousarray(rect_ref, np.int32)
    rm = np.ascontiguousarray(rect_mod, np.int32)
    n_max = int(rr[2]) * int(rr[3])
    pr = np.zeros((max(n_max, 1), 3), np.float32)
    pm = np.zeros((max(n_max, 1), 3), np.float32)
    n = lib().flr_pair_points(_p(np.ascontiguousarray(ref3, np.float32)), _p(np.ascontiguousarray(mod3, np.float32)), W, H, _p(rr), _p(rm),
                              _p(pr), _p(pm))
    if n < 0:
        return n, None, None
    return n, pr[:n].copy(), pm[:n].copy()


def icp_cloud_to_cloud_ex(pts_ref, pts_model, icp_it_thr=10, dist_mean_thr=0.5, dist_diff_thr=0.01):
    pr = np.ascontiguousarray(pts_ref, np.float32)
    pm = np.ascontiguousarray(pts_model, np.float32)
    R = np.zeros(9, np.float32)
    T = np.zeros(3, np.float32)
    ratio = C.c_float()
    dm = lib().flr_icp_cloud_to_cloud_ex(_p(pr), pr.shape[0], _p(pm), pm.shape[0], _p(R), _p(T), C.byref(ratio), icp_it_thr,
                                         C.c_float(dist_mean_thr), C.c_float(dist_diff_thr))
    return dict(dist_mean=np.float32(dm), R=R.reshape(3, 3), T=T, inlier_ratio=np.float32(ratio.value))


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
    rc = lib().flr_detection(_p(np.ascontiguousarray(model_depth)), _p(np.ascontiguousarray(ref_depth)), W, H, _p(K), _p(rm), _p(rr),
                             icp_it_thr, C.c_float(dist_mean_thr), C.c_float(dist_diff_thr), _p(r_match), _p(t_match),
                             C.c_float(d_match), _p(Tf), _p(Rf))
    return dict(rc=rc, R=Rf.reshape(3, 3), T=Tf)


def nms(t3, n_model_pts, icp_dist, th):
    t3 = np.ascontiguousarray(t3, np.float32)
    nm = np.ascontiguousarray(n_model_pts, np.int32)
    dd = np.ascontiguousarray(icp_dist, np.float32)
    out = np.zeros(max(len(nm), 1), np.int32)
    n = lib().flr_nms(_p(t3), _p(nm), _p(dd), len(nm), C.c_float(th), _p(out))
    return out[:n].copy()


def svd3_rot(cov):
    c = np.ascontiguousarray(cov, np.float32).reshape(9)
    R = np.zeros(9, np.float32)
    lib().flr_svd3_rot(_p(c), _p(R))
    return R.reshape(3, 3)


# ---- the shim's primitives, one by one (pinned on cv2 by tests/test_oracle_ref.py) ----
def add_template(det: "Detector", bgr, depth, mask=None, feature_cap: int = 4096):
    """The reference's own Detector::addTemplate on one view.  Returns (rc, headers[L*M, 7], features[n, 3], bbox[4]); rc = -1 when a
    pyramid level has too few candidates (the reference's return value)."""
    H, W = depth.shape[:2]
    b = np.ascontiguousarray(bgr, np.uint8); d = np.ascontiguousarray(depth, np.uint16)
    m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
    hdr = np.zeros((det.L * det.M, 7), np.int32)
    ft = np.zeros((feature_cap, 3), np.int32)
    nf = C.c_int(0)
    bb = np.zeros(4, np.int32)
    rc = lib().flr_detector_add_template(det._h, _p(b), _p(d), _p(m) if m is not None else None, W, H, _p(hdr), _p(ft), feature_cap, C.byref(nf), _p(bb))
    return rc, hdr, ft[:nf.value].copy(), bb


def prim_erode3(img, iterations=1):
    H, W = img.shape
    out = np.zeros((H, W), np.uint8)
    lib().flr_prim_erode3(_p(np.ascontiguousarray(img, np.uint8)), W, H, int(iterations), _p(out))
    return out


def prim_distance_c3(img):
    H, W = img.shape
    out = np.zeros((H, W), np.float32)
    lib().flr_prim_distance_c3(_p(np.ascontiguousarray(img, np.uint8)), W, H, _p(out))
    return out


def prim_gaussian7(bgr):
    H, W = bgr.shape[:2]
    out = np.empty_like(bgr)
    lib().flr_prim_gaussian7(_p(np.ascontiguousarray(bgr)), W, H, _p(out))
    return out


def prim_sobel(bgr):
    H, W = bgr.shape[:2]
    dx = np.empty((H, W, 3), np.int16)
    dy = np.empty((H, W, 3), np.int16)
    lib().flr_prim_sobel(_p(np.ascontiguousarray(bgr)), W, H, _p(dx), _p(dy))
    return dx, dy


def prim_phase_q(dx, dy):
    dx = np.ascontiguousarray(dx, np.float32).ravel()
    dy = np.ascontiguousarray(dy, np.float32).ravel()
    ang = np.empty(dx.size, np.float32)
    q = np.empty(dx.size, np.uint8)
    lib().flr_prim_phase_q(_p(dx), _p(dy), dx.size, _p(ang), _p(q))
    return ang, q


def prim_median5(img):
    H, W = img.shape
    out = np.empty((H, W), np.uint8)
    lib().flr_prim_median5(_p(np.ascontiguousarray(img)), W, H, _p(out))
    return out


def prim_pyrdown(bgr):
    H, W = bgr.shape[:2]
    out = np.empty((H // 2, W // 2, 3), np.uint8)
    lib().flr_prim_pyrdown(_p(np.ascontiguousarray(bgr)), W, H, _p(out))
    return out


def prim_resize_nn_half(img):
    H, W = img.shape
    out = np.empty((H // 2, W // 2), np.uint8)
    lib().flr_prim_resize_nn_half(_p(np.ascontiguousarray(img)), W, H, _p(out))
    return out


def prim_knn1(ref, q):
    ref = np.ascontiguousarray(ref, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    idx = np.empty(len(q), np.int32)
    dist = np.empty(len(q), np.float32)
    lib().flr_prim_knn1(_p(ref), len(ref), _p(q), len(q), _p(idx), _p(dist))
    return idx, dist


# ---- run the reference's code on the REAL OpenCV primitives (cv2 through callbacks) ----
_hook_keepalive = []
hook_calls = {}   # op -> number of times the cv2-backed callback ran (tests assert the hooks really fired)


def install_cv2_hooks(ops: Optional[Sequence[int]] = None):
    """Replace the shim's primitives by the real cv2 ones.  IPP is switched off for the duration so that cv2 runs OpenCV's
    own code (IPP's closed routines differ from it in places)."""
    import cv2
    cv2.setUseOptimized(True)
    try:
        cv2.ipp.setUseIPP(False)
    except Exception:
        pass

    def view(p, shape, dtype):
        n = int(np.prod(shape))
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)

    def hook(op, a, b, out, dims):
        hook_calls[op] = hook_calls.get(op, 0) + 1
        if op == OP_GAUSSIAN7:
            H, W, cn = dims[0], dims[1], dims[2]
            view(out, (H, W, cn), np.uint8)[...] = cv2.GaussianBlur(view(a, (H, W, cn), np.uint8), (7, 7), 0, 0, borderType=cv2.BORDER_REPLICATE)
        elif op in (OP_SOBEL_DX, OP_SOBEL_DY):
            H, W, cn = dims[0], dims[1], dims[2]
            dx, dy = (1, 0) if op == OP_SOBEL_DX else (0, 1)
            view(out, (H, W, cn), np.int16)[...] = cv2.Sobel(view(a, (H, W, cn), np.uint8), cv2.CV_16S, dx, dy, ksize=3, scale=1.0, delta=0.0,
                                                             borderType=cv2.BORDER_REPLICATE).reshape(H, W, cn)
        elif op == OP_PHASE_DEG:
            H, W = dims[0], dims[1]
            view(out, (H, W), np.float32)[...] = cv2.phase(view(a, (H, W), np.float32), view(b, (H, W), np.float32), angleInDegrees=True)
        elif op == OP_MEDIAN5:
            H, W = dims[0], dims[1]
            view(out, (H, W), np.uint8)[...] = cv2.medianBlur(view(a, (H, W), np.uint8), 5)
        elif op == OP_PYRDOWN:
            H, W, cn, dh, dw = (dims[i] for i in range(5))
            view(out, (dh, dw, cn), np.uint8)[...] = cv2.pyrDown(view(a, (H, W, cn), np.uint8), dstsize=(dw, dh)).reshape(dh, dw, cn)
        elif op == OP_RESIZE_NN:
            H, W, esz, dh, dw = (dims[i] for i in range(5))
            src = view(a, (H, W, esz), np.uint8)
            view(out, (dh, dw, esz), np.uint8)[...] = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_NEAREST).reshape(dh, dw, esz)
        elif op == OP_SVD3:
            n = dims[0]
            w, u, vt = cv2.SVDecomp(view(a, (n, n), np.float32).copy())
            o = view(out, (n + 2 * n * n,), np.float32)
            o[:n] = w.ravel()
            o[n:n + n * n] = u.ravel()
            o[n + n * n:] = vt.ravel()
        elif op == OP_KNN1:
            n_ref, nq = dims[0], dims[1]
            ref = view(a, (n_ref, 3), np.float32).copy()
            q = view(b, (nq, 3), np.float32).copy()
            index = cv2.flann_Index(ref, dict(algorithm=4, leaf_max_size=15))   # FLANN_INDEX_KDTREE_SINGLE
            idx, dist = index.knnSearch(q, 1, params=dict(checks=32, eps=0.0, sorted=True))
            view(out, (nq,), np.int32)[...] = idx.ravel().astype(np.int32)
            d = (C.c_char * (nq * 4)).from_address(out + nq * 4)
            np.frombuffer(d, dtype=np.float32, count=nq)[...] = dist.ravel().astype(np.float32)

    cb = HOOK(hook)
    _hook_keepalive.append(cb)
    for op in (range(9) if ops is None else ops):
        lib().flr_set_hook(int(op), cb)


def remove_hooks():
    for op in range(9):
        lib().flr_set_hook(op, C.cast(None, HOOK))
