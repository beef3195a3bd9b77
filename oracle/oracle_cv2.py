"""TEST INFRASTRUCTURE - NOT PRODUCT CODE.  Golden-vector generator: FEALESS hot path restated on cv2.

PARITY STATUS: the reference (rlvc/FEALESS) has no tests, golden vectors or fixtures and cannot be
compiled in this image (it needs the OpenCV 3.x C++ SDK, Eigen and librealsense2), so
its results are *unpinned by the reference itself*.  This module is the closest available anchor: it
restates the cited reference lines and evaluates every OpenCV primitive the reference calls with the
REAL OpenCV shipped in this image (cv2 4.13: GaussianBlur, Sobel, phase, convertScaleAbs,
medianBlur, pyrDown, resize, flann_Index(KDTREE_SINGLE), SVDecomp).  ``oracle/make_golden.py`` runs
it to write the fixtures under ``tests/golden/``; the OpenCV-free C restatement ``oracle/fl_oracle.c``
(which travels to the GPU box and is also the timed CPU baseline) is pinned bit-for-bit against
those fixtures for all integer stages, and within 1e-5 for ICP poses.

Only ``tests/`` and ``oracle/make_golden.py`` import this file.  It is written for clarity, not
speed: template loops are plain Python, so use small template counts.

All ``file:line`` citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import cv2
import numpy as np

f32 = np.float32

# --------------------------------------------------------------------------------------------
# the two embedded tables, regenerated from their closed forms (pinned by sha256 in the tests)
# --------------------------------------------------------------------------------------------
NORMAL_LUT_SHA256 = "729e0305a1f5975a88ebb6a9ffc28013c6bf1a2113ea3c112531f44bee7b2243"
GRANULARITY = 20  # linemod/normal_lut.i:2


def normal_lut() -> np.ndarray:
    """NORMAL_LUT[20][20][20] (linemod/normal_lut.i:4).

    Closed form found by fitting the table: the value depends only on (v2, v1) = (ny, nx) bins and is
    the one-hot code of the 45-degree sector of atan2(v2-10, v1-10) + 22.5 deg.  All 8000 bytes
    reproduce the reference table (sha256 above).
    """
    plane = np.zeros((GRANULARITY, GRANULARITY), np.uint8)
    for v2 in range(GRANULARITY):
        for v1 in range(GRANULARITY):
            a = math.degrees(math.atan2(v2 - 10, v1 - 10)) + 22.5
            plane[v2, v1] = 1 << (int(math.floor((a % 360.0) / 45.0)) % 8)
    return np.ascontiguousarray(np.broadcast_to(plane, (GRANULARITY, GRANULARITY, GRANULARITY)))


def similarity_lut() -> np.ndarray:
    """SIMILARITY_LUT[256] (linemod.cpp:970): entry [32*i + 16*half + nib] = best response of
    orientation i against the set bits of the nibble; response 4 / 2 / 1 / 0 for circular label
    distance 0 / 1 / 2 / >=3.  (This is the table the reference compiles, not stock OpenCV's.)"""
    g = (4, 2, 1, 0, 0)
    lut = np.zeros(256, np.uint8)
    for i in range(8):
        for half in range(2):
            for nib in range(16):
                best = 0
                for b in range(4):
                    if nib >> b & 1:
                        j = b + 4 * half
                        d = abs(i - j)
                        best = max(best, g[min(d, 8 - d)])
                lut[32 * i + 16 * half + nib] = best
    return lut


_NORMAL_LUT = normal_lut()
_SIM_LUT = similarity_lut()


# --------------------------------------------------------------------------------------------
# colour-gradient modality   (linemod.cpp:230-459)
# --------------------------------------------------------------------------------------------
def quantized_orientations(bgr: np.ndarray, weak_threshold: float = 10.0) -> Tuple[np.ndarray, np.ndarray]:
    """linemod.cpp:230-305 + hysteresisGradient :307-385.  Returns (quantized u8, magnitude f32)."""
    smoothed = cv2.GaussianBlur(bgr, (7, 7), 0, 0, borderType=cv2.BORDER_REPLICATE)            # :247
    dx3 = cv2.Sobel(smoothed, cv2.CV_16S, 1, 0, ksize=3, scale=1.0, delta=0.0, borderType=cv2.BORDER_REPLICATE)
    dy3 = cv2.Sobel(smoothed, cv2.CV_16S, 0, 1, ksize=3, scale=1.0, delta=0.0, borderType=cv2.BORDER_REPLICATE)
    dx3 = dx3.astype(np.int32)
    dy3 = dy3.astype(np.int32)
    mag3 = dx3 * dx3 + dy3 * dy3                                                               # :271-273
    m0, m1, m2 = mag3[..., 0], mag3[..., 1], mag3[..., 2]
    pick0 = (m0 >= m1) & (m0 >= m2)                                                            # :275
    pick1 = ~pick0 & (m1 >= m0) & (m1 >= m2)                                                   # :281
    ch = np.where(pick0, 0, np.where(pick1, 1, 2))
    ii, jj = np.indices(ch.shape)
    sdx = dx3[ii, jj, ch].astype(f32)
    sdy = dy3[ii, jj, ch].astype(f32)
    mag = mag3[ii, jj, ch].astype(f32)
    angle = cv2.phase(sdx, sdy, angleInDegrees=True)                                           # :303
    return hysteresis_gradient(mag, angle, f32(weak_threshold) * f32(weak_threshold)), mag     # :304


def hysteresis_gradient(mag: np.ndarray, angle: np.ndarray, threshold) -> np.ndarray:
    """linemod.cpp:307-385."""
    H, W = angle.shape
    # angle.convertTo(CV_8U, 16/360): saturate_cast<uchar>(cvRound(v * (float)alpha)); angle >= 0 so
    # convertScaleAbs (same kernel + abs) is the identical OpenCV primitive reachable from Python.
    q = cv2.convertScaleAbs(angle, alpha=16.0 / 360.0)                                         # :314
    q[0, :] = 0                                                                                # :318-325
    q[H - 1, :] = 0
    q[:, 0] = 0
    q[:, W - 1] = 0
    q[1:H - 1, 1:W - 1] &= 7                                                                   # :328-335
    out = np.zeros((H, W), np.uint8)
    # 3x3 histogram over the 8 bins (border values 0 vote for bin 0 exactly as in the reference)
    votes = np.zeros((8, H - 2, W - 2), np.int32)
    for dy in range(3):
        for dx in range(3):
            nb = q[dy:dy + H - 2, dx:dx + W - 2]
            for b in range(8):
                votes[b] += (nb == b)
    best = votes.argmax(axis=0)             # first (lowest) index wins ties, as :369-376
    nbest = votes.max(axis=0)
    ok = (mag[1:H - 1, 1:W - 1] > threshold) & (nbest >= 5)                                    # :346, :380
    out[1:H - 1, 1:W - 1] = np.where(ok, (1 << best).astype(np.uint8), 0)
    return out


def pyr_down_bgr(bgr: np.ndarray) -> np.ndarray:
    """ColorGradientPyramid::pyrDown, linemod.cpp:441-444."""
    h, w = bgr.shape[:2]
    return cv2.pyrDown(bgr, dstsize=(w // 2, h // 2))


def resize_nn(img: np.ndarray) -> np.ndarray:
    """resize(..., INTER_NEAREST) to (cols/2, rows/2): linemod.cpp:448, 731, 736."""
    h, w = img.shape[:2]
    return cv2.resize(img, (w // 2, h // 2), interpolation=cv2.INTER_NEAREST)


# --------------------------------------------------------------------------------------------
# depth-normal modality   (linemod.cpp:567-745)
# --------------------------------------------------------------------------------------------
def quantized_normals(depth: np.ndarray, distance_threshold: int = 2000, difference_threshold: int = 50) -> np.ndarray:
    """linemod.cpp:595-685 (vectorised; integer parts in int64, float parts in separately rounded fp32)."""
    H, W = depth.shape
    r = 5
    dst = np.zeros((H, W), np.uint8)
    if H - r - 1 <= r or W - r - 1 <= r:
        return cv2.medianBlur(dst, 5)
    d = depth.astype(np.int64)
    ys = slice(r, H - r - 1)                                                                   # :619
    xs = slice(r, W - r - 1)                                                                   # :624
    c = d[ys, xs]
    A0 = np.zeros_like(c)
    A1 = np.zeros_like(c)
    A3 = np.zeros_like(c)
    b0 = np.zeros_like(c)
    b1 = np.zeros_like(c)
    for (i, j) in ((-r, -r), (0, -r), (r, -r), (-r, 0), (r, 0), (-r, r), (0, r), (r, r)):        # :633-640
        nb = d[r + j:H - r - 1 + j, r + i:W - r - 1 + i]
        delta = nb - c
        f = (np.abs(delta) < difference_threshold).astype(np.int64)                            # :569
        A0 += f * i * i
        A1 += f * i * j
        A3 += f * j * j
        b0 += f * i * delta
        b1 += f * j * delta
    det = A0 * A3 - A1 * A1                                                                    # :643-645
    ddx = A3 * b0 - A1 * b1
    ddy = -A1 * b0 + A0 * b1
    nx = (617 * ddx).astype(f32)                                                               # :649-651
    ny = (617 * ddy).astype(f32)
    nz = (-det * c).astype(f32)
    s = np.sqrt((nx * nx + ny * ny) + nz * nz).astype(f32)                                     # :653
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = f32(1.0) / s                                                                     # :657
        nxn = nx * inv
        nyn = ny * inv
        nzn = nz * inv
        v1 = (nxn * f32(10) + f32(10))                                                         # :665-667
        v2 = (nyn * f32(10) + f32(10))
        v3 = (nzn * f32(20) + f32(20))
    pos = s > 0
    v1 = np.where(pos, v1, 0).astype(np.int64)           # C truncation; values are >= 0 here
    v2 = np.where(pos, v2, 0).astype(np.int64)
    v3 = np.where(pos, v3, 0).astype(np.int64)
    # NORMAL_LUT[20] is out of bounds in the reference when nz == 0 && s > 0 (SURVEY A.6 i); the build
    # defines that corner by clamping (the table does not depend on v3 anyway).
    lut = _NORMAL_LUT[np.clip(v3, 0, 19), np.clip(v2, 0, 19), np.clip(v1, 0, 19)]
    val = np.where(pos & (c < distance_threshold), lut, 0).astype(np.uint8)                    # :628, :655
    dst[ys, xs] = val
    return cv2.medianBlur(dst, 5)                                                              # :684


# --------------------------------------------------------------------------------------------
# spread / response maps / linear memories   (linemod.cpp:882-1088)
# --------------------------------------------------------------------------------------------
def spread(q: np.ndarray, T: int) -> np.ndarray:
    """linemod.cpp:950-965."""
    H, W = q.shape
    out = np.zeros_like(q)
    for r in range(T):
        for c in range(T):
            out[:H - r, :W - c] |= q[r:, c:]
    return out


def response_maps(sp: np.ndarray) -> np.ndarray:
    """linemod.cpp:979-1048 -> uint8 [8, H, W]."""
    lo = sp & 15
    hi = sp >> 4
    return np.stack([np.maximum(_SIM_LUT[32 * i + lo], _SIM_LUT[32 * i + 16 + hi]) for i in range(8)])


def linearize(resp: np.ndarray, T: int) -> np.ndarray:
    """linemod.cpp:1060-1088 -> uint8 [T*T, (W/T)*(H/T)]."""
    H, W = resp.shape
    assert H % T == 0 and W % T == 0                                                           # :1062-1063
    rows = []
    for r0 in range(T):
        for c0 in range(T):
            rows.append(resp[r0::T, c0::T].reshape(-1))
    return np.stack(rows)


LM_PAD = 4096  # zero bytes appended to each label's linear memory; see flat_lm()


def flat_lm(lm: np.ndarray) -> np.ndarray:
    """The reference addresses a linear memory through a raw pointer (accessLinearMemory,
    linemod.cpp:1094-1117) and reads ``template_positions`` bytes from it (:1191-1212); a feature on
    the template's last row/column (x == width is legal, cropTemplates :79-80) makes that run past
    the end of its T*T row into the next row of the same continuous Mat - defined behaviour that
    parity must keep - and, for the last row, past the buffer (undefined; defined here as zeros)."""
    return np.concatenate([lm.reshape(-1), np.zeros(LM_PAD, np.uint8)])


# --------------------------------------------------------------------------------------------
# Detector::match   (linemod.cpp:1356-1577)
# --------------------------------------------------------------------------------------------
class FrontEnd:
    """Per-frame state up to the linear-memory pyramid (linemod.cpp:1369-1416)."""

    def __init__(self, bgr, depth, T: Sequence[int], masks: Optional[Sequence[Optional[np.ndarray]]] = None,
                 weak_threshold=10.0, distance_threshold=2000, difference_threshold=50,
                 modalities: Sequence[str] = ("ColorGradient", "DepthNormal")):
        self.T = list(T)
        self.quantized: List[np.ndarray] = []   # index l*M + m
        self.spread: List[np.ndarray] = []
        self.lm: List[np.ndarray] = []          # [T*T, cells] per (l*M+m)*8 + label
        self.sizes: List[Tuple[int, int]] = []  # (W, H) per level
        M = len(modalities)
        srcs = {"ColorGradient": bgr, "DepthNormal": depth}
        cur = []
        for m, name in enumerate(modalities):
            mask = None if not masks else masks[m]
            if name == "ColorGradient":
                q, _ = quantized_orientations(srcs[name], weak_threshold)
                cur.append({"kind": name, "src": srcs[name], "q": q, "mask": mask})
            else:
                q = quantized_normals(srcs[name], distance_threshold, difference_threshold)
                cur.append({"kind": name, "q": q, "mask": mask})
        for l, Tl in enumerate(self.T):
            if l > 0:
                for st in cur:
                    if st["kind"] == "ColorGradient":                       # :434-453
                        st["src"] = pyr_down_bgr(st["src"])
                        st["q"], _ = quantized_orientations(st["src"], weak_threshold)
                    else:                                                   # :721-739
                        st["q"] = resize_nn(st["q"])
                    if st["mask"] is not None:
                        st["mask"] = resize_nn(st["mask"])
            for st in cur:
                q = st["q"]
                if st["mask"] is not None:                                  # copyTo(dst, mask) :455-459, :741-745
                    q = np.where(st["mask"] != 0, q, 0).astype(np.uint8)
                sp = spread(q, Tl)
                rm = response_maps(sp)
                self.quantized.append(q.copy())
                self.spread.append(sp)
                for j in range(8):
                    self.lm.append(linearize(rm[j], Tl))
            H, W = cur[-1]["q"].shape
            self.sizes.append((W, H))
        self.M = M


def _similarity(lms8: Sequence[np.ndarray], hdr, feats, size, T) -> np.ndarray:
    """linemod.cpp:1130-1214 for one template/modality -> uint8 [H', W']."""
    Wimg, Himg = size
    W = Wimg // T
    H = Himg // T
    width, height = int(hdr[0]), int(hdr[1])
    wf = (width - 1) // T + 1 if width - 1 >= 0 else -((1 - width) // T) + 1    # C division truncates
    hf = (height - 1) // T + 1 if height - 1 >= 0 else -((1 - height) // T) + 1
    span_x = W - wf
    span_y = H - hf
    tp = span_y * W + span_x + 1                                            # :1155
    dst = np.zeros(H * W, np.uint8)
    if tp <= 0:
        return dst.reshape(H, W)
    tp = min(tp, H * W + LM_PAD)  # (cannot exceed H*W for width,height >= 1)
    flats = [flat_lm(x) for x in lms8]
    for (x, y, label) in feats:
        if x < 0 or x >= Wimg or y < 0 or y >= Himg:                        # :1179
            continue
        cells = W * H
        base = ((y % T) * T + (x % T)) * cells + (y // T) * W + (x // T)
        n = min(tp, H * W)
        dst[:n] = dst[:n] + flats[label][base:base + n]                     # wrapping u8 add :1195, :1212
    return dst.reshape(H, W)


def _similarity_local(lms8, hdr, feats, size, T, cx, cy) -> np.ndarray:
    """linemod.cpp:1226-1300 -> uint8 [16, 16]."""
    Wimg, Himg = size
    W = Wimg // T
    cells = W * (Himg // T)
    ox = (int(cx / T) - 8) * T                                              # C truncating division :1240
    oy = (int(cy / T) - 8) * T
    dst = np.zeros((16, 16), np.uint8)
    flats = [flat_lm(x) for x in lms8]
    for (x, y, label) in feats:
        x = int(x) + ox
        y = int(y) + oy
        if x < 0 or y < 0 or x >= Wimg or y >= Himg:                        # :1257
            continue
        base = ((y % T) * T + (x % T)) * cells + (y // T) * W + (x // T)
        fl = flats[label]
        for row in range(16):
            dst[row] = dst[row] + fl[base + row * W: base + row * W + 16]
    return dst


def match(fe: FrontEnd, tset, threshold: float, class_filter: Optional[Sequence[int]] = None):
    """Detector::match + matchClass (linemod.cpp:1418-1440, 1451-1577).

    Returns (raw, final): ``raw`` is the pre-sort candidate list in the reference's emission order
    [(x, y, similarity f32, class_idx, template_id)], ``final`` the canonical sort + adjacent-unique
    (similarity desc, template_id asc, class asc, y asc, x asc; equality on x,y,sim,class).
    """
    L, M = tset.n_levels, tset.n_modalities
    T = fe.T
    raw = []
    first_of_class = {}
    for t in range(tset.n_templates):
        first_of_class.setdefault(int(tset.class_of[t]), t)
    thr = f32(threshold)
    for t in range(tset.n_templates):
        cls = int(tset.class_of[t])
        if class_filter is not None and cls not in class_filter:
            continue
        tid = t - first_of_class[cls]
        lowest = L - 1
        Tl = T[lowest]
        nf = 0
        total = None
        for m in range(M):
            hdr, feats = tset.template(t, lowest, m)
            nf += len(feats)
            s = _similarity(fe.lm[(lowest * M + m) * 8:(lowest * M + m) * 8 + 8], hdr, feats, fe.sizes[lowest], Tl)
            total = s.astype(np.uint16) if total is None else total + s                         # :1322-1338
        raw_thr = int(f32(2 * nf) + (thr / f32(100.0)) * f32(2 * nf) + f32(0.5))                 # :1487
        cands = []
        rr, cc = np.nonzero(total > raw_thr)                                                    # :1491-1506 (row-major)
        off = Tl // 2 + (Tl % 2 - 1)
        for r, c in zip(rr, cc):
            score = f32(f32(int(total[r, c])) * f32(100.0)) / f32(4 * nf) + f32(0.5)            # :1502
            cands.append([int(c) * Tl + off, int(r) * Tl + off, f32(score)])
        for l in range(L - 2, -1, -1):                                                          # :1509-1573
            Tl = T[l]
            Wl, Hl = fe.sizes[l]
            border = 8 * Tl
            off = Tl // 2 + (Tl % 2 - 1)
            h0, _ = tset.template(t, l, 0)
            max_x = Wl - int(h0[0]) - border
            max_y = Hl - int(h0[1]) - border
            for cnd in cands:
                x = cnd[0] * 2 + 1
                y = cnd[1] * 2 + 1
                x = min(max(x, border), max_x)
                y = min(max(y, border), max_y)
                nf2 = 0
                tot = np.zeros((16, 16), np.uint16)
                for m in range(M):
                    hdr, feats = tset.template(t, l, m)
                    nf2 += len(feats)
                    tot += _similarity_local(fe.lm[(l * M + m) * 8:(l * M + m) * 8 + 8], hdr, feats, (Wl, Hl), Tl, x, y)
                best, br, bc = 0, -1, -1
                flat = tot.reshape(-1)
                k = int(flat.argmax())
                if flat[k] > 0:
                    best, br, bc = int(flat[k]), k // 16, k % 16                                # first max wins :1555
                cnd[0] = (int(x / Tl) - 8 + bc) * Tl + off                                      # :1564-1566
                cnd[1] = (int(y / Tl) - 8 + br) * Tl + off
                cnd[2] = f32(f32(best) * f32(100.0)) / f32(4 * nf2)
            cands = [c for c in cands if not (c[2] < thr)]                                      # :1570-1572
        for c in cands:
            raw.append((c[0], c[1], f32(c[2]), cls, tid))
    return raw, canonical_sort_unique(raw)


def canonical_sort_unique(raw):
    """SURVEY A.5 (replaces the tie-order-dependent std::sort + std::unique of linemod.cpp:1437-1439)."""
    s = sorted(raw, key=lambda m: (-float(m[2]), m[4], m[3], m[1], m[0]))
    out = []
    for m in s:
        if out and out[-1][0] == m[0] and out[-1][1] == m[1] and out[-1][2] == m[2] and out[-1][3] == m[3]:
            continue
        out.append(m)
    return out


# --------------------------------------------------------------------------------------------
# ICP   (ICP/depth_to_3d.cpp, ICP/common.cpp, ICP/ICP.cpp, ICP/detection.cpp)
# --------------------------------------------------------------------------------------------
MAX_VALID_DEPTH = f32(900.0)  # ICP/common.cpp:264


def depth_to_3d_mm(depth16: np.ndarray, fx, fy, cx, cy) -> np.ndarray:
    """depthTo3d (depth_to_3d.cpp:99-137, 244-260) then scale_mat_vec3f(.,1000) (common.cpp:418-425)."""
    H, W = depth16.shape
    # in.convertTo(out, CV_32F, 1/1000.0): OpenCV applies the scale in float
    z = depth16.astype(f32) * f32(1.0 / 1000.0)
    z[depth16 == 0] = np.nan
    inv_fx = f32(1.0) / f32(fx)
    inv_fy = f32(1.0) / f32(fy)
    xc = (np.arange(W, dtype=f32) - f32(cx)) * inv_fx
    yc = (np.arange(H, dtype=f32) - f32(cy)) * inv_fy
    pts = np.empty((H, W, 3), f32)
    pts[..., 0] = xc[None, :] * z
    pts[..., 1] = yc[:, None] * z
    pts[..., 2] = z
    return pts * f32(1000.0)            # Vec3f *= int  ->  each component * float(1000)


def _valid(p: np.ndarray) -> np.ndarray:
    return p[..., 2] <= MAX_VALID_DEPTH  # NaN fails, common.cpp:261-266


def _seqsum(a: np.ndarray) -> f32:
    """Left-to-right fp32 accumulation starting from 0 (Vec3f +=, Matx33f +=)."""
    if a.size == 0:
        return f32(0)
    return np.cumsum(a.astype(f32), dtype=f32)[-1]


def get_mean(pts: np.ndarray) -> np.ndarray:
    """getMean, ICP.cpp:8-25."""
    c = np.zeros(3, f32)
    n = len(pts)
    if n > 0:
        for k in range(3):
            c[k] = _seqsum(pts[:, k]) / f32(n)
    return c


def transform_points(pts: np.ndarray, R: np.ndarray, T: np.ndarray) -> np.ndarray:
    """transformPoints in place, ICP.cpp:28-45: only valid points move.  Matx33f*Vec3f accumulates
    ((0 + r0*x) + r1*y) + r2*z in fp32, then + T."""
    R = R.astype(f32)
    T = T.astype(f32)
    out = pts.copy()
    v = _valid(pts)
    p = pts[v]
    q = np.empty_like(p)
    for i in range(3):
        q[:, i] = ((R[i, 0] * p[:, 0] + R[i, 1] * p[:, 1]) + R[i, 2] * p[:, 2]) + T[i]
    out[v] = q
    return out


def copy_points(pts: np.ndarray) -> np.ndarray:
    """copyPoints, ICP.cpp:48-65: invalid points become (0,0,0) (value-initialised by resize)."""
    out = np.zeros_like(pts)
    v = _valid(pts)
    out[v] = pts[v]
    return out


def l2_dist_clouds(model: np.ndarray, ref: np.ndarray, thr) -> Tuple[f32, f32]:
    """getL2distClouds, ICP.cpp:68-111 -> (ratio_inliers, dist_mean)."""
    n = len(model)
    both = _valid(ref[:n]) & _valid(model)
    d = (model[both] - ref[:n][both]).astype(f32)
    dist = np.sqrt((d.astype(np.float64) ** 2).sum(axis=1)).astype(f32)   # cv::norm(Vec3f): double accumulate
    inl = dist <= f32(thr)
    counter = int(both.sum())
    nin = int(inl.sum())
    if counter > 0:
        with np.errstate(divide="ignore", invalid="ignore"):
            dist_mean = _seqsum(dist[inl]) / f32(nin)
            ratio = f32(nin) / f32(counter)
    else:
        dist_mean = np.finfo(f32).max
        ratio = f32(0)
    return f32(ratio), f32(dist_mean)


def icp_cloud_to_cloud_ex(pts_ref: np.ndarray, pts_model: np.ndarray, icp_it_thr=10, dist_mean_thr=0.5,
                          dist_diff_thr=0.01, trace: Optional[list] = None):
    """icpCloudToCloud_Ex, ICP.cpp:617-809 -> (dist_mean, R, T, inlier_ratio, iterations)."""
    R = np.eye(3, dtype=f32)
    T = np.zeros(3, f32)
    if len(pts_model) < 3 or len(pts_ref) < 3:                                                 # :633-638
        return f32(-1), np.zeros((3, 3), f32), np.zeros(3, f32), f32(0), 0
    index = cv2.flann_Index(np.ascontiguousarray(pts_ref, f32), dict(algorithm=4, leaf_max_size=15))  # :658
    tmp = copy_points(pts_model)                                                               # :667
    ratio, dist_mean = l2_dist_clouds(tmp, pts_ref, np.finfo(f32).max)                         # :670
    dist_diff = np.finfo(f32).max
    it = 0
    dmt, ddt = f32(dist_mean_thr), f32(dist_diff_thr)
    while dist_mean > dmt and dist_diff > ddt and it < icp_it_thr:                             # :684
        it += 1
        if it == 1:                                                                            # :700-704
            cor_m = copy_points(tmp)
            cor_r = copy_points(pts_ref)
        else:                                                                                  # :708 -> :193-279
            idx, d2 = index.knnSearch(np.ascontiguousarray(tmp, f32), 1, params={})
            keep = d2[:, 0] <= f32(3) * dist_mean                                              # squared vs un-squared quirk
            cor_m = tmp[keep]
            cor_r = pts_ref[idx[keep, 0]]
        if len(cor_r) < 3 or len(cor_m) < 3:                                                   # :711-715
            it = icp_it_thr
            continue
        mc = get_mean(cor_m)                                                                   # :722-724
        rc = get_mean(cor_r)
        n = min(len(cor_m), len(cor_r))
        cov = np.zeros((3, 3), f32)                                                            # :731-735 (not centred)
        for i in range(3):
            for j in range(3):
                cov[i, j] = _seqsum(cor_m[:n, i] * cor_r[:n, j])
        w, u, vt = cv2.SVDecomp(cov)                                                           # :742
        Ropt = (vt.T.astype(f32) @ u.T.astype(f32)).astype(f32)                                # :744
        Topt = (rc - _matvec(Ropt, mc)).astype(f32)                                            # :747
        if not (np.isfinite(Ropt).all() and np.isfinite(Topt).all()):                          # :748-749
            continue
        tmp = transform_points(tmp, Ropt, Topt)                                                # :756
        dist_diff = dist_mean                                                                  # :778-780
        ratio, dist_mean = l2_dist_clouds(tmp, pts_ref, f32(3) * dist_mean)
        dist_diff = f32(dist_diff - dist_mean)
        T = (_matvec(Ropt, T) + Topt).astype(f32)                                              # :793-797
        R = _matmul(Ropt, R)
        if trace is not None:
            trace.append((it, float(dist_mean), float(dist_diff), len(cor_m)))
    return f32(dist_mean), R, T, f32(ratio), it


def _matvec(R, v):
    R = R.astype(f32)
    v = v.astype(f32)
    return np.array([(R[i, 0] * v[0] + R[i, 1] * v[1]) + R[i, 2] * v[2] for i in range(3)], f32)


def _matmul(A, B):
    A = A.astype(f32)
    B = B.astype(f32)
    C = np.zeros((3, 3), f32)
    for i in range(3):
        for j in range(3):
            C[i, j] = (A[i, 0] * B[0, j] + A[i, 1] * B[1, j]) + A[i, 2] * B[2, j]
    return C


def detection(model_depth: np.ndarray, ref_depth: np.ndarray, K_ref, rect_model, rect_ref,
              icp_it_thr=10, dist_mean_thr=0.5, dist_diff_thr=0.01, r_match=None, t_match=None, trace=None):
    """detection(), ICP/detection.cpp:11-254 (test_id == 2 branch) -> dict."""
    r_match = np.eye(3, dtype=f32) if r_match is None else np.asarray(r_match, f32)
    t_match = np.zeros(3, f32) if t_match is None else np.asarray(t_match, f32)
    fx, fy, cx, cy = K_ref
    ref3 = depth_to_3d_mm(ref_depth, fx, fy, cx, cy)                                           # :31-32, :39
    mod3 = depth_to_3d_mm(model_depth, 608.0, 608.0, 320.0, 240.0)                             # :35-36, :40
    mx, my, mw, mh = rect_model
    rx, ry, rw, rh = rect_ref
    cm = mod3[my:my + mh, mx:mx + mw].reshape(-1, 3)                                           # :43-44
    cr = ref3[ry:ry + rh, rx:rx + rw].reshape(-1, 3)
    n = min(len(cm), len(cr))
    keep = _valid(cr[:n]) & _valid(cm[:n])                                                     # :114 -> common.cpp:382-405
    pts_ref = np.ascontiguousarray(cr[:n][keep])
    pts_mod = np.ascontiguousarray(cm[:n][keep])
    m_c = get_mean(pts_mod)                                                                    # :165-166
    r_c = get_mean(pts_ref)
    t_tmp = (r_c - m_c).astype(f32)                                                            # :177
    t_init = (t_tmp + t_match).astype(f32)                                                     # :199
    pts_mod = transform_points(pts_mod, np.eye(3, dtype=f32), t_tmp)                           # :206
    dist_mean, R, T, ratio, it = icp_cloud_to_cloud_ex(pts_ref, pts_mod, icp_it_thr, dist_mean_thr,
                                                       dist_diff_thr, trace)                   # :228
    T_final = (_matvec(R, t_init) + T).astype(f32)                                             # :232-233
    R_final = _matmul(R, r_match)                                                              # :234
    return dict(R=R_final, T=T_final, dist_mean=f32(dist_mean), inlier_ratio=f32(ratio), iterations=it,
                n_points=len(pts_ref), R_icp=R, T_icp=T)


def non_maximum_suppression(objs: List[dict], th_obj_dist: float) -> List[int]:
    """nonMaximumSuppression, ICP/NMS.cpp:6-39.  objs: dicts with t(3), n_model, icp_dist; returns the
    indices of the surviving objects in emission order (mutates 'check_done')."""
    out = []
    th = f32(th_obj_dist)
    for i, o in enumerate(objs):
        if o.get("check_done"):
            continue
        win = i
        size_th = int(f32(objs[i]["n_model"]) * f32(0.85))   # (float)size * 0.85 is a double product
        size_th = int(float(f32(objs[i]["n_model"])) * 0.85)
        for j in range(i + 1, len(objs)):
            if objs[j].get("check_done"):
                continue
            d = np.asarray(objs[win]["t"], f32).astype(np.float64) - np.asarray(objs[j]["t"], f32).astype(np.float64)
            if math.sqrt(float((d * d).sum())) < float(th):
                objs[j]["check_done"] = True
                if objs[j]["n_model"] > size_th and f32(objs[j]["icp_dist"]) < f32(objs[win]["icp_dist"]):
                    win = j
        out.append(win)
    return out
