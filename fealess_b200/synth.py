"""Seeded synthetic RGB-D frames, template sets and ICP hypotheses (numpy only).

The reference ships no data (its ``data/`` directory is git-ignored), so every
parity test and the benchmark run on inputs made here.  The shapes follow SURVEY.md section 8(d):

* BGR frame: low-frequency gradient background + 20..60 random filled ellipses / convex polygons +
  sigma=2 Gaussian noise, so that a useful fraction of pixels passes the ``mag > 100`` gate and the
  5-of-9 orientation vote of ``hysteresisGradient`` (reference linemod/linemod.cpp:307-385).
* depth frame: uint16 millimetres, piecewise planar / spherical surfaces in [400, 899] mm (valid for
  LINE-MOD ``d < 2000`` linemod.cpp:628 and for ICP ``z <= 900`` ICP/common.cpp:261-266), 2..5 % zero
  holes, +-1 mm noise, never a value in (0, 50).
* templates: the data model of ``cup_linemod::Template`` (linemod/linemod.hpp:47-58) as produced by
  ``cropTemplates`` (linemod.cpp:52-96): feature coordinates in [0, width] x [0, height] inclusive,
  even offsets, width/height/offset halved per pyramid level, order [L0-M0, L0-M1, L1-M0, ...].

Everything is deterministic in (seed, arguments); the generator is PCG64 as the survey prescribes.
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional, Sequence

import numpy as np

FRAME_SEED_BASE = 0xFEA1E55

# features per modality per level (num_features /= 2 per pyrDown, linemod.cpp:437, 724)
FEATURES_PER_LEVEL = (63, 31, 15, 7)


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


# --------------------------------------------------------------------------------------------
# frames
# --------------------------------------------------------------------------------------------
def _shape_mask(rng, W, H, yy, xx):
    """One random filled ellipse or convex polygon as a boolean mask."""
    cx = rng.uniform(0, W)
    cy = rng.uniform(0, H)
    if rng.random() < 0.5:
        a = rng.uniform(0.03, 0.22) * W
        b = rng.uniform(0.03, 0.22) * W
        th = rng.uniform(0, np.pi)
        c, s = np.cos(th), np.sin(th)
        u = (xx - cx) * c + (yy - cy) * s
        v = -(xx - cx) * s + (yy - cy) * c
        return (u / a) ** 2 + (v / b) ** 2 <= 1.0
    n = int(rng.integers(3, 7))
    r = rng.uniform(0.04, 0.25) * W
    ang = np.sort(rng.uniform(0, 2 * np.pi, n))
    px = cx + r * np.cos(ang) * rng.uniform(0.6, 1.0, n)
    py = cy + r * np.sin(ang) * rng.uniform(0.6, 1.0, n)
    m = np.ones((H, W), bool)
    for i in range(n):
        x0, y0 = px[i], py[i]
        x1, y1 = px[(i + 1) % n], py[(i + 1) % n]
        m &= ((x1 - x0) * (yy - y0) - (y1 - y0) * (xx - x0)) >= 0
    return m


def make_frame(W: int = 640, H: int = 480, frame_idx: int = 0, seed: Optional[int] = None):
    """Return (bgr uint8[H,W,3], depth uint16[H,W]) for one synthetic RGB-D frame."""
    rng = _rng(FRAME_SEED_BASE + frame_idx if seed is None else seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    img = np.empty((H, W, 3), np.float32)
    for c in range(3):
        gx, gy = rng.uniform(-40, 40, 2)
        img[..., c] = rng.uniform(80, 170) + gx * (xx / W - 0.5) + gy * (yy / H - 0.5)
    depth = (rng.uniform(650, 850) + rng.uniform(-60, 60) * (xx / W - 0.5)
             + rng.uniform(-60, 60) * (yy / H - 0.5)).astype(np.float32)
    for _ in range(int(rng.integers(20, 61))):
        m = _shape_mask(rng, W, H, yy, xx)
        if not m.any():
            continue
        img[m] = rng.uniform(0, 255, 3).astype(np.float32)
        z0 = rng.uniform(430, 800)
        kind = rng.random()
        if kind < 0.6:      # tilted plane
            surf = z0 + rng.uniform(-0.5, 0.5) * (xx - xx[m].mean()) + rng.uniform(-0.5, 0.5) * (yy - yy[m].mean())
        else:               # spherical cap bulging towards the camera
            R = rng.uniform(0.2, 0.6) * W
            d2 = (xx - xx[m].mean()) ** 2 + (yy - yy[m].mean()) ** 2
            surf = z0 + 60.0 - np.sqrt(np.maximum(R * R - d2, 0.0)) * (60.0 / R)
        depth[m] = surf[m]
    img += rng.normal(0.0, 2.0, img.shape).astype(np.float32)
    bgr = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    depth += rng.integers(-1, 2, depth.shape).astype(np.float32)
    d16 = np.clip(np.rint(depth), 400, 899).astype(np.uint16)
    holes = rng.random((H, W)) < rng.uniform(0.02, 0.05)
    d16[holes] = 0
    return np.ascontiguousarray(bgr), np.ascontiguousarray(d16)


# --------------------------------------------------------------------------------------------
# templates
# --------------------------------------------------------------------------------------------
@dataclasses.dataclass
class TemplateSet:
    """Flat (C-ABI friendly) template database for ONE detector.

    headers  int32 [n_templates * L * M, 7]: width, height, offset_x, offset_y, pyramid_level,
             feature_begin, feature_count      (entry order: template-major, then level*M + modality)
    features int32 [n_features, 3]: x, y, label   (linemod.hpp:32-43)
    class_of int32 [n_templates]: class index of each template (templates of a class are contiguous,
             template_id = running index inside the class, linemod.cpp:1458)
    pose13   float32 [n_templates, 13]: 3x4 [R|t] row-major + distance (linemod.cpp:1617-1634)
    """
    n_levels: int
    n_modalities: int
    T: Sequence[int]
    class_names: List[str]
    headers: np.ndarray
    features: np.ndarray
    class_of: np.ndarray
    pose13: np.ndarray

    @property
    def n_templates(self) -> int:
        return int(self.class_of.shape[0])

    def template(self, t: int, level: int, modality: int):
        h = self.headers[(t * self.n_levels + level) * self.n_modalities + modality]
        return h, self.features[h[5]:h[5] + h[6]]

    def subset(self, idx: Sequence[int]) -> "TemplateSet":
        """Templates ``idx`` (global indices, class-contiguous order kept) re-packed; used for sharding."""
        LM = self.n_levels * self.n_modalities
        hs, fs, pos = [], [], 0
        for t in idx:
            for e in range(LM):
                h = self.headers[t * LM + e].copy()
                fs.append(self.features[h[5]:h[5] + h[6]])
                h[5] = pos
                pos += h[6]
                hs.append(h)
        return TemplateSet(self.n_levels, self.n_modalities, self.T, self.class_names,
                           np.array(hs, np.int32).reshape(-1, 7),
                           (np.concatenate(fs) if fs else np.zeros((0, 3), np.int32)).astype(np.int32),
                           self.class_of[list(idx)].copy(), self.pose13[list(idx)].copy())


def _random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], np.float32)


def _pose13(rng):
    p = np.zeros(13, np.float32)
    R = _random_rotation(rng)
    t = rng.uniform(-100, 100, 3).astype(np.float32)
    p[:12] = np.concatenate([R, t[:, None]], axis=1).reshape(-1)
    p[12] = np.float32(rng.uniform(400, 900))
    return p


def make_templates(n_templates: int, W: int = 640, H: int = 480, T: Sequence[int] = (5, 8),
                   n_modalities: int = 2, n_classes: int = 1, seed: int = 1,
                   quantized: Optional[Sequence[np.ndarray]] = None, planted_fraction: float = 0.01,
                   features_per_level: Sequence[int] = FEATURES_PER_LEVEL,
                   max_size: int = 200, min_size: int = 48) -> TemplateSet:
    """Random template set; a ``planted_fraction`` of them are cut out of ``quantized``.

    ``quantized`` is the list ``match`` returns as ``quantized_images`` (index = level*M + modality,
    one-hot uint8, linemod.cpp:1411-1412): planted templates take their features from those images so
    that they really are present in the frame; the rest have uniformly random labels/positions and
    produce (almost) no candidates at the 75 % threshold, like a real scene without the object.
    Templates of a class are contiguous and classes are dealt evenly (``obj00``, ``obj01`` ...).
    """
    rng = _rng(0x7E3A0000 + seed)
    L = len(T)
    M = n_modalities
    headers = np.zeros((n_templates * L * M, 7), np.int32)
    feats = []
    pos = 0
    class_of = (np.arange(n_templates) * n_classes // max(n_templates, 1)).astype(np.int32)
    pose13 = np.stack([_pose13(rng) for _ in range(n_templates)]) if n_templates else np.zeros((0, 13), np.float32)
    n_planted = int(round(planted_fraction * n_templates)) if quantized is not None else 0
    planted = set(rng.choice(n_templates, n_planted, replace=False).tolist()) if n_planted else set()
    # template must fit W_l - 16*T_l at every level for in-bounds similarityLocal
    lim_w = min((W >> l) - 16 * T[l] for l in range(L - 1)) if L > 1 else W - 1
    lim_h = min((H >> l) - 16 * T[l] for l in range(L - 1)) if L > 1 else H - 1
    for t in range(n_templates):
        w0 = int(rng.integers(min_size, min(max_size, lim_w, W - 2 * T[0]) + 1)) & ~1
        h0 = int(rng.integers(min_size, min(max_size, lim_h, H - 2 * T[0]) + 1)) & ~1
        ox0 = int(rng.integers(0, max(W - w0 - 1, 1))) & ~((1 << L) - 1)
        oy0 = int(rng.integers(0, max(H - h0 - 1, 1))) & ~((1 << L) - 1)
        for l in range(L):
            w, h, ox, oy = w0 >> l, h0 >> l, ox0 >> l, oy0 >> l
            nf = features_per_level[l]
            for m in range(M):
                f = None
                if t in planted:
                    q = quantized[l * M + m]
                    sub = q[oy:oy + h + 1, ox:ox + w + 1]
                    ys, xs = np.nonzero(sub)
                    if len(ys) >= nf:
                        sel = rng.choice(len(ys), nf, replace=False)
                        lab = np.log2(sub[ys[sel], xs[sel]]).astype(np.int32)
                        f = np.stack([xs[sel], ys[sel], lab], axis=1).astype(np.int32)
                if f is None:
                    cells = rng.choice((w + 1) * (h + 1), nf, replace=False)
                    f = np.stack([cells % (w + 1), cells // (w + 1), rng.integers(0, 8, nf)], axis=1).astype(np.int32)
                headers[(t * L + l) * M + m] = (w, h, ox, oy, l, pos, nf)
                feats.append(f)
                pos += nf
    features = np.concatenate(feats).astype(np.int32) if feats else np.zeros((0, 3), np.int32)
    names = ["obj%02d" % c for c in range(n_classes)]
    return TemplateSet(L, M, tuple(int(x) for x in T), names, headers, features, class_of, pose13)


# --------------------------------------------------------------------------------------------
# ICP hypotheses
# --------------------------------------------------------------------------------------------
def make_icp_pair(W: int = 640, H: int = 480, seed: int = 0, rect_wh=(100, 100),
                  fx: float = 608.0, fy: float = 608.0, cx: float = 320.0, cy: float = 240.0,
                  max_shift_mm: float = 6.0, max_rot_deg: float = 3.0, rect_xy=None):
    """A (model depth, ref depth, rect_model, rect_ref, pose13-ish) tuple for ``detection()``.

    The model depth image shows a smooth bumpy surface inside ``rect_model``; the reference depth is
    the same surface moved by a small rigid motion and re-rendered (by forward-splatting its points
    through the pinhole model ``initInternalMat`` hard-codes, ICP/common.cpp:358), shifted to
    ``rect_ref`` so the two crops are pixel-aligned the way ``Recognition`` aligns a match
    (CadReco/obj_reco_lmicp.cpp:127-132).  Depths stay in [400, 899] mm.
    """
    rng = _rng(0x1C900000 + seed)
    w, h = rect_wh
    mx = int(rng.integers(0, W - w))
    my = int(rng.integers(0, H - h))
    if rect_xy is not None:                   # caller-chosen model rect (the random draws above keep the stream position)
        mx, my = int(rect_xy[0]), int(rect_xy[1])
    # keep the two rects close: the crops are back-projected at their own pixel positions, so a large
    # offset would shear the clouds against each other (z * du / fx) beyond what ICP is meant to fix
    rx = int(np.clip(mx + rng.integers(-20, 21), 0, W - w - 1))
    ry = int(np.clip(my + rng.integers(-20, 21), 0, H - h - 1))
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    z0 = rng.uniform(520, 780)
    surf = (z0 + rng.uniform(-0.3, 0.3) * (xx - w / 2) + rng.uniform(-0.3, 0.3) * (yy - h / 2)
            + rng.uniform(5, 25) * np.sin(xx / w * rng.uniform(2, 7) + rng.uniform(0, 6))
            * np.cos(yy / h * rng.uniform(2, 7) + rng.uniform(0, 6)))
    model = np.zeros((H, W), np.uint16)
    model[my:my + h, mx:mx + w] = np.clip(np.rint(surf), 400, 899).astype(np.uint16)
    # move the surface: rotate about its centroid by a small angle and shift
    ang = np.deg2rad(rng.uniform(-max_rot_deg, max_rot_deg, 3))
    cxr, sxr = np.cos(ang[0]), np.sin(ang[0])
    cyr, syr = np.cos(ang[1]), np.sin(ang[1])
    czr, szr = np.cos(ang[2]), np.sin(ang[2])
    Rm = (np.array([[czr, -szr, 0], [szr, czr, 0], [0, 0, 1]]) @ np.array([[cyr, 0, syr], [0, 1, 0], [-syr, 0, cyr]])
          @ np.array([[1, 0, 0], [0, cxr, -sxr], [0, sxr, cxr]]))
    tm = rng.uniform(-max_shift_mm, max_shift_mm, 3)
    # dense re-render: evaluate the moved surface on a 3x supersampled grid and z-buffer splat
    ss = 3
    gy, gx = np.mgrid[0:h * ss, 0:w * ss].astype(np.float64) / ss
    gz = (z0 + (surf[np.clip(np.rint(gy).astype(int), 0, h - 1), np.clip(np.rint(gx).astype(int), 0, w - 1)] - z0))
    X = (gx + rx - cx) / fx * gz
    Y = (gy + ry - cy) / fy * gz
    P = np.stack([X.ravel(), Y.ravel(), gz.ravel()], 1)
    c = P.mean(0)
    P2 = (P - c) @ Rm.T + c + tm
    u = np.rint(P2[:, 0] / P2[:, 2] * fx + cx).astype(int)
    v = np.rint(P2[:, 1] / P2[:, 2] * fy + cy).astype(int)
    ref = np.zeros((H, W), np.float64)
    ok = (u >= 0) & (u < W) & (v >= 0) & (v < H)
    order = np.argsort(-P2[ok, 2])           # far first, so the nearest wins
    ref[v[ok][order], u[ok][order]] = P2[ok, 2][order]
    ref16 = np.where(ref > 0, np.clip(np.rint(ref), 400, 899), 0).astype(np.uint16)
    pose = _pose13(rng)
    return model, ref16, (mx, my, w, h), (rx, ry, w, h), pose
