"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's camera-motion estimator, SURVEY.md section 8(f) row 1.

Follows ``tracker/botsort.py:111-235`` (``GMC.applyFeaures``, method 'orb', downscale 2 -- what ``BoTSORT.__init__`` :286
builds) stage by stage in NumPy, so that every stage of the CUDA estimator (csrc/b2t_gmc.cu) has something to be compared
with.  The reference calls OpenCV for the arithmetic; OpenCV (4.13 in this image) is the reference's own dependency and is
used here as the pin: ``tests/test_oracle_gmc.py`` checks each restated stage against the cv2 call the reference makes
(cvtColor, resize, FastFeatureDetector(20), ORB.compute, BFMatcher.knnMatch) and the whole estimate against the UNMODIFIED
reference class imported through ``oracle/refshim.py``.

Stage                              reference line        cv2 call                       restatement          parity
gray, 1/2 scale                    :114-121              cvtColor, resize               gray_half()          bit-exact
mask (2 % border, detections)      :123-130              --                             keypoint_mask()      exact
FAST-9/16, threshold 20, 3x3 NMS   :132 (detector :20)   FastFeatureDetector.detect     fast_keypoints()     exact, same order
ORB descriptors of those points    :135 (extractor :21)  ORB_create().compute           orb_descriptors()    bit-exact up to rare rounding
                                                                                                              ties of the float blur (< 1e-5 of pixels)
2-NN Hamming matches               :149                  BFMatcher(NORM_HAMMING).knnMatch  knn2()            exact (ties: lower train index)
ratio / spatial / 2.5 sigma filter :158-185              --                             filter_matches()     exact
partial affine, RANSAC             :221                  estimateAffinePartial2D        cv2 itself (the oracle's estimator) and
                                                                                        ransac_partial_affine() = the restatement
                                                                                        the GPU kernel follows (own sampling order)

What ORB.compute does with *given* FAST keypoints (angle -1, octave 0) -- OpenCV features2d/src/orb.cpp, published
algorithm: keypoints closer than 31 px to the border are dropped; one pyramid level; the level is smoothed with a 7x7
Gaussian (sigma 2, float kernel, separable, result rounded to uint8, BORDER_REFLECT_101); bit i of the 256-bit descriptor
is ``I(p + a_i) < I(p + b_i)`` for the 256 learned point pairs of the ORB pattern rotated by the keypoint angle (-1 degree)
and rounded to integers.  The effective integer pair table (``ORB_PAIRS``) was recovered from cv2 itself by probing
``ORB.compute`` with single-pixel images (tools/extract_orb_pattern.py, deterministic, unique solution for all 256 pairs) --
OpenCV's sources are not in the image and nothing of it is copied.
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)

# 16-pixel Bresenham circle of radius 3, (dx, dy), in the order OpenCV walks it (starting at the bottom, counter-clockwise)
CIRCLE = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)]
FAST_THRESHOLD = 20          # botsort.py:20
ORB_EDGE = 31                # cv2.ORB_create() default edgeThreshold: keypoints nearer to the border are dropped by compute()
GAUSS7 = np.array([0.07015932, 0.13107488, 0.19071282, 0.21610594, 0.19071282, 0.13107488, 0.07015932], np.float32)   # getGaussianKernel(7, 2, CV_32F)


def orb_pairs():
    """(256, 2, 2) int: pair i = ((ay, ax), (by, bx)); bit i = I(p + a) < I(p + b).  Generated table, see the module docstring."""
    path = os.path.join(ROOT, "yolov7-tracker_b200", "csrc", "b2t_orb_pattern.inc")
    vals = []
    for line in open(path):
        line = line.split("//")[0].strip()
        if line:
            vals += [int(v) for v in line.replace("{", " ").replace("}", " ").split(",") if v.strip()]
    return np.array(vals, dtype=np.int64).reshape(256, 2, 2)


def gray_half(frame_bgr, downscale=2):
    """botsort.py:114-121.  cvtColor(BGR2GRAY) on uint8 is 15-bit fixed point; the exact 2x2 down-scale of cv2.resize
    (INTER_LINEAR with scale 2 = INTER_AREA) is (a + b + c + d + 2) >> 2.  Other scales: oracle/preprocess.py's resize."""
    f = frame_bgr.astype(np.int64)
    g = ((f[..., 0] * 3735 + f[..., 1] * 19235 + f[..., 2] * 9798 + (1 << 14)) >> 15).astype(np.uint8)
    if downscale <= 1:
        return g
    h, w = g.shape
    if downscale == 2 and h % 2 == 0 and w % 2 == 0:
        a = g.astype(np.int32)
        return ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    from oracle import preprocess as P
    return P.resize_linear_u8(g[..., None], (w // downscale, h // downscale))[..., 0]


def keypoint_mask(shape, detections, downscale=2):
    """botsort.py:123-130: 255 inside the central 96 % of the frame, 0 inside every detection box (tlbr / downscale, truncated).
    (Negative box corners would wrap around in the reference's NumPy slices; detections are clipped to the image upstream.)"""
    h, w = shape
    mask = np.zeros((h, w), np.uint8)
    mask[int(0.02 * h): int(0.98 * h), int(0.02 * w): int(0.98 * w)] = 255
    if detections is not None:
        for det in detections:
            t = (np.asarray(det[:4]) / downscale).astype(np.int_)
            mask[max(t[1], 0):max(t[3], 0), max(t[0], 0):max(t[2], 0)] = 0
    return mask


def fast_score_map(g, t=FAST_THRESHOLD):
    """FAST-9/16 corner score (OpenCV cornerScore<16>): max over the sixteen 9-arcs of min(v - p) and of min(p - v), minus 1;
    0 where the pixel is not a corner at threshold t.  Rows / columns nearer than 3 to the border are never corners."""
    h, w = g.shape
    a = g.astype(np.int32)
    out = np.zeros((h, w), np.int32)
    if h < 7 or w < 7:
        return out
    d = np.stack([a[3:h - 3, 3:w - 3] - a[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx] for dx, dy in CIRCLE], 0)
    d = np.concatenate([d, d[:8]], 0)
    best = np.full(d.shape[1:], -1 << 20)
    for k in range(16):
        best = np.maximum(best, np.maximum(d[k:k + 9].min(0), (-d[k:k + 9]).min(0)))
    out[3:h - 3, 3:w - 3] = np.where(best > t, best - 1, 0)
    return out


def fast_keypoints(g, mask=None, t=FAST_THRESHOLD):
    """FastFeatureDetector_create(t).detect(g, mask): corners whose score is strictly greater than all 8 neighbours' (non-corners
    score 0), THEN the mask filter; row-major order.  Returns xs, ys, scores."""
    s = fast_score_map(g, t)
    h, w = s.shape
    p = np.pad(s, 1)
    keep = s > 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx or dy:
                keep &= s > p[1 + dy:1 + dy + h, 1 + dx:1 + dx + w]
    if mask is not None:
        keep &= mask != 0
    ys, xs = np.nonzero(keep)
    return xs, ys, s[ys, xs]


def orb_blur(g):
    """7x7 Gaussian (sigma 2) as OpenCV's float filter engine applies it to a uint8 image: symmetric taps paired before the
    multiply, rows then columns, float32, round half to even."""
    h, w = g.shape
    k = GAUSS7
    p = np.pad(g, 3, mode='reflect').astype(np.int32)
    c = lambda i: p[:, 3 + i:3 + i + w]                                       # noqa: E731
    hh = k[3] * c(0).astype(np.float32)
    for j in (1, 2, 3):
        hh = hh + k[3 - j] * (c(j) + c(-j)).astype(np.float32)
    r = lambda i: hh[3 + i:3 + i + h, :]                                      # noqa: E731
    v = k[3] * r(0)
    for j in (1, 2, 3):
        v = v + k[3 - j] * (r(j) + r(-j))
    return np.rint(v).astype(np.uint8)


def orb_filter_border(xs, ys, shape, edge=ORB_EDGE):
    h, w = shape
    return (xs >= edge) & (xs < w - edge) & (ys >= edge) & (ys < h - edge)


def orb_descriptors(g, xs, ys):
    """ORB_create().compute(g, keypoints) for FAST keypoints that survive orb_filter_border: (n, 32) uint8."""
    pat = orb_pairs()
    b = orb_blur(g).astype(np.int32)
    if len(xs) == 0:
        return np.zeros((0, 32), np.uint8)
    ta = b[ys[:, None] + pat[None, :, 0, 0], xs[:, None] + pat[None, :, 0, 1]]
    tb = b[ys[:, None] + pat[None, :, 1, 0], xs[:, None] + pat[None, :, 1, 1]]
    return np.packbits((ta < tb).astype(np.uint8), axis=1, bitorder='little')


_POP = np.array([bin(i).count('1') for i in range(256)], np.int32)


def knn2(dq, dt):
    """BFMatcher(NORM_HAMMING).knnMatch(dq, dt, 2): per query the two nearest train descriptors; equal distances keep the lower
    train index first.  Returns (i1, d1, i2, d2)."""
    D = np.zeros((len(dq), len(dt)), np.int32)
    for j in range(32):
        D += _POP[dq[:, j][:, None] ^ dt[:, j][None, :]]
    q = np.arange(len(dq))
    i1 = np.argmin(D, 1)
    d1 = D[q, i1]
    D[q, i1] = 1 << 20
    i2 = np.argmin(D, 1)
    return i1, d1, i2, D[q, i2]


def filter_matches(prev_xy, cur_xy, i1, d1, d2, width, height):
    """botsort.py:158-198: ratio test 0.9, |displacement| < a quarter of the frame, one-sided 2.5 sigma test on the displacement.
    prev_xy / cur_xy: (n, 2) float (x, y).  Returns (prev_pts, cur_pts) float64 arrays in query order."""
    ok = d1.astype(np.float64) < 0.9 * d2.astype(np.float64)
    dist = prev_xy.astype(np.float64) - cur_xy[i1].astype(np.float64)
    ok &= (np.abs(dist[:, 0]) < 0.25 * width) & (np.abs(dist[:, 1]) < 0.25 * height)
    q = np.nonzero(ok)[0]
    if len(q) == 0:
        return np.zeros((0, 2)), np.zeros((0, 2))
    sd = dist[q]
    inl = np.all((sd - sd.mean(0)) < 2.5 * sd.std(0), axis=1)
    q = q[inl]
    return prev_xy[q].astype(np.float64), cur_xy[i1[q]].astype(np.float64)


# ---- the RANSAC restatement the GPU kernel follows (cv2.estimateAffinePartial2D's published scheme: minimal samples of two
# correspondences -> 4-dof similarity, inliers = reprojection error < 3 px, best = most inliers, least-squares refit on them;
# the sampling sequence is ours: OpenCV's RNG stream is not reproduced, so the two agree when they find the same inlier set)
RANSAC_HYPOTHESES = 512
RANSAC_THRESHOLD = 3.0


def lcg_pair(t, n):
    """Deterministic pair of distinct indices for hypothesis t (32-bit LCG, the kernel uses the same arithmetic)."""
    s = (t * 2654435761 + 12345) & 0xffffffff
    s = (s * 1664525 + 1013904223) & 0xffffffff
    i = (s >> 8) % n
    s = (s * 1664525 + 1013904223) & 0xffffffff
    j = (s >> 8) % (n - 1)
    if j >= i:
        j += 1
    return i, j


def similarity_from_pair(p0, p1, q0, q1):
    dx, dy = p1[0] - p0[0], p1[1] - p0[1]
    den = dx * dx + dy * dy
    if den < 1e-12:
        return None
    ux, uy = q1[0] - q0[0], q1[1] - q0[1]
    a = (dx * ux + dy * uy) / den
    b = (dx * uy - dy * ux) / den
    return a, b, q0[0] - (a * p0[0] - b * p0[1]), q0[1] - (b * p0[0] + a * p0[1])


def fit_similarity(src, dst):
    """Least-squares 4-dof similarity (what the LM refinement of estimateAffinePartial2D converges to)."""
    ms, md = src.mean(0), dst.mean(0)
    x, y = src[:, 0] - ms[0], src[:, 1] - ms[1]
    u, v = dst[:, 0] - md[0], dst[:, 1] - md[1]
    den = (x * x + y * y).sum()
    a = (x * u + y * v).sum() / den
    b = (x * v - y * u).sum() / den
    return a, b, md[0] - (a * ms[0] - b * ms[1]), md[1] - (b * ms[0] + a * ms[1])


def ransac_partial_affine(src, dst, hypotheses=RANSAC_HYPOTHESES, thr=RANSAC_THRESHOLD):
    n = len(src)
    best, best_inl = -1, None
    for t in range(hypotheses):
        i, j = lcg_pair(t, n)
        m = similarity_from_pair(src[i], src[j], dst[i], dst[j])
        if m is None:
            continue
        a, b, tx, ty = m
        ex = a * src[:, 0] - b * src[:, 1] + tx - dst[:, 0]
        ey = b * src[:, 0] + a * src[:, 1] + ty - dst[:, 1]
        inl = ex * ex + ey * ey < thr * thr
        c = int(inl.sum())
        if c > best:
            best, best_inl = c, inl
    if best < 2:
        return None
    a, b, tx, ty = fit_similarity(src[best_inl], dst[best_inl])
    return np.array([[a, -b, tx], [b, a, ty]], np.float64)


def corner_displacement(Ha, Hb, height, width):
    """max over the frame's four corners of |Ha p - Hb p| in pixels: how differently two 2 x 3 warps move the frame."""
    pts = np.array([[0, 0, 1], [width, 0, 1], [0, height, 1], [width, height, 1]], np.float64).T
    d = np.asarray(Ha, np.float64) @ pts - np.asarray(Hb, np.float64) @ pts
    return float(np.sqrt((d * d).sum(0)).max())


class GMCOracle:
    """GMC(method='orb', downscale=2).apply restated (stateful: previous key points and descriptors)."""

    def __init__(self, downscale=2, estimator="cv2"):
        self.downscale = max(1, int(downscale))
        self.prev_xy = self.prev_desc = None
        self.estimator = estimator
        self.last = {}

    def stages(self, raw_frame, detections=None):
        g = gray_half(raw_frame, self.downscale)
        mask = keypoint_mask(g.shape, detections, self.downscale)
        xs, ys, sc = fast_keypoints(g, mask)
        keep = orb_filter_border(xs, ys, g.shape)
        xs, ys = xs[keep], ys[keep]
        return g, xs, ys, orb_descriptors(g, xs, ys)

    def apply(self, raw_frame, detections=None):
        g, xs, ys, desc = self.stages(raw_frame, detections)
        h, w = g.shape
        xy = np.stack([xs, ys], 1).astype(np.float32)
        H = np.eye(2, 3)
        self.last = {"gray": g, "xy": xy, "desc": desc, "H_restated": H}
        if self.prev_xy is None:
            self.prev_xy, self.prev_desc = xy, desc
            return H
        if len(self.prev_desc) and len(desc) >= 2:
            i1, d1, i2, d2 = knn2(self.prev_desc, desc)
            src, dst = filter_matches(self.prev_xy, xy, i1, d1, d2, w, h)
            self.last.update(i1=i1, d1=d1, d2=d2, src=src, dst=dst)
            if src.shape[0] > 4:
                def full_res(est):
                    if est is None:
                        return np.eye(2, 3)
                    est = est.copy()
                    est[0, 2] *= self.downscale                      # botsort.py:224-226
                    est[1, 2] *= self.downscale
                    return est
                if self.estimator in ("cv2", "both"):
                    import cv2
                    H = full_res(cv2.estimateAffinePartial2D(src, dst, cv2.RANSAC)[0])
                if self.estimator in ("restated", "both"):
                    Hr = full_res(ransac_partial_affine(src, dst))
                    self.last["H_restated"] = Hr
                    if self.estimator == "restated":
                        H = Hr
        self.prev_xy, self.prev_desc = xy, desc
        return H
