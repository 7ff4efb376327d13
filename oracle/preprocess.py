"""Oracle of the reference's pre-processing -- TEST INFRASTRUCTURE, see oracle/__init__.py.

Restates ``TrackerLoader.__getitem__`` (tracker/tracker_dataloader.py:64-96, 'v5' / 'v7' branch) and ``_letterbox``
(:100-130): scale ratio, minimum-rectangle padding to a stride multiple, ``cv2.resize(..., INTER_LINEAR)``,
``cv2.copyMakeBorder(value=114)``, BGR -> RGB, HWC -> CHW, float32 / 255.

``cv2.resize`` is a third-party dependency of the reference (opencv-python, unpinned in requirements.txt); it IS installed in
the build container (4.13), so this restatement of its 8-bit INTER_LINEAR arithmetic is pinned against the real function
(tests/test_oracle_preprocess.py, fixtures in tests/golden/letterbox.npz written by tests/golden/make_golden_preprocess.py):
  * source coordinate fx = (float)((dx + 0.5) * scale - 0.5), left tap floor(fx), clamped at both image edges;
  * 11-bit fixed-point weights  saturate_cast<short>(w * 2048)  (round half to even);
  * horizontal pass in int32, vertical pass  ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
  * an exact 2 x 2 down-scale is INTER_AREA in OpenCV: (a + b + c + d + 2) >> 2.
"""
import numpy as np


def letterbox_geometry(shape_hw, new_shape=(1280, 1280), stride=64, auto=True, scaleup=True):
    """tracker_dataloader.py:100-126 -> dict(new_unpad=(w, h), top, bottom, left, right, ratio, dw, dh)."""
    h, w = int(shape_hw[0]), int(shape_hw[1])
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    r = min(new_shape[0] / h, new_shape[1] / w)
    if not scaleup:
        r = min(r, 1.0)
    new_unpad = int(round(w * r)), int(round(h * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = np.mod(dw, stride), np.mod(dh, stride)
    dw /= 2
    dh /= 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return dict(new_unpad=new_unpad, top=top, bottom=bottom, left=left, right=right, ratio=(r, r), dw=float(dw), dh=float(dh))


def _taps(dst, src):
    """left tap index and the two 11-bit weights of every destination coordinate (OpenCV resize.cpp, linear, 8-bit)."""
    scale = np.float64(1.0) / (np.float64(dst) / np.float64(src))      # cv::resize: scale = 1. / inv_scale, inv_scale = dsize / ssize
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    lo = s < 0
    f[lo] = 0; s[lo] = 0
    hi = s >= src - 1
    f[hi] = 0; s[hi] = src - 1
    w1 = np.rint(f * np.float32(2048)).astype(np.int64)                  # saturate_cast<short>: round half to even
    w0 = np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int64)
    return s, w0, w1


def resize_linear_u8(img, dst_wh):
    """cv2.resize(img, dst_wh, interpolation=cv2.INTER_LINEAR) for uint8 HWC images."""
    h, w = img.shape[:2]
    dw, dh = int(dst_wh[0]), int(dst_wh[1])
    src = img.astype(np.int64)
    if w == 2 * dw and h == 2 * dh:                                       # INTER_LINEAR with scale 2 x 2 == INTER_AREA (fast)
        s = src[0::2, 0::2] + src[0::2, 1::2] + src[1::2, 0::2] + src[1::2, 1::2]
        return ((s + 2) >> 2).astype(np.uint8)
    sx, a0, a1 = _taps(dw, w)
    sx1 = np.minimum(sx + 1, w - 1)
    rows = src[:, sx] * a0[None, :, None] + src[:, sx1] * a1[None, :, None]          # (h, dw, c) int, scale 2^11
    # vertical taps: the x pass clamps the weight at the edges, the y pass clamps the ROW index instead
    scale_y = np.float64(1.0) / (np.float64(dh) / np.float64(h))
    d = np.arange(dh, dtype=np.float64)
    f = ((d + 0.5) * scale_y - 0.5).astype(np.float32)
    sy = np.floor(f).astype(np.int64)
    f = (f - sy.astype(np.float32)).astype(np.float32)
    b1 = np.rint(f * np.float32(2048)).astype(np.int64)
    b0 = np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int64)
    r0 = rows[np.clip(sy, 0, h - 1)]
    r1 = rows[np.clip(sy + 1, 0, h - 1)]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def preprocess(ori_img, new_shape=(1280, 1280), stride=64, color=114):
    """tracker_dataloader.py:64-96: BGR uint8 (H, W, 3) -> float32 (3, H', W') RGB in [0, 1], plus the letterbox geometry."""
    g = letterbox_geometry(ori_img.shape[:2], new_shape, stride)
    img = ori_img
    if tuple(ori_img.shape[:2][::-1]) != tuple(g["new_unpad"]):
        img = resize_linear_u8(ori_img, g["new_unpad"])
    out = np.full((img.shape[0] + g["top"] + g["bottom"], img.shape[1] + g["left"] + g["right"], 3), color, dtype=np.uint8)
    out[g["top"]:g["top"] + img.shape[0], g["left"]:g["left"] + img.shape[1]] = img
    chw = np.ascontiguousarray(out[:, :, ::-1].transpose(2, 0, 1)).astype(np.float32)
    chw /= np.float32(255.0)
    return chw, g
