"""Oracle: dense IoU with the Fast-R-CNN "+1 pixel" convention (float64).

TEST INFRASTRUCTURE (see oracle/__init__.py).  **Parity unpinned**: this restates
the third-party ``cython_bbox.bbox_overlaps`` (PyPI ``cython_bbox``, unpinned by
the reference, absent from the image) which the reference calls at
tracker/matching.py:56-59.  Published algorithm (py-faster-rcnn ``bbox.pyx``):

    iw = min(b.x2, q.x2) - max(b.x1, q.x1) + 1
    ih = min(b.y2, q.y2) - max(b.y1, q.y1) + 1
    if iw > 0 and ih > 0:
        ua  = (b.x2-b.x1+1)*(b.y2-b.y1+1) + (q.x2-q.x1+1)*(q.y2-q.y1+1) - iw*ih
        iou = iw*ih / ua
    else 0

The evaluation order below (area_b + area_q, then - iw*ih, no FMA) is the one
the CUDA kernel uses, so the two agree bit-for-bit on identical inputs.
"""
import numpy as np


def bbox_overlaps(boxes, query_boxes):
    """(N,4),(K,4) float64 tlbr -> (N,K) float64 IoU.  Mirrors the cython_bbox signature."""
    b = np.ascontiguousarray(boxes, dtype=np.float64).reshape(-1, 4)
    q = np.ascontiguousarray(query_boxes, dtype=np.float64).reshape(-1, 4)
    n, k = b.shape[0], q.shape[0]
    out = np.zeros((n, k), dtype=np.float64)
    if n == 0 or k == 0:
        return out
    area_b = ((b[:, 2] - b[:, 0] + 1.0) * (b[:, 3] - b[:, 1] + 1.0))[:, None]
    area_q = ((q[:, 2] - q[:, 0] + 1.0) * (q[:, 3] - q[:, 1] + 1.0))[None, :]
    iw = np.minimum(b[:, None, 2], q[None, :, 2]) - np.maximum(b[:, None, 0], q[None, :, 0]) + 1.0
    ih = np.minimum(b[:, None, 3], q[None, :, 3]) - np.maximum(b[:, None, 1], q[None, :, 1]) + 1.0
    ok = (iw > 0) & (ih > 0)
    inter = iw * ih
    ua = (area_b + area_q) - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = inter / ua
    out[ok] = iou[ok]
    return out


def ious(atlbrs, btlbrs):
    """tracker/matching.py:44-61 -- zeros((N,M)) when either side is empty."""
    n, m = len(atlbrs), len(btlbrs)
    if n == 0 or m == 0:
        return np.zeros((n, m), dtype=np.float64)
    return bbox_overlaps(np.ascontiguousarray(atlbrs, dtype=np.float64),
                         np.ascontiguousarray(btlbrs, dtype=np.float64))


def iou_distance_tlbr(atlbrs, btlbrs):
    """tracker/matching.py:64-82 for raw tlbr inputs: cost = 1 - IoU."""
    return 1.0 - ious(atlbrs, btlbrs)
