"""The functions ``tracker/track.py`` imports from ``utils.general`` (:32): ``non_max_suppression``, ``scale_coords``,
``check_img_size`` (+ ``xywh2xyxy``, ``clip_coords``, ``make_divisible``).  NMS runs on the GPU kernels of
csrc/b2t_detect.cu; the coordinate helpers are the same in-place tensor arithmetic as the reference (utils/general.py:
123-127, 176-178, 265-272, 319-340)."""
import ctypes as C
import math

import torch

from . import _b2t_path  # noqa: F401
from b200track import _lib as L

_ws = {}


def make_divisible(x, divisor):
    return math.ceil(x / divisor) * divisor


def check_img_size(img_size, s=32):
    new_size = make_divisible(img_size, int(s))
    if new_size != img_size:
        print('WARNING: --img-size %g must be multiple of max stride %g, updating to %g' % (img_size, s, new_size))
    return new_size


def xywh2xyxy(x):
    y = x.clone() if isinstance(x, torch.Tensor) else x.copy()
    y[:, 0] = x[:, 0] - x[:, 2] / 2
    y[:, 1] = x[:, 1] - x[:, 3] / 2
    y[:, 2] = x[:, 0] + x[:, 2] / 2
    y[:, 3] = x[:, 1] + x[:, 3] / 2
    return y


def clip_coords(boxes, img_shape):
    boxes[:, 0].clamp_(0, img_shape[1])
    boxes[:, 1].clamp_(0, img_shape[0])
    boxes[:, 2].clamp_(0, img_shape[1])
    boxes[:, 3].clamp_(0, img_shape[0])


def scale_coords(img1_shape, coords, img0_shape, ratio_pad=None):
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    else:
        gain = ratio_pad[0][0]
        pad = ratio_pad[1]
    coords[:, [0, 2]] -= pad[0]
    coords[:, [1, 3]] -= pad[1]
    coords[:, :4] /= gain
    clip_coords(coords, img0_shape)
    return coords


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False, labels=()):
    """(B, N, 5+nc) float32 CUDA tensor -> list of (n, 6) tensors [xyxy, conf, cls] (reference :607-695, best-class path)."""
    if classes is not None or agnostic or multi_label or (labels and len(labels)):
        raise NotImplementedError("only the call tracker/track.py:239 makes (best class, class-offset NMS) is accelerated")
    if not prediction.is_cuda:
        raise L.B2TError("non_max_suppression needs a CUDA tensor: there is no CPU fallback")
    lib = L.load()
    pred = prediction.float().contiguous()
    B, N, no = pred.shape
    max_det, max_nms = 300, 30000
    key = (pred.device, B, N)
    if key not in _ws:
        _ws[key] = torch.empty(lib.b2t_nms_workspace_bytes(B, N, max_nms), dtype=torch.uint8, device=pred.device)
    out = torch.zeros((B, max_det, 6), dtype=torch.float32, device=pred.device)
    cnt = torch.zeros(B, dtype=torch.int32, device=pred.device)
    with torch.cuda.device(pred.device):
        rc = lib.b2t_nms(C.c_void_p(pred.data_ptr()), B, N, no, float(conf_thres), float(iou_thres), max_det, max_nms, N, 0, 1.0, 0.0, 0.0,
                         0.0, 0.0, C.c_void_p(_ws[key].data_ptr()), _ws[key].numel(), C.c_void_p(out.data_ptr()), C.c_void_p(cnt.data_ptr()),
                         C.c_void_p(torch.cuda.current_stream(pred.device).cuda_stream))
    if rc != 0:
        raise L.B2TError((lib.b2t_detect_last_error() or b"").decode())
    n = cnt.tolist()
    return [out[b, :n[b]] for b in range(B)]
