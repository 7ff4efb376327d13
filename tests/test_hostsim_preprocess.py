"""not-gpu: csrc/b2t_preproc.cu executed by the fiber simulator against the committed outputs of the reference's own
pre-processing (tests/golden/letterbox.npz) and against the oracle on more shapes -- bit-exact (8-bit fixed-point resize,
IEEE float / 255).  The `-m gpu` tier repeats it on the nvcc build."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "hostsim"))
from simlib import sim  # noqa: E402
from b200track.preprocess import launch_letterbox, letterbox_geometry  # noqa: E402
from oracle import preprocess as P  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "letterbox.npz")


def _run(imgs, size, stride):
    imgs = np.ascontiguousarray(imgs)
    b, h, w, _ = imgs.shape
    geo = letterbox_geometry((h, w), (size, size), stride)
    out = np.full((b, 3, geo["out_h"], geo["out_w"]), -1.0, dtype=np.float32)
    launch_letterbox(sim(), imgs.ctypes.data, b, h, w, 3 * w, geo, out.ctypes.data, None)
    return out, geo


def test_letterbox_kernel_matches_reference_golden():
    g = np.load(GOLDEN)
    for k, (h, w, size, stride) in enumerate(g["cases"]):
        out, geo = _run(g["img%d" % k][None], int(size), int(stride))
        ref = g["out%d" % k]
        assert out.shape[1:] == ref.shape, (k, out.shape, ref.shape)
        assert np.array_equal(out[0], ref), "case %d: %d values differ" % (k, int((out[0] != ref).sum()))


def test_letterbox_geometry_equals_oracle():
    for shape in [(1080, 1920), (1920, 1080), (720, 1280), (1440, 2560), (375, 1242), (37, 41), (1280, 1280), (2160, 3840)]:
        for size, stride in [(1280, 64), (640, 32), (1280, 32)]:
            a, b = letterbox_geometry(shape, (size, size), stride), P.letterbox_geometry(shape, (size, size), stride)
            assert (a["unpad_w"], a["unpad_h"]) == b["new_unpad"] and a["ratio"] == b["ratio"]
            assert (a["top"], a["bottom"], a["left"], a["right"]) == (b["top"], b["bottom"], b["left"], b["right"])


@pytest.mark.parametrize("shape,size,stride", [((54, 96), 128, 32), ((120, 67), 128, 64), ((61, 33), 96, 32), ((144, 256), 128, 64), ((20, 100), 64, 32)])
def test_letterbox_kernel_batch_vs_oracle(shape, size, stride):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    imgs = rng.integers(0, 256, (2,) + shape + (3,), dtype=np.uint8)
    out, geo = _run(imgs, size, stride)
    for b in range(2):
        ref, _ = P.preprocess(imgs[b], (size, size), stride)
        assert np.array_equal(out[b], ref)


def test_letterbox_argument_errors():
    lib = sim()
    img = np.zeros((4, 4, 3), dtype=np.uint8); out = np.zeros((3, 4, 4), dtype=np.float32)
    assert lib.b2t_letterbox(img.ctypes.data, 1, 4, 4, 8, 4, 4, 0, 0, 4, 4, 114, out.ctypes.data, None) != 0        # pitch < 3 * w
    assert lib.b2t_letterbox(img.ctypes.data, 1, 4, 4, 12, 4, 4, 1, 0, 4, 4, 114, out.ctypes.data, None) != 0       # canvas too small


@pytest.mark.parametrize("act", ["fp16", "bf16"])
@pytest.mark.parametrize("shape,size,stride", [((54, 96), 128, 32), ((120, 67), 128, 64), ((144, 256), 128, 64), ((128, 128), 128, 64)])
def test_letterbox_reorg_fused_equals_letterbox_then_reorg(shape, size, stride, act):
    """b2t_letterbox_reorg (uint8 frame -> the detector's padded ReOrg / NHWC 16-bit input) == the oracle's float canvas pushed through
    ReOrg (models/common.py:52-53) and rounded to fp16 / bf16 by torch, bit for bit; pad pixels of the row layout stay untouched."""
    import ctypes as C
    import torch
    rng = np.random.default_rng(shape[0] + 7 * shape[1])
    imgs = np.ascontiguousarray(rng.integers(0, 256, (2,) + shape + (3,), dtype=np.uint8))
    geo = letterbox_geometry(shape, (size, size), stride)
    H2, W2 = geo["out_h"] // 2, geo["out_w"] // 2
    row = W2 + 8
    out = np.full((2, H2, row, 16), 0x7777, dtype=np.uint16)
    lib = sim()
    rc = lib.b2t_letterbox_reorg(C.c_void_p(imgs.ctypes.data), 2, shape[0], shape[1], 3 * shape[1], geo["unpad_w"], geo["unpad_h"], geo["top"],
                                 geo["left"], geo["out_h"], geo["out_w"], 114, C.c_void_p(out.ctypes.data), row, 1, 1 if act == "fp16" else 0, None)
    assert rc == 0, lib.b2t_detect_last_error()
    tdt = torch.float16 if act == "fp16" else torch.bfloat16
    for b in range(2):
        canvas, _ = P.preprocess(imgs[b], (size, size), stride)                       # (3, H, W) float32
        x = torch.from_numpy(canvas)[None]
        re = torch.cat([x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]], 1)[0]      # (12, H/2, W/2)
        ref = re.permute(1, 2, 0).contiguous().to(tdt).view(torch.int16).numpy().view(np.uint16)
        assert np.array_equal(out[b, :, 1:W2 + 1, :12], ref)
        assert (out[b, :, 1:W2 + 1, 12:] == 0).all()
        assert (out[b, :, 0] == 0x7777).all() and (out[b, :, W2 + 1:] == 0x7777).all()
