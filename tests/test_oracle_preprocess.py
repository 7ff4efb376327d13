"""not-gpu: oracle/preprocess.py against (a) the committed outputs of the reference's own pre-processing
(tests/golden/letterbox.npz, written by make_golden_preprocess.py from tracker/tracker_dataloader.py) and (b) the real
cv2.resize when opencv is importable (it is in the build container) -- SURVEY.md section 8f row 2."""
import os

import numpy as np
import pytest

from oracle import preprocess as P

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "letterbox.npz")


def test_oracle_preprocess_matches_reference_golden():
    g = np.load(GOLDEN)
    for k, (h, w, size, stride) in enumerate(g["cases"]):
        got, geo = P.preprocess(g["img%d" % k], (int(size), int(size)), int(stride))
        ref = g["out%d" % k]
        assert got.shape == ref.shape, (k, got.shape, ref.shape)
        assert got.dtype == np.float32 and np.array_equal(got, ref), "case %d differs from the reference" % k
        assert got.shape[1] % int(stride) == 0 and got.shape[2] % int(stride) == 0          # minimum rectangle (auto=True)


def test_letterbox_geometry_examples():
    g = P.letterbox_geometry((1080, 1920), (1280, 1280), 64)
    assert g["new_unpad"] == (1280, 720) and (g["top"], g["bottom"], g["left"], g["right"]) == (24, 24, 0, 0)
    g = P.letterbox_geometry((1280, 1280), (1280, 1280), 64)
    assert g["new_unpad"] == (1280, 1280) and (g["top"], g["bottom"], g["left"], g["right"]) == (0, 0, 0, 0)


def test_resize_restatement_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    for (h, w, dst) in [(108, 192, (128, 72)), (54, 96, (128, 72)), (144, 256, (128, 72)), (100, 37, (53, 91)), (17, 9, (4, 8)),
                        (216, 384, (128, 72)), (50, 50, (49, 51)), (64, 64, (128, 128))]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(P.resize_linear_u8(img, dst), cv2.resize(img, dst, interpolation=cv2.INTER_LINEAR)), (h, w, dst)
