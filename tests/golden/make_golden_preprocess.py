"""Writes tests/golden/letterbox.npz: outputs of the UNMODIFIED reference pre-processing
(tracker/tracker_dataloader.py: TrackerLoader._letterbox + the BGR->RGB / CHW / float / 255 lines of __getitem__)
on small seeded images.  Build container only (needs /root/reference and opencv-python).

    python tests/golden/make_golden_preprocess.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF = os.environ.get("B2T_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_loader():
    spec = importlib.util.spec_from_file_location("ref_tracker_dataloader", os.path.join(REF, "tracker", "tracker_dataloader.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.TrackerLoader


def reference_preprocess(TrackerLoader, ori_img, height, width, stride):
    # __getitem__, 'v7' branch (tracker_dataloader.py:78-86) without the file read
    img = TrackerLoader._letterbox(None, ori_img, new_shape=(height, width), stride=stride)[0]
    img = img[:, :, ::-1].transpose(2, 0, 1)
    img = np.ascontiguousarray(img)
    t = torch.from_numpy(img).float()
    t /= 255.0
    return t.numpy()


CASES = [  # (h, w, new_size, stride)
    (108, 192, 128, 64),      # 16:9 down-scale, vertical minimum-rectangle padding
    (192, 108, 128, 64),      # portrait
    (90, 160, 256, 64),       # up-scale
    (256, 256, 128, 64),      # exact 2 x 2 down-scale: OpenCV switches to INTER_AREA
    (128, 128, 128, 64),      # no resize at all
    (75, 250, 256, 32),       # wide strip, odd sizes, stride 32
    (37, 41, 64, 64),         # tiny
]

if __name__ == "__main__":
    TL = load_loader()
    rng = np.random.default_rng(2024)
    out = {"cases": np.array(CASES, dtype=np.int64)}
    for k, (h, w, size, stride) in enumerate(CASES):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        out["img%d" % k] = img
        out["out%d" % k] = reference_preprocess(TL, img, size, size, stride)
    np.savez_compressed(os.path.join(HERE, "letterbox.npz"), **out)
    print("wrote", os.path.join(HERE, "letterbox.npz"), {k: v.shape for k, v in out.items() if k.startswith("out")})
