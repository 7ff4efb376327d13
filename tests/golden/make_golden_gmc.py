"""Build container only: runs the UNMODIFIED reference estimator (tracker/botsort.py ``GMC(method='orb', downscale=2)``, host OpenCV)
over seeded moving frames with masked detections and stores its 2 x 3 matrices in tests/golden/gmc.npz.  The frames are regenerated
from the seeds by the tests (b200track.synth.textured_frame + integer rolls: no OpenCV needed to rebuild them)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "yolov7-tracker_b200")):
    sys.path.insert(0, p)
from b200track.synth import textured_frame  # noqa: E402
from oracle import refshim  # noqa: E402

CASES = [dict(seed=301, h=360, w=640, shifts=[(0, 0), (3, -2), (5, 1), (4, 6)]),
         dict(seed=302, h=481, w=643, shifts=[(0, 0), (-2, 2), (-5, 7), (-9, 3)])]


def frames_and_dets(case):
    base = textured_frame(case["seed"], case["h"], case["w"], n_rect=300)
    rng = np.random.default_rng(case["seed"])
    x1 = rng.uniform(0, case["w"] - 100, 12); y1 = rng.uniform(0, case["h"] - 150, 12)
    dets = np.round(np.stack([x1, y1, x1 + rng.uniform(20, 90, 12), y1 + rng.uniform(40, 140, 12)], 1)).astype(np.float32)
    dets = np.concatenate([dets, np.linspace(0.9, 0.25, 12, dtype=np.float32)[:, None], np.zeros((12, 1), np.float32)], 1)
    return [np.ascontiguousarray(np.roll(base, s, (0, 1))) for s in case["shifts"]], dets


if __name__ == "__main__":
    out = {}
    for k, case in enumerate(CASES):
        gmc = refshim.load().botsort.GMC(method='orb', downscale=2)
        frames, dets = frames_and_dets(case)
        out["H%d" % k] = np.stack([gmc.apply(f, dets) for f in frames]).astype(np.float64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gmc.npz"), **out)
    print({k: v.shape for k, v in out.items()}, out["H0"][1])
