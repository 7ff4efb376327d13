"""Build container only: runs the UNMODIFIED reference extractor (tracker/reid_models/deepsort_reid.py ``Extractor`` with the
reference's own checkpoint weights/ckpt.t7, CPU) on seeded crops and stores the features in tests/golden/reid.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reid as R  # noqa: E402

REF = os.environ.get("B2T_REFERENCE_ROOT", "/root/reference")


def reference_extractor():
    sys.path.insert(0, os.path.join(REF, "tracker"))
    try:
        from reid_models.deepsort_reid import Extractor
    finally:
        sys.path.remove(os.path.join(REF, "tracker"))
        for k in [k for k in sys.modules if k.startswith("reid_models")]:
            del sys.modules[k]
    return Extractor(os.path.join(REF, "weights", "ckpt.t7"), use_cuda=False)


if __name__ == "__main__":
    ext = reference_extractor()
    crops = R.seeded_crops(5, 6)
    feats = ext(crops)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reid.npz"), features=feats.astype(np.float32), seed=5, n=6)
    print("wrote tests/golden/reid.npz", feats.shape, float(np.linalg.norm(feats, axis=1).mean()))
