"""Writes tests/golden/detector_w6.npz by running the UNMODIFIED reference detector (build container only).

The reference's ``Model('cfg/deploy/yolov7-w6.yaml')`` is built on CPU, fused (models/yolo.py:403-417), loaded
with the seeded weights of ``b200track.w6.seeded_state_dict(0)`` (the reference ships no checkpoint and its default
init yields zero detections -- SURVEY 7.2 #6) and run on a seeded 256x256 image; the reference's own
``utils.general.non_max_suppression`` produces the detections.  Stored: sampled rows of ``model(img)[0]`` and
the NMS output.  The 280 MB of weights are NOT stored: they are regenerated from the seed."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "yolov7-tracker_b200"))
from oracle import refshim                                   # noqa: E402
from b200track.w6 import seeded_state_dict                    # noqa: E402


def main():
    torch.manual_seed(0)
    model = refshim.load_detector_model()
    sd = seeded_state_dict(0)
    own = model.state_dict()
    missing = [k for k in sd if k not in own]
    assert not missing, missing[:5]
    for k, v in sd.items():
        assert own[k].shape == v.shape, (k, own[k].shape, v.shape)
    model.load_state_dict(sd, strict=False)
    g = torch.Generator().manual_seed(123)
    img = torch.rand((1, 3, 256, 256), generator=g)
    with torch.no_grad():
        pred = model(img)[0]
    sys.path.insert(0, refshim.REF_ROOT)
    try:
        import importlib
        for k in list(sys.modules):
            if k == "utils" or k.startswith("utils."):
                sys.modules.pop(k)
        general = importlib.import_module("utils.general")
        dets = general.non_max_suppression(pred.clone(), conf_thres=0.01)[0]
    finally:
        sys.path.remove(refshim.REF_ROOT)
        for k in list(sys.modules):
            if k == "utils" or k.startswith("utils.") or k == "models" or k.startswith("models."):
                sys.modules.pop(k)
    rows = np.arange(0, pred.shape[1], 7)
    np.savez_compressed(os.path.join(HERE, "detector_w6.npz"), pred_shape=np.array(pred.shape), rows=rows,
                        pred_rows=pred[0, rows].numpy(), dets=dets.numpy(), n_candidates=int((pred[0, :, 4] > 0.01).sum()),
                        torch_version=torch.__version__)
    print("pred", tuple(pred.shape), "candidates>0.01:", int((pred[0, :, 4] > 0.01).sum()), "dets", tuple(dets.shape),
          "conf range", float(dets[:, 4].min()) if len(dets) else None, float(dets[:, 4].max()) if len(dets) else None)


if __name__ == "__main__":
    main()
