"""``attempt_load`` (reference models/experimental.py:83-106) for the B200 detector.

The reference un-pickles a whole ``nn.Module`` from the checkpoint (key 'ema' or 'model').  Pickled reference modules
cannot be loaded without the reference's own class definitions, so this loader accepts
  * a ``.pt`` file holding a *state dict* of the fused reference model (``model.fuse().state_dict()``), or a dict
    with that state dict under 'model' / 'ema' / 'state_dict';
  * the string ``'seeded:<seed>[:<img_size>]'`` -- the LSUV-calibrated random weights the benches use (the reference
    ships no detector checkpoint)."""
import torch

from . import _b2t_path  # noqa: F401
from .yolo import Model


def attempt_load(weights, map_location=None):
    device = map_location if map_location is not None else "cuda:0"
    if isinstance(weights, (list, tuple)):
        if len(weights) != 1:
            raise NotImplementedError("model ensembles are out of scope")
        weights = weights[0]
    if isinstance(weights, str) and weights.startswith("seeded:"):
        from b200track.w6 import calibrated_state_dict
        parts = weights.split(":")
        sd = calibrated_state_dict(int(parts[1]), int(parts[2]) if len(parts) > 2 else 1280, device)
    else:
        ckpt = torch.load(weights, map_location="cpu", weights_only=True)
        sd = ckpt
        for key in ("ema", "model", "state_dict"):
            if isinstance(ckpt, dict) and key in ckpt and isinstance(ckpt[key], dict):
                sd = ckpt[key]
                break
    return Model(device=device).load_state_dict({k: v.float() for k, v in sd.items()})
