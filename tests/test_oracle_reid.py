"""Pins oracle/reid.py (restatement of tracker/reid_models/deepsort_reid.py) against the UNMODIFIED reference: the committed golden
features made with the reference's own checkpoint (tests/golden/reid.npz), and -- build container only -- the reference classes run
live with seeded weights and with weights/ckpt.t7.  CPU tier."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "yolov7-tracker_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from oracle import reid as R  # noqa: E402

REF = os.environ.get("B2T_REFERENCE_ROOT", "/root/reference")
HAVE_REF = os.path.exists(os.path.join(REF, "weights", "ckpt.t7"))


def _ref_module():
    sys.path.insert(0, os.path.join(REF, "tracker"))
    try:
        from reid_models import deepsort_reid as M
    finally:
        sys.path.remove(os.path.join(REF, "tracker"))
        for k in [k for k in sys.modules if k.startswith("reid_models")]:
            del sys.modules[k]
    return M


@pytest.mark.skipif(not HAVE_REF, reason="reference tree / checkpoint not present")
def test_restatement_equals_reference_net_and_preprocess():
    M = _ref_module()
    crops = R.seeded_crops(9, 5)
    ext = M.Extractor(os.path.join(REF, "weights", "ckpt.t7"), use_cuda=False)
    x_ref = ext._preprocess(crops)
    x = R.preprocess(crops)
    assert x.shape == x_ref.shape and float((x - x_ref).abs().max()) < 2e-6                  # cv2.resize on float data, restated
    sd = {k: v.float() for k, v in ext.net.state_dict().items() if v.dtype.is_floating_point}
    assert ext.net.training                                     # the quirk: the reference never switches its extractor to eval()
    with torch.no_grad():
        f = R.forward(sd, x, batch_stats=True)                   # (before the reference call: that one also updates the running buffers)
        f_ref = ext.net(x_ref)
        f_eval_ref = ext.net.eval()(x_ref)
        ext.net.train()
    assert float((f - f_ref).abs().max()) < 2e-5
    sd_after = {k: v.float() for k, v in ext.net.state_dict().items() if v.dtype.is_floating_point}
    with torch.no_grad():
        assert float((R.forward(sd_after, x, batch_stats=False) - f_eval_ref).abs().max()) < 2e-5
    # seeded weights (what the GPU tier uses) through the reference's own Net
    net = M.Net(reid=True).eval()
    seeded = R.seeded_state_dict(3)
    missing = net.load_state_dict(seeded, strict=False)
    assert not [k for k in missing.missing_keys if "classifier" not in k and "num_batches_tracked" not in k] and not missing.unexpected_keys
    with torch.no_grad():
        assert float((net(x) - R.forward(seeded, x, batch_stats=False)).abs().max()) < 2e-5
        assert float((net.train()(x) - R.forward(seeded, x, batch_stats=True)).abs().max()) < 2e-5


@pytest.mark.skipif(not HAVE_REF, reason="reference tree / checkpoint not present")
def test_golden_features_reproduce_with_the_checkpoint():
    g = np.load(os.path.join(ROOT, "tests", "golden", "reid.npz"))
    ck = torch.load(os.path.join(REF, "weights", "ckpt.t7"), map_location="cpu", weights_only=False)["net_dict"]
    sd = {k: v.float() for k, v in ck.items() if v.dtype.is_floating_point}
    with torch.no_grad():
        f = R.forward(sd, R.preprocess(R.seeded_crops(int(g["seed"]), int(g["n"]))), batch_stats=True)
    np.testing.assert_allclose(f.numpy(), g["features"], rtol=0, atol=2e-5)


def test_seeded_network_is_well_conditioned():
    """The seeded stand-in weights give distinct, unit-norm features (the GPU parity test is meaningful on them)."""
    with torch.no_grad():
        f = R.forward(R.seeded_state_dict(3), R.preprocess(R.seeded_crops(1, 6)), batch_stats=True)
    assert torch.allclose(f.norm(dim=1), torch.ones(6), atol=1e-5)
    cos = (f @ f.T).numpy()
    assert np.all(cos[~np.eye(6, dtype=bool)] < 0.999)
