"""not-gpu: the torch oracle of the detector against the committed outputs of the reference's own Model
(tests/golden/detector_w6.npz) -- pins both oracle/detector.py and the product's graph builder (w6_layers)."""
import os

import numpy as np
import pytest
import torch

from b200track.w6 import ANCHORS, STRIDES, conv_shapes, layer_channels, seeded_state_dict, w6_layers
from oracle import detector as OD

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "detector_w6.npz")


def test_graph_bookkeeping():
    layers = w6_layers()
    assert len(layers) == 119 and layers[-1][1] == "detect"
    ch = layer_channels(layers)
    assert ch[46] == 1024 and ch[47] == 512 and ch[83] == 128 and ch[117] == 1024 and ch[9] == 256 and ch[58] == 1536
    shapes = conv_shapes(layers)
    assert len(shapes) == 107                                    # 107 convs after fusion
    n_params = sum(ci * co * k * k + co for _, ci, co, k, s, _ in shapes)
    assert abs(n_params - 70.43e6) / 70.43e6 < 0.01              # 70.4 M parameters


def test_oracle_forward_and_nms_match_reference_golden():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    g = np.load(GOLDEN)
    sd = seeded_state_dict(0)
    gen = torch.Generator().manual_seed(123)
    img = torch.rand((1, 3, 256, 256), generator=gen)
    with torch.no_grad():
        pred = OD.forward(w6_layers(), sd, img, ANCHORS, STRIDES)
    assert tuple(pred.shape) == tuple(g["pred_shape"])
    np.testing.assert_allclose(pred[0, g["rows"]].numpy(), g["pred_rows"], rtol=2e-3, atol=2e-3)
    assert int((pred[0, :, 4] > 0.01).sum()) == int(g["n_candidates"])
    dets = OD.non_max_suppression(pred, conf_thres=0.01)[0].numpy()
    ref = g["dets"]
    assert dets.shape == ref.shape
    assert np.array_equal(dets[:, 5], ref[:, 5])                 # classes and order: exact
    np.testing.assert_allclose(dets[:, :5], ref[:, :5], rtol=2e-3, atol=2e-2)
