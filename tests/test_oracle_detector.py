"""not-gpu: the torch oracle of the detector against the committed outputs of the reference's own Model
(tests/golden/detector_w6.npz) -- pins both oracle/detector.py and the product's graph builder (w6_layers)."""
import os

import numpy as np
import pytest
import torch

from b200track.w6 import ANCHORS, STRIDES, conv_shapes, layer_channels, seeded_state_dict, w6_layers
from oracle import detector as OD, refshim

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


def test_stackable_1x1_pairs_and_rowpack_layout():
    """Host logic of the detector planner: the 11 ELAN blocks each open with two 1x1 convs of the same input that are adjacent
    in the block's concat; the row-packed stem weight layout is k = kh*64 + kw*16 + c with a zero fourth pixel."""
    from b200track.w6 import stackable_pairs
    pairs = stackable_pairs()
    assert len(pairs) == 11 and pairs[0] == (3, 4) and (12, 13) in pairs and (50 + 2, 50 + 3) in pairs
    layers = w6_layers()
    for a, b in pairs:
        assert layers[a][1] == layers[b][1] == "conv" and layers[a][3] == layers[b][3] and layers[b][2] == -2
    from b200track.conv import pack_conv_weight_rowpack
    w = torch.arange(64 * 12 * 9, dtype=torch.float32).reshape(64, 12, 3, 3) / 4096.0
    pk = pack_conv_weight_rowpack(w).float().reshape(64, 3, 4, 16)
    assert torch.equal(pk[:, :, :3, :12], w.to(torch.bfloat16).float().permute(0, 2, 3, 1))
    assert bool((pk[:, :, 3] == 0).all()) and bool((pk[:, :, :, 12:] == 0).all())


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


def test_fold_batchnorm_and_implicit_layers_math():
    """fold_reference_state_dict on a synthetic one-layer problem: Conv + BN(eval) and ia / m / im head == the folded convs."""
    import torch.nn.functional as F
    from b200track.w6 import conv_shapes, fold_reference_state_dict
    g = torch.Generator().manual_seed(1)
    sd = {}
    for name, cin, cout, k, s, act in conv_shapes():
        if act:
            sd[name + ".weight"] = torch.randn((cout, cin, k, k), generator=g) * 0.05
            pre = name[:-len(".conv")]
            sd[pre + ".bn.weight"] = torch.rand(cout, generator=g) + 0.5
            sd[pre + ".bn.bias"] = torch.randn(cout, generator=g) * 0.1
            sd[pre + ".bn.running_mean"] = torch.randn(cout, generator=g) * 0.1
            sd[pre + ".bn.running_var"] = torch.rand(cout, generator=g) + 0.2
        else:
            j = int(name.rsplit(".", 1)[1])
            sd["model.122.m.%d.weight" % j] = torch.randn((cout, cin, 1, 1), generator=g) * 0.05
            sd["model.122.m.%d.bias" % j] = torch.randn(cout, generator=g) * 0.1
            sd["model.122.m2.%d.weight" % j] = torch.zeros((cout, 8, 1, 1))
            sd["model.122.ia.%d.implicit" % j] = torch.randn((1, cin, 1, 1), generator=g) * 0.02
            sd["model.122.im.%d.implicit" % j] = 1 + torch.randn((1, cout, 1, 1), generator=g) * 0.02
    out = fold_reference_state_dict(sd)
    assert set(out) == {n + s for n, *_ in conv_shapes() for s in (".weight", ".bias")}
    name = "model.5.conv"                                             # a 3x3 Conv + BN
    x = torch.randn((1, 64, 9, 9), generator=g)
    pre = "model.5"
    ref = F.batch_norm(F.conv2d(x, sd[name + ".weight"], None, padding=1), sd[pre + ".bn.running_mean"], sd[pre + ".bn.running_var"],
                       sd[pre + ".bn.weight"], sd[pre + ".bn.bias"], False, 0.03, 1e-3)
    got = F.conv2d(x, out[name + ".weight"], out[name + ".bias"], padding=1)
    assert torch.allclose(got, ref, atol=1e-5)
    x = torch.randn((1, 256, 5, 5), generator=g)                      # head 0: m(ia + x) * im
    ref = F.conv2d(x + sd["model.122.ia.0.implicit"], sd["model.122.m.0.weight"], sd["model.122.m.0.bias"]) * sd["model.122.im.0.implicit"]
    got = F.conv2d(x, out["model.118.m.0.weight"], out["model.118.m.0.bias"])
    assert torch.allclose(got, ref, atol=1e-5)
    assert fold_reference_state_dict(out).keys() == out.keys()        # an already fused dict passes through


@pytest.mark.skipif(not os.path.isdir("/root/reference/cfg"), reason="needs the reference tree (build container)")
def test_fold_training_checkpoint_equals_reference_inference():
    """The reference's own TRAINING graph (cfg/training/yolov7-w6.yaml: Conv + BN, aux heads, IAuxDetect with implicit layers), with
    randomised BN statistics, evaluated by the reference in eval mode == the oracle forward on the folded deploy-graph weights."""
    from b200track.w6 import fold_reference_state_dict
    from oracle import refshim
    model = refshim.load_detector_model("cfg/training/yolov7-w6.yaml", fuse=False)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) * 0.5 + 0.75)
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.5 + 1.2)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
    img = torch.rand((1, 3, 128, 128), generator=g)
    with torch.no_grad():
        ref = model(img)[0]
    folded = fold_reference_state_dict(model.state_dict())
    with torch.no_grad():
        got = OD.forward(w6_layers(), folded, img, ANCHORS, STRIDES)
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, rtol=2e-3, atol=2e-3), float((got - ref).abs().max())


def test_pickled_reference_checkpoint_is_converted_and_loaded(tmp_path):
    """models/experimental.py:83-106 loads a pickled nn.Module.  The drop-in attempt_load refuses such a file with a message that
    names tools/export_state_dict.py; the converter (run with the reference on sys.path) writes a state dict that the drop-in
    loader folds to exactly the fused weights the reference's own fuse() produces.  Build container only (needs /root/reference)."""
    import subprocess
    import sys
    if not refshim.available():
        pytest.skip("reference tree not present")
    ckpt, fused_pt, out = tmp_path / "w6_pickled.pt", tmp_path / "w6_fused.pt", tmp_path / "w6_state.pt"
    stubs = ("import sys, types\n"
             "for n in ('matplotlib', 'matplotlib.pyplot', 'seaborn'):\n"
             "    m = types.ModuleType(n); m.use = lambda *a, **k: None; m.rc = lambda *a, **k: None; sys.modules.setdefault(n, m)\n")
    make = stubs + ("import os, copy, torch\nos.chdir(%r); sys.path.insert(0, %r)\nfrom models.yolo import Model\n"
                    "torch.manual_seed(3)\nmodel = Model('cfg/deploy/yolov7-w6.yaml', ch=3, nc=80).float().eval()\n"
                    "with torch.no_grad():\n"
                    "    for m in model.modules():\n"
                    "        if isinstance(m, torch.nn.BatchNorm2d):\n"
                    "            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5); m.weight.normal_(1, 0.1); m.bias.normal_(0, 0.1)\n"
                    "torch.save({'model': model, 'ema': None, 'epoch': -1}, %r)\n"
                    "torch.save(copy.deepcopy(model).fuse().state_dict(), %r)\n" % (refshim.REF_ROOT, refshim.REF_ROOT, str(ckpt), str(fused_pt)))
    r = subprocess.run([sys.executable, "-c", make], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "yolov7-tracker_b200")
    code = ("import sys; sys.path.insert(0, %r); from models.experimental import attempt_load\n"
            "try:\n    attempt_load(%r, map_location='cpu')\nexcept RuntimeError as e:\n    print('MSG', str(e))\n" % (pkg, str(ckpt)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "tools/export_state_dict.py" in r.stdout, r.stdout + r.stderr
    conv = stubs + "sys.argv = ['export_state_dict.py', '--reference', %r, '--weights', %r, '--out', %r]\nimport runpy; runpy.run_path(%r, run_name='__main__')\n" % (
        refshim.REF_ROOT, str(ckpt), str(out), os.path.join(root, "tools", "export_state_dict.py"))
    r = subprocess.run([sys.executable, "-c", conv], capture_output=True, text=True)        # (matplotlib / seaborn are absent from this image: stubbed)
    assert r.returncode == 0, r.stderr[-2000:]
    from b200track.w6 import fold_reference_state_dict
    folded = fold_reference_state_dict(torch.load(str(out), weights_only=True))
    ref_sd = torch.load(str(fused_pt), weights_only=True)
    assert len(folded) == 214
    for k, v in folded.items():
        assert torch.allclose(v, ref_sd[k].float(), rtol=1e-5, atol=1e-6), k


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
def test_tiny_oracle_equals_reference_model():
    """YOLOv7-tiny (cfg/deploy/yolov7-tiny.yaml: LeakyReLU convs, MP / SP pools, three-level Detect): the oracle's restatement of the
    graph against the reference's own ``Model`` (fused, eval, CPU) on the same seeded weights."""
    from b200track import tiny
    sd = tiny.seeded_state_dict(0)
    model = refshim.load_detector_model("cfg/deploy/yolov7-tiny.yaml")
    msd = model.state_dict()
    assert all(k in msd and msd[k].shape == v.shape for k, v in sd.items()), [k for k in sd if k not in msd][:3]
    model.load_state_dict(sd, strict=False)
    img = torch.rand((2, 3, 192, 256), generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        ref = model(img)[0]
        got = OD.forward(tiny.tiny_layers(), sd, img, tiny.ANCHORS, tiny.STRIDES, act="leaky", name_offset=-1)
    assert ref.shape == got.shape == (2, 3 * (24 * 32 + 12 * 16 + 6 * 8), 85)
    assert float((ref - got).abs().max()) < 2e-3 and float((ref[..., 4] - got[..., 4]).abs().max()) < 1e-5
