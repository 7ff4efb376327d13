"""-m gpu: b2t_letterbox (csrc/b2t_preproc.cu) on the nvcc build, through b200track.preprocess.Letterbox: bit-exact against
the committed outputs of the reference's own pre-processing (tracker/tracker_dataloader.py:64-130) and against the oracle at
the full 1080p -> 1280 size of BASELINE's frames.  Integer resize arithmetic + IEEE float / 255: tolerance 0."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "letterbox.npz")


def test_letterbox_matches_reference_golden():
    from b200track.preprocess import Letterbox
    g = np.load(GOLDEN)
    for k, (h, w, size, stride) in enumerate(g["cases"]):
        out, geo = Letterbox(int(size), int(stride))(g["img%d" % k])
        torch.cuda.synchronize()
        ref = g["out%d" % k]
        assert tuple(out.shape[1:]) == ref.shape
        assert np.array_equal(out[0].cpu().numpy(), ref), "case %d: %d values differ" % (k, int((out[0].cpu().numpy() != ref).sum()))


def test_letterbox_1080p_batch_matches_oracle():
    from b200track.preprocess import Letterbox
    from oracle import preprocess as P
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (2, 1080, 1920, 3), dtype=np.uint8)
    lb = Letterbox(1280, 64)
    out, geo = lb(torch.from_numpy(frames).cuda())                  # device-resident uint8 input
    torch.cuda.synchronize()
    assert tuple(out.shape) == (2, 3, 768, 1280) and (geo["top"], geo["left"]) == (24, 0)
    for b in range(2):
        ref, _ = P.preprocess(frames[b], (1280, 1280), 64)
        assert np.array_equal(out[b].cpu().numpy(), ref)
    out2, _ = lb(frames)                                            # host frames through pinned memory: same result
    torch.cuda.synchronize()
    assert torch.equal(out, out2)


def test_letterboxed_frame_through_non_square_detector():
    """uint8 frame -> b2t_letterbox (stride 128 so that both sides suit ReOrg + stride 64) -> DetectorW6 on the resulting
    384 x 512 rectangle: logits against the bf16-emulating oracle (statistically, like the 640 x 640 test), decode + NMS exactly."""
    from b200track.detector import DetectorW6
    from b200track.preprocess import Letterbox
    from b200track.w6 import ANCHORS, STRIDES, calibrated_state_dict, w6_layers
    from oracle import detector as OD
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    rng = np.random.default_rng(11)
    frame = rng.integers(0, 256, (270, 480, 3), dtype=np.uint8)
    img, geo = Letterbox(512, 128)(frame)
    assert tuple(img.shape) == (1, 3, 384, 512) and (geo["top"], geo["left"]) == (48, 0)
    sd = calibrated_state_dict(0, 512, "cuda", img=img)             # LSUV pass on THIS canvas (gray borders): the 300-row cap is hit
    det = DetectorW6(sd, batch=1, img_size=(384, 512), use_graph=False)
    pred = det.forward(img).clone()
    out, cnt = det.detect(img, post=True)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref_bf, raw_bf = OD.forward(w6_layers(), sd, img, ANCHORS, STRIDES, emulate_bf16=det.act_dtype, return_raw=True)
    assert tuple(pred.shape) == tuple(ref_bf.shape)
    for lvl, r in enumerate(raw_bf):
        got = det.raw[lvl][..., :255].reshape(1, r.shape[2], r.shape[3], 3, 85).permute(0, 3, 1, 2, 4)
        err = (got - r).abs()
        rel_rms = float((err ** 2).mean().sqrt() / r.std())
        assert float(err.mean()) < 0.12 and rel_rms < 0.12, "level %d: mean %.4f rel rms %.3f" % (lvl, err.mean(), rel_rms)
    ref = OD.post_process(OD.non_max_suppression(pred, conf_thres=0.01)[0], (384, 512))
    n = int(cnt[0])
    assert n == ref.shape[0] and n > 0
    assert torch.equal(out[0, :n, 5], ref[:, 5]) and torch.allclose(out[0, :n, :5], ref[:, :5], atol=1e-3)
    assert float(out[0, :n, 2].max()) <= 512 and float(out[0, :n, 3].max()) <= 384      # clipped to (W, H)
