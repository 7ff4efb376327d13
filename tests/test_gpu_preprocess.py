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
