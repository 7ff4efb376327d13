"""not-gpu: csrc/b2t_gmc.cu executed by the fiber simulator against oracle/gmc.py (itself pinned against cv2 and the reference's
GMC class, tests/test_oracle_gmc.py): gray image, key points and descriptors bit for bit, the estimated warp against the
restated RANSAC (same sampling sequence: 1e-9) and against cv2.estimateAffinePartial2D (its own run-to-run spread).
The `-m gpu` tier repeats it on the nvcc build at full frame sizes."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "hostsim"))
from simlib import sim, ptr  # noqa: E402
from b200track import _lib as L  # noqa: E402
from b200track import gmc as G  # noqa: E402
from b200track.synth import textured_frame  # noqa: E402
from oracle import gmc as OG  # noqa: E402


class SimGmc:
    def __init__(self, n_seq, h, w, ds=2, max_kp=4096):
        self.lib = sim()
        self.S, self.h, self.w, self.ds, self.max_kp = n_seq, h, w, ds, max_kp
        n = self.lib.b2t_gmc_workspace_bytes(n_seq, h, w, ds, max_kp)
        assert n > 0
        self.layout = G.workspace_layout(self.lib, n_seq, h, w, ds, max_kp)
        self.mem = np.zeros(n + 256, np.uint8)
        self.off = (-self.mem.ctypes.data) % 256
        self.warps = np.zeros((n_seq, 2, 3), np.float64)
        self.stat = np.zeros((n_seq, L.GMC_STAT_WORDS), np.int32)

    def estimate(self, frames, dets=None, counts=None, thresh=-np.inf):
        frames = np.ascontiguousarray(frames)
        dmax = 0 if dets is None else dets.shape[1]
        G.launch_estimate(self.lib, frames.ctypes.data, self.S, self.h, self.w, 3 * self.w, self.ds, ptr(dets), ptr(counts), dmax, thresh,
                          self.mem.ctypes.data + self.off, self.max_kp, self.warps.ctypes.data, self.stat.ctypes.data, None)
        return self.warps.copy(), self.stat.copy()

    def keypoints(self, seq):
        ws = self.mem[self.off:].tobytes()
        state = np.frombuffer(ws, np.int32, 16, seq * self.layout["stride"] + self.layout["state"])
        buf = (int(state[0]) - 1) & 1
        return G.unpack_keypoints(ws, self.layout, seq, buf, int(state[1 + buf]), self.max_kp)

    def plane(self, seq, name):
        ws = self.mem[self.off:]
        o = seq * self.layout["stride"] + self.layout[name]
        return ws[o:o + self.layout["h"] * self.layout["w"]].reshape(self.layout["h"], self.layout["w"]).copy()


def shifted(frame, dx, dy):
    return np.ascontiguousarray(np.roll(frame, (dy, dx), axis=(0, 1)))


@pytest.mark.parametrize("shape", [(200, 300), (201, 303)])
def test_gray_fast_orb_stages_equal_oracle(shape):
    h, w = shape
    f = textured_frame(21, h, w, n_rect=150)
    dets = np.zeros((1, 4, 6), np.float32)
    dets[0, 0] = [40, 30, 120, 150, 0.9, 0]
    dets[0, 1] = [200, 100, 260, 180, 0.1, 0]           # below the threshold: not masked
    dets[0, 2] = [150, 20, 190, 60, 0.5, 1]
    cnt = np.array([3], np.int32)
    g = SimGmc(1, h, w)
    _, stat = g.estimate(f[None], dets, cnt, thresh=0.2)
    orc = OG.GMCOracle()
    gray, xs, ys, desc = orc.stages(f, dets[0, [0, 2]])
    assert np.array_equal(g.plane(0, "gray"), gray)
    assert np.array_equal(g.plane(0, "blur"), OG.orb_blur(gray))
    kx, ky, kd = g.keypoints(0)
    assert len(xs) > 50 and stat[0, 0] == len(xs) and stat[0, 5] == L.GMC_FIRST_FRAME
    assert np.array_equal(kx, xs) and np.array_equal(ky, ys)
    assert np.array_equal(kd, desc)


def test_estimate_vs_oracle_two_sequences():
    h, w = 240, 360
    base = [textured_frame(31 + s, h, w, n_rect=200) for s in range(2)]
    moves = [[(0, 0), (4, -2), (-6, 2)], [(0, 0), (-2, 6), (2, 2)]]
    g = SimGmc(2, h, w)
    own = [OG.GMCOracle(estimator="restated") for _ in range(2)]
    ref = [OG.GMCOracle(estimator="cv2") for _ in range(2)]
    for k in range(3):
        frames = np.stack([shifted(base[s], *moves[s][k]) for s in range(2)])
        base = [frames[0], frames[1]]
        warps, stat = g.estimate(frames)
        for s in range(2):
            Ho, Hr = own[s].apply(frames[s]), ref[s].apply(frames[s])
            np.testing.assert_allclose(warps[s], Ho, rtol=0, atol=1e-9)
            if k:
                assert stat[s, 3] == len(own[s].last["src"]) and stat[s, 3] > 20
                assert abs(warps[s, 0, 2] - moves[s][k][0]) < 0.2 and abs(warps[s, 1, 2] - moves[s][k][1]) < 0.2      # the true shift
                assert np.abs(warps[s, :, :2] - Hr[:, :2]).max() < 1e-3 and np.abs(warps[s, :, 2] - Hr[:, 2]).max() < 0.25
            else:
                assert np.array_equal(warps[s], np.eye(2, 3)) and stat[s, 5] == L.GMC_FIRST_FRAME


def test_few_points_and_truncation_flags():
    h, w = 160, 200
    flat = np.full((1, h, w, 3), 90, np.uint8)
    g = SimGmc(1, h, w, max_kp=64)
    g.estimate(flat)
    warps, stat = g.estimate(flat)
    assert np.array_equal(warps[0], np.eye(2, 3)) and stat[0, 5] & L.GMC_FEW_POINTS and stat[0, 0] == 0
    g = SimGmc(1, 240, 360, max_kp=64)
    warps, stat = g.estimate(textured_frame(5, 240, 360, n_rect=200)[None])
    assert stat[0, 5] & L.GMC_TRUNCATED and stat[0, 0] == 64
    lib = sim()
    assert lib.b2t_gmc_workspace_bytes(1, 60, 60, 2, 64) == 0                     # too small for ORB's 31-pixel border
    assert lib.b2t_gmc_estimate(None, 1, h, w, 3 * w, 2, None, None, 0, 0.0, None, 64, None, None, None) != 0


def test_kernel_vs_committed_reference_golden():
    """The estimator kernels (simulator) against the matrices the UNMODIFIED reference class produced (tests/golden/gmc.npz): identical
    key points and matches by construction (tests above), the matrix within the spread of OpenCV's RANSAC sampling."""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_gmc import CASES, frames_and_dets
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "gmc.npz"))
    case = CASES[0]
    frames, dets = frames_and_dets(case)
    g = SimGmc(1, case["h"], case["w"], max_kp=8192)
    d = dets[None].copy(); cnt = np.array([len(dets)], np.int32)
    for i, f in enumerate(frames):
        warps, stat = g.estimate(f[None], d, cnt, thresh=-np.inf)
        ref = gold["H0"][i]
        assert OG.corner_displacement(warps[0], ref, case["h"], case["w"]) < 1.0, (i, warps[0], ref)       # px, at the frame's corners


def test_stages_property_sweep():
    """hypothesis: random frame sizes (odd / even) and contents (noise, blocky, sparse dots, saturated) -- gray image, smoothed image,
    key points and descriptors of the kernels equal the oracle's bit for bit."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=12, deadline=None, derandomize=True)
    @given(seed=st.integers(0, 10 ** 6), h=st.integers(132, 200), w=st.integers(132, 240), mode=st.sampled_from(["noise", "blocks", "dots", "saturated"]))
    def run(seed, h, w, mode):
        rng = np.random.default_rng(seed)
        if mode == "noise":
            f = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        elif mode == "blocks":
            f = np.kron(rng.integers(0, 256, (h // 6 + 1, w // 6 + 1, 3), dtype=np.uint8), np.ones((6, 6, 1), np.uint8))[:h, :w]
        elif mode == "dots":
            f = np.full((h, w, 3), 40, np.uint8)
            ys, xs = rng.integers(0, h, 300), rng.integers(0, w, 300)
            f[ys, xs] = rng.integers(150, 256, (300, 3), dtype=np.uint8)
        else:
            f = np.where(rng.random((h, w, 1)) < 0.5, 0, 255).astype(np.uint8).repeat(3, 2)
        f = np.ascontiguousarray(f)
        g = SimGmc(1, h, w, max_kp=8192)
        _, stat = g.estimate(f[None])
        gray, xs, ys, desc = OG.GMCOracle().stages(f, None)
        assert np.array_equal(g.plane(0, "gray"), gray) and np.array_equal(g.plane(0, "blur"), OG.orb_blur(gray))
        kx, ky, kd = g.keypoints(0)
        assert stat[0, 0] == len(xs) and not (stat[0, 5] & L.GMC_TRUNCATED)
        assert np.array_equal(kx, xs) and np.array_equal(ky, ys) and np.array_equal(kd, desc)
    run()
