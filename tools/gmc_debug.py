"""GPU box diagnostic: every plane of the camera-motion estimator against oracle/gmc.py."""
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "yolov7-tracker_b200")):
    sys.path.insert(0, p)
from b200track.gmc import GmcEstimator  # noqa: E402
from b200track.synth import textured_frame  # noqa: E402
from oracle import gmc as OG  # noqa: E402

src = open(os.path.join(ROOT, "tests", "test_gpu_gmc.py")).read()
ns = {}
exec(src[src.index("def _dets"):src.index('@pytest.mark.parametrize("shape", [(720')], {"np": np}, ns)
h, w = 720, 1280
for S, use_dets, thr in ((1, False, 0.2), (1, True, 0.2), (2, True, 0.2), (2, True, -1.0), (2, True, 0.5)):
    frames = np.stack([textured_frame(40 + s, h, w, n_rect=900) for s in range(S)])
    dets = np.zeros((S, 32, 6), np.float32)
    cnt = np.array([20, 32][:S], np.int32)
    for s in range(S):
        dets[s, :cnt[s]] = ns["_dets"](7 + s, int(cnt[s]), h, w)
    est = GmcEstimator(S, h, w, 2, max_kp=32768)
    dd = torch.from_numpy(dets).cuda() if use_dets else None
    cc = torch.from_numpy(cnt).cuda() if use_dets else None
    warps, stat = est.estimate(torch.from_numpy(frames).cuda(), dd, cc, det_thresh=thr)
    torch.cuda.synchronize()
    ws = est.ws.cpu().numpy()
    lay = est.layout
    for s in range(S):
        d = dets[s, :cnt[s]]
        sel = d[d[:, 4] >= np.float32(thr)] if use_dets else None
        gray, xs, ys, desc = OG.GMCOracle().stages(frames[s], sel)
        pl = lambda name: ws[s * lay["stride"] + lay[name]: s * lay["stride"] + lay[name] + lay["h"] * lay["w"]].reshape(lay["h"], lay["w"])   # noqa: E731
        sc = OG.fast_score_map(gray)
        kx, ky, kd = est.keypoints(s)
        a, b = set(zip(kx.tolist(), ky.tolist())), set(zip(xs.tolist(), ys.tolist()))
        print("S=%d dets=%s thr=%.1f seq %d: gray diff %d, blur diff %d, score diff %d | kp gpu %d oracle %d, gpu-only %d oracle-only %d | stat %s" % (
            S, use_dets, thr, s, int((pl("gray") != gray).sum()), int((pl("blur") != OG.orb_blur(gray)).sum()), int((pl("score") != sc.astype(np.uint8)).sum()),
            len(kx), len(xs), len(a - b), len(b - a), stat[s].tolist()))
        extra = sorted(a - b)[:3]
        for (x, y) in extra:
            inside = [k for k, r in enumerate(d) if int(r[0] / 2) <= x < int(r[2] / 2) and int(r[1] / 2) <= y < int(r[3] / 2)]
            print("   gpu-only point", x, y, "score", int(sc[y, x]), "inside boxes", inside, [float(d[k, 4]) for k in inside])
