"""Writes tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

Run in the build container only (``python tests/golden/make_golden.py``): /root/reference
does not exist on the GPU box, which is why the outputs are committed.  The reference modules
are imported through oracle/refshim.py (np.float alias, matplotlib stubs, lap / cython_bbox
stand-ins -- see that file for exactly what is injected; nothing in /root/reference is edited).

Fixtures (NumPy 2.3.5 / SciPy 1.18.1, recorded in every file -- SURVEY q12):
  kalman_<fmt>.npz   initiate / multi_predict / project / update of the reference Kalman
                     classes (tracker/kalman_filter.py) on seeded states.
  loop_<kind>.npz    reference SORT / ByteTrack / BoT-SORT run frame by frame on seeded
                     synthetic streams: track ids (every frame), tlwh (every 8th frame + last).
  iou_lap.npz        IoU matrices and assignments.  PARITY UNPINNED: these come from the
                     oracle's restatement of cython_bbox / lap (absent wheels), they are
                     regression values, not reference outputs.
"""
import os
import sys

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "yolov7-tracker_b200"))

from oracle import refshim, iou as oiou, lapjv as olap      # noqa: E402
from b200track.synth import make_stream, stream_digest      # noqa: E402

VERS = dict(numpy=np.__version__, scipy=scipy.__version__)

LOOPS = [  # (name, seed, n_obj, n_frames)
    ("small", 11, 40, 120),
    ("c3", 12, 300, 64),
]


def kalman_fixture(ref, name, cls, n=64, seed=5):
    rng = np.random.default_rng(seed)
    kf = cls()
    out = {}
    # measurements the way STrack.activate builds them: float32
    if name == "botsort":
        z = np.stack([rng.uniform(0, 1280, n), rng.uniform(0, 1280, n), rng.uniform(4, 200, n),
                      rng.uniform(4, 300, n)], 1).astype(np.float32)
    else:
        z = np.stack([rng.uniform(0, 1280, n), rng.uniform(0, 1280, n), rng.uniform(0.2, 2.0, n),
                      rng.uniform(4, 300, n)], 1).astype(np.float32)
    m0, c0 = zip(*[kf.initiate(zi) for zi in z])
    out["z0"] = z
    out["init_mean"] = np.stack(m0)                    # float32
    out["init_cov"] = np.stack([np.asarray(c, np.float64) for c in c0])
    out["init_cov_is_f32"] = np.array([c.dtype == np.float32 for c in c0])
    # predict from the float32 means (the frame-2 case) and from float64 means
    mp32, cp32 = kf.multi_predict(np.stack(m0), np.stack(c0))
    out["pred32_mean"], out["pred32_cov"] = mp32, cp32
    mean, cov = mp32, cp32
    steps_mean, steps_cov, steps_z, proj_m, proj_s = [], [], [], [], []
    for k in range(6):
        zk = (z + rng.normal(0, 1.0 if name == "botsort" else 0.01, z.shape) * np.array([1, 1, 0.02, 1])).astype(np.float32)
        pm, ps = zip(*[kf.project(mi, ci) for mi, ci in zip(mean, cov)])
        if name == "strongsort":
            conf = rng.uniform(0.2, 0.95, n).astype(np.float32)
            um, uc = zip(*[kf.update(mi, ci, zi, co) for mi, ci, zi, co in zip(mean, cov, zk, conf)])
            out["conf%d" % k] = conf
        else:
            um, uc = zip(*[kf.update(mi, ci, zi) for mi, ci, zi in zip(mean, cov, zk)])
        mean, cov = kf.multi_predict(np.stack(um), np.stack(uc))
        steps_z.append(zk); proj_m.append(np.stack(pm)); proj_s.append(np.stack(ps))
        steps_mean.append(np.stack(um)); steps_cov.append(np.stack(uc))
        out["pred_mean%d" % k], out["pred_cov%d" % k] = mean, cov
    out["upd_z"] = np.stack(steps_z); out["proj_mean"] = np.stack(proj_m); out["proj_cov"] = np.stack(proj_s)
    out["upd_mean"] = np.stack(steps_mean); out["upd_cov"] = np.stack(steps_cov)
    # update straight from the float32 state (an unconfirmed track's first update)
    if name != "strongsort":
        um, uc = zip(*[kf.update(mi, ci, zi) for mi, ci, zi in zip(m0, c0, steps_z[0])])
        out["upd32_mean"], out["upd32_cov"] = np.stack(um), np.stack(uc)
    # gating distance (default filter only has it)
    if hasattr(kf, "gating_distance"):
        out["gate"] = np.stack([kf.gating_distance(mean[i], cov[i], z[:8].astype(np.float64)) for i in range(16)])
    np.savez_compressed(os.path.join(HERE, "kalman_%s.npz" % name), **out, **{"ver_" + k: v for k, v in VERS.items()})
    print("kalman", name, "ok")


def run_reference(ref, kind, frames, warps):
    ref.basetrack.BaseTrack._count = 0
    if kind == "sort":
        trk = ref.basetrack.BaseTracker(refshim.Opts(kalman_format="default"))
    elif kind == "bytetrack":
        trk = ref.bytetrack.ByteTrack(refshim.Opts(kalman_format="default"))
    else:
        trk = ref.botsort.BoTSORT(refshim.Opts(kalman_format="botsort"))
        trk.gmc = refshim.FixedGMC(warps)
    img = np.zeros((4, 4, 3), np.uint8)
    res = []
    for f in frames:
        cur = trk.update(f.copy(), img)
        res.append((np.array([t.track_id for t in cur], np.int32),
                    np.array([np.asarray(t.tlwh, np.float64) for t in cur]).reshape(-1, 4),
                    np.array([float(t.cls) for t in cur], np.float32),
                    len(trk.tracked_stracks), len(trk.lost_stracks)))
    return res


def loop_fixture(ref, kind):
    out = {}
    for name, seed, n_obj, n_frames in LOOPS:
        frames, warps = make_stream(seed, n_frames, n_obj, warp_sigma=3.0 if kind == "botsort" else 0.0)
        res = run_reference(ref, kind, frames, warps)
        out[name + "_digest"] = stream_digest(frames)
        out[name + "_cfg"] = np.array([seed, n_obj, n_frames])
        out[name + "_count"] = np.array([len(r[0]) for r in res], np.int32)
        out[name + "_ids"] = np.concatenate([r[0] for r in res])
        out[name + "_cls"] = np.concatenate([r[2] for r in res])
        out[name + "_ntracked"] = np.array([r[3] for r in res], np.int32)
        out[name + "_nlost"] = np.array([r[4] for r in res], np.int32)
        keep = [i for i in range(n_frames) if i % 8 == 7 or i == n_frames - 1]
        out[name + "_tlwh_frames"] = np.array(keep, np.int32)
        out[name + "_tlwh"] = np.concatenate([res[i][1] for i in keep])
        print("loop", kind, name, "max id", out[name + "_ids"].max(), "last frame tracks", len(res[-1][0]))
    np.savez_compressed(os.path.join(HERE, "loop_%s.npz" % kind), **out, **{"ver_" + k: v for k, v in VERS.items()})


def iou_lap_fixture():
    rng = np.random.default_rng(77)
    out = {}
    a = np.round(rng.uniform(0, 600, (48, 2)))
    a = np.concatenate([a, a + np.round(rng.uniform(4, 120, (48, 2)))], 1)
    b = a[rng.permutation(48)[:40]] + np.round(rng.normal(0, 3, (40, 4)))
    # adversarial rows: touching, contained, identical, 1-px, disjoint
    adv_a = np.array([[10, 10, 20, 20], [10, 10, 20, 20], [0, 0, 100, 100], [5, 5, 5, 5], [0, 0, 10, 10], [3.5, 2.25, 9.75, 8.5]], float)
    adv_b = np.array([[21, 10, 30, 20], [20, 20, 30, 30], [40, 40, 60, 60], [5, 5, 5, 5], [11, 11, 20, 20], [3.5, 2.25, 9.75, 8.5]], float)
    out["a"], out["b"], out["adv_a"], out["adv_b"] = a, b, adv_a, adv_b
    out["iou"] = oiou.ious(a, b)
    out["adv_iou"] = oiou.ious(adv_a, adv_b)
    cost = 1.0 - out["iou"]
    for t in (0.9, 0.5, 0.7):
        _, x, y = olap.lapjv(cost, True, t)
        out["x_%02d" % int(t * 10)] = x
        out["y_%02d" % int(t * 10)] = y
    np.savez_compressed(os.path.join(HERE, "iou_lap.npz"), **out)
    print("iou_lap ok (parity unpinned)")


def main():
    ref = refshim.load()
    kalman_fixture(ref, "default", ref.kalman_filter.KalmanFilter)
    kalman_fixture(ref, "botsort", ref.kalman_filter.BoTSORTKalmanFilter)
    kalman_fixture(ref, "strongsort", ref.kalman_filter.NSAKalmanFilter)
    for kind in ("sort", "bytetrack", "botsort"):
        loop_fixture(ref, kind)
    iou_lap_fixture()


if __name__ == "__main__":
    main()
