"""The CPU arm of the bench: the reference's own CPU implementation of detect -> NMS -> associate, timed on the box's host cores.

``kind = "reference"``: the UNMODIFIED reference modules (models/yolo.py ``Model`` fused on torch-cpu fp32, utils/general.py
``non_max_suppression``, tracker/bytetrack.py ``ByteTrack.update``) run from the archive oracle/build_ref.py packs in the build
container (oracle/_ref/, unpacked into a temporary directory); oracle/refshim.py only injects what the reference needs and the
image lacks (``np.float``, matplotlib / seaborn stubs, ``lap`` / ``cython_bbox`` stand-ins: parity unpinned there).
``kind = "port"``: the oracle/ restatement, when the archive is absent.
Used by ``bench.py --impl reference`` and by the ``cpu_baseline`` leg of the default run (the only places allowed to execute
oracle/ outside tests).  One "step" = one frame (the reference is batch-1, tracker/track.py:138-179).
"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "yolov7-tracker_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def host_threads():
    """Threads for torch-cpu: the cores this process may run on (affinity / cgroup aware -- os.cpu_count() over-reports on shared
    hosts and oversubscribed convolutions ran 10x slower in round 1), capped at 32 (MKL-DNN stops scaling on the w6 shapes)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(32, n))


def synthetic_frames_u8(n, size, seed=4242):
    """uint8 BGR frames (n, size, size, 3): the same generator the GPU arm uses (bench_pipeline.make_frames)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (size, size, 3), dtype=np.uint8)
    return np.stack([np.roll(base, (2 * k, k), axis=(0, 1)) for k in range(n)])


def run_cpu_arm(sd_cpu, img_size, n_frames, warmup=1, log=None):
    """Returns {"value": frames/s, "cores", "kind", "sample", "ms_per_frame": [...]} for n_frames timed frames of one sequence."""
    import torch
    threads = host_threads()
    torch.set_num_threads(threads)
    from oracle import build_ref, refshim
    frames = synthetic_frames_u8(max(n_frames + warmup, 1), img_size)
    tmp = None
    kind = "port"
    if not refshim.available() and os.path.exists(build_ref.ARCHIVE):
        tmp = tempfile.mkdtemp(prefix="b2t_ref_")
        refshim.use_root(build_ref.unpack(tmp))
    if refshim.available():
        kind = "reference"
        model = refshim.load_detector_model()                              # models.yolo.Model(cfg/deploy/yolov7-w6.yaml).eval().fuse()
        missing = [k for k in sd_cpu if k not in model.state_dict()]
        assert not missing, missing[:3]
        model.load_state_dict(sd_cpu, strict=False)
        general = refshim.load_general()
        ns = refshim.load()
        tracker = ns.bytetrack.ByteTrack(refshim.Opts(img_size=img_size), frame_rate=30)

        def one(frame_u8):
            img = torch.from_numpy(np.ascontiguousarray(frame_u8[:, :, ::-1].transpose(2, 0, 1))).float().div_(255.0)[None]   # tracker_dataloader.py:80-86
            pred = model(img)[0]                                            # track.py:144
            out = general.non_max_suppression(pred, 0.01, 0.45)[0]           # track.py:239
            out[:, :4] = general.scale_coords(img.shape[2:], out[:, :4], frame_u8.shape, ratio_pad=None).round()   # :240
            d = out.numpy()
            d = d[(d[:, 2] - d[:, 0] >= 1) & (d[:, 3] - d[:, 1] >= 1)]      # q9: zero-size boxes give NaN Kalman states in the reference
            return tracker.update(d, frame_u8)
    else:
        from oracle import detector as OD, trackers as OT
        from b200track.w6 import ANCHORS, STRIDES, w6_layers
        layers = w6_layers()
        trk = OT.TrackerOracle("bytetrack")

        def one(frame_u8):
            img = torch.from_numpy(np.ascontiguousarray(frame_u8[:, :, ::-1].transpose(2, 0, 1))).float().div_(255.0)[None]
            pred = OD.forward(layers, sd_cpu, img, ANCHORS, STRIDES)
            det = OD.post_process(OD.non_max_suppression(pred, conf_thres=0.01)[0], (img_size, img_size))
            d = det.numpy()
            d = d[(d[:, 2] - d[:, 0] >= 1) & (d[:, 3] - d[:, 1] >= 1)]
            return trk.update(d)
    ms = []
    with torch.no_grad():
        for k in range(warmup):
            one(frames[k])
        for k in range(n_frames):
            t0 = time.perf_counter()
            one(frames[warmup + k])
            ms.append(1e3 * (time.perf_counter() - t0))
            if log:
                log("cpu arm frame %d: %.0f ms" % (k, ms[-1]))
    if tmp:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    total = sum(ms) / 1e3
    what = ("the reference's own models/yolo.py Model (torch-cpu fp32, fused) + utils/general.py non_max_suppression + tracker/bytetrack.py "
            "ByteTrack.update, unmodified, from oracle/_ref") if kind == "reference" else "oracle/ restatement of the reference's CPU path"
    return {"value": n_frames / total, "unit": "frames/s", "cores": threads, "kind": kind, "ms_per_frame": [round(v, 1) for v in ms],
            "sample": "%d frames of one %dx%d sequence after %d warm-up: %s" % (n_frames, img_size, img_size, warmup, what)}
