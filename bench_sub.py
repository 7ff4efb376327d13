"""Sub-benchmarks of the default bench line (``config.sub_benchmarks``): the BASELINE.json configurations that are not the
headline, measured in the same run on the same GPU so that the driver's one command records them.

  C3  ByteTrack full loop, ~300 detections / frame, ~250 live tracks, 4 sequences per launch, detections resident in HBM,
      L2 flushed between steps: frames/s, us per step, HBM roofline of track_step_kernel (latency-bound by construction: 1 CTA per
      sequence).
  C4  BoT-SORT (Kalman xywh + per-frame camera warp + IoU), 500 objects / frame, 8 sequences sharded over the ranks.
  C5  assignment-only sweep: N x M "+1" IoU cost matrix + exact LAP (lapjv semantics), N = M in {64 .. 2048}, 64 problems per
      launch, fp64 like the reference: GB/s of (write + read of every cost matrix) against the measured HBM peak.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "yolov7-tracker_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

STREAM_WARM = 60


def _events(torch, n):
    return [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]


def tracker_loop(torch, dev, kind, n_seq, n_obj, steps, seed0, hbm_gbs, warp_sigma=0.0, cap=1024, dmax=512, flush=None):
    """Device-resident fused tracker loop: returns frames/s, us/step and the roofline of track_step_kernel."""
    from b200track import _lib as L
    from b200track.engine import TrackEngine
    from b200track.synth import make_stream, pack_frames
    from bench import algorithmic_bytes
    n_frames = STREAM_WARM + steps
    streams = [make_stream(seed0 + s, n_frames, n_obj, warp_sigma=warp_sigma) for s in range(n_seq)]
    packed = [pack_frames(st[0], dmax) for st in streams]
    d_dets = torch.from_numpy(np.stack([p[0] for p in packed], 1)).to(dev)           # (F, S, dmax, 6)
    d_cnt = torch.from_numpy(np.stack([p[1] for p in packed], 1)).to(dev)
    d_warp = torch.from_numpy(np.stack([st[1].reshape(n_frames, 6) for st in streams], 1)).to(dev) if warp_sigma > 0 else None
    eng = TrackEngine(kind, n_seq=n_seq, dtype="f64", cap=cap, dmax=dmax, device=dev)
    rows = min(cap, 1024)
    d_out = torch.zeros((n_seq, rows, L.OUT_COLS), dtype=torch.float64, device=dev)
    d_stat = torch.zeros((n_frames, n_seq, L.STAT_WORDS), dtype=torch.int32, device=dev)
    for f in range(STREAM_WARM):
        eng.step_device(d_dets[f], d_cnt[f], d_out, d_stat[f], warps=None if d_warp is None else d_warp[f])
    torch.cuda.synchronize()
    ev = _events(torch, steps)
    for k in range(steps):
        f = STREAM_WARM + k
        if flush is not None:
            flush.zero_()                                    # L2 flush between timed iterations, outside the event pair
        ev[k][0].record()
        eng.step_device(d_dets[f], d_cnt[f], d_out, d_stat[f], warps=None if d_warp is None else d_warp[f])
        ev[k][1].record()
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(b) for a, b in ev])
    st = d_stat[STREAM_WARM:].cpu().numpy()
    assert int(st[:, :, L.STAT_ERR].max()) == 0, "tracker capacity error"
    algo = float(np.mean([algorithmic_bytes(st[k], 8) for k in range(steps)]))
    us = 1e3 * float(np.median(ms))
    return {"frames_per_s": n_seq * steps / (ms.sum() / 1e3), "us_per_step_median": us, "sequences": n_seq, "steps": steps,
            "tracked_mean": float(st[:, :, L.STAT_NTRACKED].mean()), "lost_mean": float(st[:, :, L.STAT_NLOST].mean()),
            "dets_mean": float(st[:, :, L.STAT_NHI].mean() + st[:, :, L.STAT_NLO].mean()), "ids_issued_per_seq": int(st[-1, :, L.STAT_NEXT_ID].mean()),
            "roofline": {"bound": "hbm", "kernel": "track_step_kernel<double>", "algorithmic_bytes_per_launch": algo, "achieved_GBs": algo / (us * 1e-6) / 1e9,
                         "peak_GBs": hbm_gbs, "frac": algo / (us * 1e-6) / 1e9 / hbm_gbs,
                         "note": "latency-bound: one CTA per sequence, state in L2 / shared memory (SURVEY 8d)"}}


def assignment_sweep(torch, dev, hbm_gbs, sizes=(64, 128, 256, 512, 1024, 2048), batch=64):
    """C5: batched IoU cost + exact LAP per size: us per launch and GB/s of the two kernels' algorithmic traffic."""
    import ctypes as C
    from b200track import _lib as L
    lib = L.load()
    s_ptr = lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)      # noqa: E731
    out = {}
    rng = np.random.default_rng(55)
    for n in sizes:
        b = batch if n <= 1024 else 16                        # 64 x 2048^2 x 8 B = 2 GB of cost matrices: 16 problems at the largest size
        xy = rng.uniform(0, 4000, (b, n, 2))
        wh = np.stack([rng.uniform(20, 80, (b, n)), rng.uniform(40, 160, (b, n))], -1)
        a = np.round(np.concatenate([xy, xy + wh], -1))
        perm = np.stack([rng.permutation(n) for _ in range(b)])
        bb = np.round(np.take_along_axis(a, perm[..., None], 1) + rng.normal(0, 3, (b, n, 4)))
        da, db = torch.from_numpy(a).to(dev), torch.from_numpy(bb).to(dev)
        cost = torch.empty((b, n, n), dtype=torch.float64, device=dev)
        x = torch.empty((b, n), dtype=torch.int32, device=dev); y = torch.empty((b, n), dtype=torch.int32, device=dev)
        ws = torch.empty(lib.b2t_lap_workspace_bytes(L.F64, n, n, b), dtype=torch.uint8, device=dev)

        def iou():
            L.check(lib, lib.b2t_iou_cost(L.F64, C.c_void_p(da.data_ptr()), n, C.c_void_p(db.data_ptr()), n, C.c_void_p(cost.data_ptr()), n, b, 1, s_ptr()))

        def lap(t):
            L.check(lib, lib.b2t_lap_solve(L.F64, C.c_void_p(cost.data_ptr()), n, n, n, t, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()),
                                           C.c_void_p(ws.data_ptr()), ws.numel(), b, s_ptr()))
        iou(); lap(0.9); torch.cuda.synchronize()
        reps = 5
        e = _events(torch, 3)
        e[0][0].record()
        for _ in range(reps):
            iou()
        e[0][1].record(); e[1][0].record()
        for _ in range(reps):
            lap(0.9)
        e[1][1].record(); e[2][0].record()
        for _ in range(reps):
            lap(0.5)
        e[2][1].record()
        torch.cuda.synchronize()
        t_iou, t_lap9, t_lap5 = (1e3 * p[0].elapsed_time(p[1]) / reps for p in e)
        matched = int((x >= 0).sum())
        mat = 8.0 * b * n * n
        out["N=M=%d" % n] = {"problems_per_launch": b, "iou_us": t_iou, "lap_us_thresh_0.9": t_lap9, "lap_us_thresh_0.5": t_lap5,
                             "iou_GBs": (mat + 64.0 * b * n) / (t_iou * 1e-6) / 1e9, "lap_GBs_0.9": mat / (t_lap9 * 1e-6) / 1e9,
                             "iou_frac_of_hbm": (mat + 64.0 * b * n) / (t_iou * 1e-6) / 1e9 / hbm_gbs, "lap_frac_of_hbm_0.9": mat / (t_lap9 * 1e-6) / 1e9 / hbm_gbs,
                             "matched_fraction_0.5": matched / float(b * n)}
        del cost, ws, da, db
        torch.cuda.empty_cache()
    return out


def run_all(torch, dev, rank, world, hbm_gbs, quick=False):
    """Returns the dict stored under config.sub_benchmarks (rank 0 gathers C4 over the ranks)."""
    import torch.distributed as dist
    res = {}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    if rank == 0:
        res["C3_bytetrack_4seq_tracker_only"] = tracker_loop(torch, dev, "bytetrack", 4, 300, 150 if quick else 400, 3000, hbm_gbs, flush=flush)
    # C4: 8 sequences sharded over the ranks (all 8 on this GPU when world == 1)
    s_local = max(1, 8 // world) if world <= 8 else 1
    c4 = tracker_loop(torch, dev, "botsort", s_local, 500, 100 if quick else 250, 4000 + rank * s_local, hbm_gbs, warp_sigma=3.0, cap=1152, dmax=576, flush=flush)
    if world > 1:
        t = torch.tensor([c4["us_per_step_median"], c4["frames_per_s"]], dtype=torch.float64, device=dev)
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        if rank == 0:
            worst = max(float(p[0]) for p in parts)
            c4 = dict(c4, frames_per_s=s_local * world / (worst * 1e-6), us_per_step_median=worst, sequences=s_local * world, ranks=world,
                      note="8 sequences sharded over the ranks; us = slowest rank's median step (max over ranks), frames/s = all sequences / that")
    if rank == 0:
        res["C4_botsort_500dets_8seq"] = c4
        res["C5_iou_lap_sweep_fp64"] = assignment_sweep(torch, dev, hbm_gbs, sizes=(64, 256, 1024) if quick else (64, 128, 256, 512, 1024, 2048))
    del flush
    torch.cuda.empty_cache()
    return res
