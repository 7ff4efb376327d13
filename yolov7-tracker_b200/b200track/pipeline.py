"""detect -> NMS -> associate as one pipelined object: what ``tracker/track.py:138-179`` does per frame
(model forward, non_max_suppression + scale_coords, tracker.update), for B sequences at once.

Three CUDA streams keep the B200 busy across frames:
  copy    : pinned host frames -> device (next frame's H2D overlaps the current forward)
  detect  : ReOrg + 107 tcgen05 convs + decode (CUDA graph), then NMS (second graph)
  track   : fused ByteTrack / SORT / BoT-SORT step on the NMS output, then D2H of the track rows
Frame t+1's forward runs while frame t is being associated; the only cross-frame hazards (the single input image
buffer and the single NMS output buffer) are guarded by events.  ``step()`` returns the tracks of the PREVIOUS
call (one frame of latency, same results); ``flush()`` returns the last ones.
"""
import torch

from . import _lib as L


class TrackingPipeline:
    def __init__(self, detector, engine, out_rows=512, gmc=None):
        """gmc: a ``b200track.gmc.GmcEstimator`` for the source-frame size (BoT-SORT with camera-motion compensation, reference
        botsort.py:380-382): the warp of every sequence is estimated on the GPU from the uint8 frames and the NMS output and fed
        to the tracker step without leaving the device."""
        self.det, self.eng, self.gmc = detector, engine, gmc
        if gmc is not None and gmc.S != detector.B:
            raise L.B2TError("GmcEstimator(n_seq=%d) does not match DetectorW6(batch=%d)" % (gmc.S, detector.B))
        # the fused tracker kernel indexes the NMS output as [sequence][dmax][6]: the two objects must agree on the layout
        if engine.S != detector.B or engine.dmax != detector.max_det:
            raise L.B2TError("TrackEngine(n_seq=%d, dmax=%d) does not match DetectorW6(batch=%d, max_det=%d)" % (engine.S, engine.dmax, detector.B, detector.max_det))
        dev = detector.dev
        self.dev = dev
        self.s_copy, self.s_det, self.s_trk = (torch.cuda.Stream(device=dev) for _ in range(3))
        B = detector.B
        self.t_out = torch.zeros((B, out_rows, L.OUT_COLS), dtype=torch.float64, device=dev)
        self.t_stat = torch.zeros((B, L.STAT_WORDS), dtype=torch.int32, device=dev)
        self.h_out = [torch.zeros((B, out_rows, L.OUT_COLS), dtype=torch.float64).pin_memory() for _ in range(2)]
        self.h_stat = [torch.zeros((B, L.STAT_WORDS), dtype=torch.int32).pin_memory() for _ in range(2)]
        self.ev_img_free = torch.cuda.Event()       # reorg has consumed det.img
        self.ev_img_ready = torch.cuda.Event()
        self.ev_nms_done = torch.cuda.Event()
        self.ev_trk_done = [torch.cuda.Event(), torch.cuda.Event()]
        self.ev_out_free = torch.cuda.Event()       # tracker has consumed det.out
        self.g_fwd = self.g_nms = None
        self.n = 0
        self._capture()

    def _capture(self):
        det = self.det
        torch.cuda.synchronize()
        with torch.cuda.stream(self.s_det):
            det._forward_launches(); det._nms_launch(True)                    # warm-up (also sets kernel attributes)
            torch.cuda.synchronize()
            self.g_fwd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_fwd, stream=self.s_det):
                for fn, _, name in det.ops[1:]:                               # ops[0] is the ReOrg that reads det.img
                    fn()
            self.g_nms = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_nms, stream=self.s_det):
                det._nms_launch(True)
        torch.cuda.synchronize()
        self.ev_img_free.record(self.s_det)
        self.ev_out_free.record(self.s_trk)

    def step(self, frames, warps=None):
        """frames: one frame per sequence, pinned host tensor (copied on the copy stream) or device tensor, either
          * uint8 BGR (B, h, w, 3) as cv2.imread returns them -- the letterbox / RGB / 255 / ReOrg / fp16 conversion runs on the device
            (b2t_letterbox_reorg; call ``det.set_source_frames((h, w))`` once before), 3 bytes per pixel over PCIe, or
          * float32 (B, 3, H, W) in [0, 1], the tensor the reference's dataloader produces.
        Returns (rows, stat) of the previous frame as pinned host tensors, or None on the first call."""
        det, eng = self.det, self.eng
        k = self.n & 1
        u8 = frames.dtype == torch.uint8
        if u8 and (getattr(det, "src_u8", None) is None or tuple(frames.shape) != tuple(det.src_u8.shape)):
            raise L.B2TError("uint8 frames of shape %s: call det.set_source_frames((h, w)) first" % (tuple(frames.shape),))
        # ---- input: wait until the previous ingest kernel has read the staging buffer, then copy
        with torch.cuda.stream(self.s_copy):
            self.s_copy.wait_event(self.ev_img_free)
            (det.src_u8 if u8 else det.img).copy_(frames, non_blocking=True)
            self.ev_img_ready.record(self.s_copy)
        # ---- detect
        with torch.cuda.stream(self.s_det):
            self.s_det.wait_event(self.ev_img_ready)
            if u8:
                det.ingest_u8_launch()                                         # letterbox + RGB + /255 + ReOrg + 16-bit NHWC
                if self.gmc is not None:
                    self.gmc.prepare(det.src_u8, k)                            # gray / FAST scores / smoothed image while the frame buffer is valid
            else:
                det.ops[0][0]()                                                # ReOrg + 16-bit NHWC of the float tensor
            self.ev_img_free.record(self.s_det)
            self.g_fwd.replay()
            self.s_det.wait_event(self.ev_out_free)                            # previous tracker step has read det.out
            self.g_nms.replay()
            self.ev_nms_done.record(self.s_det)
        # ---- associate + read back
        with torch.cuda.stream(self.s_trk):
            self.s_trk.wait_event(self.ev_nms_done)
            if self.gmc is not None and u8:
                # key points outside the boxes of the high-score detections (botsort.py:380), matching, RANSAC -> warps on the device
                w23, _ = self.gmc.estimate_prepared(k, det.out, det.out_count, det_thresh=float(eng.cfg.conf_thresh))
                warps = w23.view(eng.S, 6)
            eng.step_device(det.out, det.out_count, self.t_out, self.t_stat, warps=warps)
            self.ev_out_free.record(self.s_trk)
            self.h_out[k].copy_(self.t_out, non_blocking=True)
            self.h_stat[k].copy_(self.t_stat, non_blocking=True)
            self.ev_trk_done[k].record(self.s_trk)
        self.n += 1
        if self.n == 1:
            return None
        return self._collect(1 - k)

    def _collect(self, k):
        self.ev_trk_done[k].synchronize()
        err = int(self.h_stat[k][:, L.STAT_ERR].max())
        if err:
            raise L.B2TError("tracker capacity error bits 0x%x (slots / detections / edges / output rows)" % err)
        return self.h_out[k], self.h_stat[k]

    def flush(self):
        """Tracks of the last submitted frame."""
        if self.n == 0:
            return None
        return self._collect((self.n - 1) & 1)
