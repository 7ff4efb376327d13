"""Drop-in for the reference's ``tracker/botsort.py``: ``BoTSORT(opts, frame_rate=30, gamma=0.02,
use_GMC=True)``, ``multi_gmc`` and ``GMC``.

``BoTSORT.update`` (reference :313-493) runs as one fused kernel per frame (kind = botsort): Kalman
predict, ``multi_gmc`` of the pool and of the unconfirmed tracks (:380-382), three IoU associations,
births from ALL first-stage leftovers (q3), list algebra.  ``multi_gmc`` (:250-269) is also
available on its own for lists of STrack (b2t_gmc_apply).

Camera-motion ESTIMATION (``GMC.apply``, method 'orb', reference :111-235 -- SURVEY.md section 8(f) row 1)
runs on the GPU as well (``b200track/gmc.py``, csrc/b2t_gmc.cu).  ``tracker.gmc`` may be replaced by any object
with ``apply(raw_frame, detections)`` returning a 2x3 matrix."""
import numpy as np

import _b2t_path  # noqa: F401
from basetrack import BaseTrack, TrackState, STrack, BaseTracker, joint_stracks, sub_stracks, remove_duplicate_stracks  # noqa: F401
from b200track import _lib as L
from b200track import engine as _eng

import torch  # noqa: E402


class GMC:
    """``GMC(method='orb', downscale=2)`` -- what ``BoTSORT.__init__`` (reference :286) builds -- runs on the GPU estimator
    (csrc/b2t_gmc.cu through b200track/gmc.py): same key points, descriptors and matches as the reference's OpenCV calls, RANSAC
    with its own sampling sequence.  'file' and 'none' need no estimation.  'sift' and 'ecc' (unused by BoT-SORT; StrongSORT's
    ECC is outside SURVEY.md section 8) are not built and raise."""

    def __init__(self, method='orb', downscale=2, verbose=None, max_keypoints=32768):
        self.method = method
        self.downscale = max(1, int(downscale))
        self.max_keypoints = int(max_keypoints)
        self.initializedFirstFrame = False
        self._est = None
        if method == 'orb':
            pass
        elif method in ('sift', 'ecc'):
            raise NotImplementedError("GMC method %r is not built (SURVEY.md section 8f covers the ORB estimator BoT-SORT uses)" % method)
        elif method in ('file', 'files'):
            seq, ablation = verbose[0], verbose[1]
            root = 'tracker/GMC_files/MOT17_ablation' if ablation else 'tracker/GMC_files/MOTChallenge'
            for suffix in ('-FRCNN', '-DPM', '-SDP'):
                if suffix in seq:
                    seq = seq[:-len(suffix)]
            self.gmcFile = open(root + '/GMC-' + seq + '.txt', 'r')
        elif method in ('none', 'None'):
            self.method = 'none'
        else:
            raise ValueError('Error: Unknown CMC method:' + method)

    def apply(self, raw_frame, detections=None):
        if self.method == 'orb':
            return self.applyFeaures(raw_frame, detections)
        if self.method in ('file', 'files'):
            return self.applyFile(raw_frame, detections)
        return np.eye(2, 3)

    def applyFile(self, raw_frame, detections=None):
        tok = self.gmcFile.readline().split('\t')
        return np.array([[float(tok[1]), float(tok[2]), float(tok[3])],
                         [float(tok[4]), float(tok[5]), float(tok[6])]], dtype=np.float64)

    def applyFeaures(self, raw_frame, detections=None):
        """raw_frame: (H, W, 3) uint8 BGR (ndarray or tensor, host or device); detections: (n, >= 4) tlbr rows to mask out."""
        from b200track.gmc import GmcEstimator
        frame = raw_frame if isinstance(raw_frame, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(raw_frame))
        h, w = int(frame.shape[0]), int(frame.shape[1])
        if self._est is None or (self._est.h, self._est.w) != (h, w):
            self._est = GmcEstimator(1, h, w, self.downscale, self.max_keypoints)
        est = self._est
        frame = frame.to(est.dev, non_blocking=True).contiguous()[None]
        dets = None
        if detections is not None and len(detections):
            d = detections if isinstance(detections, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(detections, dtype=np.float32)))
            dets = torch.zeros((1, d.shape[0], 6), dtype=torch.float32, device=est.dev)
            dets[0, :, :4] = d[:, :4].to(est.dev)
            dets[0, :, 4] = 1.0                                   # every row handed over is masked (the caller already filtered)
        warps, stat = est.estimate(frame, dets, None, det_thresh=0.5)
        self.initializedFirstFrame = True
        H = warps[0].cpu().numpy()
        self.last_stat = stat.cpu().numpy()[0]
        if self.last_stat[5] & L.GMC_TRUNCATED and not getattr(self, '_warned', False):
            # the reference has no cap; ours keeps the first max_keypoints corners in row-major order (the top of the frame)
            print('Warning: GMC found more than %d key points; raise GMC(max_keypoints=...)' % self.max_keypoints)
            self._warned = True
        return H


def multi_gmc(stracks, H=np.eye(2, 3)):
    """Warp the Kalman state of every track in ``stracks`` (reference :250-269) on the GPU."""
    if len(stracks) == 0:
        return
    ops = _eng.ops()
    mean = ops.dev(np.asarray([st.mean.copy() for st in stracks], dtype=np.float64), torch.float64)
    cov = ops.dev(np.asarray([st.cov for st in stracks], dtype=np.float64), torch.float64)
    ops.gmc_apply(L.F64, mean, cov, H)
    mean, cov = mean.cpu().numpy(), cov.cpu().numpy()
    for st, m, c in zip(stracks, mean, cov):
        st.mean, st.cov = m, c


class BoTSORT(BaseTracker):
    _kind = 'botsort'

    def __init__(self, opts, frame_rate=30, gamma=0.02, use_GMC=True, *args, **kwargs):
        self.use_GMC = use_GMC
        super().__init__(opts, frame_rate, *args, **kwargs)
        self.use_apperance_model = False
        self.reid_model = None
        self.gamma = gamma
        self.low_conf_thresh = max(0.15, self.opts.conf_thresh - 0.3)
        self.filter_small_area = False
        self.gmc = GMC(method='orb', downscale=2, verbose=None) if use_GMC else GMC(method='none')
        self.theta_iou, self.theta_emb = 0.5, 0.25

    def _warp(self, dets, ori_img):
        if not self.use_GMC:
            return None
        det_high = dets[dets[:, 4] >= np.float32(self.det_thresh)]            # reference :380 hands over the high-score detections
        return self.gmc.apply(raw_frame=ori_img, detections=det_high)
