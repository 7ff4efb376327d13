"""Drop-in for the reference's ``tracker/botsort.py``: ``BoTSORT(opts, frame_rate=30, gamma=0.02,
use_GMC=True)``, ``multi_gmc`` and ``GMC``.

``BoTSORT.update`` (reference :313-493) runs as one fused kernel per frame (kind = botsort): Kalman
predict, ``multi_gmc`` of the pool and of the unconfirmed tracks (:380-382), three IoU associations,
births from ALL first-stage leftovers (q3), list algebra.  ``multi_gmc`` (:250-269) is also
available on its own for lists of STrack (b2t_gmc_apply).

Camera-motion ESTIMATION (``GMC.apply`` with ORB / SIFT / ECC, reference :111-235) is SURVEY.md
section 8(f) "next": it stays a host-side OpenCV call exactly as in the reference; 'file' and 'none'
need no OpenCV.  ``tracker.gmc`` may be replaced by any object with ``apply(raw_frame, detections)``
returning a 2x3 matrix."""
import numpy as np

import _b2t_path  # noqa: F401
from basetrack import BaseTrack, TrackState, STrack, BaseTracker, joint_stracks, sub_stracks, remove_duplicate_stracks  # noqa: F401
from b200track import _lib as L
from b200track import engine as _eng

import torch  # noqa: E402


class GMC:
    def __init__(self, method='orb', downscale=2, verbose=None):
        self.method = method
        self.downscale = max(1, int(downscale))
        self.prevFrame = self.prevKeyPoints = self.prevDescriptors = None
        self.initializedFirstFrame = False
        if method in ('orb', 'sift', 'ecc'):
            import cv2
            self._cv2 = cv2
            if method == 'orb':
                self.detector = cv2.FastFeatureDetector_create(20)
                self.extractor = cv2.ORB_create()
                self.matcher = cv2.BFMatcher(cv2.NORM_HAMMING)
            elif method == 'sift':
                self.detector = cv2.SIFT_create(nOctaveLayers=3, contrastThreshold=0.02, edgeThreshold=20)
                self.extractor = cv2.SIFT_create(nOctaveLayers=3, contrastThreshold=0.02, edgeThreshold=20)
                self.matcher = cv2.BFMatcher(cv2.NORM_L2)
            else:
                self.warp_mode = cv2.MOTION_EUCLIDEAN
                self.criteria = (cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 100, 1e-5)
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
        if self.method in ('orb', 'sift'):
            return self.applyFeaures(raw_frame, detections)
        if self.method == 'ecc':
            return self.applyEcc(raw_frame, detections)
        if self.method in ('file', 'files'):
            return self.applyFile(raw_frame, detections)
        return np.eye(2, 3)

    def applyFile(self, raw_frame, detections=None):
        tok = self.gmcFile.readline().split('\t')
        return np.array([[float(tok[1]), float(tok[2]), float(tok[3])],
                         [float(tok[4]), float(tok[5]), float(tok[6])]], dtype=np.float64)

    def _gray(self, raw_frame):
        cv2 = self._cv2
        frame = cv2.cvtColor(raw_frame, cv2.COLOR_BGR2GRAY)
        if self.downscale > 1:
            frame = cv2.resize(frame, (frame.shape[1] // self.downscale, frame.shape[0] // self.downscale))
        return frame

    def applyEcc(self, raw_frame, detections=None):
        cv2 = self._cv2
        frame = self._gray(raw_frame)
        if self.downscale > 1:
            frame = cv2.GaussianBlur(frame, (3, 3), 1.5)
        H = np.eye(2, 3, dtype=np.float32)
        if not self.initializedFirstFrame:
            self.prevFrame = frame.copy()
            self.initializedFirstFrame = True
            return H
        try:
            _, H = cv2.findTransformECC(self.prevFrame, frame, H, self.warp_mode, self.criteria, None, 1)
        except Exception:
            print('Warning: find transform failed. Set warp as identity')
        return H

    def applyFeaures(self, raw_frame, detections=None):
        """Host OpenCV estimation, same recipe as the reference (FAST/ORB keypoints outside detection
        boxes, kNN ratio test, spatial 2.5-sigma filter, RANSAC partial affine)."""
        cv2 = self._cv2
        frame = self._gray(raw_frame)
        height, width = frame.shape
        H = np.eye(2, 3)
        mask = np.zeros_like(frame)
        mask[int(0.02 * height): int(0.98 * height), int(0.02 * width): int(0.98 * width)] = 255
        if detections is not None:
            for det in detections:
                tlbr = (det[:4] / self.downscale).astype(np.int_)
                mask[tlbr[1]:tlbr[3], tlbr[0]:tlbr[2]] = 0
        keypoints = self.detector.detect(frame, mask)
        keypoints, descriptors = self.extractor.compute(frame, keypoints)
        if not self.initializedFirstFrame:
            self.prevFrame, self.prevKeyPoints, self.prevDescriptors = frame.copy(), keypoints, descriptors
            self.initializedFirstFrame = True
            return H
        knn = self.matcher.knnMatch(self.prevDescriptors, descriptors, 2) if descriptors is not None and self.prevDescriptors is not None else []
        cand, dists = [], []
        max_d = 0.25 * np.array([width, height])
        for pair in knn:
            if len(pair) < 2:
                continue
            m, n = pair
            if m.distance < 0.9 * n.distance:
                p0, p1 = self.prevKeyPoints[m.queryIdx].pt, keypoints[m.trainIdx].pt
                d = (p0[0] - p1[0], p0[1] - p1[1])
                if abs(d[0]) < max_d[0] and abs(d[1]) < max_d[1]:
                    dists.append(d)
                    cand.append(m)
        if len(cand):
            dists = np.asarray(dists)
            inl = np.all((dists - dists.mean(0)) < 2.5 * dists.std(0), axis=1)
            prev = np.array([self.prevKeyPoints[m.queryIdx].pt for m, ok in zip(cand, inl) if ok])
            cur = np.array([keypoints[m.trainIdx].pt for m, ok in zip(cand, inl) if ok])
            if prev.shape[0] > 4:
                est, _ = cv2.estimateAffinePartial2D(prev, cur, cv2.RANSAC)
                if est is not None:
                    H = est
                    if self.downscale > 1:
                        H[0, 2] *= self.downscale
                        H[1, 2] *= self.downscale
            else:
                print('Warning: not enough matching points')
        self.prevFrame, self.prevKeyPoints, self.prevDescriptors = frame.copy(), keypoints, descriptors
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
        if isinstance(ori_img, torch.Tensor):
            ori_img = ori_img.numpy()
        det_high = dets[dets[:, 4] >= np.float32(self.det_thresh)]
        return self.gmc.apply(raw_frame=ori_img, detections=det_high)
