"""Drop-in for the reference's ``tracker/bytetrack.py``: ``ByteTrack(opts, frame_rate=30, gamma=0.1)``
with ``update(det_results, ori_img) -> List[STrack]`` (reference :41-204), executed as one fused
kernel per frame (csrc/b2t_step.cuh, kind = bytetrack).  The appearance branch is off, as in the
reference default (``use_apperance_model = False``, :11); the ReID extractor is therefore not loaded."""
import _b2t_path  # noqa: F401
from basetrack import TrackState, STrack, BaseTracker, joint_stracks, sub_stracks, remove_duplicate_stracks  # noqa: F401


class ByteTrack(BaseTracker):
    _kind = 'bytetrack'

    def __init__(self, opts, frame_rate=30, gamma=0.1, *args, **kwargs):
        super().__init__(opts, frame_rate, *args, **kwargs)
        self.use_apperance_model = False
        self.reid_model = None
        self.gamma = gamma
        self.low_conf_thresh = max(0.15, self.opts.conf_thresh - 0.3)
        self.filter_small_area = False
