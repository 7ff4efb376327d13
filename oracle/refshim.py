"""Oracle support: import the UNMODIFIED reference modules from /root/reference.

TEST INFRASTRUCTURE (see oracle/__init__.py).  /root/reference only exists in the build
container, so this file is used by ``tests/golden/make_golden.py`` (fixture writer) and by
the ``not gpu`` tests that re-pin the oracle when the reference is present.  Nothing that
runs on the GPU box may call ``load()``.

What is injected (nothing under /root/reference is edited):
  * ``np.float`` / ``np.float_`` aliases removed in NumPy >= 1.24 / 2.0
    (tracker/matching.py:52,57 ; tracker/botsort.py:240);
  * stub modules ``matplotlib``, ``matplotlib.pyplot``, ``seaborn`` (imported, never used,
    tracker/botsort.py:10 ; utils/plots.py:11-15);
  * ``lap`` and ``cython_bbox`` stand-ins backed by oracle/lapjv.py and oracle/iou.py -- the two
    third-party wheels are absent and unpinned: parity unpinned there;
  * ``reid_models.deepsort_reid.Extractor`` stub so that ``ByteTrack.__init__`` /
    ``BoTSORT.__init__`` (bytetrack.py:12, botsort.py:278) do not load the 46 MB ReID checkpoint
    (appearance is off by default).
The reference modules are imported under their bare names (they import each other that
way), then those names are removed from ``sys.modules`` again so that the product's
same-named drop-in modules are not shadowed.
"""
import importlib
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("B2T_REFERENCE_ROOT", "/root/reference")
_BARE = ("kalman_filter", "matching", "basetrack", "bytetrack", "botsort")
_cache = {}


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "tracker"))


def use_root(path):
    """Point the shim at another copy of the reference tree (bench.py: the archive of oracle/build_ref.py unpacked into a
    temporary directory on the GPU box)."""
    global REF_ROOT
    REF_ROOT = path
    _cache.clear()


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def load():
    """Returns a namespace with the reference's tracker modules: .kalman_filter .matching
    .basetrack .bytetrack .botsort"""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    from . import iou as _iou, lapjv as _lapjv

    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "float_"):
        np.float_ = np.float64

    class _Extractor:                       # stands in for reid_models.deepsort_reid.Extractor
        def __init__(self, *a, **k):
            pass

    injected = {
        "lap": _stub("lap", lapjv=_lapjv.lapjv),
        "cython_bbox": _stub("cython_bbox", bbox_overlaps=_iou.bbox_overlaps),
        "reid_models": _stub("reid_models"),
        "reid_models.deepsort_reid": _stub("reid_models.deepsort_reid", Extractor=_Extractor),
    }
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                injected[name] = _stub(name)
    saved = {k: sys.modules.get(k) for k in list(injected) + list(_BARE)}
    for k in _BARE:
        sys.modules.pop(k, None)
    sys.modules.update(injected)
    tracker_dir = os.path.join(REF_ROOT, "tracker")
    sys.path.insert(0, tracker_dir)
    ns = types.SimpleNamespace()
    try:
        for k in _BARE:
            setattr(ns, k, importlib.import_module(k))
    finally:
        sys.path.remove(tracker_dir)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cache["ns"] = ns
    return ns


class Opts:
    """The subset of tracker/track.py's argparse namespace the trackers read (track.py:342-380)."""

    def __init__(self, conf_thresh=0.2, track_buffer=30, kalman_format="default", img_size=1280,
                 iou_thresh=0.5, reid_model_path="", dhn_path=""):
        self.conf_thresh = conf_thresh
        self.track_buffer = track_buffer
        self.kalman_format = kalman_format
        self.img_size = img_size
        self.iou_thresh = iou_thresh
        self.reid_model_path = reid_model_path
        self.dhn_path = dhn_path


class FixedGMC:
    """Replaces ``tracker.gmc`` so that BoT-SORT receives a prescribed warp per frame -- the same
    role as the reference's own ``method='file'`` path (botsort.py:237-248)."""

    def __init__(self, warps):
        self.warps = list(warps)
        self.k = 0

    def apply(self, raw_frame=None, detections=None):
        h = self.warps[self.k]
        self.k += 1
        return np.asarray(h, dtype=np.float64)


def load_detector_model(cfg_rel="cfg/deploy/yolov7-w6.yaml", fuse=True):
    """Builds the reference's own ``models.yolo.Model`` (CPU, eval, fused unless fuse=False) -- build container only.
    matplotlib / seaborn are stubbed (utils/plots.py:11-15, utils/metrics.py:5 import them, nothing on this
    path uses them)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _stub(name)
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "use"):
        sys.modules["matplotlib"].use = lambda *a, **k: None
        sys.modules["matplotlib"].rc = lambda *a, **k: None
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in ("models", "utils", "models.yolo", "models.common", "models.experimental")}
    for k in list(sys.modules):
        if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils."):
            sys.modules.pop(k)
    sys.path.insert(0, REF_ROOT)
    cwd = os.getcwd()
    try:
        os.chdir(REF_ROOT)
        yolo = importlib.import_module("models.yolo")
        model = yolo.Model(os.path.join(REF_ROOT, cfg_rel), ch=3, nc=80).float().eval()
        if fuse:
            model.fuse()
    finally:
        os.chdir(cwd)
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils."):
                sys.modules.pop(k)
        for k, v in saved_mods.items():
            if v is not None:
                sys.modules[k] = v
    return model


def load_general():
    """The reference's own ``utils.general`` module (``non_max_suppression``, ``scale_coords`` ...), imported from REF_ROOT and
    removed from ``sys.modules`` again so that the product's same-named drop-in package is not shadowed."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _stub(name)
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "use"):
        sys.modules["matplotlib"].use = lambda *a, **k: None
        sys.modules["matplotlib"].rc = lambda *a, **k: None
    saved_path = list(sys.path)
    saved = {k: v for k, v in sys.modules.items() if k == "utils" or k.startswith("utils.") or k == "models" or k.startswith("models.")}
    for k in saved:
        sys.modules.pop(k)
    sys.path.insert(0, REF_ROOT)
    try:
        general = importlib.import_module("utils.general")
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k == "utils" or k.startswith("utils.") or k == "models" or k.startswith("models."):
                sys.modules.pop(k)
        sys.modules.update(saved)
    return general
