"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT.

CPU (NumPy / SciPy / plain Python) restatement of the reference's
detect -> NMS -> associate hot path.  It exists only to
check the CUDA path:

  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import it;
  * nothing under ``yolov7-tracker_b200/`` imports it, and the product path
    raises if the CUDA library is missing instead of falling back to this.

Pinning status (see DESIGN.md "Oracle"):
  * Kalman / state machine / NMS / decode: pinned against the reference's own
    Python files executed in the build container (``oracle/refshim.py``) --
    fixtures in ``tests/golden`` written by ``tests/golden/make_golden.py``.
  * ``lap.lapjv`` and ``cython_bbox.bbox_overlaps`` are third-party wheels that
    are NOT vendored in the reference and NOT installed anywhere we can reach
    (no version is pinned by the reference either).  ``oracle/lapjv.py`` and
    ``oracle/iou.py`` restate their published algorithms: **parity unpinned**
    at exactly those two call sites (tracker/matching.py:34 and :56).
"""
