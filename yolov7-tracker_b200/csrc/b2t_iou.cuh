// b2t_iou.cuh -- "+1 pixel" IoU cost (replaces cython_bbox.bbox_overlaps as called from
// tracker/matching.py:44-82).  The expression tree is the py-faster-rcnn one restated in
// oracle/iou.py; the translation unit is compiled with --fmad=false so that fp64 results are
// bit-identical to the NumPy evaluation on identical inputs.
#pragma once
#include "b2t_prims.cuh"

namespace b2t {

template <class T> B2T_DEV T t_min(T a, T b) { return a < b ? a : b; }
template <class T> B2T_DEV T t_max(T a, T b) { return a > b ? a : b; }

// a, b: tlbr.  Returns IoU in [0, 1]; 0 when the boxes do not overlap under the +1 convention.
template <class T> B2T_DEV T iou_plus1(const T* a, const T* b) {
    const T iw = t_min(a[2], b[2]) - t_max(a[0], b[0]) + (T)1;
    if (!(iw > (T)0)) return (T)0;
    const T ih = t_min(a[3], b[3]) - t_max(a[1], b[1]) + (T)1;
    if (!(ih > (T)0)) return (T)0;
    const T area_a = (a[2] - a[0] + (T)1) * (a[3] - a[1] + (T)1);
    const T area_b = (b[2] - b[0] + (T)1) * (b[3] - b[1] + (T)1);
    const T inter = iw * ih;
    const T ua = (area_a + area_b) - inter;
    return inter / ua;
}

}  // namespace b2t
